"""Host-side data and image helpers with the reference's surface (reference util.py):
``Hdf5Iterator`` / ``iterate_hdf5`` (util.py:10-62: uint8 NHWC arrays -> shuffled float32 NCHW batches, A scaled to
[0,1] or [-1,1], paired augmentation through a shared seed), the dump/sampling helpers (util.py:69-116), and a
synthetic iterator with the same ``.N`` / ``.next()`` surface for benchmarks and tests.  Keras' ImageDataGenerator is
not a dependency: any object with Keras' ``flow(x, None, batch_size=, seed=)`` protocol works as ``imgen``;
``FlipAugmenter`` provides the flips (arbitrary-angle rotation with reflect fill is not reimplemented)."""
import struct
import zlib

import numpy as np


def _get_slices(length, bs):
    """Consecutive batch slices covering [0, length); the last one may be short (reference util.py:10-18)."""
    return [slice(b * bs, (b + 1) * bs) for b in range((length + bs - 1) // bs)]


def iterate_hdf5(imgen=None, is_a_grayscale=True, is_b_grayscale=False, is_uint8=True):
    """Generator factory of the reference (util.py:20-42).  X_arr / y_arr are array-likes indexed by slices (numpy
    arrays or h5py datasets) in NHWC layout; every pass over the data visits the batch slices in an order shuffled by
    `rnd_state` (None: in order); uint8 data are normalised (grayscale: /255, otherwise (x-127.5)/127.5); with an
    `imgen`, X and Y are augmented identically through a shared seed."""
    def _iterate_hdf5(X_arr, y_arr, bs, rnd_state=np.random.RandomState(0)):
        assert X_arr.shape[0] == y_arr.shape[0]
        while True:
            slices = _get_slices(X_arr.shape[0], bs)
            if rnd_state is not None:
                rnd_state.shuffle(slices)
            for elem in slices:
                this_X = np.asarray(X_arr[elem]).astype("float32")
                this_Y = np.asarray(y_arr[elem]).astype("float32")
                if this_X.ndim == 3:
                    this_X = this_X[..., None]
                if this_Y.ndim == 3:
                    this_Y = this_Y[..., None]
                this_X = np.ascontiguousarray(this_X.transpose(0, 3, 1, 2))          # NHWC -> NCHW
                this_Y = np.ascontiguousarray(this_Y.transpose(0, 3, 1, 2))
                if is_uint8:
                    this_X = (this_X / 255.0) if is_a_grayscale else (this_X - 127.5) / 127.5
                    this_Y = (this_Y / 255.0) if is_b_grayscale else (this_Y - 127.5) / 127.5
                if imgen is not None:
                    seed = rnd_state.randint(0, 100000) if rnd_state is not None else 0
                    this_X = next(imgen.flow(this_X, None, batch_size=bs, seed=seed))
                    this_Y = next(imgen.flow(this_Y, None, batch_size=bs, seed=seed))
                yield this_X.astype("float32"), this_Y.astype("float32")
    return _iterate_hdf5


class Hdf5Iterator(object):
    """The reference's iterator object (util.py:45-62): ``.N`` = number of examples, ``.next()`` -> (X, Y)."""

    def __init__(self, X, y, bs, imgen, is_a_grayscale, is_b_grayscale, is_uint8=True):
        assert X.shape[0] == y.shape[0]
        # a private RandomState(0) per iterator: the reference's default argument is one object shared by every
        # iterator of the process, which couples the train and validation shuffles to call order
        self.fn = iterate_hdf5(imgen, is_a_grayscale, is_b_grayscale, is_uint8)(X, y, bs, np.random.RandomState(0))
        self.N = X.shape[0]

    def __iter__(self):
        return self

    def next(self):
        return next(self.fn)

    __next__ = next


class FlipAugmenter(object):
    """Keras-free stand-in for ImageDataGenerator(horizontal_flip=, vertical_flip=) (reference experiments.py:13):
    ``flow(x, None, batch_size=, seed=)`` yields x with every sample flipped or not by a RandomState(seed), so X and Y
    passed with the same seed get the same flips."""

    def __init__(self, horizontal_flip=True, vertical_flip=True):
        self.h, self.v = horizontal_flip, vertical_flip

    def flow(self, x, y=None, batch_size=32, seed=None):
        r = np.random.RandomState(seed)
        while True:
            out = np.array(x, copy=True)
            for i in range(out.shape[0]):
                if self.h and r.rand() < 0.5:
                    out[i] = out[i][:, :, ::-1]
                if self.v and r.rand() < 0.5:
                    out[i] = out[i][:, ::-1, :]
            yield out


def convert_to_rgb(img, is_grayscale=False):
    """CHW image -> HWC in [0,1] with 3 channels; non-grayscale images are mapped
    back from [-1,1] (reference util.py:69-84)."""
    if len(img.shape) != 3:
        raise Exception("Image must have 3 dimensions (channels x height x width). Given {0}".format(len(img.shape)))
    ch = img.shape[0]
    if ch != 3 and ch != 1:
        raise Exception("Unsupported number of channels. Must be 1 or 3, given {0}.".format(ch))
    out = np.repeat(img, 3, axis=0) if ch == 1 else img
    if not is_grayscale:
        out = (out * 127.5 + 127.5) / 255.
    return np.clip(out.transpose((1, 2, 0)), 0, 1)


def compose_imgs(a, b, is_a_grayscale=True, is_b_grayscale=False):
    ap = convert_to_rgb(a, is_grayscale=is_a_grayscale)
    bp = convert_to_rgb(b, is_grayscale=is_b_grayscale)
    if ap.shape != bp.shape:
        raise Exception("A and B must have the same size. {0} != {1}".format(ap.shape, bp.shape))
    return np.concatenate([ap, bp], axis=1)


def imsave(fname, arr):
    """Write an HxWx3 (or HxW) float [0,1] / uint8 array as an 8-bit PNG (stdlib only)."""
    a = np.asarray(arr)
    if a.dtype != np.uint8:
        a = (np.clip(a, 0, 1) * 255.0 + 0.5).astype(np.uint8)
    if a.ndim == 2:
        a = np.repeat(a[:, :, None], 3, axis=2)
    h, w, _ = a.shape
    raw = b"".join(b"\x00" + a[y].tobytes() for y in range(h))

    def chunk(tag, data):
        c = struct.pack(">I", len(data)) + tag + data
        return c + struct.pack(">I", zlib.crc32(tag + data) & 0xffffffff)
    with open(fname, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)) +
                chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))


def plot_grid(out_filename, itr, out_fn, is_a_grayscale, is_b_grayscale, N=4):
    """N x N grid of [A | predict(A)] pairs (reference util.py:101-116, without matplotlib)."""
    rows = []
    for r in range(N):
        row = []
        for c in range(N):
            a, b = itr.next()
            bp = out_fn(a) if out_fn is not None else b
            row.append(compose_imgs(a[0], bp[0], is_a_grayscale=is_a_grayscale, is_b_grayscale=is_b_grayscale))
        rows.append(np.concatenate(row, axis=1))
    imsave(out_filename, np.concatenate(rows, axis=0))


def synthetic_batch(B, latent_dim, S, seed=0):
    """Seeded synthetic (Z, X, Y): Z ~ U[0,1) (reference pix2pix.py:31,206); X a smooth
    heightmap-like field in [0,1]; Y in [-1,1] (the ranges of util.py:31-36)."""
    Z = np.random.RandomState(seed).rand(B, latent_dim).astype(np.float32)
    r = np.random.RandomState(seed + 1)
    lo = max(S // 8, 1)
    base = r.rand(B, 1, lo, lo).astype(np.float32)
    X = np.kron(base, np.ones((1, 1, S // lo, S // lo), np.float32))
    X = 0.75 * X + 0.25 * r.rand(B, 1, S, S).astype(np.float32)
    X = (X ** 2).astype(np.float32)
    basey = r.rand(B, 3, lo, lo).astype(np.float32)
    Y = np.kron(basey, np.ones((1, 1, S // lo, S // lo), np.float32))
    Y = (1.5 * Y + 0.5 * r.rand(B, 3, S, S).astype(np.float32) - 1.0).clip(-1, 1).astype(np.float32)
    return Z, X, Y


class SyntheticIterator(object):
    """Same surface as the reference's Hdf5Iterator: ``.N`` and ``.next()``."""

    def __init__(self, N, batch_size, S, seed=0):
        self.N, self.bs, self.S, self.seed, self.i = N, batch_size, S, seed, 0

    def next(self):
        _, X, Y = synthetic_batch(self.bs, 1, self.S, self.seed + 7 * self.i)
        self.i = (self.i + 1) % max(self.N // self.bs, 1)
        return X, Y

    __next__ = next
