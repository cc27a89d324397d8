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


def normalise_uint8(a, is_grayscale):
    """Raw uint8 NHWC images -> the float32 NCHW batch the reference's iterator yields (util.py:28-35): the host
    restatement of hm_u8_normalize + layout change, used by the dump helpers and as the parity check of the device path."""
    a = np.asarray(a)
    if a.ndim == 3:
        a = a[..., None]
    f = np.ascontiguousarray(a.transpose(0, 3, 1, 2)).astype("float32")
    return ((f / 255.0) if is_grayscale else (f - 127.5) / 127.5).astype("float32")


def as_float_nchw(a, is_grayscale):
    """A batch as the dump helpers need it: float32 NCHW as it is, raw uint8 NHWC normalised on the host."""
    return normalise_uint8(a, is_grayscale) if np.asarray(a).dtype == np.uint8 else a


def iterate_hdf5(imgen=None, is_a_grayscale=True, is_b_grayscale=False, is_uint8=True, device_normalise=False):
    """Generator factory of the reference (util.py:20-42).  X_arr / y_arr are array-likes indexed by slices (numpy
    arrays or h5py datasets) in NHWC layout; every pass over the data visits the batch slices in an order shuffled by
    `rnd_state` (None: in order); uint8 data are normalised (grayscale: /255, otherwise (x-127.5)/127.5); with an
    `imgen`, X and Y are augmented identically through a shared seed.

    device_normalise=True (uint8 data only; not in the reference) yields the RAW uint8 NHWC batches instead -- same
    order, same augmentation, applied to the bytes -- for Pix2Pix.train_fn / loss_fn / gen_fn to normalise on the
    device (hm_u8_normalize): a quarter of the host->device bytes and no float32 pass over the batch on the host."""
    def _iterate_raw(X_arr, y_arr, bs, rnd_state):
        while True:
            slices = _get_slices(X_arr.shape[0], bs)
            if rnd_state is not None:
                rnd_state.shuffle(slices)
            for elem in slices:
                out = []
                seed = None
                for arr in (X_arr, y_arr):
                    a = np.asarray(arr[elem])
                    if a.dtype != np.uint8:
                        raise TypeError("device_normalise needs uint8 data, got %s" % a.dtype)
                    if a.ndim == 3:
                        a = a[..., None]
                    if imgen is not None:
                        if seed is None:
                            seed = rnd_state.randint(0, 100000) if rnd_state is not None else 0
                        a = next(imgen.flow(a.transpose(0, 3, 1, 2), None, batch_size=bs, seed=seed))
                        a = np.asarray(a).transpose(0, 2, 3, 1)
                        if a.dtype != np.uint8:      # an interpolating augmenter: back to the byte grid
                            a = np.clip(np.rint(a), 0, 255).astype(np.uint8)
                    out.append(np.ascontiguousarray(a))
                yield out[0], out[1]

    # Raw-byte batches are exact only under augmenters that MOVE pixels (flips, order-0 rotation: the built-in ones, marked
    # `pixel_moving`).  Any other `imgen` (e.g. a Keras ImageDataGenerator with interpolation, as the reference passes)
    # works on the normalised floats exactly as in the reference: the host float path below.
    raw = device_normalise and (imgen is None or getattr(imgen, "pixel_moving", False))

    def _iterate_hdf5(X_arr, y_arr, bs, rnd_state=np.random.RandomState(0)):
        assert X_arr.shape[0] == y_arr.shape[0]
        if raw:
            if not is_uint8:
                raise ValueError("device_normalise=True is for uint8 data (is_uint8=True)")
            for batch in _iterate_raw(X_arr, y_arr, bs, rnd_state):      # endless
                yield batch
            return
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

    def __init__(self, X, y, bs, imgen, is_a_grayscale, is_b_grayscale, is_uint8=True, device_normalise=False):
        assert X.shape[0] == y.shape[0]
        # a private RandomState(0) per iterator: the reference's default argument is one object shared by every
        # iterator of the process, which couples the train and validation shuffles to call order
        self.fn = iterate_hdf5(imgen, is_a_grayscale, is_b_grayscale, is_uint8, device_normalise)(
            X, y, bs, np.random.RandomState(0))
        self.N = X.shape[0]

    def __iter__(self):
        return self

    def next(self):
        return next(self.fn)

    __next__ = next


class FlipAugmenter(object):
    pixel_moving = True          # exact on raw uint8 data (util.iterate_hdf5, device_normalise)
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


class RotateFlipAugmenter(object):
    pixel_moving = True          # order-0 resampling: exact on raw uint8 data (util.iterate_hdf5, device_normalise)
    """Keras-free restatement of what the reference builds for training (experiments.py:13):
    ``ImageDataGenerator(horizontal_flip=True, vertical_flip=True, rotation_range=360, fill_mode="reflect")`` used through
    ``flow(x, None, batch_size=bs, seed=seed).next()`` on an NCHW batch (util.py:38-40).

    Keras is not vendored in the reference and no version is pinned, so this follows the published Keras 2.0.x
    algorithm [upstream, recalled]: the generator is seeded with `seed` (a private RandomState here; Keras seeds the
    global np.random), the batch is a random permutation of the samples (flow's shuffle=True default), and each sample
    draws theta ~ U(-rotation_range, rotation_range) degrees, is rotated about the image centre (offset H/2+0.5) by
    scipy.ndimage.affine_transform with order 0 (nearest) and the fill mode, then flipped left-right and up-down with
    probability 1/2 each.  X and Y batches passed with the same seed get the same permutation, angle and flips, which
    is all the reference relies on.  Order-0 resampling only moves pixels, so it commutes with the normalisation and
    works on raw uint8 batches as well."""

    def __init__(self, horizontal_flip=True, vertical_flip=True, rotation_range=360., fill_mode="reflect", cval=0.):
        self.h, self.v, self.rot, self.fill, self.cval = horizontal_flip, vertical_flip, rotation_range, fill_mode, cval

    def random_transform(self, x, r):
        """One CHW sample; r: the RandomState all draws come from."""
        from scipy import ndimage
        if self.rot:
            theta = np.pi / 180 * r.uniform(-self.rot, self.rot)
            if theta != 0:
                c, s_ = np.cos(theta), np.sin(theta)
                rot = np.array([[c, -s_, 0], [s_, c, 0], [0, 0, 1]])
                H, W = x.shape[1], x.shape[2]
                oy, ox = float(H) / 2 + 0.5, float(W) / 2 + 0.5
                off = np.array([[1, 0, oy], [0, 1, ox], [0, 0, 1]])
                back = np.array([[1, 0, -oy], [0, 1, -ox], [0, 0, 1]])
                m = off.dot(rot).dot(back)
                x = np.stack([ndimage.affine_transform(ch, m[:2, :2], m[:2, 2], order=0, mode=self.fill,
                                                       cval=self.cval) for ch in x], axis=0)
        if self.h and r.random_sample() < 0.5:
            x = x[:, :, ::-1]
        if self.v and r.random_sample() < 0.5:
            x = x[:, ::-1, :]
        return x

    def flow(self, x, y=None, batch_size=32, seed=None):
        x = np.asarray(x)
        n = x.shape[0]
        seen = 0
        while True:
            r = np.random.RandomState(None if seed is None else seed + seen)
            order = r.permutation(n)[:batch_size]
            yield np.stack([self.random_transform(x[j], r) for j in order], axis=0).astype(x.dtype)
            seen += 1


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
            bp = out_fn(a) if out_fn is not None else as_float_nchw(b, is_b_grayscale)
            a = as_float_nchw(a, is_a_grayscale)
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
