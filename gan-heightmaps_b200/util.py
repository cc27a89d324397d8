"""Host-side image helpers used by the trainer's dump/sampling methods
(reference util.py:69-116) and a synthetic stand-in for Hdf5Iterator
(reference util.py:45-62): objects with ``.N`` and ``.next()`` returning
``(X, Y)`` float32 NCHW batches.  The HDF5 data path itself is out of scope for
this round."""
import struct
import zlib

import numpy as np


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
