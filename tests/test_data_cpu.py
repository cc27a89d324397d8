"""The data iterator of the reference (util.py:10-62) restated in gan-heightmaps_b200/util.py: slicing, per-pass
shuffling with RandomState(0), uint8 NHWC -> float32 NCHW normalisation, paired augmentation through a shared seed."""
import os
import sys

import numpy as np

PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gan-heightmaps_b200")
if PKG not in sys.path:
    sys.path.insert(0, PKG)
import util   # noqa: E402


def _data(n=10, s=8):
    r = np.random.RandomState(1)
    return r.randint(0, 256, (n, s, s, 1)).astype(np.uint8), r.randint(0, 256, (n, s, s, 3)).astype(np.uint8)


def test_slices_cover_everything_with_a_short_tail():
    sl = util._get_slices(10, 4)
    assert [(s.start, s.stop) for s in sl] == [(0, 4), (4, 8), (8, 12)]
    assert util._get_slices(0, 4) == []


def test_normalisation_layout_and_shuffle_follow_the_reference():
    X, Y = _data()
    it = util.Hdf5Iterator(X, Y, 4, None, is_a_grayscale=True, is_b_grayscale=False)
    assert it.N == 10
    # the visiting order of one pass: the reference shuffles the slice list with RandomState(0)
    order = util._get_slices(10, 4)
    np.random.RandomState(0).shuffle(order)
    seen = []
    for sl in order:
        x, y = it.next()
        assert x.dtype == np.float32 and y.dtype == np.float32
        assert x.shape == (len(range(*sl.indices(10))), 1, 8, 8) and y.shape[1:] == (3, 8, 8)
        np.testing.assert_allclose(x, X[sl].transpose(0, 3, 1, 2).astype(np.float32) / 255.0, rtol=0, atol=1e-7)
        np.testing.assert_allclose(y, (Y[sl].transpose(0, 3, 1, 2).astype(np.float32) - 127.5) / 127.5, rtol=0, atol=1e-6)
        assert 0.0 <= x.min() and x.max() <= 1.0 and -1.0 <= y.min() and y.max() <= 1.0
        seen.append(sl.start)
    assert sorted(seen) == [0, 4, 8]
    x, _ = it.next()                       # the next pass starts with a fresh shuffle of the same generator
    assert x.shape[0] in (2, 4)


def test_grayscale_b_and_float_inputs():
    X, Y = _data()
    it = util.Hdf5Iterator(X, Y[..., :1], 5, None, is_a_grayscale=False, is_b_grayscale=True)
    x, y = it.next()
    assert -1.0 <= x.min() and x.max() <= 1.0 and 0.0 <= y.min() and y.max() <= 1.0
    itf = util.Hdf5Iterator(X.astype(np.float32), Y.astype(np.float32), 5, None, True, False, is_uint8=False)
    x, y = itf.next()
    assert x.max() > 1.5                    # untouched


def test_paired_augmentation_uses_one_seed_for_both_images():
    X, Y = _data(8, 6)
    Y3 = np.repeat(X, 3, axis=3)           # Y is a copy of X: after augmentation they must still coincide
    it = util.Hdf5Iterator(X, Y3, 4, util.FlipAugmenter(True, True), is_a_grayscale=True, is_b_grayscale=True)
    flipped = False
    for _ in range(6):
        x, y = it.next()
        np.testing.assert_allclose(np.repeat(x, 3, axis=1), y, atol=1e-7)
        flipped = True
    assert flipped


def test_synthetic_iterator_surface():
    it = util.SyntheticIterator(8, 2, 16, seed=3)
    x, y = it.next()
    assert it.N == 8 and x.shape == (2, 1, 16, 16) and y.shape == (2, 3, 16, 16)
    assert x.dtype == np.float32 and 0 <= x.min() and x.max() <= 1 and -1 <= y.min() and y.max() <= 1


def test_raw_uint8_iterator_yields_the_bytes_behind_the_float_batches():
    """device_normalise=True: same visiting order and the same paired augmentation, but raw uint8 NHWC batches whose
    host normalisation (util.normalise_uint8 = reference util.py:28-35) reproduces the float iterator bit for bit."""
    X, Y = _data(10, 8)
    for imgen in (None, util.FlipAugmenter(True, True), util.RotateFlipAugmenter(True, True, 360, "reflect")):
        itf = util.Hdf5Iterator(X, Y, 4, imgen, is_a_grayscale=True, is_b_grayscale=False)
        itr = util.Hdf5Iterator(X, Y, 4, imgen, is_a_grayscale=True, is_b_grayscale=False, device_normalise=True)
        for _ in range(5):
            xf, yf = itf.next()
            xr, yr = itr.next()
            assert xr.dtype == np.uint8 and yr.dtype == np.uint8 and xr.flags.c_contiguous and yr.flags.c_contiguous
            assert xr.shape == (xf.shape[0], 8, 8, 1) and yr.shape == (yf.shape[0], 8, 8, 3)
            np.testing.assert_array_equal(util.normalise_uint8(xr, True), xf)
            np.testing.assert_array_equal(util.normalise_uint8(yr, False), yf)
            assert util.as_float_nchw(xf, True) is xf


def test_raw_iterator_rejects_float_data():
    X, Y = _data(4, 4)
    it = util.Hdf5Iterator(X.astype(np.float32), Y, 2, None, True, False, device_normalise=True)
    try:
        it.next()
    except TypeError:
        return
    raise AssertionError("float data accepted by the raw uint8 iterator")


def test_rotate_flip_augmenter_follows_the_keras_protocol():
    """Same seed -> same permutation, angle and flips for X and Y; a rotation by a multiple of 90 degrees about the
    Keras centre (H/2+0.5) with reflect fill is a pure pixel move for the interior; no rotation and no flips is the
    identity up to the batch permutation."""
    X, _ = _data(6, 9)
    x = X.transpose(0, 3, 1, 2)
    y = np.repeat(x, 3, axis=1)
    aug = util.RotateFlipAugmenter(True, True, rotation_range=360, fill_mode="reflect")
    for seed in (0, 7, 123):
        a = next(aug.flow(x, None, batch_size=6, seed=seed))
        b = next(aug.flow(y, None, batch_size=6, seed=seed))
        assert a.shape == x.shape and a.dtype == x.dtype
        np.testing.assert_array_equal(np.repeat(a, 3, axis=1), b)
    ident = util.RotateFlipAugmenter(False, False, rotation_range=0)
    out = next(ident.flow(x, None, batch_size=6, seed=3))
    perm = np.random.RandomState(3).permutation(6)
    np.testing.assert_array_equal(out, x[perm])
    # every output pixel of a rotated sample is one of the sample's own pixel values (order-0 resampling + reflect)
    rot_only = util.RotateFlipAugmenter(False, False, rotation_range=360, fill_mode="reflect")
    r = np.random.RandomState(5)
    s = rot_only.random_transform(x[0], r)
    assert set(np.unique(s)) <= set(np.unique(x[0]))
    # batches smaller than batch_size (the last slice of a pass) come back whole
    assert next(aug.flow(x[:2], None, batch_size=4, seed=1)).shape[0] == 2


def test_get_iterators_opens_the_hdf5_layout_of_the_reference(monkeypatch):
    """experiments.get_iterators (reference experiments.py:10-18): datasets xt/yt/xv/yv of one HDF5 file, wrapped in
    Hdf5Iterator with the training augmentation.  h5py is not installed here, so a stand-in module serves numpy arrays
    under the same File(...)[name] protocol."""
    import types
    X, Y = _data(6, 8)
    store = {"xt": X, "yt": Y, "xv": X[:4], "yv": Y[:4]}
    fake = types.ModuleType("h5py")
    fake.File = lambda path, mode="r": store
    monkeypatch.setitem(sys.modules, "h5py", fake)
    monkeypatch.delenv("HMGAN_SYNTHETIC", raising=False)
    import torch  # noqa: F401  (experiments imports the trainer)
    import experiments
    for raw, dtype in (("1", np.uint8), ("0", np.float32)):
        monkeypatch.setenv("HMGAN_DEVICE_NORMALISE", raw)
        it_train, it_val = experiments.get_iterators("whatever.h5", 2, True, False, da=True)
        assert it_train.N == 6 and it_val.N == 4
        x, y = it_train.next()
        assert x.dtype == dtype and y.dtype == dtype and x.shape[0] == 2
        assert (x.shape[1:] == (8, 8, 1) and y.shape[1:] == (8, 8, 3)) if raw == "1" else \
               (x.shape[1:] == (1, 8, 8) and y.shape[1:] == (3, 8, 8))
    monkeypatch.setenv("HMGAN_DEVICE_NORMALISE", "1")
    it_train, _ = experiments.get_iterators("whatever.h5", 2, True, False, da=False)
    x, _ = it_train.next()
    order = util._get_slices(6, 2)
    np.random.RandomState(0).shuffle(order)
    np.testing.assert_array_equal(x, X[order[0]])               # no augmentation: the raw bytes of the first slice
