"""The host side of the trainer around the hot path (reference pix2pix.py:158-425), on the CPU emulation of the kernel
library: epoch loop and results.txt schema (:213-262), per-epoch dumps (:263-273), checkpoints and resume (:158-186,
:234-241), generate_gz / generate_atob / generate_interpolation / generate_interpolation_clip (:276-425)."""
import os
import struct
import sys
import zlib

import numpy as np
import pytest

import fake_hmgan
from oracle import step as S

PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gan-heightmaps_b200")
if PKG not in sys.path:
    sys.path.insert(0, PKG)
import _lib                     # noqa: E402
import util                     # noqa: E402
from test_engine_cpu import build_pair   # noqa: E402


@pytest.fixture
def cpu_backend(monkeypatch):
    monkeypatch.setattr(_lib, "call", fake_hmgan.call)
    monkeypatch.setattr(_lib, "query", fake_hmgan.query)
    monkeypatch.setattr(_lib, "load", lambda: None)


def _png_size(path):
    """(width, height, channels) of a PNG written by util.imsave; checks the signature and the IDAT stream."""
    raw = open(path, "rb").read()
    assert raw[:8] == b"\x89PNG\r\n\x1a\n"
    w, h, depth, ctype = struct.unpack(">IIBB", raw[16:26])
    assert depth == 8 and ctype in (0, 2)
    i, idat = 8, b""
    while i < len(raw):
        n, tag = struct.unpack(">I4s", raw[i:i + 8])
        if tag == b"IDAT":
            idat += raw[i + 8:i + 8 + n]
        i += 12 + n
    ch = 3 if ctype == 2 else 1
    assert len(zlib.decompress(idat)) == h * (1 + w * ch)
    return w, h, ch


def test_train_loop_writes_results_dumps_and_checkpoints(cpu_backend, tmp_path):
    cfg = S.experiment_kwargs('tiny512')
    _, m = build_pair(cfg, 'both')
    it_train = util.SyntheticIterator(2, 1, 512, seed=0)
    it_val = util.SyntheticIterator(2, 1, 512, seed=100)
    out_dir, model_dir = str(tmp_path / "out"), str(tmp_path / "models")
    p_before = m.P.get_all_param_values()
    m.train(it_train, it_val, batch_size=1, num_epochs=1, out_dir=out_dir, model_dir=model_dir, save_every=1,
            quick_run=True)
    lines = open(os.path.join(out_dir, "results.txt")).read().strip().splitlines()
    header = lines[0].split(",")
    keys = ['dcgan_gen', 'dcgan_disc', 'p2p_gen', 'p2p_recon', 'p2p_disc']
    assert header == ["epoch"] + ["train_" + k for k in keys] + ["valid_" + k for k in keys] + ["lr", "time", "mode"]
    row = lines[1].split(",")
    assert len(lines) == 2 and len(row) == len(header) and row[0] == "1" and row[-1] == "both"
    assert all(np.isfinite(float(v)) for v in row[1:-1]) and abs(float(row[11]) - 1e-3) < 1e-9
    # the training epoch moved the parameters, the validation pass (loss_fn) did not undo it
    assert any(np.abs(a - b).max() > 0 for a, b in zip(p_before, m.P.get_all_param_values()))
    # per-epoch dumps: the A|B grid, two A->B dumps, 20 generated heightmaps
    assert _png_size(os.path.join(out_dir, "out_1.png")) == (4 * 2 * 512, 4 * 512, 3)
    for d in ("dump_train", "dump_valid"):
        assert _png_size(os.path.join(out_dir, d, "0.a.png"))[:2] == (512, 512)
        assert _png_size(os.path.join(out_dir, d, "0.b.png")) == (512, 512, 3)
    assert len(os.listdir(os.path.join(out_dir, "dump_a"))) == 20
    # checkpoint of epoch 1 restores every network; a partial load touches only its half
    ck = os.path.join(model_dir, "1.model")
    assert os.path.exists(ck)
    saved = {k: n.get_all_param_values() for k, n in (("G", m.G), ("D", m.D), ("P", m.P), ("Dp", m.Dp))}
    _, m2 = build_pair(cfg, 'both', seed=5)
    g_other = m2.G.get_all_param_values()
    m2.load_model(ck, mode='p2p')
    for a, b in zip(m2.P.get_all_param_values(), saved["P"]):
        np.testing.assert_array_equal(a, b)
    for a, b in zip(m2.G.get_all_param_values(), g_other):
        np.testing.assert_array_equal(a, b)
    m2.load_model(ck)
    for k, n in (("G", m2.G), ("D", m2.D), ("P", m2.P), ("Dp", m2.Dp)):
        for a, b in zip(n.get_all_param_values(), saved[k]):
            np.testing.assert_array_equal(a, b)
    # resume: appends to results.txt after loading the given checkpoint (pix2pix.py:234-241)
    m2.train(it_train, it_val, batch_size=1, num_epochs=1, out_dir=out_dir, model_dir=None, resume=ck, quick_run=True)
    assert len(open(os.path.join(out_dir, "results.txt")).read().strip().splitlines()) == 3


def test_sampling_entry_points(cpu_backend, tmp_path):
    cfg = S.experiment_kwargs('tiny512')
    _, m = build_pair(cfg, 'both')
    m.sampler = np.random.RandomState(3).rand
    d = str(tmp_path / "gz")
    m.generate_gz(num_examples=4, batch_size=2, out_dir=d, deterministic=True)
    assert sorted(os.listdir(d)) == ["0.png", "1.png", "2.png", "3.png"] and _png_size(d + "/0.png")[:2] == (512, 512)
    it = util.SyntheticIterator(2, 1, 512, seed=1)
    d = str(tmp_path / "atob")
    m.generate_atob(it, 2, d, deterministic=True)
    assert len(os.listdir(d)) == 4
    m.generate_atob(it, 1, str(tmp_path / "gt"), dont_predict=True)          # ground truth instead of P(X)
    assert _png_size(str(tmp_path / "gt" / "0.b.png")) == (512, 512, 3)
    out = str(tmp_path / "row.png")
    m.generate_interpolation(out, mode='row')
    assert _png_size(out) == (6 * 512, 512, 3)
    d = str(tmp_path / "clip")
    m.generate_interpolation_clip(num_samples=2, batch_size=5, out_dir=d, deterministic=True, concat=True)
    files = sorted(os.listdir(d))
    assert len(files) == 25 and files[0] == "concat_0000.png" and _png_size(os.path.join(d, files[0])) == (1024, 512, 3)
    # deterministic sampling uses the running statistics and leaves them alone; the non-deterministic one moves them
    stats = [a.copy() for a in m.G.get_all_param_values()]
    Z = np.random.RandomState(0).rand(2, cfg['latent_dim']).astype(np.float32)
    a = m.z_fn_det(Z)
    for x, y in zip(stats, m.G.get_all_param_values()):
        np.testing.assert_array_equal(x, y)
    b = m.z_fn(Z)
    assert a.shape == b.shape == (2, 1, 512, 512)
    assert any(np.abs(x - y).max() > 0 for x, y in zip(stats, m.G.get_all_param_values()))


def test_loss_sync_every_n_steps_gives_the_same_epoch_row(cpu_backend, tmp_path):
    """Pix2Pix.train(loss_sync_every=N) (the loop only needs the epoch means) reads the losses back in
    groups through train_fn_async / loss_fn_async: the results.txt row equals the per-step read-back's; a model without
    the pix2pix stage (gen_fn_p2p=None) saves a checkpoint that loads again with the default mode."""
    cfg = S.experiment_kwargs('gate64')
    rows = []
    for n in (1, 3):
        _, m = build_pair(cfg, 'dcgan', with_p2p=False)
        np.random.seed(3)
        out_dir = str(tmp_path / ("out%d" % n))
        m.train(util.SyntheticIterator(4, 1, 64, seed=0), util.SyntheticIterator(4, 1, 64, seed=50), batch_size=1,
                num_epochs=1, out_dir=out_dir, model_dir=str(tmp_path / ("m%d" % n)), save_every=1, loss_sync_every=n)
        rows.append([float(v) for v in open(os.path.join(out_dir, "results.txt")).read().strip().splitlines()[1]
                     .split(",")[1:11]])
    np.testing.assert_allclose(rows[1], rows[0], rtol=1e-6, atol=0)
    m.load_model(str(tmp_path / "m3" / "1.model"))             # DCGAN-only model: the p2p lists of the file are empty
    with pytest.raises(ValueError):
        _, both = build_pair(S.experiment_kwargs('tiny512'), 'both')
        m.save_model(str(tmp_path / "dc.model"))
        both.save_model(str(tmp_path / "both.model"))
        m.load_model(str(tmp_path / "both.model"))              # holds p2p parameters this model has no network for


def test_train_warns_about_non_finite_losses_in_fast_mode(cpu_backend, tmp_path):
    """fp16 storage with a static loss scale: gradients that overflow turn the parameters (and then the losses) into
    inf / NaN, where the float32 reference would carry on.  Pix2Pix.train() must say so instead of writing rows of nan
    silently; the float32 modes keep the reference's behaviour (no warning)."""
    import warnings
    cfg = S.experiment_kwargs('tiny512')
    for precision, expect in (("fast", True), ("parity", False)):
        _, m = build_pair(cfg, 'both', precision=precision)
        nan_row = [np.float32("nan")] * 5
        m.train_fn = lambda Z, X, Y: nan_row
        m.loss_fn = lambda Z, X, Y: nan_row
        it_train = util.SyntheticIterator(2, 1, 512, seed=0)
        it_val = util.SyntheticIterator(2, 1, 512, seed=100)
        with warnings.catch_warnings(record=True) as caught:
            warnings.simplefilter("always")
            m.train(it_train, it_val, batch_size=1, num_epochs=1, out_dir=str(tmp_path / precision), quick_run=True)
        hits = [w for w in caught if issubclass(w.category, RuntimeWarning) and "non-finite losses" in str(w.message)]
        assert bool(hits) == expect, (precision, [str(w.message) for w in caught])
        if expect:
            assert "tc32" in str(hits[0].message)
