"""The real entry point on the GPU: the experiment closure's model (experiments._model, the keyword set of
test1_nobn_bilin_both, reference experiments.py:98-119) trained by Pix2Pix.train (reference pix2pix.py:187-275) for one
epoch on synthetic 512x512 pairs in fp16 fast mode, its results.txt row against the oracle's loop; the checkpoint it
writes round-trips through load_model(mode='p2p'); and the CLI itself (`python experiments.py <name> train`)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import step as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "gan-heightmaps_b200")
if PKG not in sys.path:
    sys.path.insert(0, PKG)

pytestmark = pytest.mark.gpu

KEYS = ['dcgan_gen', 'dcgan_disc', 'p2p_gen', 'p2p_recon', 'p2p_disc']


def test_experiment_model_trains_one_epoch_like_the_oracle_loop(tmp_path, monkeypatch):
    monkeypatch.setenv("HMGAN_PRECISION", "fast")
    import experiments
    import util
    from lasagne_compat import tanh, floatX
    model = experiments._model({'nf': 64, 'act': tanh, 'num_repeats': 0, 'bilinear_upsample': True}, train_mode='both')
    assert model.rt.precision == "fast" and model.rt.device.type == "cuda"
    cfg = S.experiment_kwargs('test1_nobn_bilin_both')
    om = S.OracleModel(S.build_nets(cfg, seed=2), alpha=100., opt='rmsprop', lr=1e-4, train_mode='both', lsgan=True)
    with torch.no_grad():
        om.params['D'][-1].fill_(0.6)            # live discriminator head (see test_step_gpu._set_head_bias)
    for k, net in (('G', model.G), ('D', model.D), ('P', model.P), ('Dp', model.Dp)):
        net.set_all_param_values(om.get_all_param_values(k))
    bs, n = 4, 8
    out_dir, model_dir = str(tmp_path / "out"), str(tmp_path / "models")
    np.random.seed(1)
    model.train(util.SyntheticIterator(n, bs, 512, 0), util.SyntheticIterator(n, bs, 512, 1000), batch_size=bs,
                num_epochs=1, out_dir=out_dir, model_dir=model_dir, save_every=1)
    lines = open(os.path.join(out_dir, "results.txt")).read().strip().splitlines()
    assert lines[0].split(",") == (["epoch"] + ["train_" + k for k in KEYS] + ["valid_" + k for k in KEYS]
                                   + ["lr", "time", "mode"])
    row = lines[1].split(",")
    assert row[0] == "1" and row[-1] == "both" and abs(float(row[11]) - 1e-4) < 1e-9
    ours = np.array([float(v) for v in row[1:11]])
    # the oracle's loop: same batches, same latent draws (the loop draws Z from np.random after each iterator read, and
    # the validation pass reads it_train too, reference pix2pix.py:204)
    np.random.seed(1)
    it = util.SyntheticIterator(n, bs, 512, 0)
    ref = []
    for fn in (om.train_fn, om.loss_fn):
        rec = []
        for _ in range(n // bs):
            X, Y = it.next()
            Z = floatX(np.random.rand(X.shape[0], cfg['latent_dim']))
            rec.append(fn(Z, X, Y))
        ref += list(np.mean(np.array(rec), axis=0))
    # fp16 fast mode over two updates at lr 1e-4: 2e-2 relative (measured <= 3e-3)
    np.testing.assert_allclose(ours, np.array(ref), rtol=2e-2, atol=1e-4)
    # per-epoch dumps and the checkpoint
    assert os.path.exists(os.path.join(out_dir, "out_1.png")) and os.path.exists(os.path.join(out_dir, "dump_a", "0.png"))
    ckpt = os.path.join(model_dir, "1.model")
    assert os.path.exists(ckpt)
    Xv = util.SyntheticIterator(n, 1, 512, 5).next()[0]
    before = model.gen_fn_det(Xv)
    saved_p, saved_g = model.P.get_all_param_values(), model.G.get_all_param_values()
    model.P.set_all_param_values([np.zeros_like(a) for a in saved_p])
    model.G.set_all_param_values([a + 1 for a in saved_g])
    model.load_model(ckpt, mode='p2p')                     # restores P / Dp only
    for a, b in zip(model.P.get_all_param_values(), saved_p):
        np.testing.assert_array_equal(a, b)
    for a, b in zip(model.G.get_all_param_values(), saved_g):
        np.testing.assert_array_equal(a, b + 1)
    np.testing.assert_allclose(model.gen_fn_det(Xv), before, atol=1e-6)


def test_cli_trains_and_writes_results(tmp_path):
    """`python experiments.py test1_nobn_bilin_both train` on 4 synthetic pairs, one epoch, from a scratch directory."""
    env = dict(os.environ, HMGAN_SYNTHETIC="4", HMGAN_EPOCHS="1", HMGAN_PRECISION="fast", PYTHONPATH=PKG)
    out = subprocess.run([sys.executable, os.path.join(PKG, "experiments.py"), "test1_nobn_bilin_both", "train"],
                         cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-3000:]
    res = tmp_path / "output" / "test1_nobn_bilin_both_deleteme" / "results.txt"
    lines = res.read_text().strip().splitlines()
    assert len(lines) == 2 and lines[0].startswith("epoch,train_dcgan_gen")
    vals = [float(v) for v in lines[1].split(",")[1:11]]
    assert all(np.isfinite(vals)) and lines[1].endswith(",both")
