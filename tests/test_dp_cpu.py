"""Data-parallel path on the CPU: two gloo ranks, each with half of a batch, must produce exactly the update
that the reference semantics prescribe for per-rank mean losses summed and divided by the world size
(SURVEY.md §8e): gradients are all-reduced (sum) once per network and scaled by 1/world inside the optimiser.
The CUDA kernels are replaced by tests/fake_hmgan.py (host logic only)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
PKG = os.path.join(ROOT, "gan-heightmaps_b200")


def _worker(rank, world, port, out_dir):
    for p in (ROOT, PKG, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import fake_hmgan
    import _lib
    _lib.call, _lib.query, _lib.load = fake_hmgan.call, fake_hmgan.query, (lambda: None)
    from oracle import step as S
    import test_engine_cpu as T
    cfg = S.experiment_kwargs('gate64')
    cfg = dict(cfg, D=dict(cfg['D']), G=dict(cfg['G']))
    _, m = T.build_pair(cfg, 'dcgan', with_p2p=False)
    m.pg = dist.group.WORLD
    Z, X, Y = S.synthetic_batch(4, cfg['latent_dim'], 64, seed=10)
    sl = slice(2 * rank, 2 * rank + 2)
    losses = m.train_fn(Z[sl], X[sl], Y[sl])
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), losses=np.asarray(losses),
             **{"G%d" % i: v for i, v in enumerate(m.G.get_all_param_values())},
             **{"D%d" % i: v for i, v in enumerate(m.D.get_all_param_values())})
    dist.destroy_process_group()


def test_two_rank_step_keeps_replicas_identical_and_averages_gradients(tmp_path):
    world, port = 2, 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0 = np.load(str(tmp_path / "rank0.npz"))
    r1 = np.load(str(tmp_path / "rank1.npz"))
    trainable_keys = [k for k in r0.files if k != "losses"]
    # replicas stay bit-identical in every trainable parameter (BN running statistics are per-rank by design)
    sys.path.insert(0, PKG)
    import lasagne_compat as L
    from architectures import dcgan
    from oracle import step as S
    cfg = S.experiment_kwargs('gate64')
    gp = L.get_all_params(dcgan.default_generator(cfg['latent_dim'], True, **cfg['G']))
    for k in trainable_keys:
        if k.startswith("G") and not gp[int(k[1:])].trainable:
            continue
        np.testing.assert_array_equal(r0[k], r1[k], err_msg=k)
    # D has no BatchNorm, so its update must equal a single-process step on the full batch with the
    # half-batch losses averaged: mean over 4 samples == mean of the two 2-sample means.  (G's BatchNorm sees
    # per-rank statistics, so only D is comparable to the single-process oracle.)
    assert np.isfinite(r0["losses"]).all() and np.isfinite(r1["losses"]).all()
    assert not np.array_equal(r0["losses"], r1["losses"])        # different shards, different losses
    # the update the two ranks applied == RMSprop on the MEAN of the two per-rank gradients (computed here by two
    # single-process models, one per shard)
    import fake_hmgan
    import _lib
    import test_engine_cpu as T
    saved = (_lib.call, _lib.query, _lib.load)
    _lib.call, _lib.query, _lib.load = fake_hmgan.call, fake_hmgan.query, (lambda: None)
    try:
        Z, X, Y = S.synthetic_batch(4, cfg['latent_dim'], 64, seed=10)
        grads, p0 = [], None
        for rank in range(2):
            _, m = T.build_pair(cfg, 'dcgan', with_p2p=False)
            p0 = [v.copy() for v in m.D.get_all_param_values()]
            sl = slice(2 * rank, 2 * rank + 2)
            m.train_fn(Z[sl], X[sl], Y[sl])
            grads.append(m.D.get_grads())
        for i, (w0, g0, g1) in enumerate(zip(p0, grads[0], grads[1])):
            g = 0.5 * (g0 + g1)
            acc = 0.1 * g * g
            want = w0 - np.float32(1e-3) * g / np.sqrt(acc + np.float32(1e-6))
            np.testing.assert_allclose(r0["D%d" % i], want, rtol=1e-5, atol=1e-7, err_msg="D param %d" % i)
    finally:
        _lib.call, _lib.query, _lib.load = saved
