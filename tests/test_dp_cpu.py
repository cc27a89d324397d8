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


def _worker_syncbn(rank, world, port, out_dir):
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
    om, m = T.build_pair(cfg, 'dcgan', with_p2p=False)
    m.pg = dist.group.WORLD
    m.rt.sync_bn_group = dist.group.WORLD            # what Pix2Pix(..., process_group=pg, sync_bn=True) sets
    Z, X, Y = S.synthetic_batch(4, cfg['latent_dim'], 64, seed=11)
    sl = slice(2 * rank, 2 * rank + 2)
    losses = m.train_fn(Z[sl], X[sl], Y[sl])
    out = {"losses": np.asarray(losses)}
    out.update({"gG%d" % i: v / world for i, v in enumerate(m.G.get_grads())})       # all-reduced sums -> means
    out.update({"gD%d" % i: v / world for i, v in enumerate(m.D.get_grads())})
    out.update({"sG%d" % i: v for i, v in enumerate(m.G.get_all_param_values())})
    if rank == 0:                                    # the single-process reference on the WHOLE batch
        lo = om.train_fn(Z, X, Y)
        out["olosses"] = np.asarray(lo)
        out.update({"oG%d" % i: v for i, v in enumerate(om.last_grads['G'])})
        out.update({"oD%d" % i: v for i, v in enumerate(om.last_grads['D'])})
        out.update({"osG%d" % i: v for i, v in enumerate(om.get_all_param_values('G'))})
    np.savez(os.path.join(out_dir, "sync%d.npz" % rank), **out)
    dist.destroy_process_group()


def test_sync_batchnorm_makes_two_ranks_equal_one_rank_on_the_whole_batch(tmp_path):
    """SyncBN: with BatchNorm statistics and backward reductions averaged over the process group, two
    ranks holding half of a batch each reproduce the single-process step on the whole batch: every gradient of G
    (which has BatchNorm after every layer) and D, the mean of the per-rank losses, and G's updated parameters
    including the BatchNorm running statistics (identical on both ranks)."""
    world, port = 2, 31500 + os.getpid() % 2000
    mp.spawn(_worker_syncbn, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0 = np.load(str(tmp_path / "sync0.npz"))
    r1 = np.load(str(tmp_path / "sync1.npz"))
    np.testing.assert_allclose(0.5 * (r0["losses"][:2] + r1["losses"][:2]), r0["olosses"][:2], rtol=1e-4, atol=1e-6)
    nG = len([k for k in r0.files if k.startswith("oG")])
    nD = len([k for k in r0.files if k.startswith("oD")])
    for net, n in (("G", nG), ("D", nD)):
        for i in range(n):
            a, b = r0["g%s%d" % (net, i)], r0["o%s%d" % (net, i)]
            np.testing.assert_array_equal(a, r1["g%s%d" % (net, i)])
            scale = float(np.abs(b).max()) + 1e-12
            assert float(np.abs(a - b).max()) <= 2e-3 * scale + 1e-7, (net, i, float(np.abs(a - b).max()), scale)
    nS = len([k for k in r0.files if k.startswith("osG")])
    for i in range(nS):          # parameters after the update, BatchNorm running mean / inv_std included
        np.testing.assert_array_equal(r0["sG%d" % i], r1["sG%d" % i])
        np.testing.assert_allclose(r0["sG%d" % i], r0["osG%d" % i], rtol=2e-3, atol=2e-4, err_msg="G value %d" % i)


def _worker_p2p(rank, world, port, out_dir):
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
    cfg = dict(T.TINY)
    _, m = T.build_pair(cfg, 'p2p', with_dcgan=False)
    m.pg = dist.group.WORLD
    Z, X, Y = S.synthetic_batch(4, cfg['latent_dim'], 512, seed=12)
    sl = slice(2 * rank, 2 * rank + 2)
    losses = m.train_fn(Z[sl], X[sl], Y[sl])
    out = {"losses": np.asarray(losses), "buckets": np.asarray(m.allreduce_calls)}
    out.update({"gP%d" % i: v for i, v in enumerate(m.P.get_grads())})               # all-reduced SUMS
    out.update({"gDp%d" % i: v for i, v in enumerate(m.Dp.get_grads())})
    out.update({"P%d" % i: v for i, v in enumerate(m.P.get_all_param_values())})
    out.update({"Dp%d" % i: v for i, v in enumerate(m.Dp.get_all_param_values())})
    np.savez(os.path.join(out_dir, "p2p%d.npz" % rank), **out)
    dist.destroy_process_group()


def test_two_rank_pix2pix_step_reduces_every_gradient_exactly_once(tmp_path):
    """The pix2pix networks' gradients are all-reduced in buckets issued during the backward pass (the PatchGAN's after
    its weight-gradient pass, the U-Net's decoder half while the encoder's backward pass runs, the rest at the end).
    After the step every gradient array on both ranks must be the SUM of the two shards' local gradients (computed here
    by two single-process models) -- a bucket that was skipped would hold one shard's gradient, one reduced twice four
    shards' worth -- and the trainable parameters of the replicas must be bit-identical."""
    world, port = 2, 33500 + os.getpid() % 2000
    mp.spawn(_worker_p2p, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0 = np.load(str(tmp_path / "p2p0.npz"))
    r1 = np.load(str(tmp_path / "p2p1.npz"))
    for k in r0.files:
        if k.startswith("g"):
            np.testing.assert_array_equal(r0[k], r1[k], err_msg=k)
    assert int(r0["buckets"]) == 3 and int(r1["buckets"]) == 3       # PatchGAN, U-Net decoder half, U-Net rest
    sys.path.insert(0, PKG)
    import fake_hmgan
    import _lib
    import test_engine_cpu as T
    from oracle import step as S
    saved = (_lib.call, _lib.query, _lib.load)
    _lib.call, _lib.query, _lib.load = fake_hmgan.call, fake_hmgan.query, (lambda: None)
    try:
        cfg = dict(T.TINY)
        Z, X, Y = S.synthetic_batch(4, cfg['latent_dim'], 512, seed=12)
        local = []
        for rank in range(2):
            _, m = T.build_pair(cfg, 'p2p', with_dcgan=False)
            trainable = {"P": [q.trainable for q in m.P.params], "Dp": [q.trainable for q in m.Dp.params]}
            sl = slice(2 * rank, 2 * rank + 2)
            m.train_fn(Z[sl], X[sl], Y[sl])
            local.append({"P": m.P.get_grads(), "Dp": m.Dp.get_grads()})
        for net in ("P", "Dp"):
            assert len(local[0][net]) > 4
            for i, (g0, g1) in enumerate(zip(local[0][net], local[1][net])):
                want = g0 + g1
                scale = max(float(np.abs(want).max()), 1e-12)
                np.testing.assert_allclose(r0["g%s%d" % (net, i)], want, rtol=1e-5, atol=1e-6 * scale,
                                           err_msg="%s gradient %d" % (net, i))
            # replicas: identical trainable parameters (BatchNorm running statistics are per-rank by design)
            for i, tr in enumerate(trainable[net]):
                if tr:
                    np.testing.assert_array_equal(r0["%s%d" % (net, i)], r1["%s%d" % (net, i)], err_msg="%s param %d" % (net, i))
    finally:
        _lib.call, _lib.query, _lib.load = saved
