"""Data parallelism on real GPUs: two ranks over NCCL (one process per GPU), the product path on the CUDA kernels.
Needs >= 2 visible GPUs (skipped otherwise; the round's 2-GPU visit runs it).

  * parity mode + SyncBN: two ranks with half a batch each reproduce the oracle's step on the WHOLE batch (losses,
    every gradient array, updated parameters), and the replicas are bit-identical;
  * fast mode (the benchmarked schedule: CUDA graphs with the captured, overlapped, bucketed gradient all-reduces):
    after five steps the replicas hold bit-identical parameters, and the all-reduced gradient equals the sum of the
    per-rank gradients an un-synchronised model computes on the same shard."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2,
                                 reason="needs two GPUs")]


def _setup(rank, world, port):
    for p in (ROOT, os.path.join(ROOT, "gan-heightmaps_b200"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    return dist


def _leave():
    """Results are on disk: leave without tearing the NCCL communicator down.  destroy_process_group() hangs in this
    image once CUDA graphs holding collectives of the communicator have been replayed (observed on 2 x B200 for both
    bench.py and these workers: everything had completed, the call never returned)."""
    torch.cuda.synchronize()
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)


def _worker_parity(rank, world, port, out_dir):
    dist = _setup(rank, world, port)
    from oracle import step as S
    import test_engine_cpu as T
    cfg = S.experiment_kwargs('gate64')
    om, m = T.build_pair(cfg, 'dcgan', with_p2p=False, device="cuda:%d" % rank, precision="parity")
    m.pg = dist.group.WORLD
    m.rt.sync_bn_group = dist.group.WORLD
    out = {}
    for it in range(3):
        Z, X, Y = S.synthetic_batch(8, cfg['latent_dim'], 64, seed=11 + it)
        sl = slice(4 * rank, 4 * rank + 4)
        out["losses%d" % it] = np.asarray(m.train_fn(Z[sl], X[sl], Y[sl]))
        if it == 0:
            out.update({"gG%d" % i: v / world for i, v in enumerate(m.G.get_grads())})
            out.update({"gD%d" % i: v / world for i, v in enumerate(m.D.get_grads())})
        if rank == 0:
            out["olosses%d" % it] = np.asarray(om.train_fn(Z, X, Y))
            if it == 0:
                out.update({"oG%d" % i: v for i, v in enumerate(om.last_grads['G'])})
                out.update({"oD%d" % i: v for i, v in enumerate(om.last_grads['D'])})
    out.update({"pG%d" % i: v for i, v in enumerate(m.G.get_all_param_values())})
    out.update({"pD%d" % i: v for i, v in enumerate(m.D.get_all_param_values())})
    if rank == 0:
        out.update({"opG%d" % i: v for i, v in enumerate(om.get_all_param_values('G'))})
    np.savez(os.path.join(out_dir, "parity%d.npz" % rank), **out)
    _leave()


def _worker_fast(rank, world, port, out_dir):
    dist = _setup(rank, world, port)
    from oracle import step as S
    import test_engine_cpu as T
    cfg = dict(in_shp=64, latent_dim=32, G=dict(nch=256, num_repeats=0, div=[2, 2, 4, 4]),
               D=dict(nch=64, num_repeats=0, bn=False, nonlinearity='linear', div=[1, 1, 1, 1]))
    _, m = T.build_pair(cfg, 'dcgan', with_p2p=False, device="cuda:%d" % rank, precision="fast", lr=1e-4)
    _, solo = T.build_pair(cfg, 'dcgan', with_p2p=False, device="cuda:%d" % rank, precision="fast", lr=1e-4)
    m.pg = dist.group.WORLD
    out = {}
    for it in range(5):                     # eager, eager, captured, replayed, replayed
        Z, X, Y = S.synthetic_batch(8, cfg['latent_dim'], 64, seed=21 + it)
        sl = slice(4 * rank, 4 * rank + 4)
        if it in (0, 3):                    # the un-synchronised twin starts this step from the same state
            for a, b in zip(m._nets(), solo._nets()):
                b.pflat.copy_(a.pflat)
                b.sflat.copy_(a.sflat)
                for k, v in a.opt_state.items():
                    b.opt_state[k] = v.clone()
                b._packed = False
            solo.train_fn(Z[sl], X[sl], Y[sl])
            for name, net in (("G", solo.G), ("D", solo.D)):
                out["solo%d_%s" % (it, name)] = net.gflat.detach().cpu().numpy().copy()
        out["losses%d" % it] = np.asarray(m.train_fn(Z[sl], X[sl], Y[sl]))
        if it in (0, 3):
            for name, net in (("G", m.G), ("D", m.D)):
                out["sum%d_%s" % (it, name)] = net.gflat.detach().cpu().numpy().copy()
    out["pG"] = m.G.pflat.detach().cpu().numpy()
    out["pD"] = m.D.pflat.detach().cpu().numpy()
    out["graphs"] = np.asarray([sum(1 for st in m._graphs.values() if st.get("gA") is not None or st.get("graph") is not None)])
    np.savez(os.path.join(out_dir, "fast%d.npz" % rank), **out)
    _leave()


def test_two_ranks_over_nccl_equal_the_oracle_on_the_whole_batch(tmp_path):
    world, port = 2, 32500 + os.getpid() % 2000
    mp.spawn(_worker_parity, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = np.load(str(tmp_path / "parity0.npz")), np.load(str(tmp_path / "parity1.npz"))
    for it in range(3):
        np.testing.assert_allclose(0.5 * (r0["losses%d" % it][:2] + r1["losses%d" % it][:2]), r0["olosses%d" % it][:2],
                                   rtol=1e-3, atol=1e-6)
    for net in ("G", "D"):
        n = len([k for k in r0.files if k.startswith("o" + net) and not k.startswith("op")])
        for i in range(n):
            a, b = r0["g%s%d" % (net, i)], r0["o%s%d" % (net, i)]
            np.testing.assert_array_equal(a, r1["g%s%d" % (net, i)])          # the same all-reduced sums on both ranks
            scale = float(np.abs(b).max()) + 1e-12
            assert float(np.abs(a - b).max()) <= 2e-3 * scale + 1e-7, (net, i, float(np.abs(a - b).max()), scale)
    for k in [k for k in r0.files if k.startswith("pG") or k.startswith("pD")]:
        np.testing.assert_array_equal(r0[k], r1[k])                           # bit-identical replicas
    for i in range(len([k for k in r0.files if k.startswith("opG")])):
        np.testing.assert_allclose(r0["pG%d" % i], r0["opG%d" % i], rtol=2e-3, atol=3e-4, err_msg="G value %d" % i)


def test_fast_mode_replicas_stay_identical_under_captured_overlapped_allreduce(tmp_path):
    world, port = 2, 34500 + os.getpid() % 2000
    mp.spawn(_worker_fast, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = np.load(str(tmp_path / "fast0.npz")), np.load(str(tmp_path / "fast1.npz"))
    assert int(r0["graphs"][0]) >= 1                                           # steps 3-5 ran as captured graphs
    for k in ("pG", "pD"):
        np.testing.assert_array_equal(r0[k], r1[k])
    for it in (0, 3):                                                          # eager and replayed step
        for net in ("G", "D"):
            tot = r0["sum%d_%s" % (it, net)]
            np.testing.assert_array_equal(tot, r1["sum%d_%s" % (it, net)])
            ref = r0["solo%d_%s" % (it, net)] + r1["solo%d_%s" % (it, net)]
            err = float(np.linalg.norm(tot - ref) / (np.linalg.norm(ref) + 1e-30))
            assert err <= 1e-3, (it, net, err)        # same kernels on the same shard: only the order of atomic adds differs
    for it in range(5):
        assert np.all(np.isfinite(r0["losses%d" % it]))
