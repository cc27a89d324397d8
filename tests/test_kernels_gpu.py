"""Kernel-level parity, through the C ABI: every libhmgan entry point is run on the
GPU and compared with the CPU emulation of its contract (tests/fake_hmgan.py, which
is itself pinned to the oracle by tests/test_engine_cpu.py) on the same seeded
inputs.  fp32: rtol 1e-4 of the output scale (different summation order only);
fp16 storage: 2e-3 (half rounding of inputs is shared, fp32 accumulate on both sides)."""
import ctypes as C
import os
import sys
import zlib

import numpy as np
import pytest
import torch

import fake_hmgan

PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gan-heightmaps_b200")
if PKG not in sys.path:
    sys.path.insert(0, PKG)
import _lib   # noqa: E402

pytestmark = pytest.mark.gpu
TD = {0: torch.float32, 1: torch.float16}
TOL = {0: 1e-4, 1: 2e-3}


class Both(object):
    """Holds a CPU and a CUDA copy of every tensor; runs a call on both backends."""

    def __init__(self):
        self.cpu, self.gpu = [], []

    def t(self, arr, dtype=torch.float32):
        a = torch.from_numpy(np.ascontiguousarray(arr)).to(dtype).contiguous()
        self.cpu.append(a)
        self.gpu.append(a.cuda())
        return len(self.cpu) - 1

    def run(self, name, mk):
        """mk(P) -> args, with P(i) the data pointer of tensor i on that backend."""
        fake_hmgan.call(name, *mk(lambda i: None if i is None else self.cpu[i].data_ptr()), None)
        _lib.call(name, *mk(lambda i: None if i is None else self.gpu[i].data_ptr()), None)
        torch.cuda.synchronize()

    def check(self, i, tol, what=""):
        a = self.gpu[i].cpu().double()
        b = self.cpu[i].double()
        scale = float(b.abs().max()) + 1e-30
        err = float((a - b).abs().max())
        assert err <= tol * scale, "%s: max err %.3g vs scale %.3g" % (what, err, scale)


def desc(**kw):
    d = _lib.ConvDesc()
    base = dict(dtype=0, B=2, H=8, W=8, C1=4, C2=0, up=0, kh=3, kw=3, stride=1, pad=1, transposed=0, Ho=8, Wo=8,
                Cout=8, oH=8, oW=8, os=1, ou=0, ov=0, split=8, act=0, slope=0.0, accumulate=0)
    base.update(kw)
    for k, v in base.items():
        setattr(d, k, v)
    return d


FWD_CASES = {
    "5x5_same": dict(H=12, W=10, C1=5, kh=5, kw=5, pad=2, Ho=12, Wo=10, oH=12, oW=10, Cout=7, split=7, act=1,
                     slope=0.2),
    "3x3_s2": dict(H=16, W=16, C1=6, kh=3, kw=3, stride=2, pad=1, Ho=8, Wo=8, oH=8, oW=8, Cout=9, split=9, act=1,
                   slope=0.01),
    "5x5_nearest_up_sigmoid": dict(H=6, W=6, C1=8, up=1, kh=5, kw=5, pad=2, Ho=12, Wo=12, oH=12, oW=12, Cout=1,
                                   split=1, act=3),
    "3x3_bilinear_up_concat": dict(H=5, W=5, C1=6, C2=3, up=2, kh=3, kw=3, pad=1, Ho=10, Wo=10, oH=10, oW=10,
                                   Cout=8, split=8),
    "2x2_valid": dict(H=2, W=2, C1=16, kh=2, kw=2, pad=0, Ho=1, Wo=1, oH=1, oW=1, Cout=16, split=16),
    "deconv_phase_scatter_tanh": dict(H=4, W=4, C1=8, C2=4, kh=1, kw=1, pad=0, Ho=4, Wo=4, oH=8, oW=8, os=2, ou=1,
                                      ov=0, Cout=3, split=3, act=4),
    "dense": dict(B=4, H=1, W=1, C1=100, kh=1, kw=1, pad=0, Ho=1, Wo=1, oH=1, oW=1, Cout=128, split=128),
    "wide": dict(B=1, H=20, W=20, C1=64, kh=5, kw=5, pad=2, Ho=20, Wo=20, oH=20, oW=20, Cout=96, split=96, act=2),
    "thin_in_5x5_1_64": dict(B=2, H=24, W=20, C1=1, kh=5, kw=5, pad=2, Ho=24, Wo=20, oH=24, oW=20, Cout=64, split=64,
                             act=1, slope=0.2),
    "thin_in_3x3_s2_1+3_32": dict(B=2, H=16, W=16, C1=1, C2=3, kh=3, kw=3, stride=2, pad=1, Ho=8, Wo=8, oH=8, oW=8,
                                  Cout=32, split=32, act=1, slope=0.01),
    "thin_out_5x5_64_1": dict(B=2, H=16, W=16, C1=64, kh=5, kw=5, pad=2, Ho=16, Wo=16, oH=16, oW=16, Cout=1, split=1,
                              act=3),
    "thin_out_2x2_deconv_phase_128_3": dict(B=2, H=8, W=8, C1=64, C2=64, kh=1, kw=1, pad=0, Ho=8, Wo=8, oH=16, oW=16,
                                            os=2, ou=0, ov=1, Cout=3, split=3, act=4),
    "dgrad_5x5": dict(H=12, W=10, C1=7, kh=5, kw=5, pad=2, transposed=1, Ho=12, Wo=10, oH=12, oW=10, Cout=5,
                      split=5),
    "dgrad_3x3_s2_split_acc": dict(H=8, W=8, C1=9, kh=3, kw=3, stride=2, pad=1, transposed=1, Ho=16, Wo=16, oH=16,
                                   oW=16, Cout=6, split=2, accumulate=2),
    "deconv_dgrad": dict(H=8, W=8, C1=3, kh=2, kw=2, stride=2, pad=0, Ho=4, Wo=4, oH=4, oW=4, Cout=12, split=8,
                         accumulate=1),
}


@pytest.mark.parametrize("dtype", [0, 1])
@pytest.mark.parametrize("case", sorted(FWD_CASES))
def test_conv_gather(case, dtype):
    d = desc(dtype=dtype, **FWD_CASES[case])
    r = np.random.RandomState(zlib.crc32(case.encode()) % 1000)
    Ct = d.C1 + d.C2
    bo = Both()
    x1 = bo.t(r.randn(d.B, d.H, d.W, d.C1), TD[dtype])
    x2 = bo.t(r.randn(d.B, d.H, d.W, d.C2), TD[dtype]) if d.C2 else None
    w = bo.t(r.randn(d.kh * d.kw * Ct, d.Cout) / np.sqrt(d.kh * d.kw * Ct), TD[dtype])
    bias = bo.t(r.randn(d.Cout)) if not d.transposed else None
    y = bo.t(r.randn(d.B, d.oH, d.oW, d.split), TD[dtype])
    y2 = bo.t(r.randn(d.B, d.oH, d.oW, d.Cout - d.split), TD[dtype]) if d.split < d.Cout else None
    bo.run("hm_conv_gather", lambda P: (C.byref(d), P(x1), P(x2), P(w), P(bias), P(y), P(y2)))
    bo.check(y, TOL[dtype], case)
    if y2 is not None:
        bo.check(y2, TOL[dtype], case + " (y2)")


@pytest.mark.parametrize("dtype", [0, 1])
@pytest.mark.parametrize("case", [c for c in sorted(FWD_CASES) if not FWD_CASES[c].get("transposed")])
def test_conv_wgrad(case, dtype):
    d = desc(dtype=dtype, **FWD_CASES[case])
    r = np.random.RandomState(zlib.crc32(case.encode()) % 1000 + 1)
    Ct = d.C1 + d.C2
    bo = Both()
    x1 = bo.t(r.randn(d.B, d.H, d.W, d.C1), TD[dtype])
    x2 = bo.t(r.randn(d.B, d.H, d.W, d.C2), TD[dtype]) if d.C2 else None
    dy = bo.t(r.randn(d.B, d.oH, d.oW, d.Cout), TD[dtype])
    dw = bo.t(np.zeros((d.kh * d.kw * Ct, d.Cout)))
    bo.run("hm_conv_wgrad", lambda P: (C.byref(d), P(x1), P(x2), P(dy), P(dw)))
    bo.check(dw, TOL[dtype], case)


@pytest.mark.parametrize("mode,shape", [(0, (6, 4, 5, 5)), (1, (6, 4, 3, 3)), (2, (4, 6, 2, 2)), (3, (4, 6, 2, 2)),
                                        (4, (10, 12, 1, 1))])
@pytest.mark.parametrize("dtype", [0, 1])
def test_pack_unpack(mode, shape, dtype):
    r = np.random.RandomState(mode)
    a, b, kh, kw = shape
    cout, cin = (a, b) if mode in (0, 1) else (b, a)
    if mode == 4:
        cin, cout = a, b
    bo = Both()
    w = bo.t(r.randn(*shape))
    n = a * b if mode == 2 else a * b * kh * kw
    wp = bo.t(np.zeros(n), TD[dtype])
    bo.run("hm_pack_conv_weight", lambda P: (P(w), P(wp), mode, cout, cin, kh, kw, kh - 1, 0, dtype))
    bo.check(wp, 1e-3 if dtype else 0.0, "pack mode %d" % mode)
    if mode in (0, 2, 4) and dtype == 0:
        g = bo.t(r.randn(a * b * kh * kw))
        out = bo.t(np.zeros(a * b * kh * kw))
        bo.run("hm_unpack_conv_wgrad", lambda P: (P(g), P(out), mode, cout, cin, kh, kw))
        bo.check(out, 0.0, "unpack mode %d" % mode)


@pytest.mark.parametrize("mode,shape", [(5, (64, 96, 3, 3)), (6, (64, 96, 3, 3)), (5, (40, 33, 5, 5)), (6, (40, 33, 5, 5)),
                                        (5, (128, 64, 1, 1)), (6, (96, 64, 2, 2)), (0, (64, 96, 3, 3)),
                                        (0, (40, 33, 5, 5)), (20, (64, 64, 5, 5)), (21, (48, 40, 1, 1)),
                                        (8, (32, 64, 5, 5)), (12, (64, 64, 3, 3)), (22, (64, 128, 2, 2)),
                                        (17, (64, 128, 2, 2))])
@pytest.mark.parametrize("dtype", [0, 1])
def test_tensor_core_packs(mode, shape, dtype):
    """The K-major tensor-core packs (modes 5 / 6, the phase packs 8 / 12 / 20, the dense pack 21) and the un-pack mode 0
    at layer-sized shapes against the emulation of their index maps: exact in float32, one rounding in fp16."""
    r = np.random.RandomState(mode * 7 + shape[0])
    cout, cin, kh, kw = shape
    bo = Both()
    wshape = (cin, cout) if mode == 21 else (cout, cin, kh, kw)
    w = bo.t(r.randn(*wshape))
    if mode == 0:
        if dtype == 1:
            pytest.skip("the un-pack is float32 only")
        g = bo.t(r.randn(kh * kw * cin * cout))
        out = bo.t(np.zeros(cout * cin * kh * kw))
        bo.run("hm_unpack_conv_wgrad", lambda P: (P(g), P(out), 0, cout, cin, kh, kw))
        bo.check(out, 0.0, "unpack mode 0 %r" % (shape,))
        return
    n = _lib.pack_count(mode, cout, cin, kh, kw)
    wp = bo.t(np.zeros(n), TD[dtype])
    bo.run("hm_pack_conv_weight", lambda P: (P(w), P(wp), mode, cout, cin, kh, kw, 0, 0, dtype))
    bo.check(wp, 1e-3 if dtype else (1e-6 if mode in (8, 20) else 0.0), "pack mode %d %r" % (mode, shape))


@pytest.mark.parametrize("dtype", [0, 1])
def test_all_packs_of_a_network_in_one_launch(dtype):
    """hm_pack_conv_weight_multi (an HmPackJob table in device memory, one launch) against one hm_pack_conv_weight call
    per job on the GPU: bit-identical packed copies for a mix of modes and sizes (the job table's element counts come from
    _lib.pack_count)."""
    r = np.random.RandomState(3)
    specs = [(5, 128, 64, 5, 5), (6, 128, 64, 5, 5), (15, 64, 1, 5, 5), (8, 64, 64, 5, 5), (22, 64, 128, 2, 2),
             (17, 64, 128, 2, 2), (21, 96, 40, 1, 1)]
    ws, one, multi, jobs = [], [], [], []
    for (mode, cout, cin, kh, kw) in specs:
        w = torch.from_numpy(r.randn(cout * cin * kh * kw).astype(np.float32)).cuda()
        n = _lib.pack_count(mode, cout, cin, kh, kw)
        a = torch.zeros(n, dtype=TD[dtype], device="cuda")
        b = torch.full((n,), 7.0, dtype=TD[dtype], device="cuda")
        _lib.call("hm_pack_conv_weight", w.data_ptr(), a.data_ptr(), mode, cout, cin, kh, kw, 0, 0, dtype, None)
        jobs.append((w.data_ptr(), b.data_ptr(), mode, cout, cin, kh, kw, 0, 0, dtype))
        ws.append(w), one.append(a), multi.append(b)
    raw, max_n = _lib.pack_job_table(jobs)
    tab = torch.frombuffer(bytearray(raw), dtype=torch.uint8).cuda()
    _lib.call("hm_pack_conv_weight_multi", tab.data_ptr(), len(jobs), max_n, dtype, None)
    torch.cuda.synchronize()
    for spec, a, b in zip(specs, one, multi):
        assert torch.equal(a, b), spec


@pytest.mark.parametrize("dtype", [0, 1])
@pytest.mark.parametrize("M,Cn", [(1000, 8), (4, 2048), (4099, 64), (300, 512), (77, 3)])
def test_batchnorm_chain(M, Cn, dtype):
    r = np.random.RandomState(M)
    bo = Both()
    x = bo.t(r.randn(M, Cn) * 2 + 0.5, TD[dtype])
    sums = bo.t(np.zeros(2 * Cn), torch.float64)
    bo.run("hm_bn_stats", lambda P: (P(x), dtype, M, Cn, P(sums)))
    bo.check(sums, 1e-5, "bn_stats")
    gam = r.rand(Cn) + 0.5
    gam[0] = 1e-5                                   # below the 2^-10 threshold of hm_bn_bwd_*_a: x-based fallback
    gamma, beta = bo.t(gam), bo.t(r.randn(Cn))
    rm, ri = bo.t(r.randn(Cn)), bo.t(r.rand(Cn) + 0.5)
    outs = [bo.t(np.zeros(Cn)) for _ in range(4)]
    # use the CPU sums on both sides so the comparison below isolates each kernel
    bo.gpu[sums].copy_(bo.cpu[sums])
    bo.run("hm_bn_finalize", lambda P: (P(sums), M, Cn, P(gamma), P(beta), P(rm), P(ri), 1e-4, 0.1, 1, P(outs[0]),
                                        P(outs[1]), P(outs[2]), P(outs[3])))
    for i in outs + [rm, ri]:
        bo.check(i, 2e-6, "bn_finalize")
        bo.gpu[i].copy_(bo.cpu[i])
    a = bo.t(np.zeros((M, Cn)), TD[dtype])
    bo.run("hm_bn_apply_act", lambda P: (P(x), P(a), dtype, M, Cn, P(outs[2]), P(outs[3]), 1, 0.2))
    bo.check(a, TOL[dtype], "bn_apply_act")
    bo.gpu[a].copy_(bo.cpu[a])
    da = bo.t(r.randn(M, Cn), TD[dtype])
    red = bo.t(np.zeros(2 * Cn), torch.float64)
    bo.run("hm_bn_bwd_reduce", lambda P: (P(da), P(a), P(x), dtype, M, Cn, P(outs[0]), P(outs[1]), 1, 0.2, P(red)))
    bo.check(red, 1e-5, "bn_bwd_reduce")
    bo.gpu[red].copy_(bo.cpu[red])
    dx = bo.t(np.zeros((M, Cn)), TD[dtype])
    dg, db = bo.t(np.zeros(Cn)), bo.t(np.zeros(Cn))
    bo.run("hm_bn_bwd_apply", lambda P: (P(da), P(a), P(x), P(dx), dtype, M, Cn, P(outs[0]), P(outs[1]), P(gamma),
                                         1, 0.2, P(red), P(dg), P(db)))
    bo.check(dx, TOL[dtype], "bn_bwd_apply")
    bo.check(dg, 1e-6, "dgamma")
    bo.check(db, 1e-6, "dbeta")
    # the same two passes with xhat recovered from the OUTPUT a instead of reading x (hm_bn_bwd_*_a; a channel with a
    # tiny gamma takes the x-based fallback inside the kernel): equal up to the rounding of a
    red2 = bo.t(np.zeros(2 * Cn), torch.float64)
    dx2 = bo.t(np.zeros((M, Cn)), TD[dtype])
    bo.run("hm_bn_bwd_reduce_a", lambda P: (P(da), P(a), P(x), dtype, M, Cn, P(outs[0]), P(outs[1]), P(gamma), P(beta),
                                            1, 0.2, P(red2)))
    bo.check(red2, 1e-4 if dtype == 0 else 5e-3, "bn_bwd_reduce_a")
    bo.gpu[red2].copy_(bo.cpu[red2])
    bo.run("hm_bn_bwd_apply_a", lambda P: (P(da), P(a), P(x), P(dx2), dtype, M, Cn, P(outs[0]), P(outs[1]), P(gamma),
                                           P(beta), 1, 0.2, P(red2), P(dg), P(db)))
    bo.check(dx2, 1e-4 if dtype == 0 else 5e-3, "bn_bwd_apply_a")
    # deterministic mode: scale/shift from the running statistics
    bo.run("hm_bn_finalize", lambda P: (None, M, Cn, P(gamma), P(beta), P(rm), P(ri), 1e-4, 0.1, 0, None, None,
                                        P(outs[2]), P(outs[3])))
    bo.check(outs[2], 2e-6, "bn_finalize det scale")
    bo.check(outs[3], 2e-6, "bn_finalize det shift")


@pytest.mark.parametrize("dtype", [0, 1])
def test_pool_upsample_act_layout(dtype):
    r = np.random.RandomState(7)
    B, H, W, Cn = 3, 8, 12, 5
    bo = Both()
    x = bo.t(r.randn(B, H, W, Cn), TD[dtype])
    p = bo.t(np.zeros((B, H // 2, W // 2, Cn)), TD[dtype])
    idx = bo.t(np.zeros((B, H // 2, W // 2, Cn)), torch.uint8)
    bo.run("hm_maxpool2_fwd", lambda P: (P(x), P(p), P(idx), dtype, B, H, W, Cn))
    bo.check(p, 0.0, "maxpool fwd")
    assert torch.equal(bo.gpu[idx].cpu(), bo.cpu[idx])
    dp = bo.t(r.randn(B, H // 2, W // 2, Cn), TD[dtype])
    dx = bo.t(np.zeros((B, H, W, Cn)), TD[dtype])
    bo.run("hm_maxpool2_bwd", lambda P: (P(dp), P(p), P(idx), P(dx), dtype, B, H, W, Cn, 1, 0.2, None))
    bo.check(dx, 1e-3 if dtype else 1e-6, "maxpool bwd")
    # with the fused bias gradient, on a vector-path shape (C % 8 == 0, 256 % (C/8) == 0) and on the scalar path
    for (B2, H2, W2, C2) in ((2, 8, 12, 64), (3, 8, 12, 5)):
        x2 = bo.t(r.randn(B2, H2, W2, C2), TD[dtype])
        p2 = bo.t(np.zeros((B2, H2 // 2, W2 // 2, C2)), TD[dtype])
        i2 = bo.t(np.zeros((B2, H2 // 2, W2 // 2, C2)), torch.uint8)
        bo.run("hm_maxpool2_fwd", lambda P: (P(x2), P(p2), P(i2), dtype, B2, H2, W2, C2))
        dp2 = bo.t(r.randn(B2, H2 // 2, W2 // 2, C2), TD[dtype])
        dx2 = bo.t(np.zeros((B2, H2, W2, C2)), TD[dtype])
        db2 = bo.t(np.zeros(C2))
        bo.run("hm_maxpool2_bwd", lambda P: (P(dp2), P(p2), P(i2), P(dx2), dtype, B2, H2, W2, C2, 1, 0.2, P(db2)))
        bo.check(dx2, 1e-3 if dtype else 1e-6, "maxpool bwd (+db)")
        bo.check(db2, 2e-3 if dtype else 1e-5, "maxpool bwd bias gradient")
    for mode in (1, 2):
        up = bo.t(np.zeros((B, 2 * H, 2 * W, Cn)), TD[dtype])
        bo.run("hm_upsample2_fwd", lambda P: (P(x), P(up), dtype, B, H, W, Cn, mode))
        bo.check(up, 1e-3 if dtype else 1e-6, "upsample fwd %d" % mode)
        g = bo.t(r.randn(B, 2 * H, 2 * W, Cn), TD[dtype])
        for acc in (0, 1):
            gx = bo.t(r.randn(B, H, W, Cn), TD[dtype])
            bo.run("hm_upsample2_bwd", lambda P: (P(g), P(gx), dtype, B, H, W, Cn, mode, acc))
            bo.check(gx, 2e-3 if dtype else 1e-6, "upsample bwd %d acc %d" % (mode, acc))
    for act in (1, 2, 3, 4):
        y = bo.t(r.rand(1000) * 1.6 - 0.8, TD[dtype])
        dy = bo.t(r.randn(1000), TD[dtype])
        out = bo.t(r.randn(1000), TD[dtype])
        bo.run("hm_act_bwd", lambda P: (P(dy), P(y), P(out), dtype, 1000, act, 0.2, 1))
        bo.check(out, 2e-3 if dtype else 1e-6, "act_bwd %d" % act)
    src = bo.t(r.randn(B, 3, H, W))
    dst = bo.t(np.zeros((B, H, W, 3)), TD[dtype])
    bo.run("hm_nchw_to_nhwc", lambda P: (P(src), P(dst), dtype, B, 3, H, W))
    bo.check(dst, 1e-3 if dtype else 0.0, "nchw_to_nhwc")
    back = bo.t(np.zeros((B, 3, H, W)))
    bo.run("hm_nhwc_to_nchw", lambda P: (P(dst), P(back), dtype, B, 3, H, W))
    bo.check(back, 0.0, "nhwc_to_nchw")
    flat = bo.t(r.randn(B, Cn * H * W), TD[dtype])
    perm = bo.t(np.zeros((B, H, W, Cn)), TD[dtype])
    bo.run("hm_permute", lambda P: (P(flat), P(perm), dtype, B, Cn, H, W, 0))
    bo.check(perm, 0.0, "permute")
    inv = bo.t(np.zeros((B, Cn * H * W)), TD[dtype])
    bo.run("hm_permute", lambda P: (P(perm), P(inv), dtype, B, Cn, H, W, 1))
    assert torch.equal(bo.gpu[inv].cpu(), bo.cpu[flat])
    cs = bo.t(np.zeros(Cn))
    bo.run("hm_col_sum", lambda P: (P(x), dtype, B * H * W, Cn, P(cs)))
    bo.check(cs, 1e-5, "col_sum")


@pytest.mark.parametrize("dtype", [0, 1])
def test_losses_and_optimisers(dtype):
    r = np.random.RandomState(11)
    bo = Both()
    for (R, G, out_act, target, lsgan, relu) in [(8, 16, 0, 1.0, 1, 1), (512, 1, 0, 0.0, 1, 0), (37, 4, 3, 1.0, 0, 0),
                                                 (37, 4, 3, 0.0, 0, 1)]:
        h = np.abs(r.randn(R, G)) * (r.rand(R, G) > 0.3) if relu else r.randn(R, G)
        h = bo.t(h, TD[dtype])
        dh = bo.t(r.randn(R, G), TD[dtype])
        loss = bo.t(np.array([0.25]))
        bo.run("hm_adv_loss", lambda P: (P(h), P(dh), dtype, R, G, out_act, target, lsgan, relu, 1.0, 64.0, 1,
                                         P(loss)))
        bo.check(loss, 1e-5, "adv loss")
        bo.check(dh, 2e-3 if dtype else 1e-5, "adv grad")
    n = 5000
    p, y = bo.t(r.randn(n), TD[dtype]), bo.t(r.randn(n), TD[dtype])
    for l2 in (0, 1):
        dp = bo.t(np.zeros(n), TD[dtype])
        loss = bo.t(np.zeros(1))
        bo.run("hm_recon_loss", lambda P: (P(p), P(y), P(dp), dtype, n, l2, 1.0, 100.0 * 64, 0, P(loss)))
        bo.check(loss, 1e-5, "recon loss")
        bo.check(dp, 2e-3 if dtype else 1e-6, "recon grad")
    if dtype == 0:
        n = 10001
        w, g, acc = bo.t(r.randn(n)), bo.t(r.randn(n) * 1e-3), bo.t(r.rand(n) * 1e-6)
        lr = bo.t(np.array([1e-4]))
        bo.run("hm_rmsprop", lambda P: (P(w), P(g), P(acc), n, P(lr), 0.9, 1e-6, 0.5))
        bo.check(w, 1e-6, "rmsprop p")
        bo.check(acc, 1e-6, "rmsprop acc")
        m, v = bo.t(r.randn(n) * 1e-3), bo.t(r.rand(n) * 1e-6)
        bo.run("hm_adam", lambda P: (P(w), P(g), P(m), P(v), n, P(lr), 0.9, 0.999, 1e-8, 3, 1.0))
        bo.check(w, 1e-6, "adam p")
        bo.check(m, 1e-6, "adam m")
        bo.check(v, 1e-6, "adam v")


def test_bad_arguments_return_errors_not_crashes():
    lib = _lib.load()
    d = desc(C1=0)
    assert lib.hm_conv_gather(C.byref(d), None, None, None, None, None, None, None) < 0
    assert b"hm_conv_gather" in lib.hm_last_error_string()
    assert lib.hm_bn_stats(None, 0, 10, 4, None, None) < 0
    assert lib.hm_rmsprop(None, None, None, 0, None, 0.9, 1e-6, 1.0, None) < 0
    with pytest.raises(_lib.HmError):
        _lib.call("hm_cast", None, 0, None, 0, 10, None)


@pytest.mark.parametrize("dtype", [0, 1])
def test_conv_adjoint_identity_at_full_layer_size(dtype):
    """Size-independent property at a BASELINE-size layer (64->128 5x5 @256^2, the hottest D layer):
    <conv(x), dy> == <x, dgrad(dy)> == <W, wgrad(x, dy)>  (no bias)."""
    torch.manual_seed(0)
    B, H, Ci, Co, k = 2, 256, 64, 128, 5
    td = TD[dtype]
    x = (torch.randn(B, H, H, Ci, device="cuda") * 0.5).to(td)
    dy = (torch.randn(B, H, H, Co, device="cuda") * 0.5).to(td)
    W = torch.randn(Co, Ci, k, k, device="cuda") / np.sqrt(k * k * Ci)
    wp_f = torch.empty(k * k * Ci * Co, device="cuda", dtype=td)
    wp_d = torch.empty_like(wp_f)
    _lib.call("hm_pack_conv_weight", W.data_ptr(), wp_f.data_ptr(), 0, Co, Ci, k, k, 0, 0, dtype, None)
    _lib.call("hm_pack_conv_weight", W.data_ptr(), wp_d.data_ptr(), 1, Co, Ci, k, k, 0, 0, dtype, None)
    f = desc(dtype=dtype, B=B, H=H, W=H, C1=Ci, kh=k, kw=k, pad=2, Ho=H, Wo=H, oH=H, oW=H, Cout=Co, split=Co)
    y = torch.empty(B, H, H, Co, device="cuda", dtype=td)
    _lib.call("hm_conv_gather", C.byref(f), x.data_ptr(), None, wp_f.data_ptr(), None, y.data_ptr(), None, None)
    g = desc(dtype=dtype, B=B, H=H, W=H, C1=Co, kh=k, kw=k, pad=2, transposed=1, Ho=H, Wo=H, oH=H, oW=H, Cout=Ci,
             split=Ci)
    dx = torch.empty(B, H, H, Ci, device="cuda", dtype=td)
    _lib.call("hm_conv_gather", C.byref(g), dy.data_ptr(), None, wp_d.data_ptr(), None, dx.data_ptr(), None, None)
    dwp = torch.zeros(k * k * Ci * Co, device="cuda")
    _lib.call("hm_conv_wgrad", C.byref(f), x.data_ptr(), None, dy.data_ptr(), dwp.data_ptr(), None)
    dW = torch.empty_like(W)
    _lib.call("hm_unpack_conv_wgrad", dwp.data_ptr(), dW.data_ptr(), 0, Co, Ci, k, k, None)
    torch.cuda.synchronize()
    a = float((y.double() * dy.double()).sum())
    b = float((x.double() * dx.double()).sum())
    Wq = W.to(td).double() if dtype else W.double()
    c = float((Wq * dW.double()).sum())
    tol = 2e-3 if dtype else 1e-4
    scale = float(y.double().norm() * dy.double().norm())
    assert abs(a - b) <= tol * scale and abs(a - c) <= tol * scale, (a, b, c, scale)


@pytest.mark.parametrize("dtype", [0, 1])
@pytest.mark.parametrize("Co", [1, 3])
def test_up2conv_wgrad_phases_and_fold(Co, dtype):
    """Phase-decomposed weight gradient of (nearest 2x -> 5x5) and its fold back onto the 5x5 filter (unpack mode 9)
    equal the plain weight gradient through the virtual upsampling (hm_conv_wgrad + unpack mode 0)."""
    r = np.random.RandomState(3 + Co)
    B, H, W, Ci = 2, 12, 10, 16
    d = desc(dtype=dtype, B=B, H=H, W=W, C1=Ci, up=1, kh=5, kw=5, pad=2, Ho=2 * H, Wo=2 * W, oH=2 * H, oW=2 * W,
             Cout=Co, split=Co)
    bo = Both()
    x = bo.t(r.randn(B, H, W, Ci), TD[dtype])
    dy = bo.t(r.randn(B, 2 * H, 2 * W, Co), TD[dtype])
    ph = bo.t(np.zeros(36 * Ci * Co))
    bo.run("hm_up2conv_wgrad_phases", lambda P: (C.byref(d), P(x), P(dy), P(ph)))
    bo.check(ph, TOL[dtype], "phase gradients")
    folded = bo.t(np.zeros(25 * Ci * Co))
    bo.run("hm_unpack_conv_wgrad", lambda P: (P(ph), P(folded), 9, Co, Ci, 5, 5))
    bo.check(folded, 1e-5, "fold")
    plain = bo.t(np.zeros(25 * Ci * Co))
    direct = bo.t(np.zeros(25 * Ci * Co))
    bo.run("hm_conv_wgrad", lambda P: (C.byref(d), P(x), None, P(dy), P(plain)))
    bo.run("hm_unpack_conv_wgrad", lambda P: (P(plain), P(direct), 0, Co, Ci, 5, 5))
    a, b = bo.gpu[folded].cpu().double(), bo.gpu[direct].cpu().double()
    assert float((a - b).abs().max()) <= 1e-3 * float(b.abs().max())


def test_im2col_and_s2d_relayouts():
    r = np.random.RandomState(21)
    bo = Both()
    B, H, W = 2, 9, 12
    x = bo.t(r.randn(B, H, W, 1), torch.float16)
    xc = bo.t(np.ones((B, H, W, 64)), torch.float16)
    bo.run("hm_im2col_c1", lambda P: (P(x), P(xc), B, H, W, 5, 5, 2))
    assert torch.equal(bo.gpu[xc].cpu(), bo.cpu[xc])
    for Co in (1, 3):
        dy = bo.t(r.randn(B, 2 * H, 2 * W, Co), torch.float16)
        out = bo.t(np.ones((B, H, W, 64)), torch.float16)
        bo.run("hm_s2d_pad64", lambda P: (P(dy), P(out), B, H, W, Co))
        assert torch.equal(bo.gpu[out].cpu(), bo.cpu[out])
    w = bo.t(r.randn(64, 1, 5, 5))
    wp = bo.t(np.ones(64 * 64), torch.float16)
    bo.run("hm_pack_conv_weight", lambda P: (P(w), P(wp), 11, 64, 1, 5, 5, 0, 0, 1))
    assert torch.equal(bo.gpu[wp].cpu(), bo.cpu[wp])
    g = bo.t(r.randn(9 * 16 * 64))
    o = bo.t(np.zeros(25 * 16 * 3))
    bo.run("hm_unpack_conv_wgrad", lambda P: (P(g), P(o), 10, 3, 16, 5, 5))
    bo.check(o, 1e-6, "unpack mode 10")
