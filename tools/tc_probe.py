"""Diagnostic probe (GPU): tcgen05 conv vs the SIMT gather conv on the same fp16 inputs.  Prints one line per
case with error statistics; used while bringing the tensor-core path up (the asserting version is
tests/test_tc_gpu.py)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gan-heightmaps_b200"))
import _lib  # noqa: E402

_WS = {}


def _tc_conv(dref, x1, x2, w, bias, y, y2, stream=None):
    """hm_tc_conv; through hm_tc_conv_ws with a scratch workspace (filled with NaNs first: it needs no initialisation)
    where the library asks for one (the split-K variant of the small layers)."""
    need = _lib.query("hm_tc_conv_ws_bytes", dref)
    if need > 0:
        ws = _WS.get(need)
        if ws is None:
            ws = _WS[need] = torch.empty((need + 3) // 4, dtype=torch.float32, device="cuda")
        ws.fill_(float("nan"))
        _lib.call("hm_tc_conv_ws", dref, x1, x2, w, bias, y, y2, ws.data_ptr(), ws.numel() * 4, stream)
        return
    _lib.call("hm_tc_conv", dref, x1, x2, w, bias, y, y2, stream)


CASES = [
    # name, B, H, W, C1, C2, Cout, k, pad, act
    ("1x1_64_64_w16", 2, 16, 16, 64, 0, 64, 1, 0, 0),
    ("3x3_64_64_w16", 2, 16, 16, 64, 0, 64, 3, 1, 0),
    ("5x5_64_64_w16", 2, 16, 16, 64, 0, 64, 5, 2, 1),
    ("5x5_128_64_4x4_b4", 4, 4, 4, 128, 0, 64, 5, 2, 1),
    ("5x5_64_128_w128", 1, 32, 128, 64, 0, 128, 5, 2, 1),
    ("3x3_64_64_w256", 1, 8, 256, 64, 0, 64, 3, 1, 0),
    ("5x5_128_256_w32", 2, 32, 32, 128, 0, 256, 5, 2, 2),
    ("3x3_cat128+64_512_w16", 1, 16, 16, 128, 64, 512, 3, 1, 0),
    ("2x2valid_512_512", 4, 2, 2, 512, 0, 512, 2, 0, 0),
    ("3x3_64_48_w12_ragged", 3, 10, 12, 64, 0, 48, 3, 1, 3),
    ("5x5_256_256_8x8_b3", 3, 8, 8, 256, 0, 256, 5, 2, 1),
    ("5x5_64_64_w512_rb", 1, 6, 512, 64, 0, 64, 5, 2, 1),
    ("5x5_128_128_w128_rb", 2, 20, 128, 128, 0, 128, 5, 2, 1),
    ("3x3_cat64+64_256_w256_rb", 1, 5, 256, 64, 64, 256, 3, 1, 0),
    ("5x5_64_1_w256_thin_rb", 1, 9, 256, 64, 0, 1, 5, 2, 3),
    ("5x5_64_1_w64_thin", 2, 64, 64, 64, 0, 1, 5, 2, 3),
    ("3x3_128_3_w32_thin", 2, 32, 32, 128, 0, 3, 3, 1, 4),
    ("1x1_128_12_w16_thin", 2, 16, 16, 128, 0, 12, 1, 0, 0),
    ("5x5_64_40_w16_thin", 1, 16, 16, 64, 0, 40, 5, 2, 1),
    # large enough for super tiles of 4 / 2 M tiles (S = 256/ntile while every SM still gets a work item)
    ("3x3_cat128+128_64_w256_rb_S4", 2, 160, 256, 128, 128, 64, 3, 1, 0),
    ("5x5_128_64_w256_rb_S4", 2, 160, 256, 128, 0, 64, 5, 2, 0),
    ("5x5_64_128_w256_rb_S2", 2, 160, 256, 64, 0, 128, 5, 2, 1),
    ("3x3_64_64_w64_S4", 20, 64, 64, 64, 0, 64, 3, 1, 1),
]


def desc(**kw):
    d = _lib.ConvDesc()
    for k, v in kw.items():
        setattr(d, k, v)
    return d


def run_case(name, B, H, W, C1, C2, Cout, k, pad, act, verbose=True):
    torch.manual_seed(abs(hash(name)) % 1000)
    Ct = C1 + C2
    Ho, Wo = H + 2 * pad - k + 1, W + 2 * pad - k + 1
    x1 = (torch.randn(B, H, W, C1, device="cuda")).half()
    x2 = (torch.randn(B, H, W, C2, device="cuda")).half() if C2 else None
    Wm = torch.randn(Cout, Ct, k, k, device="cuda") / np.sqrt(k * k * Ct)
    bias = torch.randn(Cout, device="cuda")
    wp = torch.empty(k * k * Ct * Cout, device="cuda", dtype=torch.float16)
    wt = torch.empty_like(wp)
    _lib.call("hm_pack_conv_weight", Wm.data_ptr(), wp.data_ptr(), 0, Cout, Ct, k, k, 0, 0, 1, None)
    _lib.call("hm_pack_conv_weight", Wm.data_ptr(), wt.data_ptr(), 5, Cout, Ct, k, k, 0, 0, 1, None)
    d = desc(dtype=1, B=B, H=H, W=W, C1=C1, C2=C2, up=0, kh=k, kw=k, stride=1, pad=pad, transposed=0, Ho=Ho, Wo=Wo,
             Cout=Cout, oH=Ho, oW=Wo, os=1, ou=0, ov=0, split=Cout, act=act, slope=0.2, accumulate=0)
    y_ref = torch.zeros(B, Ho, Wo, Cout, device="cuda", dtype=torch.float16)
    y_tc = torch.full((B, Ho, Wo, Cout), 7.0, device="cuda", dtype=torch.float16)
    p2 = x2.data_ptr() if C2 else None
    _lib.call("hm_conv_gather", C.byref(d), x1.data_ptr(), p2, wp.data_ptr(), bias.data_ptr(), y_ref.data_ptr(), None,
              None)
    _tc_conv(C.byref(d), x1.data_ptr(), p2, wt.data_ptr(), bias.data_ptr(), y_tc.data_ptr(), None, None)
    torch.cuda.synchronize()
    a, b = y_tc.float(), y_ref.float()
    err = (a - b).abs()
    scale = float(b.abs().max())
    line = "%-28s max_err %.4g  scale %.4g  rel %.3g  frac_bad %.4f  untouched %.4f" % (
        name, float(err.max()), scale, float(err.max()) / scale, float((err > 2e-2 * scale).float().mean()),
        float((a == 7.0).float().mean()))
    if verbose and float(err.max()) > 5e-3 * scale:
        bad = (err > 2e-2 * scale)
        line += "\n    bad by channel%%64 chunk(8): %s" % [round(float(bad[..., c::8].float().mean()), 3) for c in range(8)]
        line += "\n    bad by x%%8: %s" % [round(float(bad[:, :, xx::8].float().mean()), 3) for xx in range(min(8, Wo))]
        line += "\n    bad by y: %s" % [round(float(bad[:, yy].float().mean()), 3) for yy in range(min(Ho, 8))]
        line += "\n    bad by n: %s" % [round(float(bad[nn].float().mean()), 3) for nn in range(B)]
    return float(err.max()) / scale, line


UP2_CASES = [
    # name, B, H(low), W(low), Cin, Cout, act
    ("up2_64_64_w16", 2, 16, 16, 64, 64, 1),
    ("up2_128_64_w64", 1, 32, 64, 128, 64, 0),
    ("up2_64_128_w128", 1, 8, 128, 64, 128, 1),
    ("up2_64_256_w8", 3, 8, 8, 64, 256, 1),
    ("up2_64_32_w16", 2, 16, 16, 64, 32, 2),
    ("up2_64_1_thin", 2, 32, 32, 64, 1, 3),
    ("up2_128_3_thin", 1, 16, 16, 128, 3, 4),
    ("up2_512_256_4x4", 4, 4, 4, 512, 256, 1),
]


def run_up2_case(name, B, H, W, Ci, Co, act):
    """nearest-2x + 5x5 'same': tcgen05 phase-decomposed path vs the SIMT gather through the virtual upsampling."""
    torch.manual_seed(abs(hash(name)) % 1000 + 11)
    x = torch.randn(B, H, W, Ci, device="cuda").half()
    Wm = torch.randn(Co, Ci, 5, 5, device="cuda") / np.sqrt(25 * Ci)
    bias = torch.randn(Co, device="cuda")
    wp = torch.empty(25 * Ci * Co, device="cuda", dtype=torch.float16)
    w8 = torch.empty(36 * Ci * Co, device="cuda", dtype=torch.float16)
    _lib.call("hm_pack_conv_weight", Wm.data_ptr(), wp.data_ptr(), 0, Co, Ci, 5, 5, 0, 0, 1, None)
    _lib.call("hm_pack_conv_weight", Wm.data_ptr(), w8.data_ptr(), 8, Co, Ci, 5, 5, 0, 0, 1, None)
    d = desc(dtype=1, B=B, H=H, W=W, C1=Ci, C2=0, up=1, kh=5, kw=5, stride=1, pad=2, transposed=0, Ho=2 * H, Wo=2 * W,
             Cout=Co, oH=2 * H, oW=2 * W, os=1, ou=0, ov=0, split=Co, act=act, slope=0.2, accumulate=0)
    y_ref = torch.zeros(B, 2 * H, 2 * W, Co, device="cuda", dtype=torch.float16)
    y_tc = torch.full((B, 2 * H, 2 * W, Co), 7.0, device="cuda", dtype=torch.float16)
    _lib.call("hm_conv_gather", C.byref(d), x.data_ptr(), None, wp.data_ptr(), bias.data_ptr(), y_ref.data_ptr(), None,
              None)
    _tc_conv(C.byref(d), x.data_ptr(), None, w8.data_ptr(), bias.data_ptr(), y_tc.data_ptr(), None, None)
    torch.cuda.synchronize()
    a, b = y_tc.float(), y_ref.float()
    err = (a - b).abs()
    scale = float(b.abs().max())
    line = "up2 %-24s max_err %.4g  scale %.4g  rel %.3g  frac_bad %.4f  untouched %.4f" % (
        name, float(err.max()), scale, float(err.max()) / scale, float((err > 2e-2 * scale).float().mean()),
        float((a == 7.0).float().mean()))
    if float(err.max()) > 5e-3 * scale:
        bad = (err > 2e-2 * scale)
        line += "\n    bad by (oy&1, ox&1): %s" % [[round(float(bad[:, py::2, px::2].float().mean()), 3) for px in range(2)]
                                                    for py in range(2)]
        line += "\n    bad by channel%%8: %s" % [round(float(bad[..., c::8].float().mean()), 3) for c in range(min(8, Co))]
    return float(err.max()) / scale, line


S2_CASES = [
    # name, B, H, W, C1, C2, Cout
    ("s2_64_128_w32", 2, 32, 32, 64, 0, 128),
    ("s2_128_256_w128", 1, 64, 128, 128, 0, 256),
    ("s2_256_512_w16", 2, 16, 16, 256, 0, 512),
    ("s2_64+64_64_w256", 1, 16, 256, 64, 64, 64),
    ("s2_512_512_4x4", 4, 4, 4, 512, 0, 512),
    ("s2_64_64_w20_ragged", 1, 12, 20, 64, 0, 64),
]


def run_s2_case(name, B, H, W, C1, C2, Co):
    """3x3 stride-2 pad-1: tcgen05 forward / weight gradient / input gradient vs the SIMT kernels."""
    torch.manual_seed(abs(hash(name)) % 1000 + 3)
    Ct = C1 + C2
    Ho, Wo = (H + 2 - 3) // 2 + 1, (W + 2 - 3) // 2 + 1
    x1 = torch.randn(B, H, W, C1, device="cuda").half()
    x2 = torch.randn(B, H, W, C2, device="cuda").half() if C2 else None
    p2 = x2.data_ptr() if C2 else None
    Wm = torch.randn(Co, Ct, 3, 3, device="cuda") / np.sqrt(9 * Ct)
    bias = torch.randn(Co, device="cuda")
    packs = {}
    for mode, n in ((0, 9 * Ct * Co), (5, 9 * Ct * Co), (1, 9 * Ct * Co), (12, 16 * Ct * Co)):
        packs[mode] = torch.empty(n, device="cuda", dtype=torch.float16)
        _lib.call("hm_pack_conv_weight", Wm.data_ptr(), packs[mode].data_ptr(), mode, Co, Ct, 3, 3, 0, 0, 1, None)
    d = desc(dtype=1, B=B, H=H, W=W, C1=C1, C2=C2, up=0, kh=3, kw=3, stride=2, pad=1, transposed=0, Ho=Ho, Wo=Wo,
             Cout=Co, oH=Ho, oW=Wo, os=1, ou=0, ov=0, split=Co, act=1, slope=0.01, accumulate=0)
    lines = []
    worst = 0.0
    # forward
    y_ref = torch.zeros(B, Ho, Wo, Co, device="cuda", dtype=torch.float16)
    y_tc = torch.full((B, Ho, Wo, Co), 7.0, device="cuda", dtype=torch.float16)
    _lib.call("hm_conv_gather", C.byref(d), x1.data_ptr(), p2, packs[0].data_ptr(), bias.data_ptr(), y_ref.data_ptr(),
              None, None)
    _tc_conv(C.byref(d), x1.data_ptr(), p2, packs[5].data_ptr(), bias.data_ptr(), y_tc.data_ptr(), None,
              None)
    # weight gradient
    dy = torch.randn(B, Ho, Wo, Co, device="cuda").half()
    g_ref = torch.zeros(9 * Ct, Co, device="cuda")
    g_tc = torch.zeros(9 * Ct, Co, device="cuda")
    _lib.call("hm_conv_wgrad", C.byref(d), x1.data_ptr(), p2, dy.data_ptr(), g_ref.data_ptr(), None)
    wg_ok = Co % 64 == 0 and (Co <= 256 or Co % 256 == 0)
    if wg_ok:
        _lib.call("hm_tc_wgrad", C.byref(d), x1.data_ptr(), p2, dy.data_ptr(), g_tc.data_ptr(), None)
    torch.cuda.synchronize()
    pairs = [("fwd", y_tc.float(), y_ref.float())]
    if wg_ok:
        pairs.append(("wgrad", g_tc, g_ref))
    # input gradient (single source only; even sizes)
    if C2 == 0 and H % 2 == 0 and W % 2 == 0:
        dd = desc(dtype=1, B=B, H=Ho, W=Wo, C1=Co, C2=0, up=0, kh=3, kw=3, stride=2, pad=1, transposed=1, Ho=H, Wo=W,
                  Cout=Ct, oH=H, oW=W, os=1, ou=0, ov=0, split=Ct, act=0, slope=0.0, accumulate=1)
        init = torch.randn(B, H, W, Ct, device="cuda").half()
        dx_ref, dx_tc = init.clone(), init.clone()
        _lib.call("hm_conv_gather", C.byref(dd), dy.data_ptr(), None, packs[1].data_ptr(), None, dx_ref.data_ptr(), None,
                  None)
        _tc_conv(C.byref(dd), dy.data_ptr(), None, packs[12].data_ptr(), None, dx_tc.data_ptr(), None, None)
        torch.cuda.synchronize()
        pairs.append(("dgrad", dx_tc.float(), dx_ref.float()))
    for what, a, b in pairs:
        err = (a - b).abs()
        scale = float(b.abs().max())
        rel = float(err.max()) / scale
        worst = max(worst, rel)
        lines.append("s2 %-22s %-5s max_err %.4g scale %.4g rel %.3g frac_bad %.4f" % (
            name, what, float(err.max()), scale, rel, float((err > 2e-2 * scale).float().mean())))
        if rel > 5e-3 and what == "dgrad":
            bad = err > 2e-2 * scale
            lines.append("    bad by (y&1,x&1): %s" % [[round(float(bad[:, py::2, px::2].float().mean()), 3) for px in range(2)]
                                                        for py in range(2)])
    return worst, "\n".join(lines)


WGRAD_CASES = [c for c in CASES if c[6] % 64 == 0 and (c[6] <= 256 or c[6] % 256 == 0)]


def run_wgrad_case(name, B, H, W, C1, C2, Cout, k, pad, act):
    torch.manual_seed(abs(hash(name)) % 1000 + 5)
    Ct = C1 + C2
    Ho, Wo = H + 2 * pad - k + 1, W + 2 * pad - k + 1
    x1 = (torch.randn(B, H, W, C1, device="cuda")).half()
    x2 = (torch.randn(B, H, W, C2, device="cuda")).half() if C2 else None
    dy = (torch.randn(B, Ho, Wo, Cout, device="cuda")).half()
    d = desc(dtype=1, B=B, H=H, W=W, C1=C1, C2=C2, up=0, kh=k, kw=k, stride=1, pad=pad, transposed=0, Ho=Ho, Wo=Wo,
             Cout=Cout, oH=Ho, oW=Wo, os=1, ou=0, ov=0, split=Cout, act=0, slope=0.0, accumulate=0)
    ref = torch.zeros(k * k * Ct, Cout, device="cuda")
    out = torch.zeros(k * k * Ct, Cout, device="cuda")
    p2 = x2.data_ptr() if C2 else None
    _lib.call("hm_conv_wgrad", C.byref(d), x1.data_ptr(), p2, dy.data_ptr(), ref.data_ptr(), None)
    _lib.call("hm_tc_wgrad", C.byref(d), x1.data_ptr(), p2, dy.data_ptr(), out.data_ptr(), None)
    torch.cuda.synchronize()
    err = (out - ref).abs()
    scale = float(ref.abs().max())
    line = "wgrad %-28s max_err %.4g  scale %.4g  rel %.3g  frac_bad %.4f  zero_frac %.4f" % (
        name, float(err.max()), scale, float(err.max()) / scale, float((err > 2e-2 * scale).float().mean()),
        float((out == 0).float().mean()))
    if float(err.max()) > 5e-3 * scale:
        bad = (err > 2e-2 * scale).view(k * k, Ct, Cout)
        line += "\n    bad by tap: %s" % [round(float(bad[t].float().mean()), 2) for t in range(k * k)]
        line += "\n    bad by ci%%64 (8 bins): %s" % [round(float(bad[:, c::8].float().mean()), 2) for c in range(8)]
        line += "\n    bad by co (8 bins): %s" % [round(float(bad[:, :, c::8].float().mean()), 2) for c in range(8)]
        line += "\n    ratio out/ref median: %.4g" % float((out / (ref + 1e-20)).median())
    return float(err.max()) / scale, line


# ---- HM_BF16X3 ("tc32"): the same tcgen05 kernels on three-plane bf16 splits of fp32 tensors vs the SIMT fp32 kernels ----
def _split(t, layout, c1=None):
    """hm_split_bf16x3 of a float32 [..., C] tensor -> bf16 tensor with six planes along the reduction axis."""
    t = t.contiguous()
    Cn = t.shape[-1]
    rows = t.numel() // Cn
    out = torch.empty(6 * t.numel(), device="cuda", dtype=torch.bfloat16)
    _lib.call("hm_split_bf16x3", t.data_ptr(), out.data_ptr(), rows, Cn, Cn if c1 is None else c1, layout, None)
    return out


def _pack32(Wm, mode, Co, Ci, kh, kw, K, c1=None):
    """fp32 pack of `mode`, then the b-side split along its innermost extent K."""
    n = int(_lib.query("hm_pack_conv_weight_count", mode, Co, Ci, kh, kw))
    tmp = torch.empty(n, device="cuda", dtype=torch.float32)
    _lib.call("hm_pack_conv_weight", Wm.data_ptr(), tmp.data_ptr(), mode, Co, Ci, kh, kw, 0, 0, 0, None)
    return _split(tmp.view(n // K, K), 1, c1)


def _cmp32(tag, name, a, b):
    err = (a - b).abs()
    scale = float(b.abs().max())
    rel = float(err.max()) / scale
    return rel, "%s %-28s max_err %.4g scale %.4g rel %.3g" % (tag, name, float(err.max()), scale, rel)


def run_case32(name, B, H, W, C1, C2, Cout, k, pad, act, verbose=True):
    torch.manual_seed(abs(hash(name)) % 1000)
    Ct = C1 + C2
    Ho, Wo = H + 2 * pad - k + 1, W + 2 * pad - k + 1
    x1 = torch.randn(B, H, W, C1, device="cuda")
    x2 = torch.randn(B, H, W, C2, device="cuda") if C2 else None
    Wm = torch.randn(Cout, Ct, k, k, device="cuda") / np.sqrt(k * k * Ct)
    bias = torch.randn(Cout, device="cuda")
    wp = torch.empty(k * k * Ct * Cout, device="cuda")
    _lib.call("hm_pack_conv_weight", Wm.data_ptr(), wp.data_ptr(), 0, Cout, Ct, k, k, 0, 0, 0, None)
    wt = _pack32(Wm, 5, Cout, Ct, k, k, Ct, C1)
    d = desc(dtype=0, B=B, H=H, W=W, C1=C1, C2=C2, up=0, kh=k, kw=k, stride=1, pad=pad, transposed=0, Ho=Ho, Wo=Wo,
             Cout=Cout, oH=Ho, oW=Wo, os=1, ou=0, ov=0, split=Cout, act=act, slope=0.2, accumulate=0)
    y_ref = torch.zeros(B, Ho, Wo, Cout, device="cuda")
    y_tc = torch.full((B, Ho, Wo, Cout), 7.0, device="cuda")
    p2 = x2.data_ptr() if C2 else None
    _lib.call("hm_conv_gather", C.byref(d), x1.data_ptr(), p2, wp.data_ptr(), bias.data_ptr(), y_ref.data_ptr(), None, None)
    s1 = _split(x1, 0)
    s2 = _split(x2, 0) if C2 else None
    d2 = desc(**{f: getattr(d, f) for f, _ in d._fields_})
    d2.dtype, d2.C1, d2.C2 = 2, 6 * C1, 6 * C2
    _lib.call("hm_tc_conv", C.byref(d2), s1.data_ptr(), s2.data_ptr() if C2 else None, wt.data_ptr(), bias.data_ptr(),
              y_tc.data_ptr(), None, None)
    torch.cuda.synchronize()
    return _cmp32("tc32", name, y_tc, y_ref)


def run_wgrad_case32(name, B, H, W, C1, C2, Cout, k, pad, act):
    torch.manual_seed(abs(hash(name)) % 1000 + 5)
    Ct = C1 + C2
    Ho, Wo = H + 2 * pad - k + 1, W + 2 * pad - k + 1
    x1 = torch.randn(B, H, W, C1, device="cuda")
    x2 = torch.randn(B, H, W, C2, device="cuda") if C2 else None
    dy = torch.randn(B, Ho, Wo, Cout, device="cuda")
    d = desc(dtype=0, B=B, H=H, W=W, C1=C1, C2=C2, up=0, kh=k, kw=k, stride=1, pad=pad, transposed=0, Ho=Ho, Wo=Wo,
             Cout=Cout, oH=Ho, oW=Wo, os=1, ou=0, ov=0, split=Cout, act=0, slope=0.0, accumulate=0)
    ref = torch.zeros(k * k * Ct, Cout, device="cuda")
    out = torch.zeros(k * k * Ct, Cout, device="cuda")
    p2 = x2.data_ptr() if C2 else None
    _lib.call("hm_conv_wgrad", C.byref(d), x1.data_ptr(), p2, dy.data_ptr(), ref.data_ptr(), None)
    s1, s2, sd = _split(x1, 2), (_split(x2, 2) if C2 else None), _split(dy, 3)
    d2 = desc(**{f: getattr(d, f) for f, _ in d._fields_})
    d2.dtype, d2.B = 2, 6 * B
    _lib.call("hm_tc_wgrad", C.byref(d2), s1.data_ptr(), s2.data_ptr() if C2 else None, sd.data_ptr(), out.data_ptr(), None)
    torch.cuda.synchronize()
    return _cmp32("tc32 wgrad", name, out, ref)


def run_up2_case32(name, B, H, W, Ci, Co, act):
    torch.manual_seed(abs(hash(name)) % 1000 + 11)
    x = torch.randn(B, H, W, Ci, device="cuda")
    Wm = torch.randn(Co, Ci, 5, 5, device="cuda") / np.sqrt(25 * Ci)
    bias = torch.randn(Co, device="cuda")
    wp = torch.empty(25 * Ci * Co, device="cuda")
    _lib.call("hm_pack_conv_weight", Wm.data_ptr(), wp.data_ptr(), 0, Co, Ci, 5, 5, 0, 0, 0, None)
    w8 = _pack32(Wm, 8, Co, Ci, 5, 5, Ci)
    d = desc(dtype=0, B=B, H=H, W=W, C1=Ci, C2=0, up=1, kh=5, kw=5, stride=1, pad=2, transposed=0, Ho=2 * H, Wo=2 * W,
             Cout=Co, oH=2 * H, oW=2 * W, os=1, ou=0, ov=0, split=Co, act=act, slope=0.2, accumulate=0)
    y_ref = torch.zeros(B, 2 * H, 2 * W, Co, device="cuda")
    y_tc = torch.full((B, 2 * H, 2 * W, Co), 7.0, device="cuda")
    _lib.call("hm_conv_gather", C.byref(d), x.data_ptr(), None, wp.data_ptr(), bias.data_ptr(), y_ref.data_ptr(), None, None)
    sx = _split(x, 0)
    d2 = desc(**{f: getattr(d, f) for f, _ in d._fields_})
    d2.dtype, d2.C1 = 2, 6 * Ci
    _lib.call("hm_tc_conv", C.byref(d2), sx.data_ptr(), None, w8.data_ptr(), bias.data_ptr(), y_tc.data_ptr(), None, None)
    torch.cuda.synchronize()
    return _cmp32("tc32 up2", name, y_tc, y_ref)


def run_s2_case32(name, B, H, W, C1, C2, Co):
    torch.manual_seed(abs(hash(name)) % 1000 + 3)
    Ct = C1 + C2
    Ho, Wo = (H + 2 - 3) // 2 + 1, (W + 2 - 3) // 2 + 1
    x1 = torch.randn(B, H, W, C1, device="cuda")
    x2 = torch.randn(B, H, W, C2, device="cuda") if C2 else None
    p2 = x2.data_ptr() if C2 else None
    Wm = torch.randn(Co, Ct, 3, 3, device="cuda") / np.sqrt(9 * Ct)
    bias = torch.randn(Co, device="cuda")
    packs = {}
    for mode in (0, 1):
        packs[mode] = torch.empty(9 * Ct * Co, device="cuda")
        _lib.call("hm_pack_conv_weight", Wm.data_ptr(), packs[mode].data_ptr(), mode, Co, Ct, 3, 3, 0, 0, 0, None)
    w5, w12 = _pack32(Wm, 5, Co, Ct, 3, 3, Ct, C1), _pack32(Wm, 12, Co, Ct, 3, 3, Co)
    d = desc(dtype=0, B=B, H=H, W=W, C1=C1, C2=C2, up=0, kh=3, kw=3, stride=2, pad=1, transposed=0, Ho=Ho, Wo=Wo,
             Cout=Co, oH=Ho, oW=Wo, os=1, ou=0, ov=0, split=Co, act=1, slope=0.01, accumulate=0)
    y_ref = torch.zeros(B, Ho, Wo, Co, device="cuda")
    y_tc = torch.full((B, Ho, Wo, Co), 7.0, device="cuda")
    _lib.call("hm_conv_gather", C.byref(d), x1.data_ptr(), p2, packs[0].data_ptr(), bias.data_ptr(), y_ref.data_ptr(), None, None)
    s1, s2 = _split(x1, 0), (_split(x2, 0) if C2 else None)
    d2 = desc(**{f: getattr(d, f) for f, _ in d._fields_})
    d2.dtype, d2.C1, d2.C2 = 2, 6 * C1, 6 * C2
    _lib.call("hm_tc_conv", C.byref(d2), s1.data_ptr(), s2.data_ptr() if C2 else None, w5.data_ptr(), bias.data_ptr(),
              y_tc.data_ptr(), None, None)
    dy = torch.randn(B, Ho, Wo, Co, device="cuda")
    g_ref = torch.zeros(9 * Ct, Co, device="cuda")
    g_tc = torch.zeros(9 * Ct, Co, device="cuda")
    _lib.call("hm_conv_wgrad", C.byref(d), x1.data_ptr(), p2, dy.data_ptr(), g_ref.data_ptr(), None)
    pairs = [("fwd", y_tc, y_ref)]
    if Co % 64 == 0 and (Co <= 256 or Co % 256 == 0):
        t1, t2, td = _split(x1, 2), (_split(x2, 2) if C2 else None), _split(dy, 3)
        d3 = desc(**{f: getattr(d, f) for f, _ in d._fields_})
        d3.dtype, d3.B = 2, 6 * B
        _lib.call("hm_tc_wgrad", C.byref(d3), t1.data_ptr(), t2.data_ptr() if C2 else None, td.data_ptr(), g_tc.data_ptr(), None)
        pairs.append(("wgrad", g_tc, g_ref))
    if C2 == 0 and H % 2 == 0 and W % 2 == 0:
        dd = desc(dtype=0, B=B, H=Ho, W=Wo, C1=Co, C2=0, up=0, kh=3, kw=3, stride=2, pad=1, transposed=1, Ho=H, Wo=W,
                  Cout=Ct, oH=H, oW=W, os=1, ou=0, ov=0, split=Ct, act=0, slope=0.0, accumulate=1)
        init = torch.randn(B, H, W, Ct, device="cuda")
        dx_ref, dx_tc = init.clone(), init.clone()
        _lib.call("hm_conv_gather", C.byref(dd), dy.data_ptr(), None, packs[1].data_ptr(), None, dx_ref.data_ptr(), None, None)
        sdy = _split(dy, 0)
        dd2 = desc(**{f: getattr(dd, f) for f, _ in dd._fields_})
        dd2.dtype, dd2.C1 = 2, 6 * Co
        _lib.call("hm_tc_conv", C.byref(dd2), sdy.data_ptr(), None, w12.data_ptr(), None, dx_tc.data_ptr(), None, None)
        pairs.append(("dgrad", dx_tc, dx_ref))
    torch.cuda.synchronize()
    worst, lines = 0.0, []
    for what, a, b in pairs:
        rel, line = _cmp32("tc32 s2 " + what, name, a, b)
        worst = max(worst, rel)
        lines.append(line)
    return worst, "\n".join(lines)


UP2_BWD_CASES = [c for c in UP2_CASES if c[5] % 64 == 0]


def run_up2_bwd_case(name, B, H, W, Ci, Co, act, tc32=False):
    """Backward of nearest-2x + 5x5 'same': the input gradient as ONE 6x6 stride-2 convolution of dy (pack mode 20) and
    the weight gradient in phase form (hm_tc_wgrad with the forward descriptor, unpack mode 8), against the SIMT kernels
    through the virtual upsampling (high-res input gradient + hm_upsample2_bwd; gather weight gradient, unpack mode 0).
    tc32: float32 tensors, tensor-core operands as three-plane bf16 splits."""
    torch.manual_seed(abs(hash(name)) % 1000 + 17)
    dt, tdt = (0, torch.float32) if tc32 else (1, torch.float16)
    x = torch.randn(B, H, W, Ci, device="cuda").to(tdt)
    dy = torch.randn(B, 2 * H, 2 * W, Co, device="cuda").to(tdt)
    Wm = torch.randn(Co, Ci, 5, 5, device="cuda") / np.sqrt(25 * Ci)
    fwd = desc(dtype=dt, B=B, H=H, W=W, C1=Ci, C2=0, up=1, kh=5, kw=5, stride=1, pad=2, transposed=0, Ho=2 * H, Wo=2 * W,
               Cout=Co, oH=2 * H, oW=2 * W, os=1, ou=0, ov=0, split=Co, act=0, slope=0.0, accumulate=0)
    # ---- weight gradient
    ref_p = torch.zeros(25 * Ci, Co, device="cuda")
    _lib.call("hm_conv_wgrad", C.byref(fwd), x.data_ptr(), None, dy.data_ptr(), ref_p.data_ptr(), None)
    ref_w = torch.empty(Co, Ci, 5, 5, device="cuda")
    _lib.call("hm_unpack_conv_wgrad", ref_p.data_ptr(), ref_w.data_ptr(), 0, Co, Ci, 5, 5, None)
    out_p = torch.zeros(9 * Ci, 4 * Co, device="cuda")
    if tc32:
        sx, sd = _split(x, 2), _split(dy, 3)
        d2 = desc(**{f: getattr(fwd, f) for f, _ in fwd._fields_})
        d2.dtype, d2.B = 2, 6 * B
        _lib.call("hm_tc_wgrad", C.byref(d2), sx.data_ptr(), None, sd.data_ptr(), out_p.data_ptr(), None)
    else:
        _lib.call("hm_tc_wgrad", C.byref(fwd), x.data_ptr(), None, dy.data_ptr(), out_p.data_ptr(), None)
    out_w = torch.empty(Co, Ci, 5, 5, device="cuda")
    _lib.call("hm_unpack_conv_wgrad", out_p.data_ptr(), out_w.data_ptr(), 8, Co, Ci, 5, 5, None)
    # ---- input gradient
    wp1 = torch.empty(25 * Ci * Co, device="cuda", dtype=tdt)
    _lib.call("hm_pack_conv_weight", Wm.data_ptr(), wp1.data_ptr(), 1, Co, Ci, 5, 5, 0, 0, dt, None)
    dg = desc(dtype=dt, B=B, H=2 * H, W=2 * W, C1=Co, C2=0, up=0, kh=5, kw=5, stride=1, pad=2, transposed=1, Ho=2 * H,
              Wo=2 * W, Cout=Ci, oH=2 * H, oW=2 * W, os=1, ou=0, ov=0, split=Ci, act=0, slope=0.0, accumulate=0)
    gup = torch.zeros(B, 2 * H, 2 * W, Ci, device="cuda", dtype=tdt)
    _lib.call("hm_conv_gather", C.byref(dg), dy.data_ptr(), None, wp1.data_ptr(), None, gup.data_ptr(), None, None)
    ref_x = torch.zeros(B, H, W, Ci, device="cuda", dtype=tdt)
    _lib.call("hm_upsample2_bwd", gup.data_ptr(), ref_x.data_ptr(), dt, B, H, W, Ci, 1, 0, None)
    d6 = desc(dtype=dt, B=B, H=2 * H, W=2 * W, C1=Co, C2=0, up=0, kh=6, kw=6, stride=2, pad=2, transposed=0, Ho=H, Wo=W,
              Cout=Ci, oH=H, oW=W, os=1, ou=0, ov=0, split=Ci, act=0, slope=0.0, accumulate=0)
    out_x = torch.full((B, H, W, Ci), 7.0, device="cuda", dtype=tdt)
    if tc32:
        w20 = _pack32(Wm, 20, Co, Ci, 5, 5, Co)
        sdy = _split(dy, 0)
        d62 = desc(**{f: getattr(d6, f) for f, _ in d6._fields_})
        d62.dtype, d62.C1 = 2, 6 * Co
        _lib.call("hm_tc_conv", C.byref(d62), sdy.data_ptr(), None, w20.data_ptr(), None, out_x.data_ptr(), None, None)
    else:
        w20 = torch.empty(36 * Ci * Co, device="cuda", dtype=torch.float16)
        _lib.call("hm_pack_conv_weight", Wm.data_ptr(), w20.data_ptr(), 20, Co, Ci, 5, 5, 0, 0, 1, None)
        _tc_conv(C.byref(d6), dy.data_ptr(), None, w20.data_ptr(), None, out_x.data_ptr(), None, None)
    torch.cuda.synchronize()
    r1, l1 = _cmp32("up2 wgrad" + (" tc32" if tc32 else ""), name, out_w, ref_w)
    r2, l2 = _cmp32("up2 dgrad" + (" tc32" if tc32 else ""), name, out_x.float(), ref_x.float())
    return max(r1, r2), l1 + "\n" + l2


POOL_CASES = [
    # name, B, H, W, C1, C2, Cout, k, pad, act
    ("pool_5x5_64_128_w256", 2, 8, 256, 64, 0, 128, 5, 2, 1),
    ("pool_5x5_128_128_w128", 3, 6, 128, 128, 0, 128, 5, 2, 1),
    ("pool_3x3_64_64_w384", 1, 4, 384, 64, 0, 64, 3, 1, 2),
    ("pool_5x5_64_256_w128", 1, 4, 128, 64, 0, 256, 5, 2, 0),
    ("pool_3x3_cat64+64_96_w130", 2, 2, 130, 64, 64, 96, 3, 1, 1),
]


def run_pool_case(name, B, H, W, C1, C2, Cout, k, pad, act):
    """hm_tc_conv_pool (conv + bias + activation + 2x2 max-pool in the tensor-core epilogue) against hm_tc_conv followed by
    hm_maxpool2_fwd on the same data.  The pooled VALUES must be identical (rounding to fp16 is monotonic, so the max of
    the rounded values is the rounded max); the argmax may differ only where two candidates round to the same fp16 value
    (the fused kernel compares the fp32 values), so it is checked by value: the un-pooled tensor at the reported position
    equals the pooled value."""
    torch.manual_seed(abs(hash(name)) % 1000 + 23)
    Ct = C1 + C2
    x1 = torch.randn(B, H, W, C1, device="cuda").half()
    x2 = torch.randn(B, H, W, C2, device="cuda").half() if C2 else None
    p2 = x2.data_ptr() if C2 else None
    Wm = torch.randn(Cout, Ct, k, k, device="cuda") / np.sqrt(k * k * Ct)
    bias = torch.randn(Cout, device="cuda")
    wt = torch.empty(k * k * Ct * Cout, device="cuda", dtype=torch.float16)
    _lib.call("hm_pack_conv_weight", Wm.data_ptr(), wt.data_ptr(), 5, Cout, Ct, k, k, 0, 0, 1, None)
    d = desc(dtype=1, B=B, H=H, W=W, C1=C1, C2=C2, up=0, kh=k, kw=k, stride=1, pad=pad, transposed=0, Ho=H, Wo=W,
             Cout=Cout, oH=H, oW=W, os=1, ou=0, ov=0, split=Cout, act=act, slope=0.2, accumulate=0)
    assert _lib.query("hm_tc_conv_pool_supported", C.byref(d)) == 1, name
    full = torch.zeros(B, H, W, Cout, device="cuda", dtype=torch.float16)
    _lib.call("hm_tc_conv", C.byref(d), x1.data_ptr(), p2, wt.data_ptr(), bias.data_ptr(), full.data_ptr(), None, None)
    ref = torch.zeros(B, H // 2, W // 2, Cout, device="cuda", dtype=torch.float16)
    ref_i = torch.zeros(B, H // 2, W // 2, Cout, device="cuda", dtype=torch.uint8)
    _lib.call("hm_maxpool2_fwd", full.data_ptr(), ref.data_ptr(), ref_i.data_ptr(), 1, B, H, W, Cout, None)
    out = torch.full((B, H // 2, W // 2, Cout), 7.0, device="cuda", dtype=torch.float16)
    out_i = torch.full((B, H // 2, W // 2, Cout), 9, device="cuda", dtype=torch.uint8)
    _lib.call("hm_tc_conv_pool", C.byref(d), x1.data_ptr(), p2, wt.data_ptr(), bias.data_ptr(), out.data_ptr(),
              out_i.data_ptr(), None)
    torch.cuda.synchronize()
    dv = float((out.float() - ref.float()).abs().max())
    win = full.view(B, H // 2, 2, W // 2, 2, Cout).permute(0, 1, 3, 5, 2, 4).reshape(B, H // 2, W // 2, Cout, 4)
    ok_range = bool((out_i <= 3).all())
    picked = torch.gather(win, 4, out_i.clamp(max=3).long().unsqueeze(-1)).squeeze(-1)
    di = float((picked.float() - out.float()).abs().max())
    same_idx = float((out_i == ref_i).float().mean())
    bad = 0.0 if (dv == 0.0 and di == 0.0 and ok_range) else 1.0
    return bad, "pool %-28s value diff %.3g  value-at-argmax diff %.3g  argmax in range %s  same argmax %.4f" % (
        name, dv, di, ok_range, same_idx)


DC2_CASES = [
    # name, B, H, W (input grid), C1, C2, Cout, act
    ("dc2_128_3_w256_tanh", 1, 256, 256, 64, 64, 3, 4),
    ("dc2_64_1_w16", 2, 16, 16, 64, 0, 1, 0),
    ("dc2_128_3_ragged_10x12", 3, 10, 12, 128, 0, 3, 0),
]


def run_dc2_case(name, B, H, W, C1, C2, Co, act):
    """Deconv2DLayer 2x2 stride 2 through hm_tc_conv (transposed == 2, pack mode 17) and its input gradient
    (hm_s2d_pad64 + 1x1 hm_tc_conv with pack mode 18) against torch conv_transpose2d in float32 on the same fp16 data."""
    import torch.nn.functional as F
    torch.manual_seed(abs(hash(name)) % 1000)
    Ci = C1 + C2
    x1 = torch.randn(B, H, W, C1, device="cuda").half()
    x2 = torch.randn(B, H, W, C2, device="cuda").half() if C2 else None
    Wm = (torch.randn(Ci, Co, 2, 2, device="cuda") / np.sqrt(Ci)).half().float()
    bias = torch.randn(Co, device="cuda") * 0.1
    wt = torch.empty(4 * Co * Ci, device="cuda", dtype=torch.float16)
    wd = torch.empty(64 * Ci, device="cuda", dtype=torch.float16)
    _lib.call("hm_pack_conv_weight", Wm.data_ptr(), wt.data_ptr(), 17, Co, Ci, 2, 2, 0, 0, 1, None)
    _lib.call("hm_pack_conv_weight", Wm.data_ptr(), wd.data_ptr(), 18, Co, Ci, 2, 2, 0, 0, 1, None)
    d = desc(dtype=1, B=B, H=H, W=W, C1=C1, C2=C2, up=0, kh=2, kw=2, stride=2, pad=0, transposed=2, Ho=2 * H, Wo=2 * W,
             Cout=Co, oH=2 * H, oW=2 * W, os=1, ou=0, ov=0, split=Co, act=act, slope=0.2, accumulate=0)
    y = torch.full((B, 2 * H, 2 * W, Co), 7.0, device="cuda", dtype=torch.float16)
    _tc_conv(C.byref(d), x1.data_ptr(), x2.data_ptr() if C2 else None, wt.data_ptr(), bias.data_ptr(),
              y.data_ptr(), None, None)
    xin = torch.cat([x1, x2], 3) if C2 else x1
    # Lasagne Deconv2DLayer(flip_filters=False) == conv_transpose2d with the filter rotated by 180 degrees
    ref = F.conv_transpose2d(xin.float().permute(0, 3, 1, 2), Wm.flip(2, 3), bias, stride=2)
    if act == 4:
        ref = torch.tanh(ref)
    ref = ref.permute(0, 2, 3, 1)
    torch.cuda.synchronize()
    e_f = float((y.float() - ref).abs().max()) / float(ref.abs().max())
    # input gradient
    dy = torch.randn(B, 2 * H, 2 * W, Co, device="cuda").half()
    dy64 = torch.empty(B, H, W, 64, device="cuda", dtype=torch.float16)
    _lib.call("hm_s2d_pad64", dy.data_ptr(), dy64.data_ptr(), B, H, W, Co, None)
    g1 = torch.full((B, H, W, C1), 7.0, device="cuda", dtype=torch.float16)
    g2 = torch.full((B, H, W, max(C2, 1)), 7.0, device="cuda", dtype=torch.float16)
    dd = desc(dtype=1, B=B, H=H, W=W, C1=64, C2=0, up=0, kh=1, kw=1, stride=1, pad=0, transposed=0, Ho=H, Wo=W, Cout=Ci,
              oH=H, oW=W, os=1, ou=0, ov=0, split=C1, act=0, slope=0.0, accumulate=0)
    _tc_conv(C.byref(dd), dy64.data_ptr(), None, wd.data_ptr(), None, g1.data_ptr(),
              g2.data_ptr() if C2 else None, None)
    gref = F.conv2d(dy.float().permute(0, 3, 1, 2), Wm.flip(2, 3), stride=2).permute(0, 2, 3, 1)      # adjoint of the above
    torch.cuda.synchronize()
    got = torch.cat([g1, g2], 3) if C2 else g1
    e_g = float((got.float() - gref).abs().max()) / float(gref.abs().max())
    return max(e_f, e_g), "%-26s fwd %.3g  dgrad %.3g" % (name, e_f, e_g)


C1_CASES = [
    # name, B, H, W, pooled
    ("c1_pool_64", 2, 64, 64, 1),
    ("c1_pool_512", 1, 512, 512, 1),
    ("c1_pool_ragged_24x40", 3, 24, 40, 1),
    ("c1_pool_2x260", 2, 2, 260, 1),
    ("c1_dgrad_64", 2, 64, 64, 0),
    ("c1_dgrad_512", 1, 512, 512, 0),
    ("c1_dgrad_ragged_24x40", 3, 24, 40, 0),
]


def run_c1_case(name, B, H, W, pooled):
    """hm_c1s2_conv against torch (fp32 on the same fp16-rounded data).
    pooled: conv5x5(1->64, true convolution) + bias + LeakyReLU(0.2) + 2x2 max-pool (discriminator layer 1);
    plain:  input gradient of nearest-2x -> conv5x5(64->1) (generator output layer) for a given dy[B,H,W]."""
    import torch.nn.functional as F
    torch.manual_seed(abs(hash(name)) % 1000)
    if pooled:
        x = torch.randn(B, H, W, device="cuda").half()
        Wm = torch.randn(64, 1, 5, 5, device="cuda") * 0.2
        bias = torch.randn(64, device="cuda") * 0.1
        wk = torch.empty(256 * 64, device="cuda", dtype=torch.float16)
        _lib.call("hm_pack_conv_weight", Wm.data_ptr(), wk.data_ptr(), 15, 64, 1, 5, 5, 0, 0, 1, None)
        y = torch.full((B, H // 2, W // 2, 64), 7.0, device="cuda", dtype=torch.float16)
        idx = torch.full((B, H // 2, W // 2, 64), 9, device="cuda", dtype=torch.uint8)
        _lib.call("hm_c1s2_conv", x.data_ptr(), wk.data_ptr(), bias.data_ptr(), y.data_ptr(), idx.data_ptr(), B, H, W,
                  256, 1, 0.2, None)
        torch.cuda.synchronize()
        full = F.leaky_relu(F.conv2d(x.float()[:, None], Wm.half().float().flip(2, 3), bias, padding=2), 0.2)   # [B,64,H,W]
        ref = F.max_pool2d(full, 2).permute(0, 2, 3, 1)
        a = y.float()
        scale = float(ref.abs().max())
        err = float((a - ref).abs().max()) / scale
        # the argmax must point at an element that attains the pooled value (ties / fp16-near-ties may differ)
        assert int(idx.max()) <= 7, "argmax byte out of range (untouched output?)"
        # bit 2 of the byte: the stored value is on the slope-1 side of the activation (hm_c1s2_bwd takes act' from it)
        assert bool(((idx >> 2) == (y.float() >= 0).to(torch.uint8)).all()), "act' bit of the argmax byte"
        k = (idx & 3).long()
        fw = full.permute(0, 2, 3, 1).reshape(B, H // 2, 2, W // 2, 2, 64).permute(0, 1, 3, 5, 2, 4).reshape(B, H // 2, W // 2, 64, 4)
        picked = torch.gather(fw, 4, k[..., None])[..., 0]
        ierr = float((picked - ref).abs().max()) / scale
        agree = float((fw.argmax(4) == k).float().mean())
        line = "%-26s max_err/scale %.3g  picked-vs-max %.3g  argmax agreement %.4f" % (name, err, ierr, agree)
        return max(err, ierr), line
    dy = torch.randn(B, H, W, device="cuda").half()
    Wm = torch.randn(1, 64, 5, 5, device="cuda") * 0.1
    wk = torch.empty(64 * 64, device="cuda", dtype=torch.float16)
    _lib.call("hm_pack_conv_weight", Wm.data_ptr(), wk.data_ptr(), 14, 1, 64, 5, 5, 0, 0, 1, None)
    dx = torch.full((B, H // 2, W // 2, 64), 7.0, device="cuda", dtype=torch.float16)
    _lib.call("hm_c1s2_conv", dy.data_ptr(), wk.data_ptr(), None, dx.data_ptr(), None, B, H, W, 64, 0, 0.0, None)
    torch.cuda.synchronize()
    xl = torch.zeros(B, 64, H // 2, W // 2, device="cuda", requires_grad=True)
    out = F.conv2d(F.interpolate(xl, scale_factor=2, mode="nearest"), Wm.flip(2, 3), padding=2)
    out.backward(dy.float()[:, None])
    ref = xl.grad.permute(0, 2, 3, 1)
    scale = float(ref.abs().max())
    err = float((dx.float() - ref).abs().max()) / scale
    return err, "%-26s max_err/scale %.3g" % (name, err)


def run_c1wg_case(name, B, H, W):
    """hm_c1s2_wgrad + hm_unpack_conv_wgrad(mode 14): weight gradient of nearest-2x -> conv5x5(64 -> 1) (the generator's
    output layer) from dy[B,H,W] and the low-res source x[B,H/2,W/2,64], against torch autograd in float32 on the same
    fp16-rounded data.  Returns (max error / scale, line)."""
    import torch.nn.functional as F
    torch.manual_seed(abs(hash(name)) % 1000)
    dy = torch.randn(B, H, W, device="cuda").half()
    x = torch.randn(B, H // 2, W // 2, 64, device="cuda").half()
    dwk = torch.zeros(64 * 64, device="cuda")
    dw = torch.full((1, 64, 5, 5), 7.0, device="cuda")
    _lib.call("hm_c1s2_wgrad", dy.data_ptr(), x.data_ptr(), dwk.data_ptr(), B, H, W, None)
    _lib.call("hm_unpack_conv_wgrad", dwk.data_ptr(), dw.data_ptr(), 14, 1, 64, 5, 5, None)
    torch.cuda.synchronize()
    Wm = torch.zeros(1, 64, 5, 5, device="cuda", requires_grad=True)
    xu = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2, mode="nearest")
    out = F.conv2d(xu, Wm.flip(2, 3), padding=2)                  # Lasagne's Conv2DLayer convolves (flipped filter)
    out.backward(dy.float()[:, None])
    ref = Wm.grad
    scale = float(ref.abs().max())
    err = float((dw - ref).abs().max()) / scale
    return err, "%-26s max_err/scale %.3g" % (name, err)


C1B_CASES = [
    # name, B, H, W
    ("c1bwd_64", 2, 64, 64),
    ("c1bwd_512", 2, 512, 512),
    ("c1bwd_ragged_24x40", 3, 24, 40),
    ("c1bwd_2x260", 2, 2, 260),
]


def run_c1bwd_case(name, B, H, W):
    """hm_c1s2_bwd (+ fold, col2im) against the float32 adjoints of conv5x5(1->64)+bias+LeakyReLU(0.2)+max-pool on the
    same fp16 data; the forward pass (pooled values, argmax routing) comes from hm_c1s2_conv."""
    import torch.nn.functional as F
    torch.manual_seed(abs(hash(name)) % 1000)
    x = torch.randn(B, H, W, device="cuda").half()
    Wm = (torch.randn(64, 1, 5, 5, device="cuda") * 0.2).half().float()
    bias = torch.randn(64, device="cuda") * 0.1
    Hq, Wq = H // 2, W // 2
    wk = torch.empty(256 * 64, device="cuda", dtype=torch.float16)
    wk2 = torch.empty(256 * 64, device="cuda", dtype=torch.float16)
    _lib.call("hm_pack_conv_weight", Wm.data_ptr(), wk.data_ptr(), 15, 64, 1, 5, 5, 0, 0, 1, None)
    _lib.call("hm_pack_conv_weight", Wm.data_ptr(), wk2.data_ptr(), 16, 64, 1, 5, 5, 0, 0, 1, None)
    pooled = torch.empty(B, Hq, Wq, 64, device="cuda", dtype=torch.float16)
    idx = torch.empty(B, Hq, Wq, 64, device="cuda", dtype=torch.uint8)
    _lib.call("hm_c1s2_conv", x.data_ptr(), wk.data_ptr(), bias.data_ptr(), pooled.data_ptr(), idx.data_ptr(), B, H, W,
              256, 1, 0.2, None)
    g = torch.randn(B, Hq, Wq, 64, device="cuda").half()
    dwk = torch.zeros(256 * 64, device="cuda")
    u = torch.full((B, Hq, Wq, 64), 7.0, device="cuda", dtype=torch.float16)
    dx = torch.full((B, H, W), 7.0, device="cuda", dtype=torch.float16)
    dw = torch.zeros(64 * 25, device="cuda")
    db = torch.zeros(64, device="cuda")
    # the two uses of the step: weight gradient only, then input gradient only; and both at once must agree
    _lib.call("hm_c1s2_bwd", x.data_ptr(), g.data_ptr(), pooled.data_ptr(), idx.data_ptr(), None, dwk.data_ptr(), None,
              None, B, H, W, 1, 0.2, None)
    _lib.call("hm_c1s2_bwd", None, g.data_ptr(), pooled.data_ptr(), idx.data_ptr(), wk2.data_ptr(), None, u.data_ptr(),
              None, B, H, W, 1, 0.2, None)
    _lib.call("hm_c1s2_bwd_fold", dwk.data_ptr(), dw.data_ptr(), db.data_ptr(), 64, None)
    _lib.call("hm_c1s2_col2im", u.data_ptr(), dx.data_ptr(), B, H, W, None)
    dwk2 = torch.zeros(256 * 64, device="cuda")
    u2 = torch.zeros_like(u)
    _lib.call("hm_c1s2_bwd", x.data_ptr(), g.data_ptr(), pooled.data_ptr(), idx.data_ptr(), wk2.data_ptr(),
              dwk2.data_ptr(), u2.data_ptr(), None, B, H, W, 1, 0.2, None)
    # what the engine does: no pooled tensor, act' from bit 2 of the argmax bytes -- the same patch-space gradient, bit for bit
    u3 = torch.zeros_like(u)
    _lib.call("hm_c1s2_bwd", None, g.data_ptr(), None, idx.data_ptr(), wk2.data_ptr(), None, u3.data_ptr(),
              None, B, H, W, 1, 0.2, None)
    torch.cuda.synchronize()
    assert torch.equal(u3[..., :36], u[..., :36]), "hm_c1s2_bwd without the pooled tensor differs from the run with it"
    # reference in float32 with the kernel's own routing (argmax bytes) and activation derivative (sign of the pooled
    # value): full-resolution gradient, then the convolution's adjoints as GEMMs over unfolded patches
    gp = g.float() * torch.where(pooled.float() >= 0, 1.0, 0.2)
    gp = gp.half().float()
    dyf = torch.zeros(B, Hq, 2, Wq, 2, 64, device="cuda")
    k = (idx & 3).long()
    for d in range(4):
        dyf[:, :, d >> 1, :, d & 1, :] = gp * (k == d)
    dyf = dyf.reshape(B, H * W, 64)                                               # [B, L, co]
    cols = F.unfold(x.float()[:, None], 5, padding=2)                             # [B, 25, L], tap (r,s) = x[p+(r,s)-2]
    Wf = Wm.flip(2, 3).reshape(64, 25)                                            # correlation taps
    gW = torch.einsum("bkl,blc->ck", cols.double(), dyf.double()).float().reshape(64, 1, 5, 5).flip(2, 3)
    gb = dyf.sum((0, 1))
    gx = F.fold(torch.einsum("ck,blc->bkl", Wf, dyf), (H, W), 5, padding=2)[:, 0]
    def rel(a, b):
        return float((a - b).abs().max()) / float(b.abs().max())
    e_w = rel(dw.reshape(64, 1, 5, 5), gW)
    e_b = rel(db, gb)
    e_x = rel(dx.float(), gx)
    kk = torch.arange(64, device="cuda")
    live = (kk < 37)
    e_both = max(rel(dwk2.reshape(256, 64)[:, live], dwk.reshape(256, 64)[:, live]),
                 float((u2[..., :36].float() - u[..., :36].float()).abs().max()) / float(u[..., :36].float().abs().max()))
    line = "%-22s dW %.3g  db %.3g  dx %.3g  both-vs-separate %.3g" % (name, e_w, e_b, e_x, e_both)
    return max(e_w, e_b, e_x, e_both), line


def perf_c1bwd():
    B, H, W = 64, 512, 512
    x = torch.randn(B, H, W, device="cuda").half()
    g = torch.randn(B, 256, 256, 64, device="cuda").half()
    pooled = torch.randn(B, 256, 256, 64, device="cuda").half()
    idx = torch.randint(0, 4, (B, 256, 256, 64), device="cuda", dtype=torch.uint8)
    wk2 = torch.randn(256 * 64, device="cuda").half()
    dwk = torch.zeros(256 * 64, device="cuda")
    u = torch.empty(B, 256, 256, 64, device="cuda", dtype=torch.float16)
    dx = torch.empty(B, H, W, device="cuda", dtype=torch.float16)
    runs = (("D1 bwd weight gradient x64", lambda: _lib.call("hm_c1s2_bwd", x.data_ptr(), g.data_ptr(), pooled.data_ptr(), idx.data_ptr(), None, dwk.data_ptr(), None, None, 64, H, W, 1, 0.2, None), 64 * 65536 * (128 + 128 + 64) / 1e9),
            ("D1 bwd input gradient x32 (patch space)", lambda: _lib.call("hm_c1s2_bwd", None, g.data_ptr(), pooled.data_ptr(), idx.data_ptr(), wk2.data_ptr(), None, u.data_ptr(), None, 32, H, W, 1, 0.2, None), 32 * 65536 * (128 + 128 + 64 + 80) / 1e9),
            ("D1 bwd weight gradient x64, act' from idx", lambda: _lib.call("hm_c1s2_bwd", x.data_ptr(), g.data_ptr(), None, idx.data_ptr(), None, dwk.data_ptr(), None, None, 64, H, W, 1, 0.2, None), 64 * 65536 * (128 + 64) / 1e9),
            ("D1 bwd input gradient x32, act' from idx", lambda: _lib.call("hm_c1s2_bwd", None, g.data_ptr(), None, idx.data_ptr(), wk2.data_ptr(), None, u.data_ptr(), None, 32, H, W, 1, 0.2, None), 32 * 65536 * (128 + 64 + 80) / 1e9),
            ("G-out weight gradient x32 (hm_c1s2_wgrad)", lambda: _lib.call("hm_c1s2_wgrad", x.data_ptr(), g.data_ptr(), dwk.data_ptr(), 32, H, W, None), 32 * 65536 * (128 + 8) / 1e9),
            ("D1 bwd col2im x32", lambda: _lib.call("hm_c1s2_col2im", u.data_ptr(), dx.data_ptr(), 32, H, W, None), 32 * 65536 * (72 + 8) / 1e9))
    for nm, fn, gb in runs:
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print("perf %-44s %8.3f ms  %7.1f GB/s (algorithmic bytes %.2f GB)" % (nm, ms, gb / ms * 1e3, gb), flush=True)


def perf_c1():
    for pooled, nm in ((1, "D1 conv5x5(1->64)+lrelu+pool @512^2 x64"), (0, "G-out dgrad 1->64 @512^2 x32")):
        B = 64 if pooled else 32
        x = torch.randn(B, 512, 512, device="cuda").half()
        Wm = torch.randn(64, 1, 5, 5, device="cuda") if pooled else torch.randn(1, 64, 5, 5, device="cuda")
        n = 256 if pooled else 64
        wk = torch.empty(n * 64, device="cuda", dtype=torch.float16)
        _lib.call("hm_pack_conv_weight", Wm.data_ptr(), wk.data_ptr(), 15 if pooled else 14, 64 if pooled else 1,
                  1 if pooled else 64, 5, 5, 0, 0, 1, None)
        y = torch.empty(B, 256, 256, 64, device="cuda", dtype=torch.float16)
        idx = torch.empty(B, 256, 256, 64, device="cuda", dtype=torch.uint8) if pooled else None
        fn = lambda: _lib.call("hm_c1s2_conv", x.data_ptr(), wk.data_ptr(), None, y.data_ptr(),
                               idx.data_ptr() if pooled else None, B, 512, 512, n, 1 if pooled else 0, 0.2, None)
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        gb = (y.numel() * 2 + (idx.numel() if pooled else 0) + x.numel() * 2) / 1e9
        print("perf %-44s %8.3f ms  %7.1f GB/s (algorithmic bytes %.2f GB)" % (nm, ms, gb / ms * 1e3, gb), flush=True)


def perf():
    """Device time of the tensor-core kernels at the hottest DCGAN layer shapes (CUDA events, 5 launches)."""
    shapes = [("D2 64->128 @256^2 x64", 64, 256, 256, 64, 128, 5, 2),
              ("D3 128->128 @128^2 x64", 64, 128, 128, 128, 128, 5, 2),
              ("G7 64->64 @256^2 x32", 32, 256, 256, 64, 64, 5, 2),
              ("D5 128->256 @32^2 x64", 64, 32, 32, 128, 256, 5, 2),
              ("D7 256->256 @8^2 x64", 64, 8, 8, 256, 256, 5, 2)]
    if os.environ.get("HMGAN_PERF_EXTRA"):
        shapes = [("X 64->256 @256^2 x16", 16, 256, 256, 64, 256, 5, 2),
                  ("X 64->128 @256^2 x32", 32, 256, 256, 64, 128, 5, 2),
                  ("X 128->64 @256^2 x32", 32, 256, 256, 128, 64, 5, 2),
                  ("X 256->256 @128^2 x16", 16, 128, 128, 256, 256, 3, 1)]
    for (name, B, H, W, Ci, Co, k, pad) in shapes:
        x = torch.randn(B, H, W, Ci, device="cuda").half()
        dy = torch.randn(B, H, W, Co, device="cuda").half()
        wt = (torch.randn(k * k * Co * Ci, device="cuda") * 0.02).half()
        y = torch.empty(B, H, W, Co, device="cuda", dtype=torch.float16)
        dw = torch.zeros(k * k * Ci, Co, device="cuda")
        d = desc(dtype=1, B=B, H=H, W=W, C1=Ci, C2=0, up=0, kh=k, kw=k, stride=1, pad=pad, transposed=0, Ho=H, Wo=W,
                 Cout=Co, oH=H, oW=W, os=1, ou=0, ov=0, split=Co, act=1, slope=0.2, accumulate=0)
        flop = 2.0 * B * H * W * k * k * Ci * Co
        for what, fn in (("fwd  ", lambda: _tc_conv(C.byref(d), x.data_ptr(), None, wt.data_ptr(), None,
                                                      y.data_ptr(), None, None)),
                         ("wgrad", lambda: _lib.call("hm_tc_wgrad", C.byref(d), x.data_ptr(), None, dy.data_ptr(),
                                                      dw.data_ptr(), None))):
            fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            print("perf %-26s %s %8.3f ms  %7.1f TFLOP/s" % (name, what, ms, flop / ms / 1e9), flush=True)


def perf_up2():
    """Forward of the generator's last layers in phase form (nearest-2x + 5x5 as four 3x3 convolutions on the low-res grid):
    64 -> 1 @256^2 x32 (N = 4 real columns in one zero-padded 16-column tile) and 64 -> 64 @128^2 / @256^2 x32."""
    for (name, B, H, W, Ci, Co, act) in (("Gout 64->1 @256^2(low) x32", 32, 256, 256, 64, 1, 3),
                                         ("G7 64->64 @128^2(low) x32", 32, 128, 128, 64, 64, 0),
                                         ("G6 64->64 @64^2(low) x32", 32, 64, 64, 64, 64, 0)):
        x = torch.randn(B, H, W, Ci, device="cuda").half()
        w8 = (torch.randn(36 * Ci * Co, device="cuda") * 0.02).half()
        bias = torch.zeros(Co, device="cuda")
        y = torch.empty(B, 2 * H, 2 * W, Co, device="cuda", dtype=torch.float16)
        d = desc(dtype=1, B=B, H=H, W=W, C1=Ci, C2=0, up=1, kh=5, kw=5, stride=1, pad=2, transposed=0, Ho=2 * H, Wo=2 * W,
                 Cout=Co, oH=2 * H, oW=2 * W, os=1, ou=0, ov=0, split=Co, act=act, slope=0.2, accumulate=0)
        fn = lambda: _tc_conv(C.byref(d), x.data_ptr(), None, w8.data_ptr(), bias.data_ptr(), y.data_ptr(), None, None)  # noqa: E731
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        gb = (x.numel() + y.numel()) * 2 / 1e9
        print("perf %-30s %8.3f ms  %7.1f GB/s  %7.1f TFLOP/s (36-tap)" % (
            name, ms, gb / ms * 1e3, 2.0 * B * H * W * 36 * Ci * Co / ms / 1e9), flush=True)


if __name__ == "__main__":
    if sys.argv[1:] == ["perf_up2"]:
        perf_up2()
        sys.exit(0)
    if sys.argv[1:] == ["perf"]:
        perf()
        sys.exit(0)
    if sys.argv[1:] == ["dc2"]:
        for c in DC2_CASES:
            try:
                print(run_dc2_case(*c)[1], flush=True)
            except Exception as e:
                print("dc2 %-24s EXC %s" % (c[0], e), flush=True)
                break
        sys.exit(0)
    if sys.argv[1:] == ["c1"]:
        for c in C1_CASES:
            try:
                print(run_c1_case(*c)[1], flush=True)
            except Exception as e:
                print("c1 %-24s EXC %s" % (c[0], e), flush=True)
                break
        perf_c1()
        sys.exit(0)
    if sys.argv[1:] == ["c1bwd"]:
        for c in C1B_CASES:
            try:
                print(run_c1bwd_case(*c)[1], flush=True)
            except Exception as e:
                print("c1bwd %-22s EXC %s" % (c[0], e), flush=True)
                break
        perf_c1bwd()
        sys.exit(0)
    if sys.argv[1:] == ["pool"]:
        for c in POOL_CASES:
            try:
                print(run_pool_case(*c)[1], flush=True)
            except Exception as e:
                print("pool %-28s EXC %s" % (c[0], e), flush=True)
        sys.exit(0)
    if sys.argv[1:] == ["up2bwd"]:
        for c in UP2_BWD_CASES:
            for tc32 in (False, True):
                try:
                    print(run_up2_bwd_case(*c, tc32=tc32)[1], flush=True)
                except Exception as e:
                    print("up2bwd %-24s EXC %s" % (c[0], e), flush=True)
        sys.exit(0)
    if sys.argv[1:] == ["tc32"]:
        for fn, cases in ((run_case32, CASES), (run_wgrad_case32, WGRAD_CASES), (run_up2_case32, UP2_CASES),
                          (run_s2_case32, S2_CASES)):
            for c in cases:
                try:
                    print(fn(*c)[1], flush=True)
                except Exception as e:
                    print("tc32 %-28s EXC %s" % (c[0], e), flush=True)
        sys.exit(0)
    if sys.argv[1:] == ["s2"]:
        for c in S2_CASES:
            try:
                print(run_s2_case(*c)[1], flush=True)
            except Exception as e:
                print("s2 %-22s EXC %s" % (c[0], e), flush=True)
                break
        sys.exit(0)
    if sys.argv[1:] == ["up2"]:
        for c in UP2_CASES:
            try:
                print(run_up2_case(*c)[1], flush=True)
            except Exception as e:
                print("up2 %-24s EXC %s" % (c[0], e), flush=True)
                break
        sys.exit(0)
    if sys.argv[1:] == ["wgrad"]:
        for c in WGRAD_CASES:
            try:
                print(run_wgrad_case(*c)[1], flush=True)
            except Exception as e:
                print("wgrad %-28s EXC %s" % (c[0], e), flush=True)
                break
        sys.exit(0)
    sel = sys.argv[1:] or None
    for c in CASES:
        if sel and c[0] not in sel:
            continue
        try:
            rel, line = run_case(*c)
            print(line, flush=True)
        except Exception as e:     # keep going: one line per case
            print("%-28s EXC %s" % (c[0], e), flush=True)
            break
