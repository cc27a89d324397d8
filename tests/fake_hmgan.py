"""TEST INFRASTRUCTURE: a CPU emulation of the libhmgan C ABI (include/hmgan.h).

Two uses, both in tests only:
  * `-m "not gpu"` tests monkeypatch ``_lib.call`` with ``call`` below so that the
    HOST logic (graph lowering, backward ordering, gradient accumulation, the
    two-pass discriminator backward, optimiser plumbing) can be checked against
    the oracle without a GPU;
  * `-m gpu` kernel tests run each real kernel and this emulation on the same
    seeded inputs and compare.
Every function takes exactly the ctypes arguments of its C counterpart (raw
addresses, sizes) and reads/writes the memory behind them through numpy views.
It is never importable from the product package.
"""
import ctypes as C

import numpy as np
import torch
import torch.nn.functional as F

F32, F16, BF16X3 = 0, 1, 2
_NP = {F32: np.float32, F16: np.float16}


def _a(ptr, n, dtype):
    """numpy view of n elements at address ptr."""
    if ptr is None or ptr == 0:
        return None
    ptr = int(ptr) if not isinstance(ptr, int) else ptr
    nbytes = int(n) * np.dtype(dtype).itemsize
    buf = (C.c_uint8 * nbytes).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=int(n))


def _act(v, act, slope):
    if act == 1:
        return torch.where(v >= 0, v, v * slope)
    if act == 2:
        return torch.relu(v)
    if act == 3:
        return torch.sigmoid(v)
    if act == 4:
        return torch.tanh(v)
    return v


def _act_grad_from_out(y, act, slope):
    if act == 1:
        return torch.where(y >= 0, torch.ones_like(y), torch.full_like(y, slope))
    if act == 2:
        return (y > 0).to(y.dtype)
    if act == 3:
        return y * (1 - y)
    if act == 4:
        return 1 - y * y
    return torch.ones_like(y)


def _t(arr):
    return torch.from_numpy(arr.astype(np.float32))


def _bilinear2(x):      # x: [B,C,H,W]
    def up(a, axis):
        n = a.shape[axis]
        idx = torch.clamp(torch.arange(n) + 1, max=n - 1)
        nxt = a.index_select(axis, idx)
        st = torch.stack([a, 0.5 * (a + nxt)], dim=axis + 1)
        shp = list(a.shape)
        shp[axis] = 2 * n
        return st.reshape(shp)
    return up(up(x, 2), 3)


def _virtual_src(d, x1, x2):
    """[B,Ct,Hv,Wv] float32 virtual source of a forward gather."""
    B, H, W, C1, C2 = d.B, d.H, d.W, d.C1, d.C2
    np_dt = _NP[d.dtype]
    a = _t(_a(x1, B * H * W * C1, np_dt)).reshape(B, H, W, C1)
    if C2:
        b = _t(_a(x2, B * H * W * C2, np_dt)).reshape(B, H, W, C2)
        a = torch.cat([a, b], 3)
    a = a.permute(0, 3, 1, 2).contiguous()
    if d.up == 1:
        a = a.repeat_interleave(2, 2).repeat_interleave(2, 3)
    elif d.up == 2:
        a = _bilinear2(a)
    return a


def hm_conv_gather(dp, x1, x2, w, bias, y, y2, stream=None):
    d = dp._obj if hasattr(dp, "_obj") else dp
    np_dt = _NP[d.dtype]
    Ct = d.C1 + d.C2
    K = d.kh * d.kw * Ct
    wk = _t(_a(w, K * d.Cout, np_dt)).reshape(d.kh, d.kw, Ct, d.Cout)
    src = _virtual_src(d, x1, x2)
    if not d.transposed:
        out = F.conv2d(src, wk.permute(3, 2, 0, 1).contiguous(), stride=d.stride, padding=d.pad)
    else:
        nat_h = (src.shape[2] - 1) * d.stride - 2 * d.pad + d.kh
        nat_w = (src.shape[3] - 1) * d.stride - 2 * d.pad + d.kw
        out = F.conv_transpose2d(src, wk.permute(2, 3, 0, 1).contiguous(), stride=d.stride, padding=d.pad,
                                 output_padding=(max(d.Ho - nat_h, 0), max(d.Wo - nat_w, 0)))
        out = out[:, :, :d.Ho, :d.Wo]
    assert out.shape[2:] == (d.Ho, d.Wo), (out.shape, d.Ho, d.Wo)
    if bias:
        out = out + _t(_a(bias, d.Cout, np.float32)).view(1, -1, 1, 1)
    out = _act(out, d.act, d.slope).permute(0, 2, 3, 1)      # [B,Ho,Wo,Cout]
    ys = slice(d.ou, d.ou + (d.Ho - 1) * d.os + 1, d.os)
    xs = slice(d.ov, d.ov + (d.Wo - 1) * d.os + 1, d.os)
    for ptr, lo, hi, bit in ((y, 0, d.split, 1), (y2, d.split, d.Cout, 2)):
        if not ptr or hi <= lo:
            continue
        dst = _a(ptr, d.B * d.oH * d.oW * (hi - lo), np_dt).reshape(d.B, d.oH, d.oW, hi - lo)
        v = out[..., lo:hi]
        if d.accumulate & bit:
            v = v + _t(dst[:, ys, xs, :])
        dst[:, ys, xs, :] = v.numpy().astype(np_dt)
    return 0


def hm_conv_wgrad(dp, x1, x2, dy, dw, stream=None):
    d = dp._obj if hasattr(dp, "_obj") else dp
    np_dt = _NP[d.dtype]
    Ct = d.C1 + d.C2
    src = _virtual_src(d, x1, x2)
    cols = F.unfold(src, (d.kh, d.kw), padding=d.pad, stride=d.stride)          # [B, Ct*kh*kw, L]
    L = cols.shape[2]
    assert L == d.Ho * d.Wo, (L, d.Ho, d.Wo)
    cols = cols.reshape(d.B, Ct, d.kh * d.kw, L).permute(0, 2, 1, 3).reshape(d.B, d.kh * d.kw * Ct, L)
    g = _t(_a(dy, d.B * d.oH * d.oW * d.Cout, np_dt)).reshape(d.B, d.oH, d.oW, d.Cout)
    ys = slice(d.ou, d.ou + (d.Ho - 1) * d.os + 1, d.os)
    xs = slice(d.ov, d.ov + (d.Wo - 1) * d.os + 1, d.os)
    g = g[:, ys, xs, :].reshape(d.B, L, d.Cout)
    out = _a(dw, d.kh * d.kw * Ct * d.Cout, np.float32)
    out += torch.einsum("bkl,blc->kc", cols.double(), g.double()).float().numpy().reshape(-1)
    return 0


def hm_pack_conv_weight(w, wp, mode, cout, cin, kh, kw, u, v, dst_dtype, stream=None):
    n = cout * cin * kh * kw
    W = _a(w, n, np.float32)
    if mode == 0:
        src = W.reshape(cout, cin, kh, kw)[:, :, ::-1, ::-1]
        out = src.transpose(2, 3, 1, 0)                      # [r][s][ci][co]
    elif mode == 1:
        src = W.reshape(cout, cin, kh, kw)[:, :, ::-1, ::-1]
        out = src.transpose(2, 3, 0, 1)                      # [r][s][co][ci]
    elif mode == 2:
        src = W.reshape(cin, cout, kh, kw)
        out = src[:, :, kh - 1 - u, kw - 1 - v]              # [ci][co]
    elif mode == 3:
        src = W.reshape(cin, cout, kh, kw)[:, :, ::-1, ::-1]
        out = src.transpose(2, 3, 1, 0)                      # [u][v][co][ci]
    elif mode == 5:
        src = W.reshape(cout, cin, kh, kw)[:, :, ::-1, ::-1]
        out = src.transpose(2, 3, 0, 1)                      # [r][s][co][ci]
    elif mode == 8:
        Wc = W.reshape(cout, cin, 5, 5)[:, :, ::-1, ::-1]                    # Wc[co][ci][r][s] (correlation taps)
        out = np.zeros((3, 3, 4, cout, cin), np.float32)
        for py in range(2):
            for px in range(2):
                for r in range(5):
                    for s_ in range(5):
                        dy_, dx_ = (py + r - 2) // 2 + 1, (px + s_ - 2) // 2 + 1
                        out[dy_, dx_, py * 2 + px] += Wc[:, :, r, s_]
    elif mode == 12:
        Wc = W.reshape(cout, cin, 3, 3)[:, :, ::-1, ::-1]                     # correlation taps Wc[co][ci][r][s]
        out = np.zeros((2, 2, 4, cin, cout), np.float32)
        rmap = {(0, 0): 1, (1, 0): 2, (1, 1): 0}
        for py in range(2):
            for px in range(2):
                for dy_ in range(2):
                    for dx_ in range(2):
                        if (py, dy_) in rmap and (px, dx_) in rmap:
                            out[dy_, dx_, py * 2 + px] = Wc[:, :, rmap[(py, dy_)], rmap[(px, dx_)]].T
    elif mode == 11:
        Wc = W.reshape(cout, kh * kw)[:, ::-1]                    # flipped taps, row-major (r, s)
        out = np.zeros((cout, 64), np.float32)
        out[:, :kh * kw] = Wc
    elif mode == 14:
        Wc = W.reshape(cin, 5, 5)[:, ::-1, ::-1]                                 # Cout == 1: correlation taps Wc[ci][r][s]
        out = np.zeros((cin, 64), np.float32)
        for uu in range(6):
            for vv in range(6):
                py, px, dy_, dx_ = uu & 1, vv & 1, 2 - (uu >> 1), 2 - (vv >> 1)
                for r in range(5):
                    for s_ in range(5):
                        if (py + r - 2) // 2 + 1 == dy_ and (px + s_ - 2) // 2 + 1 == dx_:
                            out[:, uu * 6 + vv] += Wc[:, r, s_]
    elif mode == 21:
        out = W.reshape(cin, cout).T                                             # [co][ci]
    elif mode == 20:
        Wc = W.reshape(cout, cin, 5, 5)[:, :, ::-1, ::-1]                        # correlation taps Wc[co][ci][r][s]
        out = np.zeros((36, cin, cout), np.float32)
        for uu in range(6):
            for vv in range(6):
                py, px, dy_, dx_ = uu & 1, vv & 1, 2 - (uu >> 1), 2 - (vv >> 1)
                for r in range(5):
                    for s_ in range(5):
                        if (py + r - 2) // 2 + 1 == dy_ and (px + s_ - 2) // 2 + 1 == dx_:
                            out[uu * 6 + vv] += Wc[:, :, r, s_].T
    elif mode == 15:
        Wc = W.reshape(cout, 5, 5)[:, ::-1, ::-1]                                # Cin == 1: correlation taps Wc[co][r][s]
        out = np.zeros((4, cout, 64), np.float32)
        for dd in range(4):
            for uu in range(6):
                for vv in range(6):
                    r, s_ = uu - (dd >> 1), vv - (dd & 1)
                    if 0 <= r < 5 and 0 <= s_ < 5:
                        out[dd, :, uu * 6 + vv] = Wc[:, r, s_]
    elif mode == 16:
        Wc = W.reshape(cout, 5, 5)[:, ::-1, ::-1]
        out = np.zeros((4, 64, cout), np.float32)
        for dd in range(4):
            for uu in range(6):
                for vv in range(6):
                    r, s_ = uu - (dd >> 1), vv - (dd & 1)
                    if 0 <= r < 5 and 0 <= s_ < 5:
                        out[dd, uu * 6 + vv, :] = Wc[:, r, s_]
    elif mode == 19:
        Wc = W.reshape(cout, cin, kh, kw)[:, :, ::-1, ::-1]                        # correlation taps Wc[co][ci][r][s]
        out = np.zeros((cout, 64), np.float32)
        out[:, :kh * kw * cin] = Wc.transpose(0, 2, 3, 1).reshape(cout, kh * kw * cin)      # [co][(r,s,ci)]
    elif mode == 17:
        Wd = W.reshape(cin, cout, 2, 2)[:, :, ::-1, ::-1]                          # Wd[ci][co][u][v] = W[ci][co][1-u][1-v]
        out = Wd.transpose(2, 3, 1, 0).reshape(4 * cout, cin)                      # [(u,v,co)][ci]
    elif mode == 18:
        Wd = W.reshape(cin, cout, 2, 2)[:, :, ::-1, ::-1]
        out = np.zeros((cin, 64), np.float32)
        out[:, :4 * cout] = Wd.transpose(0, 2, 3, 1).reshape(cin, 4 * cout)        # [ci][(u,v,co)]
    elif mode == 22:
        Wd = W.reshape(cin, cout, 2, 2)[:, :, ::-1, ::-1]
        out = Wd.transpose(0, 2, 3, 1).reshape(cin, 4 * cout)                      # [ci][(u,v,co)], unpadded mode 18
    elif mode == 7:
        out = W.reshape(cout, cin, kh, kw).transpose(2, 3, 0, 1)   # [r][s][co][ci]
    elif mode == 6:
        out = W.reshape(cout, cin, kh, kw).transpose(2, 3, 1, 0)   # [r][s][ci][co]
    else:
        out = W
    out = np.ascontiguousarray(out).reshape(-1)
    _a(wp, out.size, _NP[dst_dtype])[:] = out.astype(_NP[dst_dtype])
    return 0


def hm_pack_conv_weight_multi(jobs, n_jobs, max_n, dst_dtype, stream=None):
    """HmPackJob[] in (host) memory: {w, wp, mode, cout, cin, kh, kw, u, v, pad, n}, 56 bytes each."""
    import struct
    raw = bytes(_a(jobs, n_jobs * 56, np.uint8))
    for i in range(n_jobs):
        w, wp, mode, cout, cin, kh, kw, u, v, _pad, n = struct.unpack_from("<QQiiiiiiiiq", raw, i * 56)
        assert 0 < n <= max_n
        hm_pack_conv_weight(w, wp, mode, cout, cin, kh, kw, u, v, dst_dtype)
    return 0


def hm_unpack_conv_wgrad(dwp, dw, mode, cout, cin, kh, kw, stream=None):
    n = cout * cin * kh * kw
    src = _a(dwp, n, np.float32)
    dst = _a(dw, n, np.float32)
    if mode in (8, 9, 10):
        if mode == 10:
            g3 = _a(dwp, 9 * cin * 64, np.float32).reshape(3, 3, cin, 64)[..., :4 * cout].reshape(3, 3, cin, 4, cout)
        elif mode == 8:
            g3 = _a(dwp, 36 * cin * cout, np.float32).reshape(3, 3, cin, 4, cout)
        else:
            g3 = _a(dwp, 36 * cin * cout, np.float32).reshape(4, 3, 3, cin, cout).transpose(1, 2, 3, 0, 4)
        gc = np.zeros((5, 5, cin, cout), np.float32)
        for py in range(2):
            for px in range(2):
                for r in range(5):
                    for s_ in range(5):
                        gc[r, s_] += g3[(py + r - 2) // 2 + 1, (px + s_ - 2) // 2 + 1, :, py * 2 + px, :]
        dst[:] = np.ascontiguousarray(gc.transpose(3, 2, 0, 1)[:, :, ::-1, ::-1]).reshape(-1)
        return 0
    if mode == 14:
        assert cout == 1 and kh == 5 and kw == 5
        g6 = _a(dwp, cin * 64, np.float32).reshape(cin, 64)
        gc = np.zeros((cin, 5, 5), np.float32)                                   # correlation taps [ci][r][s]
        for py in range(2):
            for px in range(2):
                for r in range(5):
                    for s_ in range(5):
                        uu = 2 * (2 - ((py + r - 2) // 2 + 1)) + py
                        vv = 2 * (2 - ((px + s_ - 2) // 2 + 1)) + px
                        gc[:, r, s_] += g6[:, uu * 6 + vv]
        dst[:] = np.ascontiguousarray(gc[:, ::-1, ::-1]).reshape(-1)
        return 0
    if mode == 17:
        ld = 64 if 4 * cout <= 64 else 4 * cout
        g4 = _a(dwp, cin * ld, np.float32).reshape(cin, ld)[:, :4 * cout].reshape(cin, 2, 2, cout)   # [ci][u][v][co]
        dst[:] = np.ascontiguousarray(g4.transpose(0, 3, 1, 2)[:, :, ::-1, ::-1]).reshape(-1)
        return 0
    if mode == 0:
        g = src.reshape(kh, kw, cin, cout).transpose(3, 2, 0, 1)[:, :, ::-1, ::-1]
    elif mode == 2:
        g = src.reshape(kh, kw, cin, cout).transpose(2, 3, 0, 1)[:, :, ::-1, ::-1]
    else:
        g = src
    dst[:] = np.ascontiguousarray(g).reshape(-1)
    return 0


def hm_bn_stats(x, dtype, M, Cn, sums, stream=None):
    v = _a(x, M * Cn, _NP[dtype]).astype(np.float64).reshape(M, Cn)
    s = _a(sums, 2 * Cn, np.float64)
    s[:Cn] += v.sum(0)
    s[Cn:] += (v * v).sum(0)
    return 0


def hm_col_sum(dy, dtype, M, Cn, db, stream=None):
    v = _a(dy, M * Cn, _NP[dtype]).astype(np.float64).reshape(M, Cn)
    _a(db, Cn, np.float32)[:] += v.sum(0).astype(np.float32)
    return 0


def hm_bn_finalize(sums, M, Cn, gamma, beta, rmean, rinv, eps, alpha, update_running, mean, inv_std, scale,
                   shift, stream=None):
    g = _a(gamma, Cn, np.float32)
    b = _a(beta, Cn, np.float32)
    if sums:
        s = _a(sums, 2 * Cn, np.float64)
        mu = s[:Cn] / M
        var = np.maximum(s[Cn:] / M - mu * mu, 0)
        m = mu.astype(np.float32)
        i = (1.0 / np.sqrt(var + np.float64(np.float32(eps)))).astype(np.float32)
        if update_running:
            rm, ri = _a(rmean, Cn, np.float32), _a(rinv, Cn, np.float32)
            a = np.float32(alpha)
            rm[:] = (np.float32(1) - a) * rm + a * m
            ri[:] = (np.float32(1) - a) * ri + a * i
    else:
        m = _a(rmean, Cn, np.float32).copy()
        i = _a(rinv, Cn, np.float32).copy()
    if mean:
        _a(mean, Cn, np.float32)[:] = m
    if inv_std:
        _a(inv_std, Cn, np.float32)[:] = i
    sc = g * i
    _a(scale, Cn, np.float32)[:] = sc
    _a(shift, Cn, np.float32)[:] = b - m * sc
    return 0


def hm_bn_apply_act(x, a, dtype, M, Cn, scale, shift, act, slope, stream=None):
    v = _t(_a(x, M * Cn, _NP[dtype])).reshape(M, Cn)
    sc = _t(_a(scale, Cn, np.float32))
    sf = _t(_a(shift, Cn, np.float32))
    out = _act(v * sc + sf, act, slope)
    _a(a, M * Cn, _NP[dtype])[:] = out.numpy().reshape(-1).astype(_NP[dtype])
    return 0


def hm_bn_bwd_reduce(da, a, x, dtype, M, Cn, mean, inv_std, act, slope, red, stream=None):
    g = _t(_a(da, M * Cn, _NP[dtype])).reshape(M, Cn)
    av = _t(_a(a, M * Cn, _NP[dtype])).reshape(M, Cn)
    xv = _t(_a(x, M * Cn, _NP[dtype])).reshape(M, Cn)
    g = g * _act_grad_from_out(av, act, slope)
    xh = (xv - _t(_a(mean, Cn, np.float32))) * _t(_a(inv_std, Cn, np.float32))
    r = _a(red, 2 * Cn, np.float64)
    r[:Cn] += g.double().sum(0).numpy()
    r[Cn:] += (g * xh).double().sum(0).numpy()
    return 0


def hm_bn_bwd_apply(da, a, x, dx, dtype, M, Cn, mean, inv_std, gamma, act, slope, red, dgamma, dbeta,
                    stream=None):
    g = _t(_a(da, M * Cn, _NP[dtype])).reshape(M, Cn)
    av = _t(_a(a, M * Cn, _NP[dtype])).reshape(M, Cn)
    xv = _t(_a(x, M * Cn, _NP[dtype])).reshape(M, Cn)
    g = g * _act_grad_from_out(av, act, slope)
    is_ = _t(_a(inv_std, Cn, np.float32))
    xh = (xv - _t(_a(mean, Cn, np.float32))) * is_
    r = _a(red, 2 * Cn, np.float64)
    r0 = torch.from_numpy(r[:Cn].astype(np.float32))
    r1 = torch.from_numpy(r[Cn:].astype(np.float32))
    out = _t(_a(gamma, Cn, np.float32)) * is_ * (g - r0 / M - xh * r1 / M)
    _a(dx, M * Cn, _NP[dtype])[:] = out.numpy().reshape(-1).astype(_NP[dtype])
    if dgamma:
        _a(dgamma, Cn, np.float32)[:] = r1.numpy()
    if dbeta:
        _a(dbeta, Cn, np.float32)[:] = r0.numpy()
    return 0


def hm_bn_bwd_reduce_a(da, a, x, dtype, M, Cn, mean, inv_std, gamma, beta, act, slope, red, stream=None):
    """Contract: the results of hm_bn_bwd_reduce up to the rounding of a (the emulation uses the x-based form)."""
    return hm_bn_bwd_reduce(da, a, x, dtype, M, Cn, mean, inv_std, act, slope, red)


def hm_bn_bwd_apply_a(da, a, x, dx, dtype, M, Cn, mean, inv_std, gamma, beta, act, slope, red, dgamma, dbeta,
                      stream=None):
    return hm_bn_bwd_apply(da, a, x, dx, dtype, M, Cn, mean, inv_std, gamma, act, slope, red, dgamma, dbeta)


def hm_act_bwd(dy, y, dx, dtype, n, act, slope, accumulate, stream=None):
    g = _t(_a(dy, n, _NP[dtype]))
    yv = _t(_a(y, n, _NP[dtype]))
    out = g * _act_grad_from_out(yv, act, slope)
    dst = _a(dx, n, _NP[dtype])
    if accumulate:
        out = out + _t(dst)
    dst[:] = out.numpy().astype(_NP[dtype])
    return 0


def hm_maxpool2_fwd(x, p, idx, dtype, B, H, W, Cn, stream=None):
    v = _t(_a(x, B * H * W * Cn, _NP[dtype])).reshape(B, H, W, Cn)
    Hp, Wp = H // 2, W // 2
    v = v[:, :Hp * 2, :Wp * 2]
    win = torch.stack([v[:, 0::2, 0::2], v[:, 0::2, 1::2], v[:, 1::2, 0::2], v[:, 1::2, 1::2]], 0)
    m, k = win.max(0)           # torch returns the first maximal index on CPU
    # enforce "first max wins"
    first = torch.zeros_like(k)
    found = torch.zeros_like(k, dtype=torch.bool)
    for j in range(4):
        hit = (win[j] == m) & ~found
        first[hit] = j
        found |= hit
    _a(p, B * Hp * Wp * Cn, _NP[dtype])[:] = m.numpy().reshape(-1).astype(_NP[dtype])
    _a(idx, B * Hp * Wp * Cn, np.uint8)[:] = first.numpy().reshape(-1).astype(np.uint8)
    return 0


def hm_maxpool2_bwd(dp, p, idx, dx, dtype, B, H, W, Cn, act, slope, db=None, stream=None):
    Hp, Wp = H // 2, W // 2
    n = B * Hp * Wp * Cn
    g = _t(_a(dp, n, _NP[dtype])).reshape(B, Hp, Wp, Cn)
    pv = _t(_a(p, n, _NP[dtype])).reshape(B, Hp, Wp, Cn)
    k = torch.from_numpy(_a(idx, n, np.uint8).astype(np.int64)).reshape(B, Hp, Wp, Cn)
    g = g * _act_grad_from_out(pv, act, slope)
    out = torch.zeros(B, H, W, Cn)
    for j, (a, b) in enumerate(((0, 0), (0, 1), (1, 0), (1, 1))):
        out[:, a::2, b::2] = torch.where(k == j, g, torch.zeros_like(g))
    _a(dx, B * H * W * Cn, _NP[dtype])[:] = out.numpy().reshape(-1).astype(_NP[dtype])
    if db:
        _a(db, Cn, np.float32)[:] += g.reshape(-1, Cn).double().sum(0).float().numpy()
    return 0


def hm_maxpool2_bwd_scaled(dp, p, idx, dx, dxs, scale, dtype, B, H, W, Cn, act, slope, db=None, stream=None):
    if dx:
        hm_maxpool2_bwd(dp, p, idx, dx, dtype, B, H, W, Cn, act, slope, None)
    sc = _t(_a(scale, B, np.float32)).view(B, 1, 1, 1)
    Hp, Wp = H // 2, W // 2
    n = B * Hp * Wp * Cn
    g = _t(_a(dp, n, _NP[dtype])).reshape(B, Hp, Wp, Cn) * _act_grad_from_out(
        _t(_a(p, n, _NP[dtype])).reshape(B, Hp, Wp, Cn), act, slope)
    k = torch.from_numpy(_a(idx, n, np.uint8).astype(np.int64)).reshape(B, Hp, Wp, Cn)
    out = torch.zeros(B, H, W, Cn)
    for j, (a, b) in enumerate(((0, 0), (0, 1), (1, 0), (1, 1))):
        out[:, a::2, b::2] = torch.where(k == j, g * sc, torch.zeros_like(g))
    _a(dxs, B * H * W * Cn, _NP[dtype])[:] = out.numpy().reshape(-1).astype(_NP[dtype])
    if db:
        _a(db, Cn, np.float32)[:] += (g * sc).reshape(-1, Cn).double().sum(0).float().numpy()
    return 0


def hm_scale_rows(src, scale, dst, dtype, R, L, stream=None):
    v = _t(_a(src, R * L, _NP[dtype])).reshape(R, L) * _t(_a(scale, R, np.float32)).view(R, 1)
    _a(dst, R * L, _NP[dtype])[:] = v.numpy().reshape(-1).astype(_NP[dtype])
    return 0


def hm_adv_loss_pair(h, dh, dhw, sw, sg, dtype, R, G, out_act, lsgan, relu_head, gscale, loss_disc, loss_gen,
                     stream=None):
    hv = _t(_a(h, R * G, _NP[dtype])).reshape(R, G)
    out = _act(hv.mean(1), out_act, 0.0)
    if lsgan:
        la, da, lb, db_ = out ** 2, 2 * out, (out - 1) ** 2, 2 * (out - 1)
    else:
        la, da, lb, db_ = -torch.log(1 - out), 1 / (1 - out), -torch.log(out), -1 / out
    _a(loss_disc, 1, np.float32)[0] += np.float32(float(la.mean()))
    _a(loss_gen, 1, np.float32)[0] += np.float32(float(lb.mean()))
    k = gscale * _act_grad_from_out(out, out_act, 0.0) / (R * G)
    a, b = da * k, db_ * k
    c = torch.where(a.abs() >= b.abs(), a, b)
    safe = torch.where(c != 0, c, torch.ones_like(c))
    _a(sw, R, np.float32)[:] = torch.where(c != 0, a / safe, torch.zeros_like(c)).numpy()
    _a(sg, R, np.float32)[:] = torch.where(c != 0, b / safe, torch.zeros_like(c)).numpy()
    mask = (hv > 0).float() if relu_head else torch.ones_like(hv)
    _a(dh, R * G, _NP[dtype])[:] = (c.view(R, 1) * mask).numpy().reshape(-1).astype(_NP[dtype])
    _a(dhw, R * G, _NP[dtype])[:] = (a.view(R, 1) * mask).numpy().reshape(-1).astype(_NP[dtype])
    return 0


def hm_upsample2_fwd(x, y, dtype, B, H, W, Cn, mode, stream=None):
    v = _t(_a(x, B * H * W * Cn, _NP[dtype])).reshape(B, H, W, Cn).permute(0, 3, 1, 2)
    out = v.repeat_interleave(2, 2).repeat_interleave(2, 3) if mode == 1 else _bilinear2(v)
    _a(y, B * 4 * H * W * Cn, _NP[dtype])[:] = out.permute(0, 2, 3, 1).numpy().reshape(-1).astype(_NP[dtype])
    return 0


def hm_upsample2_bwd(dy, dx, dtype, B, H, W, Cn, mode, accumulate, stream=None):
    g = _t(_a(dy, B * 4 * H * W * Cn, _NP[dtype])).reshape(B, 2 * H, 2 * W, Cn).permute(0, 3, 1, 2)
    x = torch.zeros(B, Cn, H, W, requires_grad=True)
    up = x.repeat_interleave(2, 2).repeat_interleave(2, 3) if mode == 1 else _bilinear2(x)
    (gx,) = torch.autograd.grad(up, x, g.contiguous())
    out = gx.permute(0, 2, 3, 1)
    dst = _a(dx, B * H * W * Cn, _NP[dtype])
    if accumulate:
        out = out + _t(dst).reshape(out.shape)
    dst[:] = out.numpy().reshape(-1).astype(_NP[dtype])
    return 0


def hm_nchw_to_nhwc(src, dst, dtype, B, Cn, H, W, stream=None):
    v = _a(src, B * Cn * H * W, np.float32).reshape(B, Cn, H, W).transpose(0, 2, 3, 1)
    _a(dst, B * Cn * H * W, _NP[dtype])[:] = np.ascontiguousarray(v).reshape(-1).astype(_NP[dtype])
    return 0


def hm_nhwc_to_nchw(src, dst, dtype, B, Cn, H, W, stream=None):
    v = _a(src, B * Cn * H * W, _NP[dtype]).reshape(B, H, W, Cn).transpose(0, 3, 1, 2)
    _a(dst, B * Cn * H * W, np.float32)[:] = np.ascontiguousarray(v).reshape(-1).astype(np.float32)
    return 0


def hm_permute(src, dst, dtype, B, Cn, H, W, inverse, stream=None):
    n = B * Cn * H * W
    s, d = _a(src, n, _NP[dtype]), _a(dst, n, _NP[dtype])
    if inverse:
        d[:] = np.ascontiguousarray(s.reshape(B, H, W, Cn).transpose(0, 3, 1, 2)).reshape(-1)
    else:
        d[:] = np.ascontiguousarray(s.reshape(B, Cn, H, W).transpose(0, 2, 3, 1)).reshape(-1)
    return 0


def hm_slice_channels(src, dst, dtype, M, Cn, c0, nc, accumulate, stream=None):
    s_ = _a(src, M * Cn, _NP[dtype]).reshape(M, Cn)[:, c0:c0 + nc].astype(np.float32)
    d_ = _a(dst, M * nc, _NP[dtype])
    out = s_.reshape(-1) + (d_.astype(np.float32) if accumulate else 0)
    d_[:] = out.astype(_NP[dtype])
    return 0


def hm_cast(src, sd, dst, dd, n, stream=None):
    _a(dst, n, _NP[dd])[:] = _a(src, n, _NP[sd]).astype(_NP[dd])
    return 0


def hm_u8_normalize(src, dst, dtype, n, tanh_range, stream=None):
    v = _a(src, n, np.uint8).astype(np.float32)
    v = (v - np.float32(127.5)) / np.float32(127.5) if tanh_range else v / np.float32(255.0)
    _a(dst, n, _NP[dtype])[:] = v.astype(_NP[dtype])
    return 0


def hm_adv_loss(h, dh, dtype, R, G, out_act, target, lsgan, relu_head, weight, gscale, accumulate, loss,
                stream=None):
    hv = _t(_a(h, R * G, _NP[dtype])).reshape(R, G)
    out = _act(hv.mean(1), out_act, 0.0)
    if lsgan:
        l = (out - target) ** 2
        dl = 2 * (out - target)
    else:
        l = -(target * torch.log(out) + (1 - target) * torch.log(1 - out))
        dl = -(target / out) + (1 - target) / (1 - out)
    _a(loss, 1, np.float32)[0] += np.float32(weight * float(l.mean()))
    if dh:
        dl = dl * _act_grad_from_out(out, out_act, 0.0)
        g = (gscale * weight * dl / (R * G)).view(R, 1).expand(R, G).clone()
        if relu_head:
            g = torch.where(hv > 0, g, torch.zeros_like(g))
        dst = _a(dh, R * G, _NP[dtype])
        if accumulate:
            g = g + _t(dst).reshape(R, G)
        dst[:] = g.numpy().reshape(-1).astype(_NP[dtype])
    return 0


def hm_recon_loss(p, y, dp, dtype, n, l2, weight, gscale, accumulate, loss, stream=None):
    e = _t(_a(p, n, _NP[dtype])) - _t(_a(y, n, _NP[dtype]))
    if l2:
        l, g = (e * e).mean(), 2 * e
    else:
        l, g = e.abs().mean(), torch.sign(e)
    _a(loss, 1, np.float32)[0] += np.float32(weight * float(l))
    if dp:
        g = g * (gscale * weight / n)
        dst = _a(dp, n, _NP[dtype])
        if accumulate:
            g = g + _t(dst)
        dst[:] = g.numpy().astype(_NP[dtype])
    return 0


def hm_rmsprop(p, g, acc, n, lr, rho, eps, gscale, stream=None):
    P, Gd, A = _a(p, n, np.float32), _a(g, n, np.float32), _a(acc, n, np.float32)
    lrv = _a(lr, 1, np.float32)[0]
    gi = Gd * np.float32(gscale)
    a = np.float32(rho) * A + (np.float32(1) - np.float32(rho)) * gi * gi
    A[:] = a
    P[:] = P - lrv * gi / np.sqrt(a + np.float32(eps))
    return 0


def hm_adam(p, g, m, v, n, lr, b1, b2, eps, t, gscale, stream=None):
    P, Gd, Mv, Vv = (_a(q, n, np.float32) for q in (p, g, m, v))
    lrv = _a(lr, 1, np.float32)[0]
    corr = np.float32(np.sqrt(1 - b2 ** t) / (1 - b1 ** t))
    gi = Gd * np.float32(gscale)
    mi = np.float32(b1) * Mv + np.float32(1 - b1) * gi
    vi = np.float32(b2) * Vv + np.float32(1 - b2) * gi * gi
    Mv[:] = mi
    Vv[:] = vi
    P[:] = P - lrv * corr * mi / (np.sqrt(vi) + np.float32(eps))
    return 0


def hm_inc_i32(counter, stream=None):
    _a(counter, 1, np.int32)[0] += 1
    return 0


def hm_adam_dev(p, g, m, v, n, lr, b1, b2, eps, t_dev, gscale, stream=None):
    return hm_adam(p, g, m, v, n, lr, b1, b2, eps, int(_a(t_dev, 1, np.int32)[0]), gscale)


def _bf16_round(a):
    """float32 array -> nearest-even bfloat16, as uint16 bit patterns."""
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint16)
    return r


def _bf16_to_f32(u16):
    return (u16.astype(np.uint32) << 16).view(np.float32)


def hm_split_bf16x3(src, dst, rows, Cn, c1, layout, stream=None):
    a = _a(src, rows * Cn, np.float32).reshape(rows, Cn)
    h = _bf16_round(a)
    r1 = a - _bf16_to_f32(h)
    m = _bf16_round(r1)
    l = _bf16_round(r1 - _bf16_to_f32(m))
    planes = (h, m, h, l, h, m) if layout & 1 else (h, h, m, h, l, m)
    out = _a(dst, 6 * rows * Cn, np.uint16)
    if layout >= 2:
        o = out.reshape(6, rows, Cn)
        for k in range(6):
            o[k] = planes[k]
    else:
        o = out.reshape(rows, 6 * Cn)
        for (a0, a1, base) in ((0, c1, 0), (c1, Cn, 6 * c1)):
            w = a1 - a0
            for k in range(6):
                if w > 0:
                    o[:, base + k * w:base + (k + 1) * w] = planes[k][:, a0:a1]
    return 0


def _bf16x3_as_f32(d, ptrs_counts):
    """HM_BF16X3 problem -> the same problem on float32 copies of its bf16 operands (keeps the arrays alive)."""
    d2 = type(d).from_buffer_copy(d)
    d2.dtype = F32
    keep = [None if not ptr else np.ascontiguousarray(_bf16_to_f32(_a(ptr, n, np.uint16))) for ptr, n in ptrs_counts]
    return d2, keep, [None if k is None else k.ctypes.data for k in keep]


def _is_up2conv(d):
    return (d.up == 1 and d.kh == 5 and d.kw == 5 and d.pad == 2 and d.stride == 1 and not d.transposed and
            d.C2 == 0 and d.Ho == 2 * d.H and d.Wo == 2 * d.W and d.split == d.Cout and not d.accumulate and
            (d.Cout % 32 == 0 or d.Cout <= 4))


def _is_dgrad_s2(d):
    return (d.transposed == 1 and d.stride == 2 and d.kh == 3 and d.kw == 3 and d.pad == 1 and not d.up and
            d.Ho == 2 * d.H and d.Wo == 2 * d.W)


def _is_deconv_d2s(d):
    return (d.transposed == 2 and d.kh == 2 and d.kw == 2 and d.stride == 2 and d.pad == 0 and not d.up
            and d.Ho == 2 * d.H and d.Wo == 2 * d.W and d.oH == d.Ho and d.oW == d.Wo and d.os == 1 and not d.ou
            and not d.ov and d.split == d.Cout and not d.accumulate)


def _tc_ok(d, wgrad):
    dtype0 = d.dtype
    if d.dtype == BF16X3:            # bf16 hi/lo splits of fp32 tensors: same shape rules as fp16
        d = type(d).from_buffer_copy(d)
        d.dtype = F16
    if wgrad and d.dtype == F16 and _is_up2_wgrad(d):
        return True
    if not wgrad and d.dtype == F16 and _is_deconv_d2s(d):
        return d.C1 % 64 == 0 and d.C2 % 64 == 0 and d.C1 > 0 and (d.Cout % 32 == 0 or d.Cout <= 4)
    if not wgrad and d.dtype == F16 and _is_dgrad_s2(d):
        return (d.C1 % 64 == 0 and d.C1 > 0 and d.C2 == 0 and d.os == 1 and not d.ou and not d.ov and d.oH == d.Ho
                and d.oW == d.Wo and d.split == d.Cout and (d.Cout % 32 == 0 or d.Cout <= 4))
    if d.stride == 2 and not d.transposed and not d.up and d.dtype == F16:
        ok = d.os == 1 and not d.ou and not d.ov and d.C1 % 64 == 0 and d.C2 % 64 == 0 and d.C1 > 0
        ok = ok and d.Ho == (d.H + 2 * d.pad - d.kh) // 2 + 1 and d.Wo == (d.W + 2 * d.pad - d.kw) // 2 + 1
        ok = ok and d.oH == d.Ho and d.oW == d.Wo
        if wgrad:
            return ok and d.Cout % 64 == 0 and (0 < d.Cout <= 256 or d.Cout % 256 == 0)
        if d.Cout % 16 or d.split % 16:
            return ok and d.split == d.Cout and 0 < d.Cout <= 256
        return ok and 0 < d.split <= d.Cout
    if not wgrad and d.dtype == F16 and _is_up2conv(d):
        return d.C1 % 64 == 0 and d.C1 > 0 and d.os == 1 and not d.ou and not d.ov and d.oH == d.Ho and d.oW == d.Wo
    ok = d.dtype == F16 and not d.transposed and not d.up and d.stride == 1 and d.os == 1 and not d.ou and not d.ov
    ragged = (not wgrad) and d.C2 == 0 and d.C1 % 8 == 0 and dtype0 == F16         # one source, TMA zero-fills the tail
    ok = ok and d.C2 % 64 == 0 and d.C1 > 0 and (d.C1 % 64 == 0 or ragged)
    ok = ok and d.Ho == d.H + 2 * d.pad - d.kh + 1 and d.Wo == d.W + 2 * d.pad - d.kw + 1
    ok = ok and d.oH == d.Ho and d.oW == d.Wo
    if wgrad:
        return ok and d.Cout % 64 == 0 and (0 < d.Cout <= 256 or d.Cout % 256 == 0)
    if d.Cout % 16 or d.split % 16:
        return ok and d.split == d.Cout and 0 < d.Cout <= 256
    ntile = [n for n in range(256, 15, -16) if d.Cout % n == 0 and d.split % n == 0]
    return ok and 0 < d.split <= d.Cout and bool(ntile)


def _pool_ok(d):
    ok = d.dtype == F16 and _tc_ok(d, False) and not d.transposed and not d.up and d.stride == 1
    ok = ok and d.split == d.Cout and not d.accumulate and d.Cout % 32 == 0
    ok = ok and d.Ho % 2 == 0 and d.Wo % 2 == 0 and d.Wo >= 128 and 2 <= d.kw <= 9 and d.act in (0, 1, 2)
    return bool(ok)


def hm_tc_conv_pool(dp, x1, x2, w_tc, bias, y_pooled, idx, stream=None):
    """Contract: hm_tc_conv followed by hm_maxpool2_fwd, the un-pooled tensor never leaving the kernel."""
    d = dp._obj if hasattr(dp, "_obj") else dp
    assert _pool_ok(d)
    full = np.zeros(d.B * d.Ho * d.Wo * d.Cout, np.float16)
    hm_tc_conv(d, x1, x2, w_tc, bias, full.ctypes.data, None)
    return hm_maxpool2_fwd(full.ctypes.data, y_pooled, idx, F16, d.B, d.Ho, d.Wo, d.Cout)


def hm_tc_conv_ws(dp, x1, x2, w_tc, bias, y, y2, ws, ws_bytes, stream=None):
    """hm_tc_conv with a workspace the emulation has no use for (it must stay zero)."""
    return hm_tc_conv(dp, x1, x2, w_tc, bias, y, y2, stream)


def hm_tc_conv(dp, x1, x2, w_tc, bias, y, y2, stream=None):
    """Same contract as hm_conv_gather, weights in the K-major pack [tap][Cout][Cin]."""
    d = dp._obj if hasattr(dp, "_obj") else dp
    assert _tc_ok(d, False)
    f16 = np.float16
    if d.dtype == BF16X3:
        nsrc = d.B * d.H * d.W
        nw = (36 if _is_up2conv(d) else 4 if d.transposed == 2 else 16 if _is_dgrad_s2(d) else d.kh * d.kw) * \
            d.Cout * (d.C1 + d.C2)
        d, _keep, (x1, x2, w_tc) = _bf16x3_as_f32(d, ((x1, nsrc * d.C1), (x2, nsrc * d.C2), (w_tc, nw)))
        dp, f16 = d, np.float32
    if _is_up2conv(d):
        # four 3x3 phase convolutions on the low-res source with the mode-8 pack, then depth-to-space
        B, H, W, Ci, Co = d.B, d.H, d.W, d.C1, d.Cout
        a = _t(_a(x1, B * H * W * Ci, f16)).reshape(B, H, W, Ci).permute(0, 3, 1, 2)
        wt = _t(_a(w_tc, 36 * Co * Ci, f16)).reshape(3, 3, 4 * Co, Ci)
        out = F.conv2d(a, wt.permute(2, 3, 0, 1).contiguous(), padding=1)          # [B,4Co,H,W]
        out = out.reshape(B, 2, 2, Co, H, W).permute(0, 4, 1, 5, 2, 3).reshape(B, 2 * H, 2 * W, Co)
        if bias:
            out = out + _t(_a(bias, Co, np.float32))
        out = _act(out, d.act, d.slope)
        _a(y, B * 4 * H * W * Co, f16)[:] = out.numpy().reshape(-1).astype(f16)
        return 0
    if d.transposed == 2:
        # Deconv2DLayer 2x2 stride 2: 1x1 convolution with N = (phase, co) (pack mode 17), then depth-to-space
        B, H, W, Ci, Co = d.B, d.H, d.W, d.C1 + d.C2, d.Cout
        a = _t(_a(x1, B * H * W * d.C1, f16)).reshape(B * H * W, d.C1)
        if d.C2:
            a = torch.cat([a, _t(_a(x2, B * H * W * d.C2, f16)).reshape(B * H * W, d.C2)], 1)
        wt = _t(_a(w_tc, 4 * Co * Ci, f16)).reshape(4 * Co, Ci)
        out = (a @ wt.t()).reshape(B, H, W, 2, 2, Co).permute(0, 1, 3, 2, 4, 5).reshape(B, 2 * H, 2 * W, Co)
        if bias:
            out = out + _t(_a(bias, Co, np.float32))
        out = _act(out, d.act, d.slope)
        _a(y, B * 4 * H * W * Co, f16)[:] = out.numpy().reshape(-1).astype(f16)
        return 0
    if _is_dgrad_s2(d):
        # 2x2-tap convolution of dy on its own grid with N = (phase, ci), then depth-to-space (pack mode 12)
        B, H, W, Co, Ci = d.B, d.H, d.W, d.C1, d.Cout
        g = _t(_a(x1, B * H * W * Co, f16)).reshape(B, H, W, Co).permute(0, 3, 1, 2)
        wt = _t(_a(w_tc, 16 * Co * Ci, f16)).reshape(2, 2, 4 * Ci, Co)
        out = F.conv2d(F.pad(g, (0, 1, 0, 1)), wt.permute(2, 3, 0, 1).contiguous())           # [B,4Ci,H,W]
        out = out.reshape(B, 2, 2, Ci, H, W).permute(0, 4, 1, 5, 2, 3).reshape(B, 2 * H, 2 * W, Ci)
        dst = _a(y, B * 4 * H * W * Ci, f16)
        if d.accumulate & 1:
            out = out + _t(dst).reshape(out.shape)
        dst[:] = out.numpy().reshape(-1).astype(f16)
        return 0
    Ct = d.C1 + d.C2
    wt = _a(w_tc, d.kh * d.kw * Ct * d.Cout, f16).reshape(d.kh * d.kw, d.Cout, Ct)
    wk = np.ascontiguousarray(wt.transpose(0, 2, 1))           # [tap][ci][co] = the gather kernel's pack
    return hm_conv_gather(dp, x1, x2, wk.ctypes.data, bias, y, y2)


def _is_up2_wgrad(d):
    return (d.up == 1 and d.kh == 5 and d.kw == 5 and d.pad == 2 and d.stride == 1 and not d.transposed and d.C2 == 0
            and d.Ho == 2 * d.H and d.Wo == 2 * d.W and d.oH == d.Ho and d.oW == d.Wo and d.os == 1 and not d.ou
            and not d.ov and d.C1 > 0 and d.C1 % 64 == 0 and d.Cout > 0 and d.Cout % 64 == 0)


def hm_tc_wgrad(dp, x1, x2, dy, dw, stream=None):
    d = dp._obj if hasattr(dp, "_obj") else dp
    assert _tc_ok(d, True)
    if _is_up2_wgrad(d):
        # phase form: dw[(tap3, ci)][(phase, co)] += gradient of the four 3x3 phase filters (unpack mode 8)
        if d.dtype == BF16X3:
            d, _keep, (x1, _x2, dy) = _bf16x3_as_f32(d, ((x1, d.B * d.H * d.W * d.C1), (None, 0),
                                                          (dy, d.B * d.Ho * d.Wo * d.Cout)))
        Ci, Co = d.C1, d.Cout
        tmp = np.zeros(36 * Ci * Co, np.float32)
        hm_up2conv_wgrad_phases(d, x1, dy, tmp.ctypes.data)                     # [phase][(tap3, ci)][co]
        out = _a(dw, 36 * Ci * Co, np.float32).reshape(9 * Ci, 4, Co)
        out += tmp.reshape(4, 9 * Ci, Co).transpose(1, 0, 2)
        return 0
    if d.dtype == BF16X3:
        nsrc = d.B * d.H * d.W
        dp, _keep, (x1, x2, dy) = _bf16x3_as_f32(d, ((x1, nsrc * d.C1), (x2, nsrc * d.C2),
                                                     (dy, d.B * d.Ho * d.Wo * d.Cout)))
    return hm_conv_wgrad(dp, x1, x2, dy, dw)


def hm_im2col_c1(x, xc, B, H, W, kh, kw, pad, stream=None):
    a = _t(_a(x, B * H * W, np.float16)).reshape(B, 1, H, W)
    cols = F.unfold(a, (kh, kw), padding=pad)                                  # [B, kh*kw, L]
    Ho, Wo = H + 2 * pad - kh + 1, W + 2 * pad - kw + 1
    assert (Ho, Wo) == (H, W)
    out = torch.zeros(B, H, W, 64)
    out[..., :kh * kw] = cols.reshape(B, kh * kw, H, W).permute(0, 2, 3, 1)
    _a(xc, B * H * W * 64, np.float16)[:] = out.numpy().reshape(-1).astype(np.float16)
    return 0


def hm_im2col_thin(x1, x2, xc, B, H, W, C1, C2, kh, kw, stride, pad, Ho, Wo, stream=None):
    a = _t(_a(x1, B * H * W * C1, np.float16)).reshape(B, H, W, C1)
    if C2:
        a = torch.cat([a, _t(_a(x2, B * H * W * C2, np.float16)).reshape(B, H, W, C2)], 3)
    Ct = C1 + C2
    cols = F.unfold(a.permute(0, 3, 1, 2), (kh, kw), padding=pad, stride=stride)        # [B, Ct*kh*kw, L], (c, r, s) order
    L = cols.shape[2]
    assert L == Ho * Wo, (L, Ho, Wo)
    cols = cols.reshape(B, Ct, kh * kw, L).permute(0, 3, 2, 1).reshape(B, Ho, Wo, kh * kw * Ct)   # (tap, c) order
    out = torch.zeros(B, Ho, Wo, 64)
    out[..., :kh * kw * Ct] = cols
    _a(xc, B * Ho * Wo * 64, np.float16)[:] = out.numpy().reshape(-1).astype(np.float16)
    return 0


def hm_c1s2_conv(x, wk, bias, y, idx, B, H, W, ncols, act, slope, stream=None):
    """A[q][u*6+v] = x[2qy-2+u][2qx-2+v]; y = act(A . wk^T + bias) (ncols 64) or its max over the 4 column groups."""
    a = _t(_a(x, B * H * W, np.float16)).reshape(B, 1, H, W)
    cols = F.unfold(a, (6, 6), padding=2, stride=2)                            # [B, 36, Hq*Wq]
    Hq, Wq = H // 2, W // 2
    wk_ = _t(_a(wk, ncols * 64, np.float16)).reshape(ncols, 64)[:, :36]
    out = torch.einsum("bkl,nk->bln", cols, wk_).reshape(B, Hq, Wq, ncols)
    bv = _t(_a(bias, 64, np.float32)) if bias else torch.zeros(64)
    if ncols == 256:
        o4 = out.reshape(B, Hq, Wq, 4, 64)
        m, k = o4.max(dim=3)
        out = m
    res = _act(out + bv, act, slope)
    r16 = res.numpy().reshape(-1).astype(np.float16)
    _a(y, B * Hq * Wq * 64, np.float16)[:] = r16
    if ncols == 256:       # bits 0-1: argmax position; bit 2: the stored value is on the slope-1 side of the activation
        side = (r16 > 0) if act == 2 else (r16 >= 0)
        _a(idx, B * Hq * Wq * 64, np.uint8)[:] = k.numpy().reshape(-1).astype(np.uint8) | (side.astype(np.uint8) << 2)
    return 0


def hm_c1s2_bwd(x, g, pooled, idx, wk2, dwk, u, img_scale, B, H, W, act, slope, stream=None):
    Hq, Wq = H // 2, W // 2
    n = B * Hq * Wq * 64
    if pooled:
        gg = _t(_a(g, n, np.float16)) * _act_grad_from_out(_t(_a(pooled, n, np.float16)), act, slope)
    else:                  # act' from bit 2 of the argmax bytes
        side = torch.from_numpy(((_a(idx, n, np.uint8) >> 2) & 1).astype(np.float32))
        neg = slope if act == 1 else (0.0 if act == 2 else 1.0)
        gg = _t(_a(g, n, np.float16)) * (side + (1 - side) * neg)
    if img_scale:
        gg = (gg.reshape(B, -1) * _t(_a(img_scale, B, np.float32)).view(B, 1)).reshape(-1)
    gg = gg.half().float().reshape(B * Hq * Wq, 64)                            # the kernel rounds g*act' to fp16
    k = torch.from_numpy((_a(idx, n, np.uint8) & 3).astype(np.int64)).reshape(B * Hq * Wq, 64)
    G4 = torch.stack([gg * (k == d) for d in range(4)], 1).reshape(B * Hq * Wq, 256)   # [(w)][(d,co)]
    if dwk:
        a = _t(_a(x, B * H * W, np.float16)).reshape(B, 1, H, W)
        cols = F.unfold(a, (6, 6), padding=2, stride=2).permute(0, 2, 1).reshape(B * Hq * Wq, 36)
        A = torch.zeros(B * Hq * Wq, 64)
        A[:, :36] = cols
        A[:, 36] = 1.0
        _a(dwk, 256 * 64, np.float32)[:] += (G4.double().t() @ A.double()).float().numpy().reshape(-1)
    if u:
        w2 = _t(_a(wk2, 256 * 64, np.float16)).reshape(4, 64, 64)             # [d][k][co]
        U = torch.einsum("wdc,dkc->wk", G4.reshape(-1, 4, 64), w2)
        _a(u, n, np.float16)[:] = U.numpy().reshape(-1).astype(np.float16)
    return 0


def hm_c1s2_wgrad(dy, x, dwk, B, H, W, stream=None):
    """dwk[ci][u*6+v] += sum_q x[q][ci] * dy[2q-2+(u,v)]  (column 36 collects sum_q x[q][ci], as the kernel's does)."""
    Hq, Wq = H // 2, W // 2
    a = _t(_a(dy, B * H * W, np.float16)).reshape(B, 1, H, W)
    cols = F.unfold(a, (6, 6), padding=2, stride=2).permute(0, 2, 1).reshape(B * Hq * Wq, 36)
    A = torch.zeros(B * Hq * Wq, 64)
    A[:, :36] = cols
    A[:, 36] = 1.0
    xs = _t(_a(x, B * Hq * Wq * 64, np.float16)).reshape(B * Hq * Wq, 64)
    _a(dwk, 64 * 64, np.float32)[:] += (xs.double().t() @ A.double()).float().numpy().reshape(-1)
    return 0


def hm_c1s2_bwd_fold(dwk, dw, db, cout, stream=None):
    D = _a(dwk, 256 * 64, np.float32).reshape(4, cout, 64)
    out = np.zeros((cout, 5, 5), np.float32)
    for a in range(5):
        for b in range(5):
            r, s_ = 4 - a, 4 - b
            for dd in range(4):
                out[:, a, b] += D[dd, :, (r + (dd >> 1)) * 6 + (s_ + (dd & 1))]
    _a(dw, cout * 25, np.float32)[:] = out.reshape(-1)
    if db:
        _a(db, cout, np.float32)[:] = D[:, :, 36].sum(0)
    return 0


def hm_c1s2_col2im(u, dx, B, H, W, stream=None):
    Hq, Wq = H // 2, W // 2
    U = _t(_a(u, B * Hq * Wq * 64, np.float16)).reshape(B, Hq * Wq, 64)[:, :, :36].permute(0, 2, 1)   # [B,36,L]
    img = F.fold(U, (H, W), (6, 6), padding=2, stride=2)                       # adjoint of the 6x6 stride-2 unfold
    _a(dx, B * H * W, np.float16)[:] = img.numpy().reshape(-1).astype(np.float16)
    return 0


def hm_s2d_pad64(dy, out, B, h, w, Co, stream=None):
    g = _a(dy, B * 4 * h * w * Co, np.float16).reshape(B, h, 2, w, 2, Co)
    o = np.zeros((B, h, w, 64), np.float16)
    o[..., :4 * Co] = g.transpose(0, 1, 3, 2, 4, 5).reshape(B, h, w, 4 * Co)
    _a(out, B * h * w * 64, np.float16)[:] = o.reshape(-1)
    return 0


def hm_up2conv_wgrad_phases(dp, x, dy, dw, stream=None):
    d = dp._obj if hasattr(dp, "_obj") else dp
    B, H, W, Ci, Co = d.B, d.H, d.W, d.C1, d.Cout
    np_dt = _NP[d.dtype]
    a = _t(_a(x, B * H * W * Ci, np_dt)).reshape(B, H, W, Ci).permute(0, 3, 1, 2)
    cols = F.unfold(a, (3, 3), padding=1).reshape(B, Ci, 9, H * W).permute(0, 2, 1, 3).reshape(B, 9 * Ci, H * W)
    g = _t(_a(dy, B * 4 * H * W * Co, np_dt)).reshape(B, 2 * H, 2 * W, Co)
    out = _a(dw, 36 * Ci * Co, np.float32).reshape(4, 9 * Ci, Co)
    for ph in range(4):
        gp = g[:, (ph >> 1)::2, (ph & 1)::2, :].reshape(B, H * W, Co)
        out[ph] += torch.einsum("bkl,blc->kc", cols.double(), gp.double()).float().numpy()
    return 0


_FUNCS = {k: v for k, v in globals().items() if k.startswith("hm_")}


def query(name, dp):
    d = dp._obj if hasattr(dp, "_obj") else dp
    if name == "hm_tc_conv_ws_bytes":
        return 0                       # the emulation never splits K
    if name == "hm_tc_conv_pool_supported":
        return int(_pool_ok(d))
    return int(_tc_ok(d, name == "hm_tc_wgrad_supported"))


def call(name, *args):
    rc = _FUNCS[name](*args)
    if rc != 0:
        raise RuntimeError("%s failed in the emulation (%d)" % (name, rc))
