"""Host logic of the 'tc32' precision (float32 storage, tcgen05 kernels on three-plane bf16 operand splits,
include/hmgan.h HM_BF16X3) on the CPU emulation: descriptor rewriting, the three split layouts, the split weight packs
(incl. ConcatLayer segments) and the float outputs -- against the float32 oracle at the north star's 1e-3."""
import os
import sys

import numpy as np
import torch

from oracle import step as S
from oracle import lasagne_ops as LO

PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gan-heightmaps_b200")
if PKG not in sys.path:
    sys.path.insert(0, PKG)

import _lib             # noqa: E402
import fake_hmgan       # noqa: E402
from test_engine_cpu import build_pair, cpu_backend, _check_grads_l2     # noqa: E402,F401


def test_split_planes_reproduce_float32():
    """h + m + l reproduces a float32 value to 2^-23 relative (three 8-bit significands); the six planes written along
    the reduction axis pair up as h.h + h.m + m.h + h.l + l.h + m.m between the a side (layouts 0, 2) and the b side
    (layouts 1, 3); ConcatLayer segments of a packed weight are split separately."""
    r = np.random.RandomState(0)
    a = (r.randn(6, 10) * np.exp(r.randn(6, 10) * 3)).astype(np.float32)

    def planes(layout):
        out = np.zeros(6 * a.size, np.uint16)
        fake_hmgan.hm_split_bf16x3(a.ctypes.data, out.ctypes.data, 6, 10, 4, layout)
        f = (out.astype(np.uint32) << 16).view(np.float32)
        if layout >= 2:
            return list(f.reshape(6, 6, 10))
        f = f.reshape(6, 60)
        return [np.concatenate([f[:, k * 4:(k + 1) * 4], f[:, 24 + k * 6:24 + (k + 1) * 6]], 1) for k in range(6)]
    for la, lb in ((0, 1), (2, 3)):
        pa, pb = planes(la), planes(lb)
        h, m, l = pa[0], pa[2], pa[4]
        assert np.all(np.abs(h + m + l - a) <= np.abs(a) * 2.0 ** -23)
        for k, (x, y) in enumerate(((h, h), (h, m), (m, h), (h, l), (l, h), (m, m))):
            np.testing.assert_array_equal(pa[k], x)
            np.testing.assert_array_equal(pb[k], y)
        prod = sum(x.astype(np.float64) * y.astype(np.float64) for x, y in zip(pa, pb))
        assert np.all(np.abs(prod - a.astype(np.float64) ** 2) <= a.astype(np.float64) ** 2 * 2.0 ** -22)


def test_tc32_wide_dcgan_step_matches_oracle(cpu_backend, monkeypatch):
    """A 64-px DCGAN whose hidden layers are 64..128 channels wide in tc32 mode: nearest-2x + 5x5 phase convolutions,
    stride-1 forward / input-gradient / weight-gradient convolutions all go through hm_tc_conv / hm_tc_wgrad on split
    operands; the losses and the discriminator's gradient arrays agree with the float32 oracle at 1e-3 (measured 2e-7 and
    2e-5), the generator's at 5e-3 in relative L2 (measured 1e-3: one re-routed max-pool tie, see _check_grads_l2)."""
    calls = {}
    real = fake_hmgan.call

    def counting(name, *a):
        calls[name] = calls.get(name, 0) + 1
        return real(name, *a)
    monkeypatch.setattr(_lib, "call", counting)
    cfg = dict(in_shp=64, latent_dim=32,
               G=dict(nch=256, num_repeats=0, div=[2, 2, 4, 4]),
               D=dict(nch=64, num_repeats=0, bn=False, nonlinearity='linear', div=[1, 1, 1, 1]))
    om, m = build_pair(cfg, 'dcgan', with_p2p=False, precision="tc32")
    assert m.rt.cd == _lib.F32 and m.rt.split and not m._single_pass
    Z, X, Y = S.synthetic_batch(2, cfg['latent_dim'], 64, seed=1)
    lo = om.train_fn(Z, X, Y)
    lm = m.train_fn(Z, X, Y)
    np.testing.assert_allclose(lm[:2], lo[:2], rtol=1e-3, atol=1e-6)
    _check_grads_l2(om, m, ('G', 'D'), 1e-3, tol_by_net={'G': 5e-3})     # one re-routed max-pool tie: see the helper
    assert calls.get("hm_tc_conv", 0) >= 10 and calls.get("hm_tc_wgrad", 0) >= 6, calls
    assert calls.get("hm_split_bf16x3", 0) >= calls["hm_tc_conv"] + 3 * calls["hm_tc_wgrad"], calls
    paths = [op.path for op in m.G.ops + m.D.ops if hasattr(op, "path")]
    assert paths.count("tcgen05") >= 6, paths


def test_tc32_unet_block_matches_oracle(cpu_backend):
    """U-Net vocabulary in tc32 mode: 3x3 stride-2 encoder convolution (TMA element strides; input gradient as the
    2x2-tap phase convolution, pack mode 12), bilinear 2x -> 3x3 convolution, and a convolution over a ConcatLayer (two
    sources: the weight split runs per segment).  Forward, weight gradients and the input gradient within 1e-4 relative
    L2 of the float32 oracle ops."""
    import lasagne_compat as LC
    import engine
    from architectures.layers import BilinearUpsample2DLayer
    r = np.random.RandomState(1)
    inp = LC.InputLayer((None, 64, 16, 16))
    c1 = LC.Conv2DLayer(inp, 128, 3, stride=2, pad='same', nonlinearity=LC.linear)                       # 128 @ 8x8
    a1 = LC.NonlinearityLayer(c1, LC.leaky_rectify)
    c2 = LC.Conv2DLayer(BilinearUpsample2DLayer(a1, 2), 64, 3, stride=1, pad='same', nonlinearity=LC.linear)   # 64 @ 16x16
    c0 = LC.Conv2DLayer(inp, 64, 3, stride=1, pad='same', nonlinearity=LC.linear)                        # skip, 64 @ 16x16
    cat = LC.NonlinearityLayer(LC.ConcatLayer([c2, c0]), LC.leaky_rectify)
    c3 = LC.Conv2DLayer(cat, 64, 3, stride=1, pad='same', nonlinearity=LC.linear)
    rt = engine.Runtime("cpu", "tc32")
    net = engine.Net(rt, c3, name="block", rng=r)
    convs = [op for op in net.ops if isinstance(op, engine.ConvOp)]
    assert all(op.tc_fwd and op.tc_wg for op in convs), [(op.tc_fwd, op.tc_wg) for op in convs]
    assert any(op.x2 is not None for op in convs) and any(op.dg2 for op in convs)
    x = r.randn(2, 64, 16, 16).astype(np.float32)
    net.ensure(2, input_grads=(0,))
    net.inputs[0].buf.copy_(torch.from_numpy(x.transpose(0, 2, 3, 1)))
    y = net.forward(2)
    vals = {id(p): torch.tensor(v, requires_grad=True) for p, v in zip(net.params, net.get_all_param_values())}

    def pv(layer):
        return [vals[id(p)] for p in layer.params]
    xt = torch.tensor(x, requires_grad=True)
    h1 = LO.leaky_rectify(LO.conv2d(xt, *pv(c1), 2, "same"), 0.01)
    h2 = LO.conv2d(LO.bilinear_upsample(h1), *pv(c2), 1, "same")
    h0 = LO.conv2d(xt, *pv(c0), 1, "same")
    ref = LO.conv2d(LO.leaky_rectify(torch.cat([h2, h0], 1), 0.01), *pv(c3), 1, "same")

    def rel(a, b):
        return np.linalg.norm((np.asarray(a) - np.asarray(b)).ravel()) / (np.linalg.norm(np.asarray(b).ravel()) + 1e-30)
    assert rel(y.numpy(), ref.detach().permute(0, 2, 3, 1).numpy()) <= 1e-4
    gy = r.randn(*ref.shape).astype(np.float32)
    ref.backward(torch.tensor(gy))
    net.out.grad.copy_(torch.from_numpy(gy.transpose(0, 2, 3, 1)))
    net.backward(0, 2, wgrad=True, input_grad=True)
    for p, g in zip([q for q in net.params if q.trainable], net.get_grads()):
        assert rel(g, vals[id(p)].grad.numpy()) <= 1e-4, (p.shape, rel(g, vals[id(p)].grad.numpy()))
    assert rel(net.inputs[0].grad[:2].numpy(), xt.grad.permute(0, 2, 3, 1).numpy()) <= 1e-4
