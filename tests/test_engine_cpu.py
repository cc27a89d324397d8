"""Host logic of the product path (graph lowering, backward schedule, simultaneous
update) checked against the oracle on the CPU.  The CUDA kernels are replaced by
tests/fake_hmgan.py (an emulation of the C ABI), so what is under test here is
everything ABOVE the C ABI; the kernels themselves are checked in the gpu tests."""
import os
import sys

import numpy as np
import pytest
import torch

import fake_hmgan
from oracle import step as S

PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gan-heightmaps_b200")
if PKG not in sys.path:
    sys.path.insert(0, PKG)

import _lib                     # noqa: E402
import lasagne_compat as L      # noqa: E402
from architectures import dcgan, p2p    # noqa: E402
from pix2pix import Pix2Pix     # noqa: E402


@pytest.fixture
def cpu_backend(monkeypatch):
    monkeypatch.setattr(_lib, "call", fake_hmgan.call)
    monkeypatch.setattr(_lib, "query", fake_hmgan.query)
    monkeypatch.setattr(_lib, "load", lambda: None)


def _nl(name):
    return {"linear": L.linear, "tanh": L.tanh, "sigmoid": L.sigmoid}[name]


def build_pair(cfg, train_mode, precision="parity", opt="rmsprop", lr=1e-3, with_p2p=True, seed=2, device="cpu",
               lsgan=True, reconstruction='l1', with_dcgan=True):
    """The same seeded weights in the oracle and in the product model."""
    which = (('G', 'D') if with_dcgan else ()) + (('P', 'Dp') if with_p2p else ())
    nets = S.build_nets(cfg, seed=seed, which=which)
    om = S.OracleModel(nets, alpha=100., opt=opt, lr=lr, train_mode=train_mode, lsgan=lsgan,
                       reconstruction=reconstruction)
    dp = dict(cfg['D'])
    dp['nonlinearity'] = _nl(dp['nonlinearity'])
    kw = dict(gen_fn_dcgan=dcgan.default_generator, disc_fn_dcgan=dcgan.default_discriminator,
              gen_params_dcgan=cfg['G'], disc_params_dcgan=dp,
              gen_fn_p2p=None, disc_fn_p2p=None, gen_params_p2p={}, disc_params_p2p={})
    if not with_dcgan:
        kw.update(gen_fn_dcgan=None, disc_fn_dcgan=None)
    if with_p2p:
        pp, dpp = dict(cfg['P']), dict(cfg['Dp'])
        pp['act'] = _nl(pp['act'])
        dpp['act'] = _nl(dpp['act'])
        kw.update(gen_fn_p2p=p2p.g_unet, disc_fn_p2p=p2p.discriminator, gen_params_p2p=pp, disc_params_p2p=dpp)
    m = Pix2Pix(in_shp=cfg['in_shp'], latent_dim=cfg['latent_dim'], is_a_grayscale=True, is_b_grayscale=False,
                lsgan=lsgan, reconstruction=reconstruction, opt=L.rmsprop if opt == "rmsprop" else L.adam,
                opt_args={'learning_rate': L.shared(L.floatX(lr))}, train_mode=train_mode, verbose=False,
                device=device, precision=precision, seed=0, **kw)
    for k, net in (('G', m.G), ('D', m.D), ('P', m.P), ('Dp', m.Dp)):
        if net is not None:
            net.set_all_param_values(om.get_all_param_values(k))
    return om, m


def _check_params(om, m, rtol, atol):
    for k, net in (('G', m.G), ('D', m.D), ('P', m.P), ('Dp', m.Dp)):
        if net is None:
            continue
        for i, (a, b) in enumerate(zip(net.get_all_param_values(), om.get_all_param_values(k))):
            np.testing.assert_allclose(a, b, rtol=rtol, atol=atol, err_msg="%s param %d" % (k, i))


def _check_grads(om, m, keys, rtol, rtol_by_net=None):
    """Per-array max error relative to that array's scale; arrays whose true gradient is zero (conv biases
    in front of a BatchNorm) are compared against the network's overall gradient scale instead."""
    nets = {'G': m.G, 'D': m.D, 'P': m.P, 'Dp': m.Dp}
    base = rtol
    for k in keys:
        rtol = (rtol_by_net or {}).get(k, base)
        ref = om.last_grads[k]
        net_scale = max(float(np.abs(b).max()) for b in ref)
        for i, (a, b) in enumerate(zip(nets[k].get_grads(), ref)):
            err = float(np.abs(a - b).max())
            assert err <= rtol * float(np.abs(b).max()) + 1e-5 * net_scale, (k, i, err, float(np.abs(b).max()))


def _check_grads_l2(om, m, keys, tol, scale=1.0, tol_by_net=None):
    """Per-array relative L2 error of the weight gradients (`scale` undoes the loss scale).  Arrays whose true gradient
    is zero (conv biases in front of a BatchNorm) are compared with the network's largest gradient norm.

    On max-pool ties: a 2x2 window whose two largest values agree to float32 rounding sends its gradient to a
    different pixel under ANY arithmetic that is not bit-identical to the oracle's.  Measured on the CPU emulation of
    the tc32 mode (64-px DCGAN, batch 2; forward values agree with the SIMT float32 path to 2e-6): ONE window of 65536
    in the discriminator's second pool flips, which moves d loss / d G(z) by 5e-4 and the generator's gradient arrays by
    1e-3 (max norm and L2 alike) while the discriminator's own arrays stay at 2e-5.  Callers therefore give the
    generator -- whose whole gradient is routed through the discriminator's pools -- a separate bound (tol_by_net)."""
    nets = {'G': m.G, 'D': m.D, 'P': m.P, 'Dp': m.Dp}
    base = tol
    for k in keys:
        tol = (tol_by_net or {}).get(k, base)
        ref = om.last_grads[k]
        net_norm = max(float(np.linalg.norm(b.ravel())) for b in ref)
        for i, (a, b) in enumerate(zip(nets[k].get_grads(), ref)):
            err = float(np.linalg.norm((a * scale - b).ravel()))
            assert err <= tol * float(np.linalg.norm(b.ravel())) + 1e-5 * net_norm, \
                (k, i, a.shape, err / (float(np.linalg.norm(b.ravel())) + 1e-30))


def test_gate64_dcgan_step_matches_oracle(cpu_backend):
    cfg = S.experiment_kwargs('gate64')
    om, m = build_pair(cfg, 'dcgan', with_p2p=False)
    for it in range(2):
        Z, X, Y = S.synthetic_batch(4, cfg['latent_dim'], 64, seed=10 + it)
        lo = om.train_fn(Z, X, Y)
        lm = m.train_fn(Z, X, Y)
        np.testing.assert_allclose(lm[:2], lo[:2], rtol=2e-4, atol=1e-6)
        if it == 0:
            _check_grads(om, m, ('G', 'D'), 2e-3)
    _check_params(om, m, rtol=2e-3, atol=2e-4)
    # forward-only entry points
    Z = np.random.RandomState(5).rand(4, cfg['latent_dim']).astype(np.float32)
    np.testing.assert_allclose(m.z_fn_det(Z), om.z_fn_det(Z), rtol=1e-3, atol=1e-4)
    np.testing.assert_allclose(m.z_fn(Z), om.z_fn(Z), rtol=1e-3, atol=1e-4)
    lo = om.loss_fn(Z, X, Y)
    lm = m.loss_fn(Z, X, Y)
    np.testing.assert_allclose(lm[:2], lo[:2], rtol=1e-3, atol=1e-6)
    _check_params(om, m, rtol=2e-3, atol=2e-4)     # BN running statistics moved identically


TINY = S.experiment_kwargs('tiny512')


@pytest.mark.parametrize("bilinear", [True, False])
def test_joint_512_step_matches_oracle(cpu_backend, bilinear):
    """Full test1_nobn_bilin_both topology (reference experiments.py:98-119) at reduced width."""
    cfg = dict(TINY)
    cfg['P'] = dict(TINY['P'], bilinear_upsample=bilinear)
    om, m = build_pair(cfg, 'both')
    Z, X, Y = S.synthetic_batch(2, cfg['latent_dim'], 512, seed=3)
    lo = om.train_fn(Z, X, Y)
    lm = m.train_fn(Z, X, Y)
    np.testing.assert_allclose(lm, lo, rtol=5e-4, atol=1e-6)
    # G's gradient arrives through D's seven 2x2 max-pools over 512x512 maps: with ~1e6 pooling windows a
    # handful have top-two values within float32 rounding of each other, the two implementations pick
    # different argmaxes there, and at this toy width (4..16 channels) each flip is a visible fraction of a
    # weight gradient.  (Fed bit-identical fake images the two d/dG(z) agree to 2e-6 everywhere except at
    # those few hundred pixels; the 64-px gate pins the same code path to 1e-6.)  D, P and Dp are tight.
    _check_grads(om, m, ('G', 'D', 'P', 'Dp'), 1e-3, {'G': 5e-2})
    X1 = X[:1]
    np.testing.assert_allclose(m.gen_fn_det(X1), om.gen_fn_det(X1), rtol=2e-3, atol=2e-4)


def test_train_modes_touch_only_their_networks(cpu_backend):
    cfg = dict(TINY)
    om, m = build_pair(cfg, 'p2p')
    before = {k: [a.copy() for a in net.get_all_param_values()] for k, net in (('G', m.G), ('D', m.D))}
    Z, X, Y = S.synthetic_batch(1, cfg['latent_dim'], 512, seed=4)
    lo = om.train_fn(Z, X, Y)
    lm = m.train_fn(Z, X, Y)
    np.testing.assert_allclose(lm, lo, rtol=5e-4, atol=1e-6)       # all five losses are still reported
    for i, (a, b) in enumerate(zip(m.D.get_all_param_values(), before['D'])):
        np.testing.assert_array_equal(a, b)
    # G's trainable parameters are untouched, its BN running statistics still moved (non-deterministic graph)
    moved = 0
    for p, a, b in zip(m.G.params, m.G.get_all_param_values(), before['G']):
        if p.trainable:
            np.testing.assert_array_equal(a, b)
        else:
            moved += int(not np.array_equal(a, b))
    assert moved > 0
    _check_params(om, m, rtol=5e-3, atol=5e-4)


def test_checkpoint_roundtrip(cpu_backend, tmp_path):
    cfg = S.experiment_kwargs('gate64')
    om, m = build_pair(cfg, 'dcgan', with_p2p=False)
    f = str(tmp_path / "1.model")
    m.save_model(f)
    vals = m.G.get_all_param_values()
    m.G.set_all_param_values([np.zeros_like(v) for v in vals])
    m.load_model(f, mode='dcgan')
    for a, b in zip(m.G.get_all_param_values(), vals):
        np.testing.assert_array_equal(a, b)
    with pytest.raises(ValueError):
        m.G.set_all_param_values(vals[:-1])


def test_param_counts_match_reference_notebook():
    # g_unet.ipynb:481 / :558 (the only known answers the reference records)
    assert L.count_params(p2p.g_unet(512, True, False, nf=64, bilinear_upsample=False)) == 22882243
    assert L.count_params(p2p.discriminator(512, True, False, nf=32)["out"]) == 391009
    g = dcgan.default_generator(1000, True, num_repeats=0, div=[2, 2, 4, 4, 8, 8, 8])
    assert len(L.get_all_params(g)) == 50 and L.count_params(g) == 14792961
    d = dcgan.default_discriminator(512, True, num_repeats=0, bn=False, nonlinearity=L.linear,
                                    div=[8, 4, 4, 4, 2, 2, 2])
    assert len(L.get_all_params(d)) == 16 and L.count_params(d) == 5129217
    assert d.output_shape == (None, 1)
    assert g.output_shape == (None, 1, 512, 512)


def test_unsupported_configurations_fail_loudly(cpu_backend):
    with pytest.raises(NotImplementedError):
        Pix2Pix(dcgan.default_generator, dcgan.default_discriminator, {'nch': 128, 'div': [2, 2, 4, 4]},
                {'nch': 128, 'div': [8, 4, 2, 1], 'nonlinearity': L.linear},    # nch != in_shp: head malformed
                None, None, {}, {}, in_shp=64, latent_dim=8, is_a_grayscale=True, is_b_grayscale=False,
                verbose=False, device="cpu", seed=0)


def test_fast_mode_lowering_uses_tensor_core_kernels_where_eligible(cpu_backend, monkeypatch):
    """Host logic of the fp16 path: 64-channel-multiple stride-1 convolutions are routed to hm_tc_conv /
    hm_tc_wgrad (forward, input gradient, weight gradient), the rest to the gather kernels; the step still
    tracks the float32 oracle (fp16 storage: 2e-2 on the losses)."""
    calls = {}
    real = fake_hmgan.call

    def counting(name, *a):
        calls[name] = calls.get(name, 0) + 1
        return real(name, *a)
    monkeypatch.setattr(_lib, "call", counting)
    cfg = dict(in_shp=64, latent_dim=32,
               G=dict(nch=256, num_repeats=0, div=[2, 2, 4, 4]),                 # 128,128,64,64 channels
               D=dict(nch=64, num_repeats=0, bn=False, nonlinearity='linear', div=[1, 1, 1, 1]))
    om, m = build_pair(cfg, 'dcgan', with_p2p=False, precision="fast")
    Z, X, Y = S.synthetic_batch(2, cfg['latent_dim'], 64, seed=1)
    lo = om.train_fn(Z, X, Y)
    lm = m.train_fn(Z, X, Y)
    np.testing.assert_allclose(lm[:2], lo[:2], rtol=2e-2, atol=1e-4)
    assert calls.get("hm_tc_conv", 0) >= 10 and calls.get("hm_tc_wgrad", 0) >= 6, calls
    # the one-channel ends: D's conv5x5(1->64)+LeakyReLU+max-pool in one pass (real+fake batch), and the input gradient of
    # G's nearest-2x -> conv5x5(64->1), both through hm_c1s2_conv
    assert calls.get("hm_c1s2_conv", 0) >= 2 and m.D.ops[0].pool_fused is not None and m.G.ops[-1].c1dg, calls
    assert calls.get("hm_c1s2_wgrad", 0) == 1 and m.G.ops[-1].c1wg, calls      # ... and that layer's weight gradient
    paths = [op.path for op in m.G.ops + m.D.ops if hasattr(op, "path")]
    # every convolution of this model is on the tensor cores, the DenseLayer included (1x1 convolution, ragged K)
    # (hm_conv_gather remains for the input gradient of the discriminator's one-channel head)
    assert set(paths) == {"tcgen05"} and m.G.ops[0].dense_tc and calls.get("hm_conv_gather", 0) <= 1, (paths, calls)


def test_generator_output_layer_weight_gradient_routes_agree(cpu_backend, monkeypatch):
    """The weight gradient of G's last layer (nearest-2x -> conv5x5, 64 -> 1): hm_c1s2_wgrad + unpack mode 14 (6x6 stride-2
    patches of dy against the low-res source) against the route it replaced (hm_s2d_pad64 + 3x3 phase weight gradient +
    unpack mode 10), from identical states: the same dy and source, so the same gradient up to float32 summation order."""
    cfg = dict(in_shp=64, latent_dim=32,
               G=dict(nch=256, num_repeats=0, div=[2, 2, 4, 4]),
               D=dict(nch=64, num_repeats=0, bn=False, nonlinearity='linear', div=[1, 1, 1, 1]))
    Z, X, Y = S.synthetic_batch(2, cfg['latent_dim'], 64, seed=1)
    grads = {}
    for knob in ("1", "0"):
        monkeypatch.setenv("HMGAN_C1WG", knob)
        om, m = build_pair(cfg, 'dcgan', with_p2p=False, precision="fast")
        vals = m.D.get_all_param_values()
        vals[-1][:] = 0.6                     # a live ReLU head: non-zero gradients
        m.D.set_all_param_values(vals)
        m.train_fn(Z, X, Y)
        assert m.G.ops[-1].c1wg == (knob == "1")
        tr = [q for q in m.G.params if q.trainable]
        iw = max(i for i, q in enumerate(tr) if q.kind == "W")
        assert tuple(tr[iw].shape) == (1, 64, 5, 5)
        grads[knob] = m.G.get_grads()[iw]
    a, b = grads["1"].ravel().astype(np.float64), grads["0"].ravel().astype(np.float64)
    assert np.linalg.norm(b) > 0
    assert np.linalg.norm(a - b) <= 1e-5 * np.linalg.norm(b), np.linalg.norm(a - b) / np.linalg.norm(b)


def test_fast_mode_stride2_layers_route_to_tensor_cores(cpu_backend, monkeypatch):
    """pix2pix encoder / PatchGAN 3x3 stride-2 convolutions with 64-multiple channels: forward and weight gradient
    through hm_tc_conv / hm_tc_wgrad with TMA element strides, input gradient as the 2x2-tap phase convolution
    (pack mode 12).  A two-level U-Net-like chain is checked against the float32 oracle ops."""
    import lasagne_compat as LC
    import engine
    from oracle import lasagne_ops as LO
    r = np.random.RandomState(0)
    inp = LC.InputLayer((None, 64, 16, 16))
    c1 = LC.Conv2DLayer(inp, 128, 3, stride=2, pad='same', nonlinearity=LC.linear)
    a1 = LC.NonlinearityLayer(c1, LC.leaky_rectify)
    c2 = LC.Conv2DLayer(a1, 64, 3, stride=2, pad='same', nonlinearity=LC.linear)
    rt = engine.Runtime("cpu", "fast", loss_scale=1.0)
    net = engine.Net(rt, c2, name="s2", rng=r)
    ops = [op for op in net.ops if isinstance(op, engine.ConvOp)]
    assert all(op.tc_fwd and op.tc_wg for op in ops) and ops[1].dg2
    x = r.randn(2, 64, 16, 16).astype(np.float32)
    net.ensure(2)
    net.inputs[0].buf.copy_(torch.from_numpy(x.transpose(0, 2, 3, 1)).half())
    out = net.forward(2)
    W1, b1, W2, b2 = [torch.tensor(v) for v in net.get_all_param_values()]
    xt = torch.tensor(x, requires_grad=True)
    W1.requires_grad_(True), W2.requires_grad_(True)
    ref = LO.conv2d(LO.leaky_rectify(LO.conv2d(xt, W1, b1, 2, "same"), 0.01), W2, b2, 2, "same")
    np.testing.assert_allclose(out.float().numpy(), ref.detach().permute(0, 2, 3, 1).numpy(), rtol=2e-2, atol=2e-2)
    gy = r.randn(*ref.shape).astype(np.float32)
    ref.backward(torch.tensor(gy))
    net.out.grad.copy_(torch.from_numpy(gy.transpose(0, 2, 3, 1)).half())
    net.backward(0, 2, wgrad=True)
    g = net.get_grads()
    # relative L2 error: a handful of leaky-rectify inputs within fp16 rounding of zero change sign between the two
    # runs, which moves single terms of a weight gradient (max-norm comparisons would see those)
    for a, b in ((g[0], W1.grad.numpy()), (g[2], W2.grad.numpy())):
        assert np.linalg.norm((a - b).ravel()) <= 2e-2 * np.linalg.norm(b.ravel())


def test_fast_mode_unet_bottleneck_deconv_is_a_dense_tensor_core_gemm(cpu_backend, monkeypatch):
    """The U-Net's bottleneck (reference architectures/p2p.py:193-198): conv 2x2 'valid' (2x2 -> 1x1), LeakyReLU, then
    Deconv2DLayer 2x2 stride 1 (1x1 -> 2x2).  In fast mode the deconvolution of the single input pixel runs as 1x1
    tensor-core GEMMs -- forward with pack mode 17 and the bias tiled over the four output positions, input gradient with
    pack mode 22, weight gradient + unpack mode 17 (leading dimension 4*Cout) -- instead of four SIMT gathers per pass.
    Forward and every gradient against the float32 oracle ops (2e-2 in relative L2), and the route against the SIMT one."""
    import lasagne_compat as LC
    import engine
    from oracle import lasagne_ops as LO
    res = {}
    for knob in ("1", "0"):
        monkeypatch.setenv("HMGAN_DC1", knob)
        r = np.random.RandomState(5)
        inp = LC.InputLayer((None, 64, 2, 2))
        c9 = LC.NonlinearityLayer(LC.Conv2DLayer(inp, 128, 2, stride=1, pad='valid', nonlinearity=LC.linear), LC.leaky_rectify)
        d1 = LC.TransposedConv2DLayer(c9, 64, 2, stride=1, nonlinearity=LC.linear)
        rt = engine.Runtime("cpu", "fast", loss_scale=1.0)
        net = engine.Net(rt, d1, name="neck", rng=r)
        convs = [op for op in net.ops if isinstance(op, engine.ConvOp)]
        assert convs[-1].kind == "deconv" and convs[-1].dc1 == (knob == "1") and convs[-1].path == ("tcgen05" if knob == "1" else "simt")
        x = r.randn(3, 64, 2, 2).astype(np.float32)
        net.ensure(3)
        net.inputs[0].buf.copy_(torch.from_numpy(x.transpose(0, 2, 3, 1)).half())
        vals = net.get_all_param_values()
        vals[3][:] = r.randn(*vals[3].shape).astype(np.float32) * 0.1            # a non-zero bias (tiled over the positions)
        net.set_all_param_values(vals)
        y = net.forward(3)
        W1, b1, W2, b2 = [torch.tensor(v, requires_grad=True) for v in net.get_all_param_values()]
        xt = torch.tensor(x)
        ref = LO.deconv2d(LO.leaky_rectify(LO.conv2d(xt, W1, b1, 1, "valid"), 0.01), W2, b2, 1)
        assert tuple(ref.shape) == (3, 64, 2, 2)
        np.testing.assert_allclose(y.float().numpy(), ref.detach().permute(0, 2, 3, 1).numpy(), rtol=2e-2, atol=2e-2)
        gy = r.randn(*ref.shape).astype(np.float32)
        ref.backward(torch.tensor(gy))
        net.out.grad.copy_(torch.from_numpy(gy.transpose(0, 2, 3, 1)).half())
        net.backward(0, 3, wgrad=True)
        g = net.get_grads()
        for a, b in zip(g, (W1.grad, b1.grad, W2.grad, b2.grad)):
            e = np.linalg.norm((a - b.numpy()).ravel()) / (np.linalg.norm(b.numpy().ravel()) + 1e-30)
            assert e <= 2e-2, (a.shape, e)
        res[knob] = [y.float().numpy().copy()] + [a.copy() for a in g]
    for a, b in zip(res["1"], res["0"]):
        assert np.linalg.norm((a - b).ravel()) <= 2e-3 * np.linalg.norm(b.ravel())


def test_fast_mode_unet_ends_route_to_tensor_cores(cpu_backend):
    """The pix2pix ends in fast mode: a thin-source 3x3 stride-2 convolution over a ConcatLayer of a 1- and a 3-channel
    image (PatchGAN layer 1, reference architectures/p2p.py:279-285) runs as hm_im2col_thin + 1x1 tensor-core GEMMs, and
    a Deconv2DLayer 2x2 stride 2 -> 3 channels with tanh over a ConcatLayer (U-Net output layer, p2p.py:272-275) as ONE
    hm_tc_conv launch (pack mode 17) with gradients over hm_s2d_pad64.  Forward and every gradient against the float32
    oracle ops (fp16 storage: 2e-2 in relative L2 norm)."""
    import lasagne_compat as LC
    import engine
    from oracle import lasagne_ops as LO
    r = np.random.RandomState(3)
    ia, ib = LC.InputLayer((None, 1, 16, 16)), LC.InputLayer((None, 3, 16, 16))
    c1 = LC.Conv2DLayer(LC.ConcatLayer([ia, ib]), 64, 3, stride=2, pad='same', nonlinearity=LC.linear)      # 4 -> 64 @8x8
    c2 = LC.Conv2DLayer(LC.NonlinearityLayer(c1, LC.leaky_rectify), 64, 3, stride=1, pad='same', nonlinearity=LC.linear)
    cat = LC.NonlinearityLayer(LC.ConcatLayer([c2, c1]), LC.leaky_rectify)                                   # 128 ch
    out = LC.NonlinearityLayer(LC.TransposedConv2DLayer(cat, 3, 2, stride=2, nonlinearity=LC.linear), LC.tanh)
    rt = engine.Runtime("cpu", "fast", loss_scale=1.0)
    net = engine.Net(rt, out, input_layers=[ia, ib], name="ends", rng=r)
    convs = [op for op in net.ops if isinstance(op, engine.ConvOp)]
    assert convs[0].colk and convs[0].x2 is not None and convs[-1].dc2 and convs[-1].x2 is not None
    A, Bm = r.randn(2, 1, 16, 16).astype(np.float32), r.randn(2, 3, 16, 16).astype(np.float32)
    net.ensure(2, input_grads=(1,))
    net.inputs[0].buf.copy_(torch.from_numpy(A.transpose(0, 2, 3, 1)).half())
    net.inputs[1].buf.copy_(torch.from_numpy(Bm.transpose(0, 2, 3, 1)).half())
    y = net.forward(2)
    W1, b1, W2, b2, W3, b3 = [torch.tensor(v, requires_grad=True) for v in net.get_all_param_values()]
    ta, tb = torch.tensor(A), torch.tensor(Bm, requires_grad=True)
    h1 = LO.conv2d(torch.cat([ta, tb], 1), W1, b1, 2, "same")
    h2 = LO.conv2d(LO.leaky_rectify(h1, 0.01), W2, b2, 1, "same")
    ref = torch.tanh(LO.deconv2d(LO.leaky_rectify(torch.cat([h2, h1], 1), 0.01), W3, b3, 2))
    np.testing.assert_allclose(y.float().numpy(), ref.detach().permute(0, 2, 3, 1).numpy(), rtol=2e-2, atol=2e-2)
    gy = r.randn(*ref.shape).astype(np.float32)
    ref.backward(torch.tensor(gy))
    net.out.grad.copy_(torch.from_numpy(gy.transpose(0, 2, 3, 1)).half())
    net.backward(0, 2, wgrad=True, input_grad=True)

    def rel(a, b):
        return np.linalg.norm((a - b).ravel()) / (np.linalg.norm(b.ravel()) + 1e-30)
    g = net.get_grads()
    for a, b in zip(g, (W1.grad, b1.grad, W2.grad, b2.grad, W3.grad, b3.grad)):
        assert rel(a, b.numpy()) <= 2e-2, (a.shape, rel(a, b.numpy()))
    gb = net.inputs[1].grad[:2].float().numpy()
    assert rel(gb, tb.grad.permute(0, 2, 3, 1).numpy()) <= 2e-2


def test_single_pass_discriminator_backward_equals_two_passes(cpu_backend, monkeypatch):
    """Host logic of the weighted single backward pass through D (hm_adv_loss_pair; pix2pix.py:107-108 share D(G(z)))
    against the reference's two separate passes, on the CPU emulation: same losses, and every gradient of D and G
    equal to fp16-rounding noise (both schedules store the same tensors in fp16; only per-sample scalars move)."""
    cfg = dict(in_shp=64, latent_dim=32,
               G=dict(nch=256, num_repeats=0, div=[2, 2, 4, 4]),
               D=dict(nch=64, num_repeats=0, bn=False, nonlinearity='linear', div=[1, 1, 1, 1]))
    Z, X, Y = S.synthetic_batch(2, cfg['latent_dim'], 64, seed=4)
    res = {}
    for sp in ("1", "0"):
        monkeypatch.setenv("HMGAN_SINGLE_PASS_D", sp)
        _, m = build_pair(cfg, 'dcgan', with_p2p=False, precision="fast")
        assert m._single_pass == (sp == "1")
        vals = m.D.get_all_param_values()
        vals[-1][:] = 0.6                      # D(.) ~ 0.6: neither per-sample factor is negligible
        m.D.set_all_param_values(vals)
        res[sp] = (m.train_fn(Z, X, Y), m.D.get_grads(), m.G.get_grads(), [q for q in m.G.params if q.trainable])
    np.testing.assert_allclose(res["1"][0][:2], res["0"][0][:2], rtol=1e-4, atol=1e-6)
    for k, idx in (("D", 1), ("G", 2)):
        for i, (a, b) in enumerate(zip(res["1"][idx], res["0"][idx])):
            if k == "G" and res["0"][3][i].kind == "b" and i < len(res["0"][idx]) - 1:
                continue                       # biases in front of a BatchNorm: zero gradient in both
            rel = float(np.linalg.norm((a - b).ravel()) / (np.linalg.norm(b.ravel()) + 1e-30))
            assert rel <= 2e-2, (k, i, a.shape, rel)


def test_step_split_into_two_parts_equals_the_whole_step(cpu_backend):
    """The host path replays the step as two CUDA graphs (part 1 = what needs only Z, part 2 = the rest).  On the CPU
    the same split, run eagerly, must give bit-identical losses and parameters to the unsplit step."""
    cfg = S.experiment_kwargs('gate64')
    Z, X, Y = S.synthetic_batch(4, cfg['latent_dim'], 64, seed=6)
    outs = []
    for split in (False, True):
        _, m = build_pair(cfg, 'dcgan', with_p2p=False)
        Zd, Xd = torch.from_numpy(Z), torch.from_numpy(X)
        for _ in range(2):
            if split:
                m._step_eager(Zd, Xd, None, True, part=1)
                losses = m._step_eager(Zd, Xd, None, True, part=2)
            else:
                losses = m._step_eager(Zd, Xd, None, True)
        outs.append((losses.clone().numpy(), m.G.get_all_param_values(), m.D.get_all_param_values()))
    np.testing.assert_array_equal(outs[0][0], outs[1][0])
    for a, b in zip(outs[0][1] + outs[0][2], outs[1][1] + outs[1][2]):
        np.testing.assert_array_equal(a, b)


def test_batched_weight_packs_equal_separate_packs(cpu_backend, monkeypatch):
    """HMGAN_BATCH_PACKS=1 defers a network's weight packs into one hm_pack_conv_weight_multi launch (job table in
    device memory): same step result as the separate launches."""
    cfg = S.experiment_kwargs('gate64')
    Z, X, Y = S.synthetic_batch(4, cfg['latent_dim'], 64, seed=8)
    outs = []
    for flag in ("0", "1"):
        monkeypatch.setenv("HMGAN_BATCH_PACKS", flag)
        _, m = build_pair(cfg, 'dcgan', with_p2p=False)
        losses = [m.train_fn(Z, X, Y) for _ in range(2)]
        outs.append((np.asarray(losses), m.D.get_all_param_values()))
        assert bool(m.rt._pack_tables) == (flag == "1")
    np.testing.assert_array_equal(outs[0][0], outs[1][0])
    for a, b in zip(outs[0][1], outs[1][1]):
        np.testing.assert_array_equal(a, b)


def test_raw_uint8_batches_give_the_step_of_the_host_normalised_batches(cpu_backend):
    """train_fn / gen_fn fed raw uint8 NHWC data (util.Hdf5Iterator(device_normalise=True)) normalise on the device
    (hm_u8_normalize); losses, updated parameters and P(X) must equal, bit for bit in parity mode, those of the float32
    NCHW batches the reference's iterator computes on the host (util.py:28-35)."""
    import util
    cfg = dict(TINY)
    r = np.random.RandomState(3)
    Xu = r.randint(0, 256, (2, 512, 512, 1)).astype(np.uint8)
    Yu = r.randint(0, 256, (2, 512, 512, 3)).astype(np.uint8)
    Z = np.random.RandomState(1).rand(2, cfg['latent_dim']).astype(np.float32)
    outs = []
    for raw in (False, True):
        _, m = build_pair(cfg, 'both')
        X, Y = (Xu, Yu) if raw else (util.normalise_uint8(Xu, True), util.normalise_uint8(Yu, False))
        losses = [m.train_fn(Z, X, Y) for _ in range(2)]
        params = [a for net in (m.G, m.D, m.P, m.Dp) for a in net.get_all_param_values()]
        outs.append((np.array(losses), params, m.gen_fn_det(X[:1]), m.gen_fn(Xu[:1, :, :, 0] if raw else X[:1])))
    np.testing.assert_array_equal(outs[0][0], outs[1][0])
    for a, b in zip(outs[0][1], outs[1][1]):
        np.testing.assert_array_equal(a, b)
    np.testing.assert_array_equal(outs[0][2], outs[1][2])
    np.testing.assert_array_equal(outs[0][3], outs[1][3])
    with pytest.raises(ValueError):
        m.train_fn(Z, Xu[:, :256], Yu)          # wrong image size


def test_reference_default_objective_adam_and_cross_entropy(cpu_backend):
    """The Pix2Pix constructor's own defaults (reference pix2pix.py:24-31): opt=adam, lsgan=False (binary
    cross-entropy on sigmoid discriminators), here with reconstruction='l2' -- the experiments override all three, so
    nothing else exercises hm_adam, the cross-entropy branch of hm_adv_loss and the L2 branch of hm_recon_loss against
    the oracle.  Two steps (Adam's bias correction depends on the step count)."""
    cfg = dict(TINY)
    cfg['D'] = dict(TINY['D'], nonlinearity='sigmoid')
    cfg['Dp'] = dict(TINY['Dp'], act='sigmoid')
    om, m = build_pair(cfg, 'both', opt="adam", lr=2e-4, lsgan=False, reconstruction='l2')
    for it in range(2):
        Z, X, Y = S.synthetic_batch(2, cfg['latent_dim'], 512, seed=30 + it)
        lo = om.train_fn(Z, X, Y)
        lm = m.train_fn(Z, X, Y)
        np.testing.assert_allclose(lm, lo, rtol=1e-3, atol=1e-6)
    # Adam's first steps are ~lr * sign(g) whatever |g| is, so elements whose gradient is rounding noise (G's, through the
    # max-pool argmax ties described above) move by +-lr in either implementation: compare robustly -- the second
    # step's losses above already agree to 1e-3 -- by the fraction of elements further apart than half a step
    for k, net in (('G', m.G), ('D', m.D), ('P', m.P), ('Dp', m.Dp)):
        a = np.concatenate([x.ravel() for x in om.get_all_param_values(k)])
        b = np.concatenate([x.ravel() for x in net.get_all_param_values()])
        far = float((np.abs(a - b) > 1e-4).mean())
        assert far < (0.03 if k == 'G' else 2e-3), (k, far)


def test_num_repeats_adds_convolutions_per_level(cpu_backend):
    """num_repeats=1 (reference dcgan.py:21-22,41-42: an extra conv -> [BN] -> LReLU block per resolution; 0 in every
    experiment) lowers and trains like the oracle.  G's gradient goes through twice as many layers of D here; its
    tolerance is the joint test's (float32 rounding decides a few max-pool argmax ties differently)."""
    cfg = S.experiment_kwargs('gate64')
    cfg['G'] = dict(cfg['G'], num_repeats=1)
    cfg['D'] = dict(cfg['D'], num_repeats=1)
    om, m = build_pair(cfg, 'dcgan', with_p2p=False)
    assert len(m.G.get_all_param_values()) == 2 + 4 + 6 * 8 + 2 and len(m.D.get_all_param_values()) == 2 * 8 + 2
    Z, X, Y = S.synthetic_batch(4, cfg['latent_dim'], 64, seed=6)
    np.testing.assert_allclose(m.train_fn(Z, X, Y)[:2], om.train_fn(Z, X, Y)[:2], rtol=1e-4, atol=1e-6)
    _check_grads(om, m, ('G', 'D'), 1e-3, {'G': 1e-2})


def test_fast_mode_pool_in_the_conv_epilogue(cpu_backend):
    """Conv2DLayer + LeakyReLU + MaxPool2DLayer(2) with rows of >= 128 pixels (the DCGAN discriminator's layers 2 and 3 at
    512x512, reference architectures/dcgan.py:42-47) lower to hm_tc_conv_pool: the un-pooled activation has no buffer,
    forward and every gradient agree with the float32 oracle ops (fp16 storage: 2e-2 relative L2)."""
    import lasagne_compat as LC
    import engine
    from oracle import lasagne_ops as LO
    r = np.random.RandomState(5)
    inp = LC.InputLayer((None, 64, 4, 128))
    c1 = LC.Conv2DLayer(inp, 64, 3, pad='same', nonlinearity=LC.LeakyRectify(0.2))
    p1 = LC.MaxPool2DLayer(c1, 2)
    c2 = LC.Conv2DLayer(p1, 64, 3, pad='same', nonlinearity=LC.linear)
    rt = engine.Runtime("cpu", "fast", loss_scale=1.0)
    net = engine.Net(rt, c2, name="pool", rng=r)
    convs = [op for op in net.ops if isinstance(op, engine.ConvOp)]
    pools = [op for op in net.ops if isinstance(op, engine.PoolOp)]
    assert convs[0].pool_tc is pools[0] and pools[0].fused_tc and convs[1].pool_tc is None
    x = r.randn(2, 64, 4, 128).astype(np.float32)
    net.ensure(2, input_grads=(0,))
    assert convs[0].out.buf is None and convs[0].out.grad is not None
    net.inputs[0].buf.copy_(torch.from_numpy(x.transpose(0, 2, 3, 1)).half())
    y = net.forward(2)
    W1, b1, W2, b2 = [torch.tensor(v, requires_grad=True) for v in net.get_all_param_values()]
    xt = torch.tensor(x, requires_grad=True)
    ref = LO.conv2d(LO.max_pool(LO.leaky_rectify(LO.conv2d(xt, W1, b1, 1, "same"), 0.2)), W2, b2, 1, "same")

    def rel(a, b):
        return np.linalg.norm((np.asarray(a) - np.asarray(b)).ravel()) / (np.linalg.norm(np.asarray(b).ravel()) + 1e-30)
    assert rel(y.float().numpy(), ref.detach().permute(0, 2, 3, 1).numpy()) <= 2e-2
    gy = r.randn(*ref.shape).astype(np.float32)
    ref.backward(torch.tensor(gy))
    net.out.grad.copy_(torch.from_numpy(gy.transpose(0, 2, 3, 1)).half())
    net.backward(0, 2, wgrad=True, input_grad=True)
    for a, b in zip(net.get_grads(), (W1.grad, b1.grad, W2.grad, b2.grad)):
        assert rel(a, b.numpy()) <= 2e-2, (a.shape, rel(a, b.numpy()))
    assert rel(net.inputs[0].grad[:2].float().numpy(), xt.grad.permute(0, 2, 3, 1).numpy()) <= 2e-2
