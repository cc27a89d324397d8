"""Step-level parity on the GPU: the product path (Pix2Pix.train_fn & co. ->
engine -> libhmgan kernels) against the oracle on the same seeded weights and
inputs.  float32 'parity' mode: 1e-3 relative (BASELINE.json north_star);
fp16 'fast' mode: tolerances stated per test."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import step as S

PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gan-heightmaps_b200")
if PKG not in sys.path:
    sys.path.insert(0, PKG)

from test_engine_cpu import TINY, build_pair, _check_grads, _check_grads_l2, _check_params   # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("precision", ["parity", "tc32"])
def test_gate64_dcgan_step_parity(precision):
    """BASELINE.json configs[0], in both float32-grade modes: 'parity' (SIMT fp32 kernels) and 'tc32' (the tcgen05
    kernels of the benchmarked path on three-plane bf16 operand splits, 7 of the 11 convolutions at these widths)."""
    cfg = S.experiment_kwargs('gate64')
    om, m = build_pair(cfg, 'dcgan', with_p2p=False, device="cuda", precision=precision)
    if precision == "tc32":
        paths = [op.path for op in m.G.ops + m.D.ops if hasattr(op, "path")]
        assert paths.count("tcgen05") >= 7, paths
    for it in range(3):
        Z, X, Y = S.synthetic_batch(4, cfg['latent_dim'], 64, seed=10 + it)
        lo = om.train_fn(Z, X, Y)
        lm = m.train_fn(Z, X, Y)
        np.testing.assert_allclose(lm[:2], lo[:2], rtol=1e-3, atol=1e-6)
        if it == 0:
            _check_grads(om, m, ('G', 'D'), 1e-3)
    _check_params(om, m, rtol=2e-3, atol=2e-4)
    Z = np.random.RandomState(5).rand(4, cfg['latent_dim']).astype(np.float32)
    np.testing.assert_allclose(m.z_fn_det(Z), om.z_fn_det(Z), rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(m.z_fn(Z), om.z_fn(Z), rtol=1e-3, atol=1e-5)
    assert m.rt.launches > 0


@pytest.mark.parametrize("bilinear", [True, False])
def test_joint_512_step_parity(bilinear):
    """test1_nobn_bilin_both topology (all four networks, 512x512) at reduced width."""
    cfg = dict(TINY)
    cfg['P'] = dict(TINY['P'], bilinear_upsample=bilinear)
    om, m = build_pair(cfg, 'both', device="cuda")
    Z, X, Y = S.synthetic_batch(2, cfg['latent_dim'], 512, seed=3)
    lo = om.train_fn(Z, X, Y)
    lm = m.train_fn(Z, X, Y)
    np.testing.assert_allclose(lm, lo, rtol=1e-3, atol=1e-6)
    _check_grads(om, m, ('G', 'D', 'P', 'Dp'), 1e-3, {'G': 5e-2})   # see tests/test_engine_cpu.py on G
    np.testing.assert_allclose(m.gen_fn_det(X[:1]), om.gen_fn_det(X[:1]), rtol=2e-3, atol=2e-4)


@pytest.mark.parametrize("precision", ["parity", "tc32"])
def test_full_width_dcgan_forward_and_losses_parity(precision):
    """BASELINE configs[1] architecture (full widths, 512x512, z=1000) at batch 2: G(z) and the two DCGAN
    losses against the oracle at 1e-3; range properties of the outputs.  tc32: 15 of the 17 convolutions run on the
    tcgen05 kernels (all but the one-channel ends)."""
    cfg = S.experiment_kwargs('test1_nobn_bilin_both')
    om, m = build_pair(cfg, 'dcgan', with_p2p=False, device="cuda", precision=precision)
    if precision == "tc32":
        paths = [op.path for op in m.G.ops + m.D.ops if hasattr(op, "path")]
        assert paths.count("tcgen05") >= 15, paths
    _set_head_bias(om, m, 0.6)
    Z, X, Y = S.synthetic_batch(2, cfg['latent_dim'], 512, seed=1)
    lo = om.loss_fn(Z, X, Y)
    lm = m.loss_fn(Z, X, Y)
    np.testing.assert_allclose(lm[:2], lo[:2], rtol=1e-3, atol=1e-6)
    gz = m.z_fn_det(Z)
    assert gz.shape == (2, 1, 512, 512) and gz.min() > 0 and gz.max() < 1
    np.testing.assert_allclose(gz, om.z_fn_det(Z), rtol=1e-3, atol=1e-5)


def _set_head_bias(om, m, value):
    """At the Glorot initialisation the ReLU head of the full-width DCGAN discriminator is dead on these inputs (D(.) = 0,
    both losses exactly 1, every gradient exactly 0): a head bias puts D(.) near `value`, where the step is informative."""
    vals = m.D.get_all_param_values()
    vals[-1][:] = value
    m.D.set_all_param_values(vals)
    with torch.no_grad():
        om.params['D'][-1].fill_(value)


# Gradient bounds of the full-width step tests.  The discriminator's max-pools and LeakyReLU kinks make the per-array
# gradient error grow like the SQUARE ROOT of the forward perturbation: every unit whose two candidates (or whose sign)
# are closer than the perturbation re-routes its whole gradient, the number of such units is proportional to the
# perturbation, and the L2 error to the root of that number.  Measured on B200 against the oracle (tools/dbg_tc32.py,
# profiles/r2_parity_sensitivity.txt; DCGAN 512x512 full width, relative L2 of the worst weight array of G / of D):
#     SIMT float32 'parity' (forward agrees to ~1e-6)     batch 2:  5.0e-3 / 8.4e-4
#     'tc32' on tcgen05   (forward agrees to ~1e-5)        batch 2:  2.9e-2 / 6.5e-3      batch 8: 2.3e-2 / 3.6e-3
#     'fast' fp16 on tcgen05 (forward agrees to ~5e-4)     batch 2:  1.8e-1 / 4.6e-2      batch 8: 1.5e-1 / 2.1e-2
# while every convolution kernel by itself agrees with the SIMT float32 kernel to 2.5e-5 (tests/test_tc_gpu.py) and a
# wrong kernel moves these numbers to 0.7 - 1.0.  The bounds below are 2x the measured values.
FULL_WIDTH_BOUNDS = {
    #            losses per step (rtol)   G grads  D grads
    "tc32": ((1e-3, 2e-3, 1e-2), 6e-2, 1e-2),
    "fast": ((1e-3, 5e-3, 5e-2), 0.3, 5e-2),
}


@pytest.mark.parametrize("precision", ["tc32", "fast"])
def test_full_width_dcgan_train_steps_track_oracle(precision):
    """The benchmarked configuration (BASELINE configs[1]: DCGAN 512x512, z=1000, full widths) at batch 8, three train_fn
    steps against the oracle on the path bench.py times -- 'fast' (fp16 tcgen05, CUDA graphs, single-pass discriminator
    backward, side streams) -- and on the same kernels in float32-grade 'tc32' mode.  Losses per step, every
    weight-gradient array of the first step (relative L2; biases in front of a BatchNorm have a zero true gradient and
    are compared with the network's gradient norm) within the bounds justified above, and the direction of every weight
    array's accumulated update after three RMSprop steps."""
    cfg = S.experiment_kwargs('test1_nobn_bilin_both')
    om, m = build_pair(cfg, 'dcgan', with_p2p=False, device="cuda", precision=precision, lr=1e-4)
    _set_head_bias(om, m, 0.6)
    loss_tol, g_tol, d_tol = FULL_WIDTH_BOUNDS[precision]
    p0 = {k: [a.copy() for a in om.get_all_param_values(k)] for k in ('G', 'D')}
    for it in range(3):
        Z, X, Y = S.synthetic_batch(8, cfg['latent_dim'], 512, seed=10 + it)
        lo = om.train_fn(Z, X, Y)
        lm = m.train_fn(Z, X, Y)
        assert np.all(np.isfinite(lm))
        np.testing.assert_allclose(lm[:2], lo[:2], rtol=loss_tol[it], atol=1e-6)
        if it == 0:
            _check_grads_l2(om, m, ('G', 'D'), d_tol, scale=1.0 / m.rt.loss_scale, tol_by_net={'G': g_tol})
    # updated parameters: RMSprop turns a gradient into a step of ~lr*sqrt(10) times its SIGN, so weights whose gradient
    # is smaller than the gradient error may step the other way; the comparable quantity is the direction of the
    # accumulated update of each weight array (cosine with the oracle's)
    for k, net in (('G', m.G), ('D', m.D)):
        for a, b, a0, q in zip(net.get_all_param_values(), om.get_all_param_values(k), p0[k], net.params):
            if not (q.trainable and q.kind == "W"):
                continue
            ua, ub = (a - a0).ravel().astype(np.float64), (b - a0).ravel().astype(np.float64)
            cos = float(ua @ ub / (np.linalg.norm(ua) * np.linalg.norm(ub) + 1e-30))
            assert cos >= UPDATE_COS[precision], (k, a.shape, cos)


UPDATE_COS = {"tc32": 0.98, "fast": 0.95}      # measured on B200: 0.996 and 0.987 (generator, worst array)


def test_full_width_joint_step_parity_on_tensor_cores():
    """BASELINE configs[2] architecture (all four networks at full width, 512x512, bilinear U-Net) at batch 2 in tc32
    mode: 34 of the 40 convolutions on the tcgen05 kernels (forward, input gradient, weight gradient; stride-2,
    concat + bilinear, phase-decomposed and row-box variants).  The five losses and P(x) against the oracle at 1e-3.
    Weight gradients (relative L2): PatchGAN <= 5e-3 (measured 7e-4); U-Net <= 0.1 -- the step is ill-conditioned at
    batch 2 (the 1x1 / 2x2 bottleneck BatchNorms see 2..8 values per channel): the SIMT float32 'parity' mode measures
    3.4e-2 against the same oracle, tc32 2.3e-2 (profiles/r2_parity_sensitivity.txt)."""
    cfg = S.experiment_kwargs('test1_nobn_bilin_both')
    om, m = build_pair(cfg, 'both', device="cuda", precision="tc32")
    paths = [op.path for net in (m.G, m.D, m.P, m.Dp) for op in net.ops if hasattr(op, "path")]
    assert paths.count("tcgen05") >= 34, paths
    _set_head_bias(om, m, 0.6)
    Z, X, Y = S.synthetic_batch(2, cfg['latent_dim'], 512, seed=2)
    np.testing.assert_allclose(m.gen_fn_det(X[:1]), om.gen_fn_det(X[:1]), rtol=1e-3, atol=1e-4)
    lo = om.train_fn(Z, X, Y)
    lm = m.train_fn(Z, X, Y)
    np.testing.assert_allclose(lm, lo, rtol=1e-3, atol=1e-6)
    _check_grads_l2(om, m, ('Dp', 'P'), 5e-3, tol_by_net={'P': 0.1})


def test_fast_mode_tracks_parity_mode():
    """fp16 storage / fp32 accumulate against the float32 oracle on the 64-px gate over three training steps
    (the two trajectories drift apart slowly): losses within 5e-2 relative (2e-3 absolute for the small
    generator loss), G(z) within 2e-2 absolute (values in (0,1))."""
    cfg = S.experiment_kwargs('gate64')
    om, m = build_pair(cfg, 'dcgan', with_p2p=False, device="cuda", precision="fast")
    for it in range(3):
        Z, X, Y = S.synthetic_batch(4, cfg['latent_dim'], 64, seed=10 + it)
        lo = om.train_fn(Z, X, Y)
        lm = m.train_fn(Z, X, Y)
        assert np.all(np.isfinite(lm))
        np.testing.assert_allclose(lm[:2], lo[:2], rtol=5e-2, atol=2e-3)
    Z = np.random.RandomState(5).rand(4, cfg['latent_dim']).astype(np.float32)
    np.testing.assert_allclose(m.z_fn_det(Z), om.z_fn_det(Z), atol=2e-2)


def test_p2p_mode_leaves_dcgan_untouched_and_device_api_matches_host_api():
    cfg = dict(TINY)
    om, m = build_pair(cfg, 'p2p', device="cuda")
    before = [a.copy() for a in m.D.get_all_param_values()]
    Z, X, Y = S.synthetic_batch(1, cfg['latent_dim'], 512, seed=4)
    lm = m.loss_fn(Z, X, Y)
    dev = m.step_device(torch.from_numpy(Z).cuda(), torch.from_numpy(X).cuda(), torch.from_numpy(Y).cuda(),
                        train=False).cpu().numpy()
    np.testing.assert_allclose(dev[2:], np.asarray(lm)[2:], rtol=1e-5)     # G's BN statistics moved in between
    m.train_fn(Z, X, Y)
    for a, b in zip(m.D.get_all_param_values(), before):
        np.testing.assert_array_equal(a, b)


def test_fast_mode_tensor_core_step_tracks_oracle():
    """A 64-px DCGAN whose hidden layers are 64/128 channels wide, so that forward, input-gradient and
    weight-gradient convolutions all run on the tcgen05 kernels.  Against the float32 oracle: losses within
    2e-2; D's weight gradients within 3e-2 of each array's scale; G's weight gradients within 0.2 in relative
    L2 norm.  (G's gradient is routed through D's four max-pools: with fp16 activations the argmax of a
    near-tied 2x2 window differs from the float32 run in a few percent of the windows, which re-routes that
    share of the gradient -- measured 4-5 % per pooling layer in the CPU emulation of the same arithmetic,
    while every non-pooling op tracks float32 to ~3e-4.)"""
    cfg = dict(in_shp=64, latent_dim=32,
               G=dict(nch=256, num_repeats=0, div=[2, 2, 4, 4]),
               D=dict(nch=64, num_repeats=0, bn=False, nonlinearity='linear', div=[1, 1, 1, 1]))
    om, m = build_pair(cfg, 'dcgan', with_p2p=False, device="cuda", precision="fast")
    paths = [op.path for op in m.G.ops + m.D.ops if hasattr(op, "path")]
    assert paths.count("tcgen05") >= 6, paths
    assert m._single_pass and m.D.ops[0].pool_fused is not None and m.G.ops[-1].c1dg
    Z, X, Y = S.synthetic_batch(4, cfg['latent_dim'], 64, seed=1)
    lo = om.train_fn(Z, X, Y)
    lm = m.train_fn(Z, X, Y)
    np.testing.assert_allclose(lm[:2], lo[:2], rtol=2e-2, atol=1e-4)
    scale = 1.0 / m.rt.loss_scale
    ref = om.last_grads['D']
    net_scale = max(float(np.abs(b).max()) for b in ref)
    for i, (a, b) in enumerate(zip(m.D.get_grads(), ref)):
        err = float(np.abs(a * scale - b).max())
        assert err <= 3e-2 * float(np.abs(b).max()) + 1e-3 * net_scale, ('D', i, err, float(np.abs(b).max()))
    for i, (a, b, p) in enumerate(zip(m.G.get_grads(), om.last_grads['G'], [q for q in m.G.params if q.trainable])):
        if p.kind == "b" and i < len(om.last_grads['G']) - 1:
            continue                     # biases in front of a BatchNorm: true gradient is zero
        rel = float(np.linalg.norm((a * scale - b).ravel()) / (np.linalg.norm(b.ravel()) + 1e-30))
        assert rel <= 0.2, ('G', i, rel)


def test_fast_mode_full_width_joint_step_tracks_oracle():
    """BASELINE configs[2] architecture (all four networks at full width, 512x512, bilinear U-Net) at batch 2 in
    fp16 fast mode -- tensor-core forward / input-gradient / weight-gradient kernels incl. the stride-2, the
    concat+bilinear and the super-tile row-box layers -- against the float32 oracle: the five losses within 3e-3.

    Weight gradients, relative L2 error per array.  PatchGAN (no BatchNorm): <= 2e-2.  U-Net: the step is
    ill-conditioned at batch 2 -- the 1x1 / 2x2 bottleneck BatchNorms see 2..8 values per channel, so inv_std
    amplifies any rounding.  Measured on the ORACLE alone (float32, only X rounded to fp16, /tmp experiment recorded
    in DESIGN.md): the decoder arrays move by 0.001 (dconv9), 0.007, 0.023, 0.05, 0.064, 0.079, 0.097, 0.12, 0.13
    (dconv1) and every encoder array by 0.6-0.7.  fp16 storage of every activation is a larger perturbation than
    that, so the bounds are: the seven arrays nearest the loss <= 0.2 (measured 0.002 ... 0.13), the rest <= 0.5
    (measured 0.15 ... 0.23; a wrong kernel gives 0.7-1.0, as the row-box super-tile bug did)."""
    cfg = S.experiment_kwargs('test1_nobn_bilin_both')
    om, m = build_pair(cfg, 'both', device="cuda", precision="fast")
    Z, X, Y = S.synthetic_batch(2, cfg['latent_dim'], 512, seed=2)
    lo = om.train_fn(Z, X, Y)
    lm = m.train_fn(Z, X, Y)
    assert np.all(np.isfinite(lm))
    np.testing.assert_allclose(lm, lo, rtol=3e-3, atol=1e-4)
    scale = 1.0 / m.rt.loss_scale
    for k, net in (('Dp', m.Dp), ('P', m.P)):
        tr = [q for q in net.params if q.trainable]
        rows = [(i, a, b, q) for i, (a, b, q) in enumerate(zip(net.get_grads(), om.last_grads[k], tr)) if q.kind == "W"]
        for j, (i, a, b, q) in enumerate(rows):
            rel = float(np.linalg.norm((a * scale - b).ravel()) / (np.linalg.norm(b.ravel()) + 1e-30))
            bound = 2e-2 if k == 'Dp' else (0.2 if j >= len(rows) - 7 else 0.5)
            assert rel <= bound, (k, i, q.shape, rel, bound)
    paths = [op.path for op in m.P.ops + m.Dp.ops if hasattr(op, "path")]
    assert paths.count("tcgen05") >= 16, paths


def test_single_pass_discriminator_backward_equals_two_passes_at_full_width(monkeypatch):
    """BASELINE configs[1] architecture (512x512, full width) at batch 2, fast mode: the weighted single backward pass
    through D (hm_adv_loss_pair: disc-loss weights on the weight gradients, gen-loss weights on dG(z)) against the
    reference's two separate passes (pix2pix.py:107-108,131-135) on the same weights and inputs.  Same kernels, same
    fp16 storage; only the scalar per-sample factors move, so every gradient array agrees to 5e-2 in relative L2 norm
    (fp16 rounding of differently-scaled intermediates; measured 1.5e-2 .. 2.6e-2 on the first layer's array depending on
    whether the small layers split K) and the losses to 1e-5."""
    cfg = S.experiment_kwargs('test1_nobn_bilin_both')
    Z, X, Y = S.synthetic_batch(2, cfg['latent_dim'], 512, seed=3)
    res = {}
    for sp in ("1", "0"):
        monkeypatch.setenv("HMGAN_SINGLE_PASS_D", sp)
        _, m = build_pair(cfg, 'dcgan', with_p2p=False, device="cuda", precision="fast")
        assert m._single_pass == (sp == "1")
        # a head bias of 0.6 puts D(.) near 0.6, where neither per-sample factor is negligible (at the Glorot
        # initialisation D(G(z)) ~ 1e-6 and the fake samples hardly touch D's weight gradient in either scheme)
        vals = m.D.get_all_param_values()
        vals[-1][:] = 0.6
        m.D.set_all_param_values(vals)
        lm = m.train_fn(Z, X, Y)
        res[sp] = (lm, m.D.get_grads(), m.G.get_grads(), [q for q in m.G.params if q.trainable])
        del m
        torch.cuda.empty_cache()
    np.testing.assert_allclose(res["1"][0][:2], res["0"][0][:2], rtol=1e-3, atol=1e-6)
    for k, idx in (("D", 1), ("G", 2)):
        for i, (a, b) in enumerate(zip(res["1"][idx], res["0"][idx])):
            if k == "G" and res["0"][3][i].kind == "b" and i < len(res["0"][idx]) - 1:
                continue                 # biases in front of a BatchNorm: the true gradient is zero, both are noise
            rel = float(np.linalg.norm((a - b).ravel()) / (np.linalg.norm(b.ravel()) + 1e-30))
            assert rel <= 5e-2, (k, i, a.shape, rel)


def _sync_state(src, dst):
    """dst <- src: master parameters, BatchNorm running statistics and optimiser state of every network."""
    for a, b in zip(src._nets(), dst._nets()):
        b.pflat.copy_(a.pflat)
        b.sflat.copy_(a.sflat)
        for k, v in a.opt_state.items():
            if k not in b.opt_state:
                b.opt_state[k] = v.clone()
            else:
                b.opt_state[k].copy_(v)
        b._packed = False


@pytest.mark.parametrize("case", ["gate64_fast", "gate64_parity", "tiny512_both_fast", "wide64_fast"])
def test_multi_stream_schedule_changes_nothing(case, monkeypatch):
    """The default schedule -- weight gradients on a side stream (engine.Runtime.wgrad_stream), D(x) beside G's forward
    pass (Runtime.fork), D's weight gradients under G's backward pass -- against the single-stream schedule
    (HMGAN_WGRAD_STREAM=0 HMGAN_FORK=0): same kernels, same inputs, only issued on several streams.
    Two models run side by side over five steps (eager, eager, captured, replayed, replayed); BEFORE every step the
    side-stream model receives the other one's complete state, so that every step starts from identical parameters and
    what is compared is one step's losses and gradient vectors -- a GAN step with RMSprop amplifies the run-to-run
    noise of atomically reduced gradients within a few steps.  A missing cross-stream dependency would show up as
    an O(1) relative error of some gradient array.  Bounds: float32 'parity' mode (side stream only) 1e-5 of the array
    norm, the atomics' rounding noise -- measured 0 to 1.4e-7, i.e. the two schedules are bit-identical up to the order
    of atomic adds.  fp16 'fast' mode 5e-2: with the fork D runs as two half-batch launches, whose small layers pick a
    different split-K factor than the full batch, i.e. another (deterministic) summation order; the few activations that
    round to the neighbouring fp16 value re-route max-pool ties downstream (measured 1.4e-2 on the generator's arrays,
    the same sensitivity tests/test_step_gpu.py documents for fp16 against float32)."""
    import torch
    wide = dict(in_shp=64, latent_dim=32, G=dict(nch=256, num_repeats=0, div=[2, 2, 4, 4]),
                D=dict(nch=64, num_repeats=0, bn=False, nonlinearity='linear', div=[1, 1, 1, 1]))
    cfg, mode, p2p, B, px, prec = {
        "gate64_fast": (S.experiment_kwargs('gate64'), 'dcgan', False, 4, 64, "fast"),
        "gate64_parity": (S.experiment_kwargs('gate64'), 'dcgan', False, 4, 64, "parity"),
        "tiny512_both_fast": (S.experiment_kwargs('tiny512'), 'both', True, 2, 512, "fast"),
        "wide64_fast": (wide, 'dcgan', False, 4, 64, "fast"),
    }[case]
    monkeypatch.setenv("HMGAN_WGRAD_STREAM", "0")
    monkeypatch.setenv("HMGAN_FORK", "0")
    _, m0 = build_pair(cfg, mode, with_p2p=p2p, device="cuda", precision=prec)
    monkeypatch.setenv("HMGAN_WGRAD_STREAM", "1")
    monkeypatch.setenv("HMGAN_FORK", "1")
    _, m1 = build_pair(cfg, mode, with_p2p=p2p, device="cuda", precision=prec)
    assert m1.rt.wgrad_stream() is not None and m0.rt.wgrad_stream() is None
    tol = 1e-5 if prec == "parity" else 5e-2
    for it in range(5):
        _sync_state(m0, m1)
        Z, X, Y = S.synthetic_batch(B, cfg['latent_dim'], px, seed=20 + it)
        l0, l1 = m0.train_fn(Z, X, Y), m1.train_fn(Z, X, Y)
        np.testing.assert_allclose(l1, l0, rtol=1e-5 if prec == "parity" else 2e-3, atol=1e-6)
        for n0, n1 in zip(m0._nets(), m1._nets()):
            for i, (a, b) in enumerate(zip(n0.get_grads(), n1.get_grads())):
                na = float(np.linalg.norm(a.ravel()))
                err = float(np.linalg.norm((a - b).ravel()))
                gmax = max(float(np.linalg.norm(g.ravel())) for g in n0.get_grads())
                assert err <= tol * na + 1e-6 * gmax, (case, it, n0.name, i, a.shape, err / (na + 1e-30))
    torch.cuda.synchronize()


def test_default_objective_adam_cross_entropy_l2():
    """The constructor's default objective (reference pix2pix.py:58-59: Adam, binary cross-entropy) with the L2
    reconstruction, on the joint topology at toy widths: two steps against the oracle at 1e-3.  Adam's step count lives
    in device memory (hm_adam_dev), so the step is captured in a CUDA graph like the RMSprop one: steps 3-5 replay it."""
    cfg = dict(TINY)
    cfg['D'] = dict(TINY['D'], nonlinearity='sigmoid')
    cfg['Dp'] = dict(TINY['Dp'], act='sigmoid')
    om, m = build_pair(cfg, 'both', opt="adam", lr=2e-4, lsgan=False, reconstruction='l2', device="cuda")
    assert m._graphs_ok
    for it in range(5):
        Z, X, Y = S.synthetic_batch(2, cfg['latent_dim'], 512, seed=30 + it)
        np.testing.assert_allclose(m.train_fn(Z, X, Y), om.train_fn(Z, X, Y), rtol=1e-3 if it < 2 else 5e-3, atol=1e-6)
    assert any(st.get("gA") is not None or st.get("graph") is not None for st in m._graphs.values())
    assert int(m.G.opt_state["t"].item()) == 5


def test_replayed_graphs_see_updated_weights_and_survive_reallocation(monkeypatch):
    """Captured CUDA graphs against the eager schedule (HMGAN_CUDA_GRAPHS=0) over the call sequence of Pix2Pix.train():
    three train_fn calls (the third is captured), three loss_fn calls (captured too), a train_fn replay, then a loss_fn
    replay -- which must see the weights of the LAST update (the packed compute-dtype copies are refreshed inside the
    step, not by a host flag) -- then a z_fn call with a larger batch that reallocates G's buffers, after which the
    train_fn graph must be re-captured instead of replaying into freed memory."""
    cfg = S.experiment_kwargs('gate64')
    res = {}
    for graphs in ("1", "0"):
        monkeypatch.setenv("HMGAN_CUDA_GRAPHS", graphs)
        _, m = build_pair(cfg, 'dcgan', with_p2p=False, device="cuda", precision="fast")
        out = []
        batches = [S.synthetic_batch(4, cfg['latent_dim'], 64, seed=40 + i) for i in range(10)]
        for i in range(3):
            out.append(m.train_fn(*batches[i]))
        for i in range(3, 6):
            out.append(m.loss_fn(*batches[i]))
        out.append(m.train_fn(*batches[6]))
        out.append(m.loss_fn(*batches[7]))
        Zbig = np.random.RandomState(1).rand(16, cfg['latent_dim']).astype(np.float32)
        gz = m.z_fn_det(Zbig)
        out.append(m.train_fn(*batches[8]))
        out.append(m.loss_fn(*batches[9]))
        res[graphs] = (np.array(out)[:, :2], gz)
        del m
        torch.cuda.empty_cache()
    # fp16 fast mode, atomically reduced weight gradients: the two schedules agree to run-to-run noise
    np.testing.assert_allclose(res["1"][0], res["0"][0], rtol=5e-3, atol=1e-5)
    np.testing.assert_allclose(res["1"][1], res["0"][1], atol=3e-2)      # G(z) in (0,1) after five updates at lr 1e-3


def test_async_loss_readback_equals_per_step_readback():
    """train_fn_async / loss_fn_async (no host synchronisation per step; Pix2Pix.train(loss_sync_every=N)) against
    train_fn / loss_fn from identical states: eager, captured and replayed steps, losses read back only at the end."""
    cfg = S.experiment_kwargs('gate64')
    _, m0 = build_pair(cfg, 'dcgan', with_p2p=False, device="cuda", precision="fast", lr=1e-4)
    _, m1 = build_pair(cfg, 'dcgan', with_p2p=False, device="cuda", precision="fast", lr=1e-4)
    sync, pend = [], []
    for it in range(6):
        Z, X, Y = S.synthetic_batch(4, cfg['latent_dim'], 64, seed=60 + it)
        fn0, fn1 = (m0.train_fn, m1.train_fn_async) if it % 3 != 2 else (m0.loss_fn, m1.loss_fn_async)
        sync.append(fn0(Z, X, Y))
        pend.append(fn1(Z, X, Y))
        assert isinstance(pend[-1], torch.Tensor) and pend[-1].is_cuda
    got = torch.stack(pend).cpu().numpy()
    np.testing.assert_allclose(got[:, :2], np.array(sync)[:, :2], rtol=2e-3, atol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["p2p", "both"])
def test_host_path_graphs_equal_the_eager_schedule(mode, monkeypatch):
    """The host path (train_fn from numpy batches) replays the step as several CUDA graphs around the X / Y upload:
    G's forward pass at once, D(x) -- or P(X) for a pix2pix-only model -- once X has landed, the rest once Y has.
    Against the eager schedule (HMGAN_CUDA_GRAPHS=0) in fast mode (the fork of D(x) exists there), two models side by side
    over six training steps (the last four replayed) and a loss_fn call; BEFORE every call the graph model receives the
    eager one's complete state, so that each call starts from identical parameters (left to themselves two correct
    schedules drift apart by up to 5e-3 within six steps: atomically reduced weight gradients, amplified by the
    max-pool / RMSprop sensitivity).  Both sides launch the same kernels on the same data: losses within 1e-4, every
    gradient array within 1e-3 of its norm (order of the atomic adds); a mis-ordered upload or graph shows up at O(1).
    (Batch 2: at batch 1 the BatchNorm behind G's DenseLayer sees one sample, G's activations are constants, every
    later BatchNorm has zero variance and inv_std = 1/sqrt(eps) = 100, and seven such layers overflow fp16 gradients --
    a degenerate configuration in the reference too, where G(z) no longer depends on z.)"""
    cfg = dict(TINY)
    kw = dict(device="cuda", lr=1e-4, precision="fast", with_dcgan=(mode == "both"))
    monkeypatch.setenv("HMGAN_CUDA_GRAPHS", "0")
    _, m0 = build_pair(cfg, mode, **kw)
    monkeypatch.setenv("HMGAN_CUDA_GRAPHS", "1")
    _, m1 = build_pair(cfg, mode, **kw)
    assert m1._graphs_ok and not m0._graphs_ok and m1.have_dcgan == (mode == "both")
    for it in range(7):
        _sync_state(m0, m1)
        Z, X, Y = S.synthetic_batch(2, cfg['latent_dim'], 512, seed=70 + it)
        if it < 6:
            l0, l1 = m0.train_fn(Z, X, Y), m1.train_fn(Z, X, Y)
        else:
            l0, l1 = m0.loss_fn(Z, X, Y), m1.loss_fn(Z, X, Y)
        np.testing.assert_allclose(l1, l0, rtol=1e-4, atol=1e-6, err_msg="call %d" % it)
        if it < 6:
            for n0, n1 in zip(m0._nets(), m1._nets()):
                gmax = max(float(np.linalg.norm(g.ravel())) for g in n0.get_grads())
                for i, (a_, b_) in enumerate(zip(n0.get_grads(), n1.get_grads())):
                    err = float(np.linalg.norm((a_ - b_).ravel()))
                    assert np.isfinite(a_).all() and np.isfinite(b_).all(), (mode, it, n0.name, i)
                    assert err <= 1e-3 * float(np.linalg.norm(a_.ravel())) + 1e-5 * gmax, (mode, it, n0.name, i, a_.shape, err)
    # the graphs this test is about exist: D(x) on its own stream for the joint model, P(X) for the pix2pix-only one
    slot = "gC" if mode == "both" else "gP"
    assert any(k[0] == "host" and v.get(slot) is not None for k, v in m1._graphs.items() if isinstance(k, tuple))
    torch.cuda.synchronize()


@pytest.mark.gpu
def test_unet_bottleneck_deconv_on_tensor_cores_equals_the_simt_route(monkeypatch):
    """The U-Net bottleneck at its real width (reference architectures/p2p.py:193-198: conv 2x2 'valid' 512 -> 512 on a
    2x2 map, LeakyReLU, Deconv2DLayer 2x2 stride 1 512 -> 512 on the 1x1 map), batch 16, fast mode: the deconvolution as
    1x1 tcgen05 GEMMs with N = (u, v, co) = 2048 columns (pack modes 17 / 22, unpack mode 17 with leading dimension 2048)
    against the four-SIMT-gathers route (HMGAN_DC1=0) on the same fp16 data: output, input gradient (through the conv's
    weight gradient) and both layers' gradients within 3e-3 in relative L2; and against the float32 oracle ops (2e-2)."""
    import lasagne_compat as LC
    import engine
    from oracle import lasagne_ops as LO
    res = {}
    for knob in ("1", "0"):
        monkeypatch.setenv("HMGAN_DC1", knob)
        r = np.random.RandomState(5)
        inp = LC.InputLayer((None, 512, 2, 2))
        c9 = LC.NonlinearityLayer(LC.Conv2DLayer(inp, 512, 2, stride=1, pad='valid', nonlinearity=LC.linear), LC.leaky_rectify)
        d1 = LC.TransposedConv2DLayer(c9, 512, 2, stride=1, nonlinearity=LC.linear)
        rt = engine.Runtime("cuda", "fast", loss_scale=1.0)
        net = engine.Net(rt, d1, name="neck", rng=r)
        convs = [op for op in net.ops if isinstance(op, engine.ConvOp)]
        assert convs[-1].dc1 == (knob == "1")
        x = r.randn(16, 512, 2, 2).astype(np.float32)
        net.ensure(16)
        net.inputs[0].buf.copy_(torch.from_numpy(x.transpose(0, 2, 3, 1)).half())
        vals = net.get_all_param_values()
        vals[3][:] = r.randn(*vals[3].shape).astype(np.float32) * 0.1
        net.set_all_param_values(vals)
        y = net.forward(16)
        gy = r.randn(16, 512, 2, 2).astype(np.float32)
        net.out.grad.copy_(torch.from_numpy(gy.transpose(0, 2, 3, 1)).half())
        net.backward(0, 16, wgrad=True)
        torch.cuda.synchronize()
        res[knob] = [y.float().cpu().numpy().copy()] + [a.copy() for a in net.get_grads()]
        if knob == "1":
            W1, b1, W2, b2 = [torch.tensor(v, requires_grad=True) for v in net.get_all_param_values()]
            ref = LO.deconv2d(LO.leaky_rectify(LO.conv2d(torch.tensor(x), W1, b1, 1, "valid"), 0.01), W2, b2, 1)
            ref.backward(torch.tensor(gy))
            refs = [ref.detach().permute(0, 2, 3, 1).numpy()] + [t.grad.numpy() for t in (W1, b1, W2, b2)]
            for a, b in zip(res[knob], refs):
                e = np.linalg.norm((a - b).ravel()) / (np.linalg.norm(b.ravel()) + 1e-30)
                assert e <= 2e-2, (a.shape, e)
        del net
    for a, b in zip(res["1"], res["0"]):
        e = np.linalg.norm((a - b).ravel()) / (np.linalg.norm(b.ravel()) + 1e-30)
        assert e <= 3e-3, (a.shape, e)
