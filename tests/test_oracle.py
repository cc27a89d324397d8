"""Oracle self-checks: two independent restatements agree, and the only known
answers the reference records (g_unet.ipynb:481,558) hold."""
import numpy as np
import torch

from oracle import lasagne_ops as L
from oracle import networks as N
from oracle import numpy_ref as R
from oracle import step as S

t64 = lambda a: torch.tensor(a, dtype=torch.float64)


def test_conv_matches_first_principles():
    r = np.random.RandomState(0)
    for (k, s, pad, H) in [(5, 1, "same", 9), (3, 2, "same", 8), (2, 1, "valid", 2), (3, 1, "same", 6)]:
        x = r.randn(2, 3, H, H)
        W = r.randn(4, 3, k, k)
        b = r.randn(4)
        a = L.conv2d(t64(x), t64(W), t64(b), s, pad).numpy()
        np.testing.assert_allclose(a, R.conv2d(x, W, b, s, pad), rtol=1e-10, atol=1e-10)


def test_conv_is_true_convolution_not_correlation():
    x = np.zeros((1, 1, 5, 5)); x[0, 0, 2, 2] = 1.0
    W = np.arange(9.).reshape(1, 1, 3, 3)
    y = L.conv2d(t64(x), t64(W), t64(np.zeros(1)), 1, "same").numpy()[0, 0]
    # an impulse convolved with W reproduces W itself (not its flip)
    np.testing.assert_allclose(y[1:4, 1:4], W[0, 0])


def test_deconv_matches_first_principles():
    r = np.random.RandomState(1)
    for (k, s, H) in [(2, 2, 4), (2, 1, 1), (2, 2, 1)]:
        x = r.randn(2, 3, H, H)
        W = r.randn(3, 5, k, k)
        b = r.randn(5)
        a = L.deconv2d(t64(x), t64(W), t64(b), s).numpy()
        np.testing.assert_allclose(a, R.deconv2d(x, W, b, s), rtol=1e-10, atol=1e-10)


def test_bilinear_known_answer_and_restatements():
    x = np.array([1., 2., 4., 8.]).reshape(1, 1, 1, 4)
    y = L.bilinear_upsample(t64(np.repeat(x, 4, 2)), 2).numpy()
    np.testing.assert_allclose(y[0, 0, 0], [1, 1.5, 2, 3, 4, 6, 8, 8])
    r = np.random.RandomState(2).randn(2, 3, 5, 7)
    np.testing.assert_allclose(L.bilinear_upsample(t64(r), 2).numpy(), R.bilinear_upsample2(r),
                               rtol=1e-12, atol=1e-12)


def test_bn_pool_upscale():
    r = np.random.RandomState(3)
    x = r.randn(4, 3, 6, 6)
    beta, gamma = r.randn(3), r.randn(3)
    y, nm, ns = L.batch_norm(t64(x), t64(beta), t64(gamma), t64(np.zeros(3)), t64(np.ones(3)), False)
    yr, m, s = R.batch_norm_train(x, beta, gamma)
    np.testing.assert_allclose(y.numpy(), yr, rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(nm.numpy(), 0.1 * m, rtol=1e-10)
    np.testing.assert_allclose(ns.numpy(), 0.9 + 0.1 * s, rtol=1e-10)      # inv_std is averaged
    np.testing.assert_allclose(L.max_pool(t64(x)).numpy(), R.max_pool2(x))
    np.testing.assert_allclose(L.upscale2d(t64(x)).numpy(), R.upscale2(x))


def test_rmsprop_first_step():
    g = torch.tensor([0.5, -2.0]); p = torch.zeros(2)
    p1, acc = L.rmsprop_update(p, g, torch.zeros(2), 1e-4)
    np.testing.assert_allclose(p1.numpy(), (-1e-4 * g / torch.sqrt(0.1 * g * g + 1e-6)).numpy(), rtol=1e-6)


def _count(params):
    return sum(int(np.prod(p.shape)) for p in params)


def test_known_parameter_counts():
    rng = np.random.RandomState(0)
    p, _ = N.g_unet_init(rng, 512, True, False, nf=64, bilinear_upsample=False)
    assert _count(p) == 22882243                       # g_unet.ipynb:481
    p, _ = N.patch_discriminator_init(rng, 512, True, False, nf=32)
    assert _count(p) == 391009                         # g_unet.ipynb:558
    cfg = S.experiment_kwargs('test1_nobn_bilin_both')
    nets = S.build_nets(cfg)
    assert [len(nets[k][0]) for k in ('G', 'D', 'P', 'Dp')] == [50, 16, 104, 10]
    assert [_count(nets[k][0]) for k in ('G', 'D', 'P', 'Dp')] == [14792961, 5129217, 35088323, 1556161]


def test_gate64_shapes_and_head_nonneg():
    cfg = S.experiment_kwargs('gate64')
    m = S.OracleModel(S.build_nets(cfg), train_mode='dcgan')
    Z, X, Y = S.synthetic_batch(4, 100, 64)
    gz = m.z_fn_det(Z)
    assert gz.shape == (4, 1, 64, 64) and gz.min() > 0 and gz.max() < 1
    d, _ = N.discriminator_forward(m.params['D'], torch.tensor(X), **cfg['D'])
    assert d.shape == (4, 1) and float(d.min()) >= 0.0


def test_unet_shapes_small():
    # the 512 assert in p2p.py:137 is about the 9 stride-2 levels; 512 -> 1x1 at conv9
    rng = np.random.RandomState(0)
    p, meta = N.g_unet_init(rng, 512, True, False, nf=4, bilinear_upsample=True)
    x = torch.rand(1, 1, 512, 512)
    y, upd = N.g_unet_forward([torch.tensor(a) for a in p], x, act='tanh', bilinear_upsample=True)
    assert y.shape == (1, 3, 512, 512)
    assert len(upd) == 2 * 17
    q, _ = N.patch_discriminator_init(rng, 512, True, False, nf=4)
    d, _ = N.patch_discriminator_forward([torch.tensor(a) for a in q], x, y, act='linear')
    assert d.shape == (1, 1, 16, 16)
