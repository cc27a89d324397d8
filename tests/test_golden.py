"""Golden vectors (tests/golden/gate64.npz and joint_tiny512.npz, written by oracle/make_golden.py).
CPU: the oracle still reproduces them.  GPU: the CUDA path matches them (float32
parity mode, tolerance 1e-3 relative as BASELINE.json's north_star states)."""
import os
import sys

import numpy as np
import pytest

from oracle import make_golden
from oracle import step as S

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gate64.npz"))
JOINT = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "joint_tiny512.npz"))
PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gan-heightmaps_b200")
if PKG not in sys.path:
    sys.path.insert(0, PKG)


def test_oracle_reproduces_golden():
    out = make_golden.gate64()
    np.testing.assert_allclose(out['losses'], GOLD['losses'], rtol=1e-5)
    np.testing.assert_allclose(out['gz_det'], GOLD['gz_det'], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(out['bn0_inv_std'], GOLD['bn0_inv_std'], rtol=1e-4)


@pytest.mark.gpu
def test_cuda_path_matches_golden():
    from test_engine_cpu import build_pair
    cfg = S.experiment_kwargs('gate64')
    _, m = build_pair(cfg, 'dcgan', with_p2p=False, device="cuda")
    p0 = {k: n.get_all_param_values() for k, n in (('G', m.G), ('D', m.D))}
    losses = []
    for it in range(GOLD['losses'].shape[0]):
        Z, X, Y = S.synthetic_batch(4, cfg['latent_dim'], 64, seed=10 + it)
        losses.append(m.train_fn(Z, X, Y))
    np.testing.assert_allclose(np.asarray(losses)[:, :2], GOLD['losses'][:, :2], rtol=1e-3)
    np.testing.assert_allclose(m.z_fn_det(GOLD['z_test']), GOLD['gz_det'], rtol=1e-3, atol=1e-5)
    for k, n in (('G', m.G), ('D', m.D)):
        p1 = n.get_all_param_values()
        norms = np.asarray([np.linalg.norm((a - b).ravel()) for a, b in zip(p1, p0[k])])
        np.testing.assert_allclose(norms, GOLD['upd_norm_' + k], rtol=2e-2, atol=1e-6)
    np.testing.assert_allclose(m.G.get_all_param_values()[4], GOLD['bn0_mean'], rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(m.G.get_all_param_values()[5], GOLD['bn0_inv_std'], rtol=1e-3)


def test_oracle_reproduces_joint_golden():
    out = make_golden.joint_tiny512()
    np.testing.assert_allclose(out['losses'], JOINT['losses'], rtol=1e-5)
    np.testing.assert_allclose(out['px_det_sample'], JOINT['px_det_sample'], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(out['px_det_moments'], JOINT['px_det_moments'], rtol=1e-5)
    for k in ('G', 'D', 'P', 'Dp'):
        np.testing.assert_allclose(out['upd_norm_' + k], JOINT['upd_norm_' + k], rtol=1e-3, atol=1e-7)


@pytest.mark.gpu
def test_cuda_joint_step_matches_golden():
    """All four networks at 512x512 (toy widths), one 'both' step, then the deterministic P(X): same tolerances as the
    oracle comparison in tests/test_step_gpu.py."""
    from test_engine_cpu import build_pair
    cfg = S.experiment_kwargs('tiny512')
    _, m = build_pair(cfg, 'both', device="cuda")
    Z, X, Y = S.synthetic_batch(2, cfg['latent_dim'], 512, seed=3)
    np.testing.assert_allclose(m.train_fn(Z, X, Y), JOINT['losses'][0], rtol=1e-3, atol=1e-6)
    px = m.gen_fn_det(X[:1])
    np.testing.assert_allclose(px[:, :, ::16, ::16], JOINT['px_det_sample'], rtol=2e-3, atol=2e-4)
    np.testing.assert_allclose([px.mean(), px.std()], JOINT['px_det_moments'], rtol=2e-3, atol=5e-4)
