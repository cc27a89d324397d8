"""Generates tests/golden/*.npz from the oracle.  TEST INFRASTRUCTURE.

PARITY UNPINNED: the reference cannot run in this environment and ships no
numerical fixtures, so these vectors are produced by the oracle
itself (torch-CPU float32; cross-checked against the independent float64 numpy
restatement in tests/test_oracle.py).  They pin the oracle against regressions
and give the CUDA path a committed, box-independent target.

    python -m oracle.make_golden            # rewrites tests/golden/
"""
import os

import numpy as np

from . import step as S

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def gate64(steps=3):
    """BASELINE.json configs[0]: DCGAN-only, 64x64, batch 4, z=100, LSGAN, rmsprop lr 1e-3."""
    cfg = S.experiment_kwargs('gate64')
    m = S.OracleModel(S.build_nets(cfg, seed=2, which=('G', 'D')), opt='rmsprop', lr=1e-3, train_mode='dcgan',
                      lsgan=True)
    p0 = {k: m.get_all_param_values(k) for k in ('G', 'D')}
    losses = []
    for it in range(steps):
        Z, X, Y = S.synthetic_batch(4, cfg['latent_dim'], 64, seed=10 + it)
        losses.append(m.train_fn(Z, X, Y))
    Zt = np.random.RandomState(99).rand(4, cfg['latent_dim']).astype(np.float32)
    out = dict(losses=np.asarray(losses, np.float32), gz_det=m.z_fn_det(Zt), z_test=Zt)
    for k in ('G', 'D'):
        p1 = m.get_all_param_values(k)
        out['upd_norm_' + k] = np.asarray([np.linalg.norm((a - b).ravel()) for a, b in zip(p1, p0[k])], np.float64)
        out['sum_' + k] = np.asarray([a.astype(np.float64).sum() for a in p1], np.float64)
    # BN running statistics of the first generator BN after `steps` steps
    out['bn0_mean'] = m.get_all_param_values('G')[4]
    out['bn0_inv_std'] = m.get_all_param_values('G')[5]
    return out


def joint_tiny512(steps=1):
    """The joint 'both' step (reference experiments.py:98-119 topology: G, D, P with bilinear up-sampling, Dp; alpha=100,
    L1 + LSGAN, rmsprop) at toy widths on 2 synthetic 512x512 pairs: the five losses, a 16x16-strided sample and the
    moments of the deterministic P(X), update norms of all four parameter sets."""
    cfg = S.experiment_kwargs('tiny512')
    m = S.OracleModel(S.build_nets(cfg, seed=2), alpha=100., opt='rmsprop', lr=1e-3, train_mode='both', lsgan=True)
    keys = ('G', 'D', 'P', 'Dp')
    p0 = {k: m.get_all_param_values(k) for k in keys}
    losses = []
    for it in range(steps):
        Z, X, Y = S.synthetic_batch(2, cfg['latent_dim'], 512, seed=3 + it)
        losses.append(m.train_fn(Z, X, Y))
    _, X, _ = S.synthetic_batch(2, cfg['latent_dim'], 512, seed=3)
    px = m.gen_fn_det(X[:1])
    out = dict(losses=np.asarray(losses, np.float32), px_det_sample=px[:, :, ::16, ::16].copy(),
               px_det_moments=np.asarray([px.astype(np.float64).mean(), px.astype(np.float64).std()], np.float64))
    for k in keys:
        p1 = m.get_all_param_values(k)
        out['upd_norm_' + k] = np.asarray([np.linalg.norm((a - b).ravel()) for a, b in zip(p1, p0[k])], np.float64)
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, fn in (("gate64", gate64), ("joint_tiny512", joint_tiny512)):
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **fn())
        print("wrote", os.path.join(OUT, name + ".npz"))


if __name__ == "__main__":
    main()
