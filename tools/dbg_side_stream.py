"""GPU diagnostic: one training step with the weight-gradient side stream off / on from identical states, compared
buffer by buffer in backward order (first divergence = the racing op).  Also off vs off (run-to-run noise baseline).
usage: python tools/dbg_side_stream.py [wide64|gate64] [fast|parity]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tests'))
sys.path.insert(0, ROOT)
import numpy as np      # noqa: E402
import torch            # noqa: E402
from oracle import step as S                 # noqa: E402
from test_engine_cpu import build_pair       # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "wide64"
prec = sys.argv[2] if len(sys.argv) > 2 else "fast"
WIDE64 = dict(in_shp=64, latent_dim=32,
              G=dict(nch=256, num_repeats=0, div=[2, 2, 4, 4]),
              D=dict(nch=64, num_repeats=0, bn=False, nonlinearity='linear', div=[1, 1, 1, 1]))
cfg = WIDE64 if name == "wide64" else S.experiment_kwargs('gate64')


def build(side):
    os.environ["HMGAN_WGRAD_STREAM"] = side
    _, m = build_pair(cfg, 'dcgan', with_p2p=False, device="cuda", precision=prec)
    return m


def rel(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / (b.norm() + 1e-30))


Z, X, Y = S.synthetic_batch(4, cfg['latent_dim'], 64, seed=20)
for tag, sides in (("off vs off", ("0", "0")), ("off vs ON", ("0", "1"))):
    ma, mb = build(sides[0]), build(sides[1])
    la, lb = ma.train_fn(Z, X, Y), mb.train_fn(Z, X, Y)
    torch.cuda.synchronize()
    print("== %s  (%s, %s)  losses %s | %s" % (tag, name, prec, la[:2], lb[:2]))
    for key in ("D", "G"):
        na, nb = getattr(ma, key), getattr(mb, key)
        n = 8 if key == "D" else 4
        for va, vb in zip(reversed(na.vals), reversed(nb.vals)):
            if va.kind == "buf" and va.grad is not None:
                lo = 4 if key == "D" else 0
                print("  %s val %2d %-16s grad rel %.2e  buf rel %.2e" % (key, va.vid, tuple(va.shape), rel(va.grad[lo:n], vb.grad[lo:n]),
                                                                    rel(va.buf[:n], vb.buf[:n])))
        tr = [q for q in na.params if q.trainable]
        for i, (a, b, q) in enumerate(zip(na.get_grads(), nb.get_grads(), tr)):
            print("  %s param %2d %-5s %-18s grad rel %.2e" % (key, i, q.kind, q.shape, float(
                np.linalg.norm((a - b).ravel()) / (np.linalg.norm(b.ravel()) + 1e-30))))
    del ma, mb
    torch.cuda.empty_cache()
