"""GPU diagnostic: the same joint step in parity (fp32 SIMT) and fast (fp16 tcgen05) mode, compared tensor by tensor
(forward buffers, then gradients) for one network.  usage: python tools/dbg_fast_vs_parity.py [P|Dp|G|D] [batch]"""
import os
import sys

os.environ["HMGAN_CUDA_GRAPHS"] = "0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tests'))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from oracle import step as S
from test_engine_cpu import build_pair

which = sys.argv[1] if len(sys.argv) > 1 else 'P'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cfg = S.experiment_kwargs('test1_nobn_bilin_both')
_, mp = build_pair(cfg, 'both', device="cuda", precision="parity")
_, mf = build_pair(cfg, 'both', device="cuda", precision="fast")
Z, X, Y = S.synthetic_batch(B, cfg['latent_dim'], 512, seed=2)
lp = mp.train_fn(Z, X, Y)
lf = mf.train_fn(Z, X, Y)
print("losses parity", lp)
print("losses fast  ", lf)
np_, nf = getattr(mp, which), getattr(mf, which)
sc = 1.0 / mf.rt.loss_scale


def rel(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / (b.norm() + 1e-30))


prod = {}
for op in nf.ops:
    prod[op.out.vid] = op
for vp, vf in zip(np_.vals, nf.vals):
    if vf.kind not in ("buf", "input") or vf.buf is None:
        continue
    op = prod.get(vf.vid)
    tag = type(op).__name__ + ("/" + getattr(op, "path", "") if op is not None else "") if op is not None else "input"
    n = B if which in ('P', 'G') else 2 * B
    line = "val %3d %-6s %-18s %-18s fwd rel %.4f" % (vf.vid, vf.kind, tag, tuple(vf.shape), rel(vf.buf[:n], vp.buf[:n]))
    if vf.grad is not None and vp.grad is not None:
        lo = 0 if which in ('P', 'G') else B
        line += "   grad rel %.4f" % rel(vf.grad[lo:n] * sc, vp.grad[lo:n])
    print(line)
tr = [q for q in nf.params if q.trainable]
for i, (a, b, q) in enumerate(zip(nf.get_grads(), np_.get_grads(), tr)):
    r = float(np.linalg.norm((a * sc - b).ravel()) / (np.linalg.norm(b.ravel()) + 1e-30))
    print("param %3d %-5s %-20s rel %.4f" % (i, q.kind, q.shape, r))
