"""GPU diagnostic: step-level error of a precision mode against the oracle, per config.
usage: python tools/dbg_tc32.py [precision=tc32] [configs: wide64 gate64 dcgan512 joint512 ...] [B=2]
Prints losses, per-network worst / median relative-L2 gradient error, updated-parameter error and the share of
convolutions on the tcgen05 path."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tests'))
sys.path.insert(0, ROOT)
import numpy as np      # noqa: E402
import torch            # noqa: E402
from oracle import step as S                 # noqa: E402
from test_engine_cpu import build_pair       # noqa: E402

args = [a for a in sys.argv[1:] if "=" not in a]
kv = dict(a.split("=") for a in sys.argv[1:] if "=" in a)
precs = kv.get("precision", "tc32").split(",")
B = int(kv.get("B", 2))
steps = int(kv.get("steps", 2))
head_bias = float(kv["bias"]) if "bias" in kv else None     # D's head bias: at the Glorot init its ReLU head is dead
WIDE64 = dict(in_shp=64, latent_dim=32,
              G=dict(nch=256, num_repeats=0, div=[2, 2, 4, 4]),
              D=dict(nch=64, num_repeats=0, bn=False, nonlinearity='linear', div=[1, 1, 1, 1]))
for name, prec in [(n, p) for n in (args or ["wide64", "gate64", "dcgan512", "joint512"]) for p in precs]:
    if name == "wide64":
        cfg, mode, p2p, size = WIDE64, 'dcgan', False, 64
    elif name == "gate64":
        cfg, mode, p2p, size = S.experiment_kwargs('gate64'), 'dcgan', False, 64
    elif name == "dcgan512":
        cfg, mode, p2p, size = S.experiment_kwargs('test1_nobn_bilin_both'), 'dcgan', False, 512
    else:
        cfg, mode, p2p, size = S.experiment_kwargs('test1_nobn_bilin_both'), 'both', True, 512
    om, m = build_pair(cfg, mode, with_p2p=p2p, device="cuda", precision=prec, lr=float(kv.get("lr", 1e-3)))
    if head_bias is not None:
        vals = m.D.get_all_param_values()
        vals[-1][:] = head_bias
        m.D.set_all_param_values(vals)
        with torch.no_grad():
            om.params['D'][-1].fill_(head_bias)
    sc = 1.0 / m.rt.loss_scale
    nets = [('G', m.G), ('D', m.D)] + ([('P', m.P), ('Dp', m.Dp)] if p2p else [])
    paths = [op.path for _, n in nets for op in n.ops if hasattr(op, "path")]
    print("== %s precision=%s B=%d: %d of %d convolutions on tcgen05" % (name, prec, B, paths.count("tcgen05"), len(paths)),
          flush=True)
    p_init = {k: [a.copy() for a in om.get_all_param_values(k)] for k, _ in nets}
    for it in range(steps):
        Z, X, Y = S.synthetic_batch(B, cfg['latent_dim'], size, seed=10 + it)
        lo, lm = om.train_fn(Z, X, Y), m.train_fn(Z, X, Y)
        print("  step %d losses oracle %s" % (it, ["%.6g" % v for v in lo]))
        print("  step %d losses ours   %s  max rel %.2e" % (it, ["%.6g" % v for v in lm], max(
            abs(a - b) / (abs(b) + 1e-12) for a, b in zip(lm, lo) if b != 0)), flush=True)
        if it == 0:
            for k, net in nets:
                tr = [q for q in net.params if q.trainable]
                rels = []
                for a, b, q in zip(net.get_grads(), om.last_grads[k], tr):
                    nb = np.linalg.norm(b.ravel())
                    rels.append((float(np.linalg.norm((a * sc - b).ravel()) / (nb + 1e-30)), q.kind, q.shape, float(nb)))
                w = [r for r in rels if r[1] == "W"]
                print("  grads %-2s W arrays: worst %.2e median %.2e | all: %s" % (
                    k, max(r[0] for r in w), float(np.median([r[0] for r in w])),
                    " ".join("%.1e" % r[0] for r in rels)), flush=True)
    for k, net in nets:
        cs = []
        for a, b, a0, q in zip(net.get_all_param_values(), om.get_all_param_values(k), p_init[k], net.params):
            if q.trainable and q.kind == "W":
                ua, ub = (a - a0).ravel().astype(np.float64), (b - a0).ravel().astype(np.float64)
                cs.append(float(ua @ ub / (np.linalg.norm(ua) * np.linalg.norm(ub) + 1e-30)))
        print("  params %-2s after %d steps: cosine of the accumulated update per W array: min %.3f | %s" % (
            k, steps, min(cs), " ".join("%.3f" % c for c in cs)))
    if p2p:
        Zx = X[:1]
        a, b = m.gen_fn_det(Zx), om.gen_fn_det(Zx)
        print("  gen_fn_det max abs err %.2e (range %.2f)" % (float(np.abs(a - b).max()), float(np.abs(b).max())))
    Zz = np.random.RandomState(5).rand(2, cfg['latent_dim']).astype(np.float32)
    a, b = m.z_fn_det(Zz), om.z_fn_det(Zz)
    print("  z_fn_det max abs err %.2e" % float(np.abs(a - b).max()), flush=True)
    del m, om
    torch.cuda.empty_cache()
