"""One small training step of every kernel family for compute-sanitizer (memcheck / racecheck / initcheck):
the 64-px gate DCGAN and a 64-px DCGAN whose layers are wide enough for the tcgen05 kernels, fp16 fast mode, eager (no
CUDA graphs), two steps each.   usage: compute-sanitizer --tool memcheck python tools/sanitize_step.py"""
import os
import sys

os.environ["HMGAN_CUDA_GRAPHS"] = "0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tests'))
sys.path.insert(0, ROOT)
import numpy as np      # noqa: E402
import torch            # noqa: E402
from oracle import step as S                 # noqa: E402
from test_engine_cpu import build_pair       # noqa: E402

WIDE64 = dict(in_shp=64, latent_dim=32,
              G=dict(nch=256, num_repeats=0, div=[2, 2, 4, 4]),
              D=dict(nch=64, num_repeats=0, bn=False, nonlinearity='linear', div=[1, 1, 1, 1]))
for name, cfg in (("gate64", S.experiment_kwargs('gate64')), ("wide64", WIDE64)):
    for prec in ("fast", "tc32"):
        _, m = build_pair(cfg, 'dcgan', with_p2p=False, device="cuda", precision=prec)
        for it in range(2):
            Z, X, Y = S.synthetic_batch(4, cfg['latent_dim'], 64, seed=it)
            losses = m.train_fn(Z, X, Y)
        torch.cuda.synchronize()
        print("%s %s: losses %s, %d launches" % (name, prec, [float(v) for v in losses[:2]], m.rt.launches), flush=True)
        del m
print("sanitize_step done")
