"""Per-call device time of one eager training step (CUDA events around every C-ABI call, synchronised; use for
SHARES, not absolutes).  usage: python tools/op_times.py [batch] [workload]"""
import collections
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gan-heightmaps_b200"))
os.environ["HMGAN_CUDA_GRAPHS"] = "0"
import bench      # noqa: E402
import engine     # noqa: E402
import _lib       # noqa: E402
from util import synthetic_batch   # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
wl = sys.argv[2] if len(sys.argv) > 2 else "dcgan"
m = bench.build_model(wl, "cuda:0", "fast")
bench.liven_head(m, 0.6)
Z, X, Y = synthetic_batch(B, 1000, 512, seed=100)
Zd, Xd, Yd = (torch.from_numpy(t).cuda() for t in (Z, X, Y))
for _ in range(2):
    m.step_device(Zd, Xd, Yd, True)
torch.cuda.synchronize()

records = []
ctx = {"tag": ""}
orig_call = engine.Runtime.call


def timed_call(self, name, *args):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    orig_call(self, name, *args)
    e1.record()
    torch.cuda.synchronize()
    records.append((ctx["tag"], name, e0.elapsed_time(e1)))


def wrap(cls, meth, phase):
    orig = getattr(cls, meth)

    def f(self, *a, **k):
        old = ctx["tag"]
        desc = type(self).__name__
        if isinstance(self, engine.ConvOp):
            desc = "conv %dx%d %d->%d @%dx%d%s" % (self.kh, self.kw, self.Cin, self.Cout, self.out.shape[0], self.out.shape[1],
                                                   " up" if self.up else "")
        ctx["tag"] = "%s.%s %s" % (self.net.name, phase, desc)
        try:
            return orig(self, *a, **k)
        finally:
            ctx["tag"] = old
    setattr(cls, meth, f)


for cls in (engine.ConvOp, engine.BNActOp, engine.PoolOp, engine.PermuteOp):
    wrap(cls, "fwd", "fwd")
    wrap(cls, "bwd", "bwd")
    wrap(cls, "pack", "pack")
engine.Runtime.call = timed_call
m.step_device(Zd, Xd, Yd, True)
engine.Runtime.call = orig_call
tot = sum(r[2] for r in records)
print("total %.2f ms over %d calls (eager, serialised)" % (tot, len(records)))
agg = collections.OrderedDict()
for tag, name, ms in records:
    k = (tag, name)
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += ms
for (tag, name), (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print("%7.3f ms %5.1f%% n=%2d  %-46s %s" % (ms, 100 * ms / tot, n, tag, name))
byname = collections.Counter()
for tag, name, ms in records:
    byname[name] += ms
print("-- by entry point")
for name, ms in byname.most_common(14):
    print("%7.3f ms %5.1f%%  %s" % (ms, 100 * ms / tot, name))
