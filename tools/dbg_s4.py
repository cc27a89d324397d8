"""GPU diagnostic for the row-box kernel with super tiles of 4: which K slices contribute for each sub-tile."""
import sys, os, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tools')); sys.path.insert(0, os.path.join(ROOT, 'gan-heightmaps_b200'))
import torch, numpy as np
import tc_probe as T
import _lib
B, H, W, C1, Cout, k, pad = 2, 160, 256, 128, 64, 5, 2
torch.manual_seed(1)
Wm = torch.randn(Cout, C1, k, k, device="cuda") / np.sqrt(k * k * C1)
bias = torch.zeros(Cout, device="cuda")
wp = torch.empty(k * k * C1 * Cout, device="cuda", dtype=torch.float16); wt = torch.empty_like(wp)
_lib.call("hm_pack_conv_weight", Wm.data_ptr(), wp.data_ptr(), 0, Cout, C1, k, k, 0, 0, 1, None)
_lib.call("hm_pack_conv_weight", Wm.data_ptr(), wt.data_ptr(), 5, Cout, C1, k, k, 0, 0, 1, None)
d = T.desc(dtype=1, B=B, H=H, W=W, C1=C1, C2=0, up=0, kh=k, kw=k, stride=1, pad=pad, transposed=0, Ho=H, Wo=W, Cout=Cout,
           oH=H, oW=W, os=1, ou=0, ov=0, split=Cout, act=0, slope=0.2, accumulate=0)
full = torch.randn(B, H, W, C1, device="cuda").half()
for sl in range(8):
    x1 = torch.zeros_like(full)
    x1[..., sl * 16:(sl + 1) * 16] = full[..., sl * 16:(sl + 1) * 16]
    yr = torch.zeros(B, H, W, Cout, device="cuda", dtype=torch.float16); yt = torch.zeros_like(yr)
    _lib.call("hm_conv_gather", C.byref(d), x1.data_ptr(), None, wp.data_ptr(), bias.data_ptr(), yr.data_ptr(), None, None)
    _lib.call("hm_tc_conv", C.byref(d), x1.data_ptr(), None, wt.data_ptr(), bias.data_ptr(), yt.data_ptr(), None, None)
    torch.cuda.synchronize()
    a, b = yt.float(), yr.float()
    out = []
    for yp in range(2):
        for xh in range(2):
            aa = a[:, yp::2, xh * 128:(xh + 1) * 128].reshape(-1); bb = b[:, yp::2, xh * 128:(xh + 1) * 128].reshape(-1)
            out.append("%.3f/%.3f" % (float((aa * bb).sum() / (aa.norm() * bb.norm() + 1e-20)), float(aa.norm() / bb.norm())))
    print("channels [%3d,%3d): corr/normratio per sub-tile i=0..3: %s" % (sl * 16, sl * 16 + 16, "  ".join(out)), flush=True)
