"""Lowering of Lasagne-style layer graphs to programs of sm_100a kernel launches.

The reference hands its layer graphs to Theano (``lasagne.layers.get_output`` +
``theano.function``, reference pix2pix.py:91-101,142-147), which compiles them to
CorrMM / cuDNN calls.  Here ``Net`` plays that role: it walks a
``lasagne_compat`` graph once, emits a list of ops over NHWC device buffers, and
runs them forwards and backwards by calling the C-ABI kernel library
(include/hmgan.h) through ``_lib.call``.  PyTorch only owns the device memory and
the stream; every arithmetic operation on the path is one of our kernels.

Graph canonicalisation (all exact, elementwise ops commute with concatenation):
  * NonlinearityLayer(ConcatLayer([a, b]))  ->  concat(act(a), act(b)); the U-Net
    skip tensor act(BN(conv_k)) is then shared with the encoder's own activation
    (reference architectures/p2p.py:146-147 and :213-214 use the same slope);
  * BatchNormLayer + NonlinearityLayer      ->  one normalise+activate pass;
  * Conv/Dense/Deconv + NonlinearityLayer   ->  activation in the conv epilogue;
  * Upscale2DLayer / BilinearUpsample2DLayer / ConcatLayer are never materialised
    in parity mode: the consumer convolution gathers through them;
  * Pool2D(average, full extent) + Reshape(-1,1) + Nonlinearity (the DCGAN
    discriminator head, reference architectures/dcgan.py:50-56) is evaluated
    inside the adversarial-loss kernel.
"""
import ctypes as C
import os

import numpy as np
import torch

import _lib
import lasagne_compat as L
from architectures.layers import BilinearUpsample2DLayer

ACT = _lib.ACT


def _ptr(t):
    return None if t is None else t.data_ptr()


class Runtime(object):
    """Device, compute dtype and stream shared by the networks of one model."""

    def __init__(self, device="cuda", precision="parity", loss_scale=None):
        self.device = torch.device(device)
        if precision not in ("parity", "fast", "tc32"):
            raise ValueError("precision must be 'parity' (fp32, SIMT kernels), 'tc32' (fp32 storage, the tcgen05 kernels on "
                             "three-plane bf16 operand splits: float32-grade) or 'fast' (fp16 storage, fp32 accumulate)")
        self.precision = precision
        self.cd = _lib.F16 if precision == "fast" else _lib.F32
        self.tdtype = torch.float16 if precision == "fast" else torch.float32
        # tc: GEMM-shaped convolutions run on the tcgen05 kernels; split: their operands are three-plane bf16 splits of the
        # fp32 tensors (HM_BF16X3, csrc/split_bf16.cu) and their results fp32 -- the mode the 1e-3 parity gate runs in
        self.tc = precision in ("fast", "tc32")
        self.split = precision == "tc32"
        self.tc_dtype = torch.bfloat16 if self.split else self.tdtype
        self._scratch = {}
        if loss_scale is None:
            loss_scale = 1024.0 if precision == "fast" else 1.0
        self.loss_scale = float(loss_scale)
        self.launches = 0
        # SyncBN: a torch.distributed process group over which every BatchNorm layer averages its batch
        # statistics (forward) and its two backward reductions, so that N ranks with B samples each compute exactly the
        # BatchNorm of one rank with N*B samples.  None (default) = per-rank statistics, the DDP convention.
        self.sync_bn_group = None
        self._pack_jobs = None
        self._pack_tables = {}
        # lanes: independent parts of the step run on their own streams (Runtime.fork): "main", "aux" (D(x) beside G's
        # forward pass), "p2p" (the pix2pix half of a joint step beside the DCGAN half).  Per-lane state -- the
        # weight-gradient side stream, scratch workspaces -- is keyed by the lane NAME, which is the same during the
        # eager warm-up calls and under CUDA-graph capture.
        self.lane = "main"
        self._lane_streams = {}
        self._wgrad_streams = {}
        self._wgrad_forked = {}
        self._wgrad_side = (self.device.type == "cuda" and not self.split
                            and os.environ.get("HMGAN_WGRAD_STREAM", "1") != "0")
        self._splitk = self.device.type == "cuda" and os.environ.get("HMGAN_TC_SPLITK", "1") != "0"
        # the weighted max-pool gradient copies and D1's weight gradient on the weight-gradient stream (A/B knob)
        # (Measured and left out, profiles/r2_variants_ab.txt visits r2q / r2r: running the per-sample-weighted max-pool copy
        # and D layer 1's weight gradient on the weight-gradient stream as well -- 12.23 / 12.28 ms against 12.06 / 12.08 ms
        # per DCGAN step: the deferred copy re-reads the pooled gradient and argmax, and that stream is already the longer
        # one while D's backward pass runs.)
        self._tc_ws = {}
        self._fork_ok = (self.device.type == "cuda" and precision == "fast"
                         and os.environ.get("HMGAN_FORK", "1") != "0")
        _lib.load()

    class _Fork(object):
        def __init__(self, rt, lane):
            self.rt, self.lane = rt, lane

        def __enter__(self):
            rt = self.rt
            st = rt._lane_streams.get(self.lane)
            if st is None:
                st = rt._lane_streams[self.lane] = torch.cuda.Stream(rt.device)
            st.wait_stream(torch.cuda.current_stream(rt.device))      # everything issued so far is visible
            self.ctx = torch.cuda.stream(st)
            self.ctx.__enter__()
            self.prev, rt.lane = rt.lane, self.lane
            return self

        def __exit__(self, *exc):
            self.rt.lane = self.prev
            return self.ctx.__exit__(*exc)

    def fork(self, lane="aux"):
        """`with rt.fork(lane):` issues the enclosed launches on that lane's stream, ordered after everything already
        issued on the current stream; join(lane) makes the current stream wait for them.  Works eagerly and under
        CUDA-graph capture (cross-stream edges of the captured graph).  Only independent work may go there."""
        return Runtime._Fork(self, lane)

    def join(self, lane="aux"):
        st = self._lane_streams.get(lane)
        if st is not None:
            torch.cuda.current_stream(self.device).wait_stream(st)

    def allreduce_mean(self, t):
        """Average a small device tensor over the SyncBN group (sum, then 1/world: gloo has no AVG)."""
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.sync_bn_group)
        t.div_(dist.get_world_size(self.sync_bn_group))

    @property
    def stream(self):
        if self.device.type == "cuda":
            return torch.cuda.current_stream(self.device).cuda_stream
        return None

    def wgrad_stream(self):
        """Side stream for the weight-gradient half of a convolution's backward pass, or None.

        On by default on CUDA (HMGAN_WGRAD_STREAM=0 switches it off; measured on B200: DCGAN step -1.2 ms, results
        identical to the single-stream schedule, tests/test_step_gpu.py): dW and dX of a layer both need only dY, and
        the small layers' kernels fill a fraction of the 148 SMs, so issuing dW on a second stream lets it run beside
        the input-gradient chain of the layers below.  Net.backward joins the stream before it returns, so the fork is
        invisible to callers and is captured into the step's CUDA graph as ordinary cross-stream dependencies."""
        if not self._wgrad_side:
            return None
        st = self._wgrad_streams.get(self.lane)
        if st is None:
            st = self._wgrad_streams[self.lane] = torch.cuda.Stream(self.device)
        return st

    def mark_wgrad_forked(self):
        self._wgrad_forked[self.lane] = True

    def side_if_forked(self):
        """This lane's weight-gradient stream if work is pending on it (un-joined), else None."""
        return self._wgrad_streams.get(self.lane) if self._wgrad_forked.get(self.lane) else None

    def join_wgrad_stream(self):
        if self._wgrad_forked.get(self.lane):
            torch.cuda.current_stream(self.device).wait_stream(self._wgrad_streams[self.lane])
            self._wgrad_forked[self.lane] = False

    def call(self, name, *args):
        if self._pack_jobs is not None and name == "hm_pack_conv_weight":
            self._pack_jobs.append(args)             # deferred: one hm_pack_conv_weight_multi launch per network
            return
        self.launches += 1
        if name == "hm_tc_conv" and self._splitk:
            # split-K for the layers that fill only a few SMs (include/hmgan.h): one fp32 scratch workspace per LANE
            # (= stream role, see fork(): calls on a stream are serialised); sized during the eager warm-up calls, never
            # while a graph is captured
            need = _lib.query("hm_tc_conv_ws_bytes", args[0])
            if need > 0:
                key = self.lane
                ws = self._tc_ws.get(key)
                if ws is None or ws.numel() * 4 < need:
                    ws = self._tc_ws[key] = self.empty(((need + 3) // 4,), torch.float32)
                self.launches += 1               # the finishing pass
                _lib.call("hm_tc_conv_ws", *args, ws.data_ptr(), ws.numel() * 4, self.stream)
                return
        _lib.call(name, *args, self.stream)

    # -- tensor-core entry points (fast: fp16 tensors as they are; tc32: three-plane bf16 splits of the fp32 tensors) -------- #
    def scratch(self, key, nbytes):
        """Reusable device scratch (launches on one stream are ordered, so one buffer per role serves every layer).
        Grows only during the eager warm-up steps, never while a CUDA graph is captured."""
        t = self._scratch.get(key)
        if t is None or t.numel() < nbytes:
            t = self._scratch[key] = torch.empty((int(nbytes),), dtype=torch.uint8, device=self.device)
        return t.data_ptr()

    def tc_desc(self, d, wgrad=False):
        """The descriptor the tcgen05 entry points see: unchanged in fast mode; in tc32 mode the sources carry three
        bf16 planes per fp32 channel, six along the reduction axis (forward / input gradient: the channels; weight
        gradient: the batch)."""
        if not self.split:
            return d
        d2 = _lib.ConvDesc.from_buffer_copy(d)
        d2.dtype = _lib.BF16X3
        if wgrad:
            d2.B = 6 * d.B
        else:
            d2.C1, d2.C2 = 6 * d.C1, 6 * d.C2
        return d2

    def tc_supported(self, d, wgrad=False):
        return bool(_lib.query("hm_tc_wgrad_supported" if wgrad else "hm_tc_conv_supported",
                               C.byref(self.tc_desc(d, wgrad))))

    def tc_conv(self, d, x1, x2, w, bias, y, y2):
        """hm_tc_conv on the layer's tensors (device pointers)."""
        if self.split:
            rows = d.B * d.H * d.W
            s1 = self.scratch("x1", rows * 6 * d.C1 * 2)
            self.call("hm_split_bf16x3", x1, s1, rows, d.C1, d.C1, 0)
            x1 = s1
            if d.C2:
                s2 = self.scratch("x2", rows * 6 * d.C2 * 2)
                self.call("hm_split_bf16x3", x2, s2, rows, d.C2, d.C2, 0)
                x2 = s2
        self.call("hm_tc_conv", C.byref(self.tc_desc(d)), x1, x2, w, bias, y, y2)

    def tc_wgrad(self, d, x1, x2, dy, dwp):
        """hm_tc_wgrad on the layer's tensors (device pointers)."""
        if self.split:
            rows = d.B * d.H * d.W
            s1 = self.scratch("x1", rows * 6 * d.C1 * 2)
            self.call("hm_split_bf16x3", x1, s1, rows, d.C1, d.C1, 2)
            x1 = s1
            if d.C2:
                s2 = self.scratch("x2", rows * 6 * d.C2 * 2)
                self.call("hm_split_bf16x3", x2, s2, rows, d.C2, d.C2, 2)
                x2 = s2
            orow = d.B * d.Ho * d.Wo
            sd = self.scratch("dy", orow * 6 * d.Cout * 2)
            self.call("hm_split_bf16x3", dy, sd, orow, d.Cout, d.Cout, 3)
            dy = sd
        self.call("hm_tc_wgrad", C.byref(self.tc_desc(d, True)), x1, x2, dy, dwp)

    def tc_pack(self, w, dst, mode, cout, cin, kh, kw, K, c1=None):
        """K-major tensor-core weight pack (hm_pack_conv_weight modes 5/6/8/12/...; K = innermost extent of the pack).
        tc32: packed in fp32 first, then split along K into the six b-side planes (per ConcatLayer segment when c1 < K)."""
        if not self.split:
            self.call("hm_pack_conv_weight", w, dst, mode, cout, cin, kh, kw, 0, 0, self.cd)
            return
        n = _lib.pack_count(mode, cout, cin, kh, kw)
        tmp = self.scratch("pack", n * 4)
        self.call("hm_pack_conv_weight", w, tmp, mode, cout, cin, kh, kw, 0, 0, _lib.F32)
        self.call("hm_split_bf16x3", tmp, dst, n // K, K, K if c1 is None else c1, 1)

    def begin_pack_batch(self):
        if self.split:              # the split reads the fp32 pack from a shared scratch buffer: keep program order
            return
        self._pack_jobs = []

    def end_pack_batch(self):
        """Issue the deferred weight packs as ONE launch (the job table lives in device memory and is cached: the
        pointers of a network's packs never change, so CUDA-graph capture sees no allocation or copy)."""
        jobs, self._pack_jobs = self._pack_jobs, None
        if not jobs:
            return
        # measured on B200: no gain over the ~35 separate launches inside the captured graph (14.6 vs 14.5 ms/step),
        # so batching is opt-in
        if len(jobs) < 3 or os.environ.get("HMGAN_BATCH_PACKS", "0") != "1":
            for a in jobs:
                self.call("hm_pack_conv_weight", *a)
            return
        key = tuple(jobs)
        ent = self._pack_tables.get(key)
        if ent is None:
            raw, max_n = _lib.pack_job_table(jobs)
            tab = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(self.device)
            ent = self._pack_tables[key] = (tab, len(jobs), max_n)
        tab, n_jobs, max_n = ent
        self.call("hm_pack_conv_weight_multi", tab.data_ptr(), n_jobs, max_n, jobs[0][9])

    def empty(self, shape, dtype=None):
        return torch.empty(shape, dtype=dtype or self.tdtype, device=self.device)

    def zeros(self, shape, dtype=None):
        return torch.zeros(shape, dtype=dtype or self.tdtype, device=self.device)


class Val(object):
    """A tensor of the lowered program.  kind: 'input' | 'buf' | 'up' | 'cat'.
    shape is per-sample (H, W, C)."""

    def __init__(self, vid, kind, shape, srcs=(), up=0):
        self.vid, self.kind, self.shape, self.srcs, self.up = vid, kind, tuple(shape), tuple(srcs), up
        self.buf = None
        self.grad = None
        self.grad_w = None               # per-sample-weighted copy of grad (single-pass discriminator backward)
        self.consumers = 0
        self.grad_is_preact = False      # consumer hands back d/d(pre-activation)
        self.want_grad = kind == "buf"
        self.gw = False                  # gradient written during the current backward pass

    def b(self, lo, hi):
        return self.buf[lo:hi]

    def g(self, lo, hi):
        return self.grad[lo:hi]

    def take_acc(self):
        acc = 1 if self.gw else 0
        self.gw = True
        return acc


def _resolve_src(v):
    """conv source -> (x1, x2, up_mode) with x1/x2 physical values."""
    up = 0
    if v.kind == "up":
        up = v.up
        v = v.srcs[0]
    if v.kind == "cat":
        a, b = v.srcs
        if a.kind not in ("buf", "input") or b.kind not in ("buf", "input"):
            raise NotImplementedError("nested virtual tensors")
        return a, b, up
    if v.kind not in ("buf", "input"):
        raise NotImplementedError("nested virtual tensors")
    return v, None, up


# --------------------------------------------------------------------------- #
# ops
# --------------------------------------------------------------------------- #
class ConvOp(object):
    """Conv2DLayer / DenseLayer / TransposedConv2DLayer with its bias and epilogue
    activation.  kind in {'conv', 'dense', 'deconv'}."""

    def __init__(self, net, kind, layer, src, out, act):
        self.net, self.kind, self.layer, self.out, self.act = net, kind, layer, out, act
        self.x1, self.x2, self.up = _resolve_src(src)
        self.src = src
        self.W, self.bias = layer.params
        self.C1 = self.x1.shape[2]
        self.C2 = self.x2.shape[2] if self.x2 is not None else 0
        self.Cin = self.C1 + self.C2
        self.Cout = out.shape[2]
        if kind == "dense":
            self.kh = self.kw = 1
            self.stride, self.pad = 1, 0
        else:
            self.kh, self.kw = layer.filter_size
            if layer.stride[0] != layer.stride[1]:
                raise NotImplementedError("anisotropic stride")
            self.stride = layer.stride[0]
            self.pad = layer.pad[0] if kind == "conv" else 0
            if kind == "conv" and layer.pad[0] != layer.pad[1]:
                raise NotImplementedError("anisotropic padding")
        sh = 1 if self.up else 0
        self.Hv, self.Wv = self.x1.shape[0] << sh, self.x1.shape[1] << sh     # virtual source grid
        if kind == "deconv":
            if self.up:
                raise NotImplementedError("upsampling into a transposed convolution")
            if not (self.stride == self.kh == self.kw or (self.Hv == 1 and self.Wv == 1)):
                raise NotImplementedError("overlapping transposed convolution (stride < kernel on a >1x1 input)")
        self.K = self.kh * self.kw * self.Cin
        self.wp_f = self.wp_d = self.dwp = self.gup = None
        # tensor-core (tcgen05/TMA) eligibility, fast mode only; everything else runs the SIMT gather kernels
        self.tc_fwd = self.tc_dg = self.tc_wg = False
        self.wt_f = self.wt_d = self.x1u = self.x2u = None
        rt = net.rt
        self.up2 = False      # nearest-2x + 5x5 evaluated as four 3x3 phase convolutions on the low-res source
        self.dg2_cat = False
        self.gcat = None
        self.dg2 = False      # input gradient of a 3x3 stride-2 conv as a 2x2-tap phase convolution of dy (pack mode 12)
        self.dg6 = False      # input gradient of nearest-2x + 5x5 as a 6x6 stride-2 convolution of dy (pack mode 20)
        self.wg8 = False      # weight gradient of nearest-2x + 5x5 as the gradient of its four 3x3 phase filters
        if rt.tc and kind == "conv" and self.stride in (1, 2):
            if self.up == _lib.UP_NEAREST2 and self.x2 is None and self.stride == 1:
                # (thin outputs take the fp16-only hm_s2d_pad64 route for their weight gradient: fast mode only)
                self.up2 = (not rt.split or self.Cout > 4) and rt.tc_supported(self._fwd_desc(rt, 1))
            # ... and its input gradient as ONE 6x6 stride-2 convolution of dy that lands on the low-res source grid
            # (pack mode 20): 36 taps on H x W pixels instead of 25 taps on 2H x 2W plus the adjoint of the upsampling
            self.dg6 = (self.up2 and self.kh == 5 and os.environ.get("HMGAN_DG6", "1") != "0"
                        and rt.tc_supported(self._dg6_desc(rt, 1, 0)))
            self.tc_fwd = self.up2 or rt.tc_supported(self._tc_fwd_desc(rt, 1))
            self.tc_wg = rt.tc_supported(self._tc_fwd_desc(rt, 1), wgrad=True)
            # ... and its weight gradient in phase form: gradient of the four 3x3 phase filters on the low-res source
            # against the strided phases of dy (hm_tc_wgrad with the forward descriptor, unpack mode 8): no
            # materialised 2x copy of the source, 36 instead of 100 low-res taps
            self.wg8 = (self.up2 and self.Cout % 64 == 0 and os.environ.get("HMGAN_WG8", "1") != "0"
                        and rt.tc_supported(self._fwd_desc(rt, 1), wgrad=True))
            if self.stride == 1:
                self.tc_dg = rt.tc_supported(self._tc_dgrad_desc(rt, 1, 0))
            elif not self.up:
                dd = self._dgrad_desc(rt, 1, 0)
                dd.split = dd.Cout
                ok = rt.tc_supported(dd)
                if self.x2 is None:
                    self.tc_dg = self.dg2 = ok
                else:
                    self.dg2_cat = ok        # gradient of the whole (thin) concat on the tensor cores, then sliced
        # DenseLayer (the generator's first layer, z[B,1000] -> 8192) as a 1x1 tensor-core convolution over [B,1,1,in]
        # (pack mode 21; the SIMT gather spent 0.17 ms on 0.5 GFLOP there); its weight gradient stays on the SIMT kernel
        self.dense_tc = False
        if rt.precision == "fast" and kind == "dense" and self.x2 is None and self.Cin % 8 == 0:
            self.dense_tc = rt.tc_supported(self._dense_desc(rt, 1))
            self.tc_fwd = self.dense_tc
        # one-channel input (first discriminator layer): im2col to 64 "tap channels", then a 1x1 tensor-core GEMM
        self.col1 = (rt.precision == "fast" and kind == "conv" and self.Cin == 1 and self.x2 is None and not self.up
                     and self.stride == 1 and self.kh * self.kw <= 64 and 2 * self.pad == self.kh - 1
                     and self.kh == self.kw and self.Cout % 64 == 0 and self.Cout <= 256)
        if self.col1:
            self.tc_fwd = True
        # other thin-source convolutions (<= 4 input channels, e.g. the PatchGAN's 4-channel concat and the U-Net's first
        # layer, 3x3 stride 2): hm_im2col_thin to 64 "tap-channels", then 1x1 tensor-core GEMMs (forward, weight gradient)
        self.colk = (rt.precision == "fast" and kind == "conv" and not self.col1 and self.Cin <= 4 and not self.up
                     and self.stride in (1, 2) and self.kh * self.kw * self.Cin <= 64 and self.Cout % 64 == 0
                     and self.Cout <= 256)
        if self.colk:
            self.tc_fwd = True
        # Deconv2DLayer 2x2 stride 2 (no overlap) as ONE tensor-core launch: 1x1 convolution with N = (phase, co) and a
        # depth-to-space epilogue; its gradients through hm_s2d_pad64 (dy regrouped to (phase, co) channels)
        self.dc2 = False
        if (rt.precision == "fast" and kind == "deconv" and self.kh == 2 and self.kw == 2 and self.stride == 2
                and not self.up and 4 * self.Cout <= 64 and self.C1 % 64 == 0 and self.C2 % 64 == 0):
            self.dc2 = bool(_lib.query("hm_tc_conv_supported", C.byref(self._dc2_desc(rt, 1))))
            self.tc_fwd = self.dc2
        # Deconv2DLayer 2x2 on a 1x1 input (the U-Net's bottleneck, p2p.py:197-198): every output position sees the one input
        # pixel, so the layer IS a dense layer in -> (u, v, co) and runs as 1x1 tensor-core GEMMs (forward with pack mode
        # 17 and a bias tiled over the four positions, input gradient with pack mode 22, weight gradient + unpack mode 17)
        self.dc1 = False
        if (rt.precision == "fast" and kind == "deconv" and not self.dc2 and self.kh == 2 and self.kw == 2
                and self.Hv == 1 and self.Wv == 1 and self.x2 is None and self.C1 % 64 == 0
                and (4 * self.Cout) % 256 == 0 and os.environ.get("HMGAN_DC1", "1") != "0"):
            self.dc1 = (rt.tc_supported(self._dc1_desc(rt, 1, True)) and rt.tc_supported(self._dc1_desc(rt, 1, False))
                        and rt.tc_supported(self._dc1_desc(rt, 1, True), wgrad=True))
            self.tc_fwd = self.dc1
        self.path = "tcgen05" if self.tc_fwd else "simt"
        # hm_c1s2_conv (in-kernel im2col of a one-channel image): (a) this layer + its 2x2 max-pool in one pass, set by
        # Net when a PoolOp consumes the output (pool_fused); (b) the input gradient of nearest-2x -> 5x5 -> one channel
        self.pool_fused = None
        self.pool_tc = None           # PoolOp whose 2x2 max-pool runs in this convolution's tensor-core epilogue
        self.db_done = False
        self.bias_grad_zero = False
        self.wk = None
        self.c1dg = (rt.precision == "fast" and kind == "conv" and self.up == _lib.UP_NEAREST2 and self.x2 is None
                     and self.Cout == 1 and self.Cin == 64 and self.kh == 5 and self.kw == 5 and self.pad == 2
                     and self.stride == 1)
        # input gradient of a thin-output stride-1 convolution (dy has <= 4 channels): forward-form gather of dy,
        # which the library serves with its thin-input kernel
        self.fw_dg = kind == "conv" and self.stride == 1 and self.Cout <= 4 and not self.tc_dg

    # -- buffers ------------------------------------------------------------ #
    def alloc(self, rt, B):
        n = self.K * self.Cout
        if self.wp_f is None:
            self.wp_f = rt.empty((n,))
            self.wp_d = rt.empty((n,)) if self.kind != "dense" else None
            self.dwp = rt.empty((n,), torch.float32)
        if self.up and self.src.srcs[0].kind != "input" and not self.c1dg and not self.dg6:
            self.gup = rt.empty((B, self.Hv, self.Wv, self.Cin))
        if self.dc2:
            self.dy64 = rt.empty((B, self.x1.shape[0], self.x1.shape[1], 64))
            if self.wt_f is None:
                self.wt_f = rt.empty((4 * self.Cout * self.Cin,))
                self.wt_d = rt.empty((64 * self.Cin,))
                self.dwp = rt.empty((64 * self.Cin,), torch.float32)
        if self.dc1 and self.wt_f is None:
            self.wt_f = rt.empty((4 * self.Cout * self.Cin,))
            self.wt_d = rt.empty((4 * self.Cout * self.Cin,))
            self.bias4 = rt.empty((4, self.Cout), torch.float32)
        if self.c1dg and self.wk is None:
            self.wk = rt.empty((64 * 64,))
        if self.pool_fused is not None:
            if self.wk is None:
                self.wk = rt.empty((256 * 64,))
                self.wk2 = rt.empty((256 * 64,))
                self.dwk = rt.zeros((256 * 64,), torch.float32)
            self.ubuf = rt.empty((B, self.Hv // 2, self.Wv // 2, 64))      # patch-space input gradient
            return          # no im2col tensor, no packed-gradient buffers: forward and backward are hm_c1s2_* calls
        n = self.K * self.Cout
        tcm = 6 if rt.split else 1        # tc32: six bf16 planes (h m h l h m) per packed fp32 weight
        if self.col1 or self.colk:
            self.xc = rt.empty((B, self.out.shape[0], self.out.shape[1], 64))
            if self.wt_f is None:
                self.wt_f = rt.empty((64 * self.Cout,))
                self.dwp = rt.empty((64 * self.Cout,), torch.float32)
        if self.tc_fwd and self.wt_f is None:
            self.wt_f = rt.empty((tcm * (36 * self.Cin * self.Cout if self.up2 else n),), rt.tc_dtype)
        if self.dg2_cat:
            self.gcat = rt.empty((B, self.Hv, self.Wv, self.Cin))
            if self.wt_d is None:
                self.wt_d = rt.empty((tcm * 16 * self.Cin * self.Cout,), rt.tc_dtype)
        if self.dg6 and self.wt_d is None:
            self.wt_d = rt.empty((tcm * 36 * self.Cin * self.Cout,), rt.tc_dtype)
        elif self.tc_dg and self.wt_d is None:
            self.wt_d = rt.empty((tcm * (16 * self.Cin * self.Cout if self.dg2 else n),), rt.tc_dtype)
        self.thin_up2_wg = self.up2 and self.Cout <= 4           # weight gradient of the thin phase-decomposed layer
        # ... which for the generator's last layer (64 -> 1) is the c1s2 gather with the roles swapped: 6x6 stride-2
        # patches of the one-channel dy against the low-res source rows, nothing materialised (hm_c1s2_wgrad)
        self.c1wg = self.thin_up2_wg and self.c1dg and os.environ.get("HMGAN_C1WG", "1") != "0"
        if self.c1wg:
            if self.dwp is None or self.dwp.numel() < 64 * 64:
                self.dwp = rt.empty((64 * 64,), torch.float32)
        elif self.thin_up2_wg:
            # tensor cores: s2d(dy) zero-padded to 64 channels against the low-res source (3x3 taps)
            self.dy64 = rt.empty((B, self.x1.shape[0], self.x1.shape[1], 64))
            if self.dwp is None or self.dwp.numel() < 9 * self.Cin * 64:
                self.dwp = rt.empty((9 * self.Cin * 64,), torch.float32)
        if self.wg8 and (self.dwp is None or self.dwp.numel() < 36 * self.Cin * self.Cout):
            self.dwp = rt.empty((36 * self.Cin * self.Cout,), torch.float32)
        if self.up and ((self.tc_fwd and not self.up2) or (self.tc_wg and not self.thin_up2_wg and not self.wg8)):
            # the tensor-core kernels read dense NHWC tiles through TMA: materialise the 2x resampling once
            self.x1u = rt.empty((B, self.Hv, self.Wv, self.C1))
            self.x2u = rt.empty((B, self.Hv, self.Wv, self.C2)) if self.x2 is not None else None

    def pack(self, rt):
        """master (Lasagne layout, fp32) -> packed [K][Cout] copies in the compute dtype."""
        w = self.net.pview(self.W)
        if self.kind == "dense" and self.dense_tc:
            rt.call("hm_pack_conv_weight", _ptr(w), _ptr(self.wt_f), 21, self.Cout, self.Cin, 1, 1, 0, 0, rt.cd)
        elif self.kind == "dense":
            rt.call("hm_pack_conv_weight", _ptr(w), _ptr(self.wp_f), 4, self.Cout, self.Cin, 1, 1, 0, 0, rt.cd)
        elif self.kind == "conv":
            if self.pool_fused is not None:
                rt.call("hm_pack_conv_weight", _ptr(w), _ptr(self.wk), 15, self.Cout, 1, 5, 5, 0, 0, rt.cd)
                rt.call("hm_pack_conv_weight", _ptr(w), _ptr(self.wk2), 16, self.Cout, 1, 5, 5, 0, 0, rt.cd)
                return
            if self.col1:
                rt.call("hm_pack_conv_weight", _ptr(w), _ptr(self.wt_f), 11, self.Cout, 1, self.kh, self.kw, 0, 0, rt.cd)
            elif self.colk:
                rt.call("hm_pack_conv_weight", _ptr(w), _ptr(self.wt_f), 19, self.Cout, self.Cin, self.kh, self.kw, 0, 0,
                        rt.cd)
            elif not self.tc_fwd:
                rt.call("hm_pack_conv_weight", _ptr(w), _ptr(self.wp_f), 0, self.Cout, self.Cin, self.kh, self.kw,
                        0, 0, rt.cd)
            else:
                rt.tc_pack(_ptr(w), _ptr(self.wt_f), 8 if self.up2 else 5, self.Cout, self.Cin, self.kh, self.kw,
                           self.Cin, self.C1)
            if self.c1dg:
                rt.call("hm_pack_conv_weight", _ptr(w), _ptr(self.wk), 14, 1, self.Cin, 5, 5, 0, 0, rt.cd)
            if self.dg6:
                rt.tc_pack(_ptr(w), _ptr(self.wt_d), 20, self.Cout, self.Cin, 5, 5, self.Cout)
            elif self.dg2_cat:
                rt.tc_pack(_ptr(w), _ptr(self.wt_d), 12, self.Cout, self.Cin, self.kh, self.kw, self.Cout)
            elif not self.tc_dg:
                rt.call("hm_pack_conv_weight", _ptr(w), _ptr(self.wp_d), 7 if self.fw_dg else 1, self.Cout, self.Cin,
                        self.kh, self.kw, 0, 0, rt.cd)
            else:
                rt.tc_pack(_ptr(w), _ptr(self.wt_d), 12 if self.dg2 else 6, self.Cout, self.Cin, self.kh, self.kw,
                           self.Cout)
        elif self.dc2:
            rt.call("hm_pack_conv_weight", _ptr(w), _ptr(self.wt_f), 17, self.Cout, self.Cin, 2, 2, 0, 0, rt.cd)
            rt.call("hm_pack_conv_weight", _ptr(w), _ptr(self.wt_d), 18, self.Cout, self.Cin, 2, 2, 0, 0, rt.cd)
        elif self.dc1:
            rt.call("hm_pack_conv_weight", _ptr(w), _ptr(self.wt_f), 17, self.Cout, self.Cin, 2, 2, 0, 0, rt.cd)
            rt.call("hm_pack_conv_weight", _ptr(w), _ptr(self.wt_d), 22, self.Cout, self.Cin, 2, 2, 0, 0, rt.cd)
            self.bias4.copy_(self.net.pview(self.bias).view(1, self.Cout).expand(4, self.Cout))
        else:
            per = self.Cin * self.Cout
            for u in range(self.kh):
                for v in range(self.kw):
                    t = u * self.kw + v
                    rt.call("hm_pack_conv_weight", _ptr(w), _ptr(self.wp_f[t * per:]), 2, self.Cout, self.Cin,
                            self.kh, self.kw, u, v, rt.cd)
            rt.call("hm_pack_conv_weight", _ptr(w), _ptr(self.wp_d), 3, self.Cout, self.Cin, self.kh, self.kw,
                    0, 0, rt.cd)

    # -- descriptors -------------------------------------------------------- #
    def _fwd_desc(self, rt, n, tap=None):
        d = _lib.ConvDesc()
        d.dtype = rt.cd
        d.B, d.H, d.W, d.C1, d.C2 = n, self.x1.shape[0], self.x1.shape[1], self.C1, self.C2
        d.up = self.up
        d.transposed = 0
        d.Cout = self.Cout
        d.split = self.Cout
        d.act = ACT[self.act.name]
        d.slope = self.act.slope
        d.accumulate = 0
        oH, oW = self.out.shape[0], self.out.shape[1]
        d.oH, d.oW = oH, oW
        if self.kind == "deconv":
            u, v = tap
            d.kh = d.kw = 1
            d.stride, d.pad = 1, 0
            d.Ho, d.Wo = self.Hv, self.Wv
            d.os, d.ou, d.ov = self.stride, u, v
        else:
            d.kh, d.kw, d.stride, d.pad = self.kh, self.kw, self.stride, self.pad
            d.Ho, d.Wo = oH, oW
            d.os, d.ou, d.ov = 1, 0, 0
        return d

    def _dc2_desc(self, rt, n):
        """Deconv2DLayer 2x2 stride 2, all four phases (hm_tc_conv, transposed == 2): H x W is the INPUT grid."""
        d = _lib.ConvDesc()
        d.dtype = rt.cd
        d.B, d.H, d.W, d.C1, d.C2, d.up = n, self.x1.shape[0], self.x1.shape[1], self.C1, self.C2, 0
        d.kh = d.kw = 2
        d.stride, d.pad, d.transposed = 2, 0, 2
        d.Ho, d.Wo, d.Cout = self.out.shape[0], self.out.shape[1], self.Cout
        d.oH, d.oW, d.os, d.ou, d.ov = d.Ho, d.Wo, 1, 0, 0
        d.split = self.Cout
        d.act, d.slope = ACT[self.act.name], self.act.slope
        d.accumulate = 0
        return d

    def _dc2_1x1_desc(self, rt, n, c1, c2, cout, split=None, acc=0):
        """1x1 convolution on the deconv's INPUT grid (weight / input gradient against s2d(dy))."""
        d = _lib.ConvDesc()
        d.dtype = rt.cd
        d.B, d.H, d.W, d.C1, d.C2, d.up = n, self.x1.shape[0], self.x1.shape[1], c1, c2, 0
        d.kh = d.kw = 1
        d.stride, d.pad, d.transposed = 1, 0, 0
        d.Ho, d.Wo, d.Cout = d.H, d.W, cout
        d.oH, d.oW, d.os, d.ou, d.ov = d.H, d.W, 1, 0, 0
        d.split = cout if split is None else split
        d.act, d.slope = 0, 0.0
        d.accumulate = acc
        return d

    def _tc_fwd_desc(self, rt, n):
        """Forward descriptor over the MATERIALISED (already upsampled) source, for the tensor-core kernels."""
        d = self._fwd_desc(rt, n)
        d.H, d.W, d.up = self.Hv, self.Wv, 0
        return d

    def _col1_desc(self, rt, n):
        """The 1x1 convolution over the im2col tensor (64 tap-channels) that stands for a thin-source conv."""
        d = self._fwd_desc(rt, n)
        d.H, d.W = self.out.shape[0], self.out.shape[1]
        d.C1, d.C2, d.kh, d.kw, d.pad, d.up, d.stride = 64, 0, 1, 1, 0, 0, 1
        return d

    def _tc_dgrad_desc(self, rt, n, acc):
        """Input gradient of a stride-1 convolution as a forward correlation of dy with the pack of mode 6."""
        d = _lib.ConvDesc()
        d.dtype = rt.cd
        d.B, d.H, d.W = n, self.out.shape[0], self.out.shape[1]
        d.C1, d.C2, d.up = self.Cout, 0, 0
        d.kh, d.kw, d.stride, d.pad = self.kh, self.kw, 1, self.kh - 1 - self.pad
        d.transposed = 0
        d.Ho, d.Wo, d.Cout = self.Hv, self.Wv, self.Cin
        d.oH, d.oW, d.os, d.ou, d.ov = self.Hv, self.Wv, 1, 0, 0
        d.split = self.C1
        d.act, d.slope = 0, 0.0
        d.accumulate = acc
        return d

    def _dense_desc(self, rt, n):
        d = _lib.ConvDesc()
        d.dtype = rt.cd
        d.B, d.H, d.W, d.C1, d.C2, d.up = n, 1, 1, self.Cin, 0, 0
        d.kh = d.kw = 1
        d.stride, d.pad, d.transposed = 1, 0, 0
        d.Ho, d.Wo, d.Cout = 1, 1, self.Cout
        d.oH, d.oW, d.os, d.ou, d.ov = 1, 1, 1, 0, 0
        d.split = self.Cout
        d.act, d.slope = ACT[self.act.name], self.act.slope
        d.accumulate = 0
        return d

    def _dc1_desc(self, rt, n, fwd):
        """The 2x2 deconvolution of a 1x1 input as a 1x1 convolution over [n,1,1,C]: forward Cin -> (u,v,co) columns, input
        gradient (u,v,co) -> Cin."""
        d = _lib.ConvDesc()
        d.dtype = rt.cd
        cin, cout = (self.Cin, 4 * self.Cout) if fwd else (4 * self.Cout, self.Cin)
        d.B, d.H, d.W, d.C1, d.C2, d.up = n, 1, 1, cin, 0, 0
        d.kh = d.kw = 1
        d.stride, d.pad, d.transposed = 1, 0, 0
        d.Ho, d.Wo, d.Cout = 1, 1, cout
        d.oH, d.oW, d.os, d.ou, d.ov = 1, 1, 1, 0, 0
        d.split = cout
        d.act, d.slope = (ACT[self.act.name], self.act.slope) if fwd else (0, 0.0)
        d.accumulate = 0
        return d

    def _dg6_desc(self, rt, n, acc):
        """Input gradient of (nearest-2x -> 5x5 'same' conv) as a forward 6x6 stride-2 pad-2 convolution of dy
        [n, 2H, 2W, Cout] onto the low-res source grid [n, H, W, Cin] (weights: pack mode 20)."""
        d = _lib.ConvDesc()
        d.dtype = rt.cd
        d.B, d.H, d.W = n, self.out.shape[0], self.out.shape[1]
        d.C1, d.C2, d.up = self.Cout, 0, 0
        d.kh, d.kw, d.stride, d.pad = 6, 6, 2, 2
        d.transposed = 0
        d.Ho, d.Wo, d.Cout = self.x1.shape[0], self.x1.shape[1], self.Cin
        d.oH, d.oW, d.os, d.ou, d.ov = d.Ho, d.Wo, 1, 0, 0
        d.split = self.Cin
        d.act, d.slope = 0, 0.0
        d.accumulate = acc
        return d

    def _srcs(self, rt, lo, hi, materialised):
        """Device pointers of the two sources for rows [lo,hi): the tensors themselves, or their 2x copies."""
        if materialised and self.up:
            return _ptr(self.x1u[lo:hi]), (_ptr(self.x2u[lo:hi]) if self.x2 is not None else None)
        return _ptr(self.x1.b(lo, hi)), (_ptr(self.x2.b(lo, hi)) if self.x2 is not None else None)

    def _dgrad_desc(self, rt, n, acc):
        d = _lib.ConvDesc()
        d.dtype = rt.cd
        d.B, d.H, d.W = n, self.out.shape[0], self.out.shape[1]
        d.C1, d.C2, d.up = self.Cout, 0, 0
        d.kh, d.kw, d.stride = self.kh, self.kw, self.stride
        d.Ho, d.Wo, d.Cout = self.Hv, self.Wv, self.Cin
        d.oH, d.oW, d.os, d.ou, d.ov = self.Hv, self.Wv, 1, 0, 0
        d.split = self.C1
        d.act, d.slope = 0, 0.0
        d.accumulate = acc
        if self.kind == "deconv":
            d.transposed, d.pad = 0, 0          # input gradient of a transposed conv = strided conv of dy
        else:
            d.transposed, d.pad = 1, self.pad
        return d

    # -- execution ---------------------------------------------------------- #
    def fwd(self, rt, lo, hi, det):
        n = hi - lo
        x1 = _ptr(self.x1.b(lo, hi))
        x2 = _ptr(self.x2.b(lo, hi)) if self.x2 is not None else None
        bias = _ptr(self.net.pview(self.bias))
        y = _ptr(self.out.b(lo, hi)) if self.out.buf is not None else None
        if self.dense_tc:
            rt.call("hm_tc_conv", C.byref(self._dense_desc(rt, n)), x1, None, _ptr(self.wt_f), bias, y, None)
        elif self.dc2:
            rt.call("hm_tc_conv", C.byref(self._dc2_desc(rt, n)), x1, x2, _ptr(self.wt_f), bias, y, None)
        elif self.dc1:
            rt.call("hm_tc_conv", C.byref(self._dc1_desc(rt, n, True)), x1, None, _ptr(self.wt_f), _ptr(self.bias4), y,
                    None)
        elif self.kind == "deconv":
            per = self.Cin * self.Cout
            for u in range(self.kh):
                for v in range(self.kw):
                    d = self._fwd_desc(rt, n, (u, v))
                    rt.call("hm_conv_gather", C.byref(d), x1, x2, _ptr(self.wp_f[(u * self.kw + v) * per:]),
                            bias, y, None)
        elif self.pool_fused is not None:
            # conv + bias + activation + 2x2 max-pool in one pass; the full-resolution activation is never written
            pool = self.pool_fused
            rt.call("hm_c1s2_conv", x1, _ptr(self.wk), bias, _ptr(pool.out.b(lo, hi)), _ptr(pool.idx[lo:hi]), n,
                    self.Hv, self.Wv, 256, ACT[self.act.name], self.act.slope)
        elif self.col1 or self.colk:
            if self.col1:
                rt.call("hm_im2col_c1", x1, _ptr(self.xc[lo:hi]), n, self.Hv, self.Wv, self.kh, self.kw, self.pad)
            else:
                rt.call("hm_im2col_thin", x1, x2, _ptr(self.xc[lo:hi]), n, self.Hv, self.Wv, self.C1, self.C2, self.kh,
                        self.kw, self.stride, self.pad, self.out.shape[0], self.out.shape[1])
            d = self._col1_desc(rt, n)
            rt.call("hm_tc_conv", C.byref(d), _ptr(self.xc[lo:hi]), None, _ptr(self.wt_f), bias, y, None)
        elif self.pool_tc is not None:
            # conv + bias + activation + 2x2 max-pool in the tcgen05 epilogue: the un-pooled activation is never written
            pool = self.pool_tc
            rt.call("hm_tc_conv_pool", C.byref(self._tc_fwd_desc(rt, n)), x1, x2, _ptr(self.wt_f), bias,
                    _ptr(pool.out.b(lo, hi)), _ptr(pool.idx[lo:hi]))
        elif self.up2:
            self._xu_valid = False
            d = self._fwd_desc(rt, n)
            rt.tc_conv(d, x1, None, _ptr(self.wt_f), bias, y, None)
        elif self.tc_fwd:
            if self.up:
                self._xu_valid = True
                for (x, xu, c) in ((self.x1, self.x1u, self.C1), (self.x2, self.x2u, self.C2)):
                    if x is not None:
                        rt.call("hm_upsample2_fwd", _ptr(x.b(lo, hi)), _ptr(xu[lo:hi]), rt.cd, n, x.shape[0],
                                x.shape[1], c, self.up)
            u1, u2 = self._srcs(rt, lo, hi, True)
            d = self._tc_fwd_desc(rt, n)
            rt.tc_conv(d, u1, u2, _ptr(self.wt_f), bias, y, None)
        else:
            d = self._fwd_desc(rt, n)
            rt.call("hm_conv_gather", C.byref(d), x1, x2, _ptr(self.wp_f), bias, y, None)

    def _bwd_pool_fused(self, rt, lo, hi, wgrad, input_grad):
        """conv5x5(1->64)+activation+max-pool backward straight from the pooled tensor's gradient (hm_c1s2_bwd):
        the full-resolution gradient of the un-pooled activation is never written."""
        n = hi - lo
        pool = self.pool_fused
        g, idx = _ptr(pool.out.g(lo, hi)), _ptr(pool.idx[lo:hi])
        # act' comes from bit 2 of the argmax bytes hm_c1s2_conv wrote: the pooled tensor is not read again (-0.54 GB per pass)
        signbits = os.environ.get("HMGAN_C1_SIGNBITS", "1") != "0"
        pl = None if signbits else _ptr(pool.out.b(lo, hi))
        t1 = self.x1.want_grad or (input_grad and self.x1.kind == "input" and self.x1.grad is not None)
        act, slope = ACT[self.act.name], self.act.slope
        x1 = _ptr(self.x1.b(lo, hi))
        net = self.net
        ws = _ptr(net.wscale) if net.wscale is not None else None
        ia, ib = net.ig_range if (net.ig_range is not None and t1) else (lo, hi)
        same = (ia, ib) == (lo, hi) and ws is None

        def dw_part():          # weight / bias gradient: reads g, the argmax bytes and x
            self.dwk.zero_()
            rt.call("hm_c1s2_bwd", x1, g, pl, idx, None, _ptr(self.dwk), None, ws, n, self.Hv, self.Wv, act, slope)
            rt.call("hm_c1s2_bwd_fold", _ptr(self.dwk), _ptr(net.gview(self.W)), _ptr(net.gview(self.bias)), self.Cout)
        if wgrad and t1 and same:                          # one launch produces both
            self.dwk.zero_()
            rt.call("hm_c1s2_bwd", x1, g, pl, idx, _ptr(self.wk2), _ptr(self.dwk), _ptr(self.ubuf[lo:hi]), None, n,
                    self.Hv, self.Wv, act, slope)
            rt.call("hm_c1s2_bwd_fold", _ptr(self.dwk), _ptr(net.gview(self.W)), _ptr(net.gview(self.bias)), self.Cout)
        else:
            if wgrad:
                dw_part()
            if t1:
                rt.call("hm_c1s2_bwd", None, _ptr(pool.out.g(ia, ib)), None if signbits else _ptr(pool.out.b(ia, ib)),
                        _ptr(pool.idx[ia:ib]),
                        _ptr(self.wk2), None, _ptr(self.ubuf[ia:ib]), None, ib - ia, self.Hv, self.Wv, act, slope)
        if t1:
            if self.x1.take_acc():
                raise NotImplementedError("accumulating into the source gradient of the fused first layer")
            rt.call("hm_c1s2_col2im", _ptr(self.ubuf[ia:ib]), _ptr(self.x1.g(ia, ib)), ib - ia, self.Hv, self.Wv)

    def bwd(self, rt, lo, hi, wgrad, input_grad):
        if self.pool_fused is not None:
            return self._bwd_pool_fused(rt, lo, hi, wgrad, input_grad)
        n = hi - lo
        g = self.out.g(lo, hi)
        if self.act.name != "linear" and not self.out.grad_is_preact:
            rt.call("hm_act_bwd", _ptr(g), _ptr(self.out.b(lo, hi)), _ptr(g), rt.cd, g.numel(),
                    ACT[self.act.name], self.act.slope, 0)
        x1 = _ptr(self.x1.b(lo, hi))
        x2 = _ptr(self.x2.b(lo, hi)) if self.x2 is not None else None
        g_plain = g
        if wgrad and self.net.wscale is not None:
            g = self.out.grad_w[lo:hi]        # weight and bias gradients see the per-sample-weighted gradient
        if wgrad:
            side = None if self.dc2 else rt.wgrad_stream()     # (dc2: the input gradient reads the wgrad block's s2d copy)
            if side is None:
                self._bwd_wgrad(rt, lo, hi, g, x1, x2)
            else:
                side.wait_stream(torch.cuda.current_stream(rt.device))      # dY (and its activation backward) is ready
                with torch.cuda.stream(side):
                    self._bwd_wgrad(rt, lo, hi, g, x1, x2)
                rt.mark_wgrad_forked()
        self._bwd_dgrad(rt, lo, hi, g_plain, wgrad, input_grad)

    def _bwd_wgrad(self, rt, lo, hi, g, x1, x2):
        """Weight and bias gradient of the layer from dY = g (reads g, the layer's sources and per-layer scratch only)."""
        n = hi - lo
        self.dwp.zero_()
        if self.dc2:
            # dW[ci][(phase,co)] = x^T . s2d(dy): a 1x1 tensor-core weight gradient on the input grid
            h, w = self.x1.shape[0], self.x1.shape[1]
            rt.call("hm_s2d_pad64", _ptr(g), _ptr(self.dy64[lo:hi]), n, h, w, self.Cout)
            d = self._dc2_1x1_desc(rt, n, self.C1, self.C2, 64)
            rt.call("hm_tc_wgrad", C.byref(d), x1, x2, _ptr(self.dy64[lo:hi]), _ptr(self.dwp))
            mode = 17
        elif self.dc1:
            d = self._dc1_desc(rt, n, True)
            d.act = 0
            rt.call("hm_tc_wgrad", C.byref(d), x1, None, _ptr(g), _ptr(self.dwp))      # dwp[ci][(u,v,co)]
            mode = 17
        elif self.kind == "deconv":
            per = self.Cin * self.Cout
            for u in range(self.kh):
                for v in range(self.kw):
                    d = self._fwd_desc(rt, n, (u, v))
                    rt.call("hm_conv_wgrad", C.byref(d), x1, x2, _ptr(g),
                            _ptr(self.dwp[(u * self.kw + v) * per:]))
            mode = 2
        elif self.col1 or self.colk:
            d = self._col1_desc(rt, n)
            rt.call("hm_tc_wgrad", C.byref(d), _ptr(self.xc[lo:hi]), None, _ptr(g), _ptr(self.dwp))
            mode = 0              # rows [0, kh*kw*Cin) of the [64][Cout] result are the packed gradient
        elif self.c1wg:
            rt.call("hm_c1s2_wgrad", _ptr(g), x1, _ptr(self.dwp), n, self.Hv, self.Wv)
            mode = 14
        elif self.thin_up2_wg:
            # dW of (nearest-2x -> 5x5 -> few channels): per output phase a 3x3 weight gradient on the low-res
            # source against the phase's strided slice of dy, folded back onto the 5x5 filter (unpack mode 9)
            h, w = self.x1.shape[0], self.x1.shape[1]
            rt.call("hm_s2d_pad64", _ptr(g), _ptr(self.dy64[lo:hi]), n, h, w, self.Cout)
            d = self._fwd_desc(rt, n)
            d.up, d.kh, d.kw, d.pad = 0, 3, 3, 1
            d.Ho, d.Wo, d.oH, d.oW, d.Cout, d.split = h, w, h, w, 64, 64
            rt.call("hm_tc_wgrad", C.byref(d), x1, None, _ptr(self.dy64[lo:hi]), _ptr(self.dwp))
            mode = 10
        elif self.wg8:
            rt.tc_wgrad(self._fwd_desc(rt, n), x1, None, _ptr(g), _ptr(self.dwp))
            mode = 8
        elif self.tc_wg and (not self.up or self.x1u is not None):
            if self.up and not getattr(self, "_xu_valid", False):   # forward did not materialise the 2x copy
                for (x, xu, c) in ((self.x1, self.x1u, self.C1), (self.x2, self.x2u, self.C2)):
                    if x is not None:
                        rt.call("hm_upsample2_fwd", _ptr(x.b(lo, hi)), _ptr(xu[lo:hi]), rt.cd, n, x.shape[0],
                                x.shape[1], c, self.up)
                self._xu_valid = True
            u1, u2 = self._srcs(rt, lo, hi, True)
            d = self._tc_fwd_desc(rt, n)
            rt.tc_wgrad(d, u1, u2, _ptr(g), _ptr(self.dwp))
            mode = 0
        else:
            d = self._fwd_desc(rt, n)
            rt.call("hm_conv_wgrad", C.byref(d), x1, x2, _ptr(g), _ptr(self.dwp))
            mode = 4 if self.kind == "dense" else 0
        rt.call("hm_unpack_conv_wgrad", _ptr(self.dwp), _ptr(self.net.gview(self.W)), mode, self.Cout,
                self.Cin, self.kh, self.kw)
        if self.db_done:              # the max-pool backward already summed the bias gradient
            self.db_done = False
        elif self.bias_grad_zero:
            self.net.gview(self.bias).zero_()
        else:
            M = n * self.out.shape[0] * self.out.shape[1]
            db = self.net.gview(self.bias)
            db.zero_()
            rt.call("hm_col_sum", _ptr(g), rt.cd, M, self.Cout, _ptr(db))

    def _bwd_dgrad(self, rt, lo, hi, g, wgrad, input_grad):
        """Input gradient(s) of the layer from dY = g."""
        n = hi - lo
        t1 = self.x1.want_grad or (input_grad and self.x1.kind == "input" and self.x1.grad is not None)
        t2 = self.x2 is not None and (self.x2.want_grad or (input_grad and self.x2.kind == "input"
                                                             and self.x2.grad is not None))
        if not (t1 or t2):
            return
        if self.kind == "dense":
            raise NotImplementedError("input gradient of a DenseLayer (only ever fed by the latent input)")
        if self.dc2 and (t1 or t2):
            # dx[q][ci] = sum_(phase,co) dy[2q+phase][co] W[ci][co][phase]: 1x1 convolution of s2d(dy) (pack mode 18),
            # channels [0,C1) to the first source's gradient, the rest to the second's
            if not wgrad:
                rt.call("hm_s2d_pad64", _ptr(g), _ptr(self.dy64[lo:hi]), n, self.x1.shape[0], self.x1.shape[1], self.Cout)
            acc = (self.x1.take_acc() if t1 else 0) | ((self.x2.take_acc() << 1) if t2 else 0)
            d = self._dc2_1x1_desc(rt, n, 64, 0, self.Cin, split=self.C1, acc=acc)
            rt.call("hm_tc_conv", C.byref(d), _ptr(self.dy64[lo:hi]), None, _ptr(self.wt_d), None,
                    _ptr(self.x1.g(lo, hi)) if t1 else None, _ptr(self.x2.g(lo, hi)) if t2 else None)
        elif self.dc1:
            d = self._dc1_desc(rt, n, False)
            d.accumulate = self.x1.take_acc()
            rt.call("hm_tc_conv", C.byref(d), _ptr(g), None, _ptr(self.wt_d), None, _ptr(self.x1.g(lo, hi)), None)
        elif self.c1dg and t1 and not self.x1.gw:
            # dy[B,2H,2W,1] -> dx[B,H,W,64] in one pass (four 3x3 phase filters as a 6x6 stride-2 gather of dy)
            self.x1.gw = True
            rt.call("hm_c1s2_conv", _ptr(g), _ptr(self.wk), None, _ptr(self.x1.g(lo, hi)), None, n, self.Hv, self.Wv,
                    64, 0, 0.0)
        elif self.dg6 and t1:
            d = self._dg6_desc(rt, n, self.x1.take_acc())
            rt.tc_conv(d, _ptr(g), None, _ptr(self.wt_d), None, _ptr(self.x1.g(lo, hi)), None)
        elif self.up:
            # gradient on the virtual (2x) grid, then the adjoint of the resampling
            if self.gup is None:
                self.gup = rt.empty((self.net.B, self.Hv, self.Wv, self.Cin))
            gu = self.gup[lo:hi]
            flat = gu.view(-1)
            n1 = n * self.Hv * self.Wv * self.C1
            y1 = flat[:n1] if t1 else None
            y2 = flat[n1:] if t2 else None
            if self.tc_dg:
                d = self._tc_dgrad_desc(rt, n, 0)
                rt.tc_conv(d, _ptr(g), None, _ptr(self.wt_d), None, _ptr(y1), _ptr(y2))
            else:
                d = self._tc_dgrad_desc(rt, n, 0) if self.fw_dg else self._dgrad_desc(rt, n, 0)
                rt.call("hm_conv_gather", C.byref(d), _ptr(g), None, _ptr(self.wp_d), None, _ptr(y1), _ptr(y2))
            for (x, y, c) in ((self.x1, y1, self.C1), (self.x2, y2, self.C2)):
                if y is None:
                    continue
                rt.call("hm_upsample2_bwd", _ptr(y), _ptr(x.g(lo, hi)), rt.cd, n, x.shape[0], x.shape[1], c,
                        self.up, x.take_acc())
        else:
            acc = (self.x1.take_acc() if t1 else 0) | ((self.x2.take_acc() << 1) if t2 else 0)
            y1 = _ptr(self.x1.g(lo, hi)) if t1 else None
            y2 = _ptr(self.x2.g(lo, hi)) if t2 else None
            if self.dg2_cat:
                d = self._dgrad_desc(rt, n, 0)
                d.split = d.Cout
                gc = self.gcat[lo:hi]
                rt.tc_conv(d, _ptr(g), None, _ptr(self.wt_d), None, _ptr(gc), None)
                M = n * self.Hv * self.Wv
                if t1:
                    rt.call("hm_slice_channels", _ptr(gc), y1, rt.cd, M, self.Cin, 0, self.C1, acc & 1)
                if t2:
                    rt.call("hm_slice_channels", _ptr(gc), y2, rt.cd, M, self.Cin, self.C1, self.C2, (acc >> 1) & 1)
            elif self.tc_dg:
                d = self._dgrad_desc(rt, n, acc) if self.dg2 else self._tc_dgrad_desc(rt, n, acc)
                rt.tc_conv(d, _ptr(g), None, _ptr(self.wt_d), None, y1, y2)
            else:
                d = self._tc_dgrad_desc(rt, n, acc) if self.fw_dg else self._dgrad_desc(rt, n, acc)
                rt.call("hm_conv_gather", C.byref(d), _ptr(g), None, _ptr(self.wp_d), None, y1, y2)


class BNActOp(object):
    """BatchNormLayer (+ the NonlinearityLayer that follows it)."""

    def __init__(self, net, layer, x, out, act):
        self.net, self.layer, self.x, self.out, self.act = net, layer, x, out, act
        self.beta, self.gamma, self.mean, self.inv_std = layer.params
        self.Cn = x.shape[2]
        self.st = None

    def alloc(self, rt, B):
        if self.st is None:
            self.sums = rt.zeros((2 * self.Cn,), torch.float64)
            self.red = rt.zeros((2 * self.Cn,), torch.float64)
            self.st = rt.zeros((4, self.Cn), torch.float32)     # batch mean, inv_std, scale, shift

    def pack(self, rt):
        pass

    def fwd(self, rt, lo, hi, det):
        n = hi - lo
        M = n * self.x.shape[0] * self.x.shape[1]
        net = self.net
        bm, bi, sc, sf = (self.st[i] for i in range(4))
        rm, ri = net.sview(self.mean), net.sview(self.inv_std)
        if det:
            sums = None
        else:
            self.sums.zero_()
            rt.call("hm_bn_stats", _ptr(self.x.b(lo, hi)), rt.cd, M, self.Cn, _ptr(self.sums))
            if rt.sync_bn_group is not None:       # mean over ranks of [sum x, sum x^2]: equal shards, so sums / M is global
                rt.allreduce_mean(self.sums)
            sums = _ptr(self.sums)
        rt.call("hm_bn_finalize", sums, M, self.Cn, _ptr(net.pview(self.gamma)), _ptr(net.pview(self.beta)),
                _ptr(rm), _ptr(ri), self.layer.epsilon, self.layer.alpha, 0 if det else 1,
                _ptr(bm), _ptr(bi), _ptr(sc), _ptr(sf))
        rt.call("hm_bn_apply_act", _ptr(self.x.b(lo, hi)), _ptr(self.out.b(lo, hi)), rt.cd, M, self.Cn,
                _ptr(sc), _ptr(sf), ACT[self.act.name], self.act.slope)

    def bwd(self, rt, lo, hi, wgrad, input_grad):
        n = hi - lo
        M = n * self.x.shape[0] * self.x.shape[1]
        net = self.net
        bm, bi = self.st[0], self.st[1]
        da, a, x = self.out.g(lo, hi), self.out.b(lo, hi), self.x.b(lo, hi)
        self.red.zero_()
        # fast mode: xhat from the layer's output instead of reading x a third time (hm_bn_bwd_*_a, include/hmgan.h)
        from_a = rt.precision == "fast" and os.environ.get("HMGAN_BN_FROM_A", "1") != "0"
        gam, bet = _ptr(net.pview(self.gamma)), _ptr(net.pview(self.beta))
        if from_a:
            rt.call("hm_bn_bwd_reduce_a", _ptr(da), _ptr(a), _ptr(x), rt.cd, M, self.Cn, _ptr(bm), _ptr(bi), gam, bet,
                    ACT[self.act.name], self.act.slope, _ptr(self.red))
        else:
            rt.call("hm_bn_bwd_reduce", _ptr(da), _ptr(a), _ptr(x), rt.cd, M, self.Cn, _ptr(bm), _ptr(bi),
                    ACT[self.act.name], self.act.slope, _ptr(self.red))
        if rt.sync_bn_group is not None:
            # [sum g, sum g*xhat] averaged over ranks: dx then uses the global-batch means, and d gamma / d beta come
            # out as (global sum) / world -- what the later sum all-reduce + 1/world of the gradients expects
            rt.allreduce_mean(self.red)
        assert self.x.consumers == 1
        self.x.gw = True
        if from_a:
            rt.call("hm_bn_bwd_apply_a", _ptr(da), _ptr(a), _ptr(x), _ptr(self.x.g(lo, hi)), rt.cd, M, self.Cn,
                    _ptr(bm), _ptr(bi), gam, bet, ACT[self.act.name], self.act.slope,
                    _ptr(self.red), _ptr(net.gview(self.gamma)) if wgrad else None,
                    _ptr(net.gview(self.beta)) if wgrad else None)
            return
        rt.call("hm_bn_bwd_apply", _ptr(da), _ptr(a), _ptr(x), _ptr(self.x.g(lo, hi)), rt.cd, M, self.Cn,
                _ptr(bm), _ptr(bi), gam, ACT[self.act.name], self.act.slope,
                _ptr(self.red), _ptr(net.gview(self.gamma)) if wgrad else None,
                _ptr(net.gview(self.beta)) if wgrad else None)


class PoolOp(object):
    """MaxPool2DLayer(2).  Its backward is fused with the backward of the monotonic
    activation that the producing convolution applied in its epilogue."""

    def __init__(self, net, x, out, act):
        self.net, self.x, self.out, self.act = net, x, out, act
        self.idx = None
        self.fused = False        # forward AND backward inside the producing convolution (hm_c1s2_*)
        self.fused_tc = False     # forward inside the producing convolution's epilogue (hm_tc_conv_pool)
        self.prod = None          # the ConvOp that wrote x (set by Net)

    def alloc(self, rt, B):
        H, W, Cn = self.out.shape
        self.idx = torch.empty((B, H, W, Cn), dtype=torch.uint8, device=rt.device)

    def pack(self, rt):
        pass

    def fwd(self, rt, lo, hi, det):
        if self.fused or self.fused_tc:            # written by the producing convolution
            return
        H, W, Cn = self.x.shape
        rt.call("hm_maxpool2_fwd", _ptr(self.x.b(lo, hi)), _ptr(self.out.b(lo, hi)), _ptr(self.idx[lo:hi]),
                rt.cd, hi - lo, H, W, Cn)

    def bwd(self, rt, lo, hi, wgrad, input_grad):
        if self.fused:            # the producing convolution's backward reads the pooled gradient itself (hm_c1s2_bwd)
            return
        H, W, Cn = self.x.shape
        assert self.x.consumers == 1
        self.x.gw = True
        db = None
        if wgrad and self.prod is not None and self.prod.kind == "conv":
            # the scattered gradient is exactly what the producing convolution sums for its bias gradient
            dbt = self.net.gview(self.prod.bias)
            dbt.zero_()
            db = _ptr(dbt)
            self.prod.db_done = True
        if self.net.wscale is not None:
            args = (_ptr(self.out.g(lo, hi)), _ptr(self.out.b(lo, hi)), _ptr(self.idx[lo:hi]))
            tail = (rt.cd, hi - lo, H, W, Cn, ACT[self.act.name], self.act.slope)
            rt.call("hm_maxpool2_bwd_scaled", *args, _ptr(self.x.g(lo, hi)), _ptr(self.x.grad_w[lo:hi]),
                    _ptr(self.net.wscale), *tail, db)
            return
        rt.call("hm_maxpool2_bwd", _ptr(self.out.g(lo, hi)), _ptr(self.out.b(lo, hi)), _ptr(self.idx[lo:hi]),
                _ptr(self.x.g(lo, hi)), rt.cd, hi - lo, H, W, Cn, ACT[self.act.name], self.act.slope, db)


class PermuteOp(object):
    """ReshapeLayer((-1, C, H, W)) of a flat feature vector: NCHW order -> NHWC buffer."""

    def __init__(self, net, x, out):
        self.net, self.x, self.out = net, x, out

    def alloc(self, rt, B):
        pass

    def pack(self, rt):
        pass

    def fwd(self, rt, lo, hi, det):
        H, W, Cn = self.out.shape
        rt.call("hm_permute", _ptr(self.x.b(lo, hi)), _ptr(self.out.b(lo, hi)), rt.cd, hi - lo, Cn, H, W, 0)

    def bwd(self, rt, lo, hi, wgrad, input_grad):
        H, W, Cn = self.out.shape
        assert self.x.consumers == 1
        self.x.gw = True
        rt.call("hm_permute", _ptr(self.out.g(lo, hi)), _ptr(self.x.g(lo, hi)), rt.cd, hi - lo, Cn, H, W, 1)


# --------------------------------------------------------------------------- #
# network
# --------------------------------------------------------------------------- #
class Net(object):
    """One lowered network: parameters (flat fp32 master copies in Lasagne's
    get_all_params order), program, buffers."""

    def __init__(self, rt, out_layer, input_layers=None, name="net", rng=None):
        self.rt, self.name = rt, name
        self.layers = L.get_all_layers(out_layer)
        if input_layers is None:
            input_layers = [l for l in self.layers if isinstance(l, L.InputLayer)]
        self.input_layers = list(input_layers)
        self.vals, self.ops, self._memo = [], [], {}
        self.inputs = []
        for il in self.input_layers:
            s = il.shape
            shape = (1, 1, s[1]) if len(s) == 2 else (s[2], s[3], s[1])
            v = self._new("input", shape)
            self._memo[(id(il), "linear", 0.0)] = v
            self.inputs.append(v)
        self.head = None
        body = self._strip_head(out_layer)
        self.out = self._lower(body, None)
        if self.out.kind != "buf":
            raise NotImplementedError("network output must be a materialised tensor")
        if self.head is not None and self.head["relu_head"]:
            self.out.grad_is_preact = True
        # pooled activations hand back pre-activation gradients
        for op in self.ops:
            if isinstance(op, PoolOp) and op.x.consumers == 1 and op.act.name != "linear":
                op.x.grad_is_preact = True
        # parameters
        self.params = L.get_all_params(out_layer)
        self._loc = {}
        nt = ns = 0
        for p in self.params:
            if p.trainable:
                self._loc[id(p)] = ("p", nt)
                nt += p.size
            else:
                self._loc[id(p)] = ("s", ns)
                ns += p.size
        self.n_trainable, self.n_stats = nt, ns
        self.pflat = rt.zeros((max(nt, 1),), torch.float32)
        self.gflat = rt.zeros((max(nt, 1),), torch.float32)
        self.sflat = rt.zeros((max(ns, 1),), torch.float32)
        self.opt_state = {}
        self.B = 0
        self.generation = 0
        self._packed = False
        self.single_pass = False
        self.wscale = self.ig_range = None
        if rng is not None:
            self.set_all_param_values([L.init_param(rng, p) for p in self.params])

    # -- lowering ----------------------------------------------------------- #
    def _new(self, kind, shape, srcs=(), up=0):
        v = Val(len(self.vals), kind, shape, srcs, up)
        self.vals.append(v)
        for s in srcs:
            s.consumers += 1
        return v

    def _strip_head(self, out_layer):
        """Detect NonlinearityLayer(ReshapeLayer(Pool2DLayer(avg, full extent)(conv)), f)."""
        l = out_layer
        if isinstance(l, L.NonlinearityLayer) and isinstance(l.input_layer, L.ReshapeLayer) and \
                isinstance(l.input_layer.input_layer, L.Pool2DLayer) and \
                l.input_layer.input_layer.mode.startswith("average"):
            pool = l.input_layer.input_layer
            conv = pool.input_layer
            s = conv.output_shape
            if pool.pool_size != (s[2], s[3]) or s[1] != 1 or l.input_layer.shape != (-1, 1):
                raise NotImplementedError(
                    "discriminator head: the average pool (size %r, derived from nch) must cover the %dx%d "
                    "feature map, i.e. nch == in_shp (reference architectures/dcgan.py:51-52)"
                    % (pool.pool_size, s[2], s[3]))
            f = L.as_nonlinearity(l.nonlinearity)
            if f.name not in ("linear", "sigmoid"):
                raise NotImplementedError("head nonlinearity %r" % f)
            cf = getattr(conv, "nonlinearity", L.linear)
            if cf.name not in ("linear", "rectify"):
                raise NotImplementedError("head convolution nonlinearity %r" % cf)
            self.head = dict(G=s[2] * s[3], out_act=f.name, relu_head=cf.name == "rectify")
            return conv
        return out_layer

    @staticmethod
    def _compose(inner, outer):
        if outer is None or outer.name == "linear":
            return inner
        if inner.name == "linear":
            return outer
        raise NotImplementedError("two stacked nonlinearities %r, %r" % (inner, outer))

    def _lower(self, layer, act):
        a = act or L.linear
        key = (id(layer), a.name, a.slope)
        if key in self._memo:
            return self._memo[key]
        v = self._lower_uncached(layer, a)
        self._memo[key] = v
        return v

    def _lower_uncached(self, layer, act):
        if isinstance(layer, L.InputLayer):
            if act.name != "linear":
                raise NotImplementedError("nonlinearity applied directly to a network input")
            raise ValueError("input layer %r was not declared" % layer)
        if isinstance(layer, L.NonlinearityLayer):
            return self._lower(layer.input_layer, self._compose(layer.nonlinearity, act))
        if isinstance(layer, L.DropoutLayer):
            if layer.p > 0:
                raise NotImplementedError("dropout is disabled in every experiment of the reference "
                                          "(SURVEY.md appendix A); p>0 is not implemented")
            return self._lower(layer.input_layer, act)
        if isinstance(layer, L.ConcatLayer):
            srcs = [self._lower(i, act) for i in layer.input_layers]
            if len(srcs) != 2:
                raise NotImplementedError("concatenation of %d tensors" % len(srcs))
            s = layer.output_shape
            return self._new("cat", (s[2], s[3], s[1]), srcs)
        if isinstance(layer, (L.Upscale2DLayer, BilinearUpsample2DLayer)):
            if act.name != "linear":
                raise NotImplementedError("nonlinearity after an upsampling layer")
            src = self._lower(layer.input_layer, None)
            s = layer.output_shape
            mode = _lib.UP_NEAREST2 if isinstance(layer, L.Upscale2DLayer) else _lib.UP_BILINEAR2
            return self._new("up", (s[2], s[3], s[1]), [src], up=mode)
        if isinstance(layer, L.BatchNormLayer):
            x = self._lower(layer.input_layer, None)
            if x.kind != "buf":
                raise NotImplementedError("BatchNorm of a virtual tensor")
            out = self._new("buf", x.shape, [x])
            self.ops.append(BNActOp(self, layer, x, out, act))
            if self.rt.precision == "fast" and x.consumers == 1:
                # batch-statistics BatchNorm removes the per-channel mean: the bias of the convolution that feeds ONLY
                # this layer has an exactly zero gradient (the reference computes rounding noise there); skip the pass
                for o in self.ops:
                    if isinstance(o, ConvOp) and o.out is x:
                        o.bias_grad_zero = True
            return out
        if isinstance(layer, L.ReshapeLayer):
            if act.name != "linear":
                raise NotImplementedError("nonlinearity after a reshape")
            x = self._lower(layer.input_layer, None)
            shp = layer.shape
            if len(shp) != 4 or x.shape[:2] != (1, 1) or shp[1] * shp[2] * shp[3] != x.shape[2]:
                raise NotImplementedError("reshape %r" % (shp,))
            out = self._new("buf", (shp[2], shp[3], shp[1]), [x])
            self.ops.append(PermuteOp(self, x, out))
            return out
        if isinstance(layer, L.Pool2DLayer):
            if layer.mode != "max" or layer.pool_size != (2, 2):
                raise NotImplementedError("pooling mode %r size %r (every experiment uses 2x2 max pooling)"
                                          % (layer.mode, layer.pool_size))
            if act.name != "linear":
                raise NotImplementedError("nonlinearity after a pooling layer")
            x = self._lower(layer.input_layer, None)
            if x.kind != "buf":
                raise NotImplementedError("pooling of a virtual tensor")
            s = layer.output_shape
            out = self._new("buf", (s[2], s[3], s[1]), [x])
            # which activation produced x (for the fused backward)?
            prod = [o for o in self.ops if getattr(o, "out", None) is x]
            pact = prod[0].act if prod and isinstance(prod[0], ConvOp) else L.linear
            if pact.name not in ("linear", "leaky_rectify", "rectify"):
                pact = L.linear
            pop = PoolOp(self, x, out, pact)
            self.ops.append(pop)
            cv = prod[0] if prod and isinstance(prod[0], ConvOp) else None
            pop.prod = cv
            if (cv is not None and cv.col1 and cv.Cout == 64 and cv.kh == 5 and cv.pad == 2 and x.consumers == 1
                    and cv.act.name in ("linear", "leaky_rectify", "rectify") and x.shape[0] % 2 == 0
                    and x.shape[1] % 2 == 0):
                cv.pool_fused, pop.fused = pop, True
                x.fused_away = True          # the un-pooled activation and its gradient are never materialised
            elif (cv is not None and self.rt.precision == "fast" and cv.kind == "conv" and cv.tc_fwd and not cv.up
                  and not cv.col1 and not cv.colk and cv.stride == 1 and x.consumers == 1
                  and os.environ.get("HMGAN_POOL_TC", "1") != "0"
                  and bool(_lib.query("hm_tc_conv_pool_supported", C.byref(cv._tc_fwd_desc(self.rt, 1))))):
                cv.pool_tc, pop.fused_tc = pop, True
                x.no_buf = True              # only the GRADIENT of the un-pooled activation exists (hm_maxpool2_bwd writes it)
            return out
        if isinstance(layer, (L.Conv2DLayer, L.TransposedConv2DLayer, L.DenseLayer)):
            src = self._lower(layer.input_layer, None)
            s = layer.output_shape
            if isinstance(layer, L.DenseLayer):
                kind, shape = "dense", (1, 1, s[1])
                if src.shape[:2] != (1, 1):
                    raise NotImplementedError("DenseLayer on a spatial tensor")
            else:
                kind = "conv" if isinstance(layer, L.Conv2DLayer) else "deconv"
                shape = (s[2], s[3], s[1])
            out = self._new("buf", shape, [src])
            self.ops.append(ConvOp(self, kind, layer, src, out, self._compose(layer.nonlinearity, act)))
            return out
        raise NotImplementedError("layer %r is not on the hot path" % layer)

    # -- parameters --------------------------------------------------------- #
    def _view(self, flat, p):
        _, off = self._loc[id(p)]
        return flat[off:off + p.size]

    def pview(self, p):
        return self._view(self.pflat if p.trainable else self.sflat, p)

    def sview(self, p):
        return self._view(self.sflat, p)

    def gview(self, p):
        return self._view(self.gflat, p)

    def get_all_param_values(self):
        return [self.pview(p).detach().cpu().numpy().reshape(p.shape).copy() for p in self.params]

    def set_all_param_values(self, values):
        if len(values) != len(self.params):
            raise ValueError("mismatch: got %d values to set %d parameters" % (len(values), len(self.params)))
        for p, v in zip(self.params, values):
            v = np.asarray(v, dtype=np.float32)
            if v.shape != p.shape:
                raise ValueError("mismatch: parameter has shape %r but value to set has shape %r"
                                 % (p.shape, v.shape))
            self.pview(p).copy_(torch.from_numpy(np.ascontiguousarray(v).reshape(-1)))
        self._packed = False

    def param_offset(self, op_index):
        """Offset in the flat trainable vector of the first parameter owned by ops[op_index:] (parameters are laid out in
        layer order, so the gradients of ops[op_index:] are the tail [offset, n_trainable))."""
        offs = []
        for op in self.ops[op_index:]:
            for p in (getattr(op, "W", None), getattr(op, "bias", None), getattr(op, "beta", None),
                      getattr(op, "gamma", None)):
                if p is not None and p.trainable:
                    offs.append(self._loc[id(p)][1])
        return min(offs) if offs else self.n_trainable

    def get_grads(self):
        return [self.gview(p).detach().cpu().numpy().reshape(p.shape).copy() for p in self.params if p.trainable]

    # -- buffers / execution ------------------------------------------------ #
    def ensure(self, B, input_grads=()):
        rt = self.rt
        if B > self.B:
            for v in self.vals:
                if v.kind == "buf" and getattr(v, "fused_away", False):
                    continue
                if v.kind == "buf":
                    v.buf = None if getattr(v, "no_buf", False) else rt.empty((B,) + v.shape)
                    v.grad = rt.empty((B,) + v.shape)
                    if self.single_pass and any(isinstance(o, ConvOp) and o.out is v for o in self.ops):
                        v.grad_w = rt.empty((B,) + v.shape)
                elif v.kind == "input":
                    v.buf = rt.empty((B,) + v.shape)
            for op in self.ops:
                op.alloc(rt, B)
            self.B = B
            self._packed = False
            self.generation += 1          # captured CUDA graphs hold the old buffers' addresses (Pix2Pix re-captures)
        for i in input_grads:
            v = self.inputs[i]
            if v.grad is None or v.grad.shape[0] < self.B:
                v.grad = rt.empty((self.B,) + v.shape)

    def pack(self):
        self.rt.begin_pack_batch()
        try:
            for op in self.ops:
                op.pack(self.rt)
        finally:
            self.rt.end_pack_batch()
        self._packed = True

    def forward(self, B, lo=0, hi=None, deterministic=False):
        """Inputs must already be in self.inputs[i].buf[lo:hi]."""
        hi = B if hi is None else hi
        if not self._packed:
            self.pack()
        for op in self.ops:
            op.fwd(self.rt, lo, hi, deterministic)
        return self.out.buf[lo:hi]

    def backward(self, lo, hi, wgrad=True, input_grad=False, wscale=None, ig_range=None, join=True, after_op=None):
        """self.out.grad[lo:hi] must hold d loss / d out.

        wscale (fp32 device vector, one weight per sample of [lo,hi)) switches on the single-pass weighted mode of a
        network that supports it (single_pass_ok): the input-gradient chain runs on the plain gradients, every
        weight/bias gradient on the weighted copies (Val.grad_w; self.out.grad_w must hold the weighted d loss / d out),
        and the network-input gradient is only produced for samples ig_range = (a, b).

        join=False leaves the weight-gradient side stream (Runtime.wgrad_stream) un-joined when the call returns.
        after_op: {op index: callable}, called once the backward launches of ops[index:] have all been issued (the
        gradients of their parameters -- the flat range [param_offset(index), n) -- are then complete in stream order)."""
        for v in self.vals:
            v.gw = False
        self.out.gw = True
        self.wscale, self.ig_range = wscale, ig_range
        try:
            for i in range(len(self.ops) - 1, -1, -1):
                self.ops[i].bwd(self.rt, lo, hi, wgrad, input_grad)
                if after_op is not None and i in after_op:
                    after_op[i]()
        finally:
            self.wscale, self.ig_range = None, None
            if join:       # join=False: the caller joins (Runtime.join_wgrad_stream) before it touches this network's
                self.rt.join_wgrad_stream()      # gradients or buffers again; weight gradients keep running meanwhile

    def single_pass_ok(self):
        """True if the weighted single-pass backward is implemented for this program: a scalar head, a first layer
        fused with its pool (hm_c1s2_*), then only [convolution -> 2x2 max-pool] pairs and the head convolution."""
        if self.rt.precision != "fast" or self.head is None or not self.ops:
            return False
        if not (isinstance(self.ops[0], ConvOp) and self.ops[0].pool_fused is not None):
            return False
        for i, op in enumerate(self.ops):
            if isinstance(op, PoolOp):
                Cn = op.x.shape[2]
                if not op.fused and (op.prod is None or Cn % 8 or 256 % (Cn // 8)):
                    return False
            elif isinstance(op, ConvOp):
                last = i == len(self.ops) - 1
                if op.kind != "conv" or op.x2 is not None or op.up:
                    return False
                if not last and not (i + 1 < len(self.ops) and isinstance(self.ops[i + 1], PoolOp)
                                     and self.ops[i + 1].x is op.out):
                    return False
                if last and op.act.name != "linear" and not op.out.grad_is_preact:
                    return False
            else:
                return False
        return True

    def enable_single_pass(self):
        self.single_pass = True
        self.B = 0                      # (re)allocate with the weighted gradient copies

    def apply_update(self, opt, lr_dev, gscale, hyper):
        """lasagne.updates.rmsprop / adam over the whole flat parameter vector."""
        rt, n = self.rt, self.n_trainable
        if n == 0:
            return
        if opt == "rmsprop":
            if "acc" not in self.opt_state:
                self.opt_state["acc"] = rt.zeros((n,), torch.float32)
            rt.call("hm_rmsprop", _ptr(self.pflat), _ptr(self.gflat), _ptr(self.opt_state["acc"]), n,
                    _ptr(lr_dev), hyper.get("rho", 0.9), hyper.get("epsilon", 1e-6), gscale)
        elif opt == "adam":
            if "m" not in self.opt_state:
                self.opt_state["m"] = rt.zeros((n,), torch.float32)
                self.opt_state["v"] = rt.zeros((n,), torch.float32)
                # the step count lives in device memory, so a captured CUDA graph keeps counting on every replay
                self.opt_state["t"] = rt.zeros((1,), torch.int32)
            rt.call("hm_inc_i32", _ptr(self.opt_state["t"]))
            rt.call("hm_adam_dev", _ptr(self.pflat), _ptr(self.gflat), _ptr(self.opt_state["m"]),
                    _ptr(self.opt_state["v"]), n, _ptr(lr_dev), hyper.get("beta1", 0.9),
                    hyper.get("beta2", 0.999), hyper.get("epsilon", 1e-8), _ptr(self.opt_state["t"]), gscale)
        else:
            raise NotImplementedError("optimiser %r" % opt)
        self._packed = False
