"""Two-stage DCGAN / pix2pix trainer with the reference's object surface
(reference pix2pix.py:19-425), executing on B200 through the hmgan kernel library.

What Theano compiled from the symbolic graph (pix2pix.py:87-147) is written out
here as an explicit step: four lowered networks (engine.Net), the five losses,
gradients of four losses with respect to four disjoint parameter sets evaluated
at the OLD parameters, and one simultaneous update.  The six callables keep
their names and their numpy NCHW float32 calling convention:

    train_fn(Z, X, Y) -> [dcgan_gen, dcgan_disc, p2p_gen, p2p_recon, p2p_disc]
    loss_fn(Z, X, Y)  -> same, without parameter updates (BatchNorm running
                         averages still move: the graph is non-deterministic,
                         pix2pix.py:92,99)
    gen_fn / gen_fn_det (X) -> P(X);   z_fn / z_fn_det (Z) -> G(z)

Differences from the reference, all additive:
  * discriminators see real and fake samples as one 2B batch (they have no
    BatchNorm in any experiment), which is arithmetically the same as two
    applications with shared weights;
  * gen_fn_p2p=None builds a DCGAN-only model (the 64-px gate of BASELINE.json);
  * keyword-only extras: device, precision ('parity' fp32 | 'fast' fp16), seed,
    process_group (data-parallel gradient all-reduce over NCCL), sync_bn (BatchNorm
    statistics over the whole process group instead of per rank);
  * X / Y may also be RAW uint8 NHWC batches, as the HDF5 file stores them
    (util.Hdf5Iterator(..., device_normalise=True)): the iterator's /255 or
    (x-127.5)/127.5 (reference util.py:33-35) then runs on the device
    (hm_u8_normalize) and the upload is a quarter of the float32 bytes.
"""
import gzip
import os
import pickle
from time import time

import numpy as np
import torch

import _lib
import engine
import lasagne_compat as L
from lasagne_compat import adam, floatX, shared
from util import convert_to_rgb, plot_grid, imsave, as_float_nchw

_ptr = engine._ptr


class Pix2Pix(object):
    def _print_network(self, l_out):
        for layer in L.get_all_layers(l_out):
            print(layer, layer.output_shape, "" if not hasattr(layer, 'nonlinearity') else layer.nonlinearity)
        print("# learnable params:", L.count_params(l_out, trainable=True))

    def __init__(self,
                 gen_fn_dcgan, disc_fn_dcgan,
                 gen_params_dcgan, disc_params_dcgan,
                 gen_fn_p2p, disc_fn_p2p,
                 gen_params_p2p, disc_params_p2p,
                 in_shp, latent_dim, is_a_grayscale, is_b_grayscale,
                 alpha=100, opt=adam, opt_args=None,
                 train_mode='both', reconstruction='l1', sampler=np.random.rand, lsgan=False, verbose=True,
                 device="cuda", precision=None, seed=None, process_group=None, loss_scale=None, sync_bn=False):
        assert train_mode in ['dcgan', 'p2p', 'both']
        assert reconstruction in ['l1', 'l2']
        if opt_args is None:
            opt_args = {'learning_rate': shared(floatX(1e-3))}
        self.is_a_grayscale = is_a_grayscale
        self.is_b_grayscale = is_b_grayscale
        self.latent_dim = latent_dim
        self.sampler = sampler
        self.in_shp = in_shp
        self.verbose = verbose
        self.train_mode = train_mode
        self.alpha = float(alpha)
        self.reconstruction = reconstruction
        self.lsgan = lsgan
        if precision is None:
            precision = os.environ.get("HMGAN_PRECISION", "parity")
        self.rt = rt = engine.Runtime(device, precision, loss_scale)
        self.pg = process_group
        if sync_bn:
            if process_group is None:
                raise ValueError("sync_bn=True needs a process_group")
            rt.sync_bn_group = process_group
        # the reference draws its Glorot weights from the global, unseeded np.random
        rng = np.random.RandomState(seed) if seed is not None else np.random.mtrand._rand
        self.have_dcgan = gen_fn_dcgan is not None
        self.have_p2p = gen_fn_p2p is not None
        dcgan_gen = dcgan_disc = p2p_gen = p2p_disc = None
        if self.have_dcgan:
            dcgan_gen = gen_fn_dcgan(latent_dim, is_a_grayscale, **gen_params_dcgan)
            dcgan_disc = disc_fn_dcgan(in_shp, is_a_grayscale, **disc_params_dcgan)
        if self.have_p2p:
            p2p_gen = gen_fn_p2p(in_shp, is_a_grayscale, is_b_grayscale, **gen_params_p2p)
            p2p_disc = disc_fn_p2p(in_shp, is_a_grayscale, is_b_grayscale, **disc_params_p2p)
        if verbose:
            for tag, net in (("dcgan gen:", dcgan_gen), ("dcgan disc:", dcgan_disc), ("p2p gen:", p2p_gen),
                             ("p2p disc:", p2p_disc["out"] if p2p_disc else None)):
                if net is not None:
                    print(tag)
                    self._print_network(net)
        self.dcgan = {'gen': dcgan_gen, 'disc': dcgan_disc}
        self.p2p = {'gen': p2p_gen, 'disc': p2p_disc["out"] if p2p_disc else None}
        self.G = self.D = self.P = self.Dp = None
        if self.have_dcgan:
            self.G = engine.Net(rt, dcgan_gen, name="dcgan_gen", rng=rng)
            self.D = engine.Net(rt, dcgan_disc, name="dcgan_disc", rng=rng)
        if self.have_p2p:
            self.P = engine.Net(rt, p2p_gen, name="p2p_gen", rng=rng)
            self.Dp = engine.Net(rt, p2p_disc["out"], input_layers=p2p_disc["inputs"], name="p2p_disc", rng=rng)
        self.nets = {'dcgan': {'gen': self.G, 'disc': self.D}, 'p2p': {'gen': self.P, 'disc': self.Dp}}
        for tag, net in (("disc_fn_dcgan", self.D), ("disc_fn_p2p", self.Dp)):
            # real and fake samples go through a discriminator as ONE 2B batch; BatchNorm statistics would then be taken
            # over the merged batch, whereas the reference applies the network twice (pix2pix.py:94-95,98,101), each with
            # its own batch statistics.  No experiment builds such a discriminator (bn=False everywhere): refuse it.
            if net is not None and any(isinstance(op, engine.BNActOp) for op in net.ops):
                raise NotImplementedError("%s with bn=True: BatchNorm inside a discriminator is not implemented (real "
                                          "and fake batches would share statistics)" % tag)
        # one weighted backward pass through the DCGAN discriminator instead of two (hm_adv_loss_pair, include/hmgan.h)
        self._single_pass = (self.D is not None and os.environ.get("HMGAN_SINGLE_PASS_D", "1") != "0"
                             and self.D.single_pass_ok())
        if self._single_pass:
            self.D.enable_single_pass()
        self._wscale = None
        # optimiser
        if not isinstance(opt, L._Optimiser):
            raise TypeError("opt must be lasagne_compat.rmsprop or lasagne_compat.adam")
        self.opt = opt.name
        hyper = dict(opt.defaults)
        hyper.update({k: v for k, v in opt_args.items() if k != 'learning_rate'})
        self.opt_hyper = hyper
        lr = opt_args.get('learning_rate', opt.defaults['learning_rate'])
        if not hasattr(lr, 'get_value'):
            lr = shared(floatX(lr))
        self.lr = lr
        self._lr_dev = rt.zeros((1,), torch.float32)
        self._lr_host = None
        self.losses = rt.zeros((5,), torch.float32)
        self.train_keys = ['dcgan_gen', 'dcgan_disc', 'p2p_gen', 'p2p_recon', 'p2p_disc']
        self._stage = {}
        self._graphs = {}
        self._pending, self._reduced = [], set()
        self._graphs_ok = rt.device.type == "cuda" and os.environ.get("HMGAN_CUDA_GRAPHS", "1") != "0"
        self.train_fn = lambda Z, X, Y: self._step_host(Z, X, Y, True)
        self.loss_fn = lambda Z, X, Y: self._step_host(Z, X, Y, False)
        self.train_fn_async = lambda Z, X, Y: self._step_host_async(Z, X, Y, True)      # device tensors, no host sync
        self.loss_fn_async = lambda Z, X, Y: self._step_host_async(Z, X, Y, False)
        self.gen_fn = lambda X: self._gen_p2p(X, False)
        self.gen_fn_det = lambda X: self._gen_p2p(X, True)
        self.z_fn = lambda Z: self._gen_dcgan(Z, False)
        self.z_fn_det = lambda Z: self._gen_dcgan(Z, True)

    # ------------------------------------------------------------------ #
    # host <-> device staging
    # ------------------------------------------------------------------ #
    def _to_dev(self, name, a):
        """numpy float32 (or raw uint8 image data) -> device staging buffer of the same type (reused across steps)."""
        src = self._host_tensor(a)
        t = self._stage.get(name)
        if t is None or t.shape != src.shape or t.dtype != src.dtype:
            t = self.rt.empty(tuple(src.shape), src.dtype)
            self._stage[name] = t
        t.copy_(src, non_blocking=True)
        return t

    def _load_nchw(self, src, dst, B, Cn, H, W, gray=True):
        """Image batch on the device -> NHWC compute-dtype buffer dst[:B].  float32: the reference's NCHW convention.
        uint8: raw NHWC data as the HDF5 file stores them (reference util.py:28-35), normalised here instead of on the
        host (`gray`: x/255, otherwise (x-127.5)/127.5)."""
        if src.dtype == torch.uint8:
            if src.numel() != B * Cn * H * W:
                raise ValueError("uint8 image batch of %d elements, expected [%d,%d,%d,%d] NHWC"
                                 % (src.numel(), B, H, W, Cn))
            self.rt.call("hm_u8_normalize", _ptr(src), _ptr(dst), self.rt.cd, src.numel(), 0 if gray else 1)
            return
        self.rt.call("hm_nchw_to_nhwc", _ptr(src), _ptr(dst), self.rt.cd, B, Cn, H, W)

    def _store_nchw(self, src, B, Cn, H, W):
        out = self.rt.empty((B, Cn, H, W), torch.float32)
        self.rt.call("hm_nhwc_to_nchw", _ptr(src), _ptr(out), self.rt.cd, B, Cn, H, W)
        return out

    def _copy(self, src, dst):
        self.rt.call("hm_cast", _ptr(src), self.rt.cd, _ptr(dst), self.rt.cd, src.numel())

    def _sync_lr(self):
        v = float(self.lr.get_value())
        if v != self._lr_host:
            self._lr_dev.fill_(v)
            self._lr_host = v

    # ------------------------------------------------------------------ #
    # the step (reference pix2pix.py:87-147)
    # ------------------------------------------------------------------ #
    def _allreduce_async(self, flat):
        """Sum all-reduce of (a slice of) a network's flat gradient over the data-parallel group, issued NOW behind
        everything already launched -- including the weight gradients still running on the side stream -- and awaited
        only before the update, so it overlaps the backward passes that follow (NCCL over NVLink)."""
        if self.pg is None:
            return
        import torch.distributed as dist
        rt = self.rt
        side = rt.side_if_forked()
        if side is not None:
            side.wait_stream(torch.cuda.current_stream(rt.device))
            with torch.cuda.stream(side):
                work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.pg, async_op=True)
        else:
            work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.pg, async_op=True)
        self._pending.append(work)
        self.allreduce_calls = getattr(self, "allreduce_calls", 0) + 1       # diagnostic: buckets issued so far
        base = flat._base if flat._base is not None else flat
        for net in self._nets():
            if net.gflat is base:
                self._reduced.add(id(net))

    def _update_on_side_stream(self, net, ls):
        """Optimiser step + weight re-packs of `net` on the weight-gradient side stream, behind the network's last weight
        gradient (and its all-reduce), while the current stream goes on with work that no longer reads this network's
        weights or gradients.  Returns False (nothing done) when there is no side stream."""
        rt = self.rt
        side = rt.side_if_forked()
        if side is None:
            return False
        self._sync_lr()
        world = 1
        side.wait_stream(torch.cuda.current_stream(rt.device))     # gradients produced on the current stream
        with torch.cuda.stream(side):
            if self.pg is not None:
                import torch.distributed as dist
                world = dist.get_world_size(self.pg)
                for work in self._pending:                     # this network's all-reduce(s): the side stream waits
                    work.wait()
                self._pending = []
            net.apply_update(self.opt, self._lr_dev, 1.0 / (ls * world), self.opt_hyper)
            net.pack()
        return True

    def _adv(self, net, h, dh, target, slot, gscale):
        """adv_loss(out, target).mean() of pix2pix.py:102-110 on a (half) batch of
        discriminator outputs; optionally its gradient."""
        head = net.head or dict(G=1, out_act="linear", relu_head=False)
        R = h.numel() // head["G"]
        self.rt.call("hm_adv_loss", _ptr(h), _ptr(dh), self.rt.cd, R, head["G"], _lib.ACT[head["out_act"]],
                     float(target), 1 if self.lsgan else 0, 1 if head["relu_head"] else 0, 1.0, gscale, 0,
                     _ptr(self.losses[slot:]))

    def step_device(self, Zd, Xd, Yd, train=True):
        """One train_fn / loss_fn evaluation on float32 NCHW DEVICE tensors (Yd may be None for a model without the
        pix2pix stage); the five losses stay on the device in self.losses (no host synchronisation).

        On a CUDA device the 200-650 kernel launches of a step are captured once into a CUDA graph per
        (batch size, train flag) -- after two eager warm-up calls that size every buffer -- and replayed
        afterwards, so the host cost of a step is one graph launch (all buffers are static, tensor maps and
        descriptors are kernel arguments, the learning rate is read from device memory).  Set
        HMGAN_CUDA_GRAPHS=0 to always run eagerly.  Invariants that keep replays correct: (i) the packed compute-dtype
        weight copies are refreshed INSIDE the step right after the update, and by the host before any launch when
        set_all_param_values / a reallocation invalidated them (_ensure_packed), so no graph depends on a host flag;
        (ii) a graph is dropped and re-captured when any network reallocated its buffers after the capture
        (Net.generation); (iii) Adam's step count lives in device memory."""
        if not self._graphs_ok:
            return self._step_eager(Zd, Xd, Yd, train)
        key = (tuple(Zd.shape), tuple(Xd.shape), Xd.dtype, tuple(Yd.shape) if Yd is not None else None,
               Yd.dtype if Yd is not None else None, bool(train))
        st = self._graph_state(key, "graph")
        if st["graph"] is None:
            st["calls"] += 1
            if st["calls"] <= 2:
                return self._step_eager(Zd, Xd, Yd, train)
            # capture: static input buffers, then record the step on torch's capture stream
            st["Z"], st["X"], st["Y"] = Zd.clone(), Xd.clone(), (Yd.clone() if Yd is not None else None)
            self._sync_lr()
            self._ensure_packed()
            torch.cuda.synchronize(self.rt.device)
            g = torch.cuda.CUDAGraph()
            l0 = self.rt.launches
            with torch.cuda.graph(g):
                self._step_eager(st["Z"], st["X"], st["Y"], train)
            st["launches"] = self.rt.launches - l0
            st["graph"] = g
            st["gen"] = self._generations()
        for dst, src in ((st["Z"], Zd), (st["X"], Xd), (st["Y"], Yd)):
            if dst is not None and dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        self._sync_lr()
        self._ensure_packed()
        st["graph"].replay()
        self.rt.launches += st["launches"]
        return self.losses

    def _nets(self):
        return [n for n in (self.G, self.D, self.P, self.Dp) if n is not None]

    def _generations(self):
        return tuple(n.generation for n in self._nets())

    def _graph_state(self, key, slot):
        """Capture state of one (shapes, train flag) key.  A graph holds raw addresses of the networks' buffers: when a
        network has reallocated since the capture (Net.ensure with a larger batch from ANY entry point) the graph is
        dropped and captured again after two fresh eager calls."""
        st = self._graphs.get(key)
        if st is not None and st.get(slot) is not None and st.get("gen") != self._generations():
            st = None
        if st is None:
            st = self._graphs[key] = {"calls": 0, slot: None}
        return st

    def _ensure_packed(self):
        """Refresh the packed weight copies of any network whose master parameters changed outside a step
        (set_all_param_values, load_model) or whose buffers were reallocated; eager launches, never captured."""
        for n in self._nets():
            if not n._packed and n.B > 0:
                n.pack()

    def _step_eager(self, Zd, Xd, Yd, train=True, part=0):
        """part 0: the whole step.  part 1: only what depends on Z alone (G's forward pass); part 3: D(x), which depends
        on X alone; part 4 (a pix2pix-only model): P(X), which depends on X alone; part 2: everything else (given the
        other parts).  The host path captures the parts as separate CUDA graphs so that the X/Y upload overlaps part 1,
        D(x) runs beside the rest of it and the texture batch Y lands under P's forward pass (_step_host_overlapped)."""
        rt = self.rt
        B = int(Xd.shape[0])
        S = self.in_shp
        ls = rt.loss_scale
        ca = 1 if self.is_a_grayscale else 3
        fork = part == 0 and self.have_dcgan and rt._fork_ok      # D(x) beside G(z) on the auxiliary stream
        if part in (0, 1):
            self.losses.zero_()
            if self.have_dcgan:
                self.G.ensure(B)
                self.D.ensure(2 * B, input_grads=(0,))
                if fork:
                    # D(x) needs only X: it runs beside G's forward pass, whose 4x4..64x64 layers fill a fraction of the
                    # SMs.  (Without BatchNorm D treats every sample independently, so two half-batch passes equal one.)
                    with rt.fork():
                        self._load_nchw(Xd, self.D.inputs[0].buf, B, ca, S, S, self.is_a_grayscale)
                        self.D.forward(2 * B, 0, B)                         # D(x)            :94
                rt.call("hm_cast", _ptr(Zd), _lib.F32, _ptr(self.G.inputs[0].buf), rt.cd, Zd.numel())
                self.G.forward(B)                                           # G(z)            :92
            if part == 1:
                return self.losses
        if part == 3:
            self._load_nchw(Xd, self.D.inputs[0].buf, B, ca, S, S, self.is_a_grayscale)
            self.D.forward(2 * B, 0, B)                                     # D(x)            :94
            return self.losses
        if part == 4:
            self._p2p_part(Xd, Yd, B, train, phase=1)                       # P(X)            :99
            return self.losses
        p2p_phase = 2 if (part == 2 and not self.have_dcgan) else 0         # P(X) came from part 4
        upd = []
        # a joint step: the pix2pix half is independent of the DCGAN half until the update -- it runs on its own lane
        # (stream), so that its many HBM-bound passes (17 BatchNorm layers, resampling) overlap the DCGAN half's
        # tensor-core kernels and vice versa; the halves meet again before the update
        p2p_done, p2p_upd = False, []
        if self.have_p2p and self.have_dcgan and rt._fork_ok and part in (0, 2) and \
                os.environ.get("HMGAN_FORK_P2P", "1") != "0":
            with rt.fork("p2p"):
                p2p_upd = self._p2p_part(Xd, Yd, B, train)
            p2p_done = True
        if self.have_dcgan:
            G, D = self.G, self.D
            do = train and self.train_mode in ('both', 'dcgan')
            gz = G.out.buf[:B]
            self._copy(gz, D.inputs[0].buf[B:2 * B])
            if fork or (part == 2 and rt._fork_ok):
                if fork:
                    rt.join()                                               # D(x) done (and D's packs, if it made them)
                D.forward(2 * B, B, 2 * B)                                  # D(G(z))         :95
            else:
                self._load_nchw(Xd, D.inputs[0].buf, B, ca, S, S, self.is_a_grayscale)
                D.forward(2 * B)                                            # D(x), D(G(z))   :94-95
            h = D.out.buf[:2 * B]
            dh = D.out.grad if do else None
            self._adv(D, h[:B], dh[:B] if do else None, 1., 1, ls)          # disc_loss_dcgan :108
            if do and self._single_pass:
                # D(G(z)) enters disc_loss_dcgan (target 0) and gen_loss_dcgan (target 1): per sample the two backward
                # passes through D differ by a scalar, so ONE pass carries both (weights sw on dW, sg on dG(z))
                head = D.head
                if self._wscale is None or self._wscale.numel() != 3 * B:
                    self._wscale = torch.ones(3 * B, dtype=torch.float32, device=rt.device)   # [1]*B | sw | sg
                ws = self._wscale
                dhw = D.out.grad_w
                R = h[B:].numel() // head["G"]
                self._copy(dh[:B], dhw[:B])
                rt.call("hm_adv_loss_pair", _ptr(h[B:]), _ptr(dh[B:2 * B]), _ptr(dhw[B:2 * B]), _ptr(ws[B:]),
                        _ptr(ws[2 * B:]), rt.cd, R, head["G"], _lib.ACT[head["out_act"]], 1 if self.lsgan else 0,
                        1 if head["relu_head"] else 0, ls, _ptr(self.losses[1:]), _ptr(self.losses[0:]))
                # D's weight gradients stay on the side stream (un-joined) while G's backward pass runs: nothing below
                # touches D's buffers again before the join at the end of G.backward
                D.backward(0, 2 * B, wgrad=True, input_grad=True, wscale=ws[:2 * B], ig_range=(B, 2 * B), join=False)
                self._allreduce_async(D.gflat)                  # under G's backward pass
                early = self._update_on_side_stream(D, ls)      # ... and so are D's update and weight re-packs
                n_in = D.inputs[0].grad[B:2 * B].numel() // B
                rt.call("hm_scale_rows", _ptr(D.inputs[0].grad[B:2 * B]), _ptr(ws[2 * B:]), _ptr(G.out.grad[:B]),
                        rt.cd, B, n_in)
                # G's flat gradient in two buckets: everything above the dense layer's block as soon as it is complete
                # (the dense layer's 8.2M-element gradient, the head of the flat vector, is produced last)
                hook = None
                if self.pg is not None and len(G.ops) > 4:
                    k = 3
                    off = G.param_offset(k)
                    if 0 < off < G.n_trainable:
                        hook = {k: lambda: self._allreduce_async(G.gflat[off:])}
                G.backward(0, B, wgrad=True, after_op=hook)
                self._allreduce_async(G.gflat[:off] if hook else G.gflat)
                upd += [G] if early else [G, D]
            else:
                self._adv(D, h[B:], dh[B:2 * B] if do else None, 0., 1, ls)
                if do:
                    D.backward(0, 2 * B, wgrad=True, input_grad=False)
                self._adv(D, h[B:], dh[B:2 * B] if do else None, 1., 0, ls)     # gen_loss_dcgan  :107
                if do:
                    D.backward(B, 2 * B, wgrad=False, input_grad=True)
                    self._copy(D.inputs[0].grad[B:2 * B], G.out.grad[:B])
                    G.backward(0, B, wgrad=True)
                    upd += [G, D]
        if self.have_p2p and not p2p_done:
            upd += self._p2p_part(Xd, Yd, B, train, phase=p2p_phase)
        elif p2p_done:
            rt.join("p2p")
            upd += p2p_upd
        if upd:
            self._sync_lr()
            world = 1
            if self.pg is not None:
                import torch.distributed as dist
                world = dist.get_world_size(self.pg)
                for net in upd:                                  # data-parallel: sum of per-rank mean gradients
                    if id(net) not in self._reduced:
                        self._allreduce_async(net.gflat)
                for work in self._pending:                       # the current stream waits for every reduction
                    work.wait()
                self._pending, self._reduced = [], set()
            for net in upd:
                net.apply_update(self.opt, self._lr_dev, 1.0 / (ls * world), self.opt_hyper)
                net.pack()       # packed copies follow the master weights inside the step (and inside its CUDA graph)
                # (tried: G's packs at the START of the next step, beside D(x), instead of here where nothing else is left
                # to run -- 12.17 / 12.20 vs 12.16 / 12.25 ms on one box: no gain, removed)
        return self.losses

    def _p2p_part(self, Xd, Yd, B, train, phase=0):
        """The pix2pix half of the step (pix2pix.py:98-101,110-121): P and Dp forward, the three losses, their backward
        passes.  Returns the networks to update.  Independent of the DCGAN half until the update.  phase 1 = only P's
        forward pass (all that can run before Y has arrived), phase 2 = the rest, phase 0 = both."""
        rt = self.rt
        S = self.in_shp
        ls = rt.loss_scale
        ca = 1 if self.is_a_grayscale else 3
        upd = []
        P, Dp = self.P, self.Dp
        do = train and self.train_mode in ('both', 'p2p')
        P.ensure(B)
        Dp.ensure(2 * B, input_grads=(1,))
        cb = 1 if self.is_b_grayscale else 3
        if phase in (0, 1):
            self._load_nchw(Xd, P.inputs[0].buf, B, ca, S, S, self.is_a_grayscale)
            px = P.forward(B)                                           # P(X)            :99
            if phase == 1:
                return upd
        else:
            px = P.out.buf[:B]
        a_in, b_in = Dp.inputs
        self._load_nchw(Xd, a_in.buf, B, ca, S, S, self.is_a_grayscale)
        self._copy(a_in.buf[:B], a_in.buf[B:2 * B])
        self._load_nchw(Yd, b_in.buf, B, cb, S, S, self.is_b_grayscale)
        self._copy(px, b_in.buf[B:2 * B])
        h = Dp.forward(2 * B)                                           # Dp(X,Y), Dp(X,P(X)) :98,101
        dh = Dp.out.grad if do else None
        self._adv(Dp, h[:B], dh[:B] if do else None, 1., 4, ls)         # disc_loss_p2p   :121
        self._adv(Dp, h[B:2 * B], dh[B:2 * B] if do else None, 0., 4, ls)
        if do:
            Dp.backward(0, 2 * B, wgrad=True, input_grad=False)
            self._allreduce_async(Dp.gflat)                             # data-parallel: under the rest of this half
        self._adv(Dp, h[B:2 * B], dh[B:2 * B] if do else None, 1., 2, ls)   # gen_loss_p2p :110
        dpx = None
        if do:
            Dp.backward(B, 2 * B, wgrad=False, input_grad=True)
            dpx = b_in.grad[B:2 * B]
        # recon_loss :112-115 ; gradient weight alpha (gen_total_loss_p2p :117)
        rt.call("hm_recon_loss", _ptr(px), _ptr(b_in.buf[:B]), _ptr(dpx), rt.cd, px.numel(),
                1 if self.reconstruction == 'l2' else 0, 1.0, ls * self.alpha, 1, _ptr(self.losses[3:]))
        if do:
            self._copy(dpx, P.out.grad[:B])
            # data-parallel: the U-Net's flat gradient (92 MB) in two buckets -- the decoder half (the tail of the flat
            # vector, produced first) goes out while the encoder's backward pass runs
            hook, off = None, 0
            if self.pg is not None and len(P.ops) > 8:
                k = min(range(1, len(P.ops)), key=lambda i: abs(P.param_offset(i) - P.n_trainable // 2))
                off = P.param_offset(k)
                if 0 < off < P.n_trainable:
                    hook = {k: lambda: self._allreduce_async(P.gflat[off:])}
            P.backward(0, B, wgrad=True, after_op=hook)
            self._allreduce_async(P.gflat[:off] if hook else P.gflat)
            upd += [P, Dp]
        return upd

    def _host_tensor(self, a):
        """float32 host tensor of a numpy array / tensor; raw uint8 image data stay uint8 (normalised on the device)."""
        if isinstance(a, torch.Tensor):
            return a if a.dtype in (torch.float32, torch.uint8) else a.float()
        a = np.asarray(a)
        if a.dtype == np.uint8:
            return torch.from_numpy(np.ascontiguousarray(a))
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))

    def _step_host_overlapped(self, Z, X, Y, train):
        """train_fn / loss_fn from HOST inputs with the step captured as several graphs: part 1 (G's forward pass, needs
        only Z) starts at once while X and then Y are uploaded on a copy stream; D(x) (or, without a DCGAN, P(X)) starts
        as soon as X has landed; part 2 waits for Y and for those.  Returns None until the graphs exist (two eager calls
        size every buffer first)."""
        Zs, Xs = self._host_tensor(Z), self._host_tensor(X)
        Ys = self._host_tensor(Y) if self.have_p2p else None
        key = ("host", tuple(Zs.shape), tuple(Xs.shape), Xs.dtype, tuple(Ys.shape) if Ys is not None else None,
               Ys.dtype if Ys is not None else None, bool(train))
        st = self._graph_state(key, "gA")
        dev = self.rt.device
        if st["gA"] is None:
            st["calls"] += 1
            if st["calls"] <= 2:
                return None
            st["Z"], st["X"] = Zs.to(dev), Xs.to(dev)
            st["Y"] = Ys.to(dev) if Ys is not None else None
            self._sync_lr()
            self._ensure_packed()
            torch.cuda.synchronize(dev)
            l0 = self.rt.launches
            gA, gB = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(gA):
                self._step_eager(st["Z"], st["X"], st["Y"], train, part=1)
            st["gC"] = st["gP"] = None
            if self.have_dcgan and self.rt._fork_ok:        # D(x) as its own graph, replayed beside G's forward pass
                st["gC"] = torch.cuda.CUDAGraph()
                self.rt.lane = "aux"
                try:
                    with torch.cuda.graph(st["gC"], pool=gA.pool()):
                        self._step_eager(st["Z"], st["X"], st["Y"], train, part=3)
                finally:
                    self.rt.lane = "main"
            elif self.have_p2p and not self.have_dcgan:     # P(X) as its own graph: Y is uploaded under it
                st["gP"] = torch.cuda.CUDAGraph()
                with torch.cuda.graph(st["gP"], pool=gA.pool()):
                    self._step_eager(st["Z"], st["X"], st["Y"], train, part=4)
            with torch.cuda.graph(gB, pool=gA.pool()):
                self._step_eager(st["Z"], st["X"], st["Y"], train, part=2)
            st["launches"] = self.rt.launches - l0
            st["gA"], st["gB"] = gA, gB
            st["gen"] = self._generations()
            st["side"], st["aux"] = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
            st["evX"], st["evY"], st["evC"] = torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()
        self._sync_lr()
        self._ensure_packed()
        main = torch.cuda.current_stream(dev)
        if st.get("done") is not None:              # X / Y of the previous step are free once its second graph has finished
            st["side"].wait_event(st["done"])       # (a no-op after a host synchronisation; needed by *_fn_async)
        st["Z"].copy_(Zs, non_blocking=True)
        st["gA"].replay()
        with torch.cuda.stream(st["side"]):         # the copy stream: X, then Y right behind it
            st["X"].copy_(Xs, non_blocking=True)
            st["evX"].record(st["side"])
            if Ys is not None:
                st["Y"].copy_(Ys, non_blocking=True)
            st["evY"].record(st["side"])
        if st["gC"] is not None:
            with torch.cuda.stream(st["aux"]):      # D(x) as soon as X has landed, beside G's forward pass and Y's upload
                st["aux"].wait_event(st["evX"])
                st["gC"].replay()
                st["evC"].record(st["aux"])
        if st["gP"] is not None:
            main.wait_event(st["evX"])
            st["gP"].replay()                       # P(X) while Y is still on its way
        main.wait_event(st["evY"])
        if st["gC"] is not None:
            main.wait_event(st["evC"])
        st["gB"].replay()
        if st.get("done") is None:
            st["done"] = torch.cuda.Event()
        st["done"].record(main)
        self.rt.launches += st["launches"]
        return self.losses

    def _step_host(self, Z, X, Y, train):
        if self._graphs_ok and os.environ.get("HMGAN_OVERLAP_H2D", "1") != "0":
            out = self._step_host_overlapped(Z, X, Y, train)
            if out is not None:
                return [np.float32(v) for v in out.cpu().numpy()]
        Zd = self._to_dev("Z", Z)
        Xd = self._to_dev("X", X)
        Yd = self._to_dev("Y", Y) if self.have_p2p else None      # a DCGAN-only model never reads the texture batch
        losses = self.step_device(Zd, Xd, Yd, train).cpu().numpy()
        return [np.float32(v) for v in losses]

    def _step_host_async(self, Z, X, Y, train):
        """train_fn / loss_fn without the per-step host synchronisation: the step is enqueued and a DEVICE copy of the five
        losses is returned (read it with .cpu() whenever convenient; the next call may be issued at once).  What the
        reference's loop does after every batch -- np.mean over the epoch's losses (pix2pix.py:201-212) -- only needs
        the values at the end of the epoch: Pix2Pix.train(loss_sync_every=N) uses this."""
        if self._graphs_ok and os.environ.get("HMGAN_OVERLAP_H2D", "1") != "0":
            out = self._step_host_overlapped(Z, X, Y, train)
            if out is not None:
                return out.clone()
        Zd = self._to_dev("Z", Z)
        Xd = self._to_dev("X", X)
        Yd = self._to_dev("Y", Y) if self.have_p2p else None
        return self.step_device(Zd, Xd, Yd, train).clone()

    def _gen_dcgan(self, Z, det):
        if not self.have_dcgan:
            raise RuntimeError("this model was built without a DCGAN")
        Zd = self._to_dev("Z", Z)
        B = Zd.shape[0]
        G = self.G
        G.ensure(B)
        self.rt.call("hm_cast", _ptr(Zd), _lib.F32, _ptr(G.inputs[0].buf), self.rt.cd, Zd.numel())
        out = G.forward(B, deterministic=det)
        H, W, Cn = G.out.shape
        return self._store_nchw(out, B, Cn, H, W).cpu().numpy()

    def _gen_p2p(self, X, det):
        if not self.have_p2p:
            raise RuntimeError("this model was built without a pix2pix stage")
        Xd = self._to_dev("X", X)
        if Xd.dtype == torch.uint8:                # raw NHWC data ([B,S,S] or [B,S,S,C])
            B, S = int(Xd.shape[0]), int(Xd.shape[1])
            ca = int(Xd.shape[3]) if Xd.dim() == 4 else 1
        else:
            B, ca, S, _ = Xd.shape
        P = self.P
        P.ensure(B)
        self._load_nchw(Xd, P.inputs[0].buf, B, ca, S, S, self.is_a_grayscale)
        out = P.forward(B, deterministic=det)
        H, W, Cn = P.out.shape
        return self._store_nchw(out, B, Cn, H, W).cpu().numpy()

    # ------------------------------------------------------------------ #
    # checkpoints (reference pix2pix.py:158-186): gzip-pickle of
    # get_all_param_values lists
    # ------------------------------------------------------------------ #
    def save_model(self, filename):
        def vals(net):
            return net.get_all_param_values() if net is not None else []
        with gzip.open(filename, "wb") as g:
            pickle.dump({'dcgan': {'gen': vals(self.G), 'disc': vals(self.D)},
                         'p2p': {'gen': vals(self.P), 'disc': vals(self.Dp)}}, g, pickle.HIGHEST_PROTOCOL)

    def load_model(self, filename, mode='both'):
        assert mode in ['both', 'dcgan', 'p2p']
        with gzip.open(filename) as g:
            dd = pickle.load(g, encoding='latin1')      # reference checkpoints are Python-2 pickles
        for stage, ok in (('dcgan', mode in ('both', 'dcgan')), ('p2p', mode in ('both', 'p2p'))):
            for role in ('gen', 'disc'):
                net, vals = self.nets[stage][role], dd[stage][role]
                if not ok:
                    continue
                if net is None:           # a model built without this stage (gen_fn_p2p=None / gen_fn_dcgan=None)
                    if len(vals):
                        raise ValueError("checkpoint holds %s/%s parameters but this model was built without that "
                                         "network (use mode=%r)" % (stage, role, 'dcgan' if stage == 'p2p' else 'p2p'))
                    continue
                net.set_all_param_values(vals)

    # ------------------------------------------------------------------ #
    # training loop (reference pix2pix.py:187-275)
    # ------------------------------------------------------------------ #
    def train(self, it_train, it_val, batch_size, num_epochs, out_dir, model_dir=None, save_every=10,
              resume=False, quick_run=False, loss_sync_every=None):
        """loss_sync_every (not in the reference; default 1 or $HMGAN_LOSS_SYNC_EVERY): read the losses back every N steps
        instead of after every step -- the epoch means are the same numbers, the host just stops waiting for each step."""
        if loss_sync_every is None:
            loss_sync_every = int(os.environ.get("HMGAN_LOSS_SYNC_EVERY", "1"))

        def _loop(fn, itr):
            rec = [[] for _ in range(len(self.train_keys))]
            pend = []
            fn_async = {self.train_fn: self.train_fn_async, self.loss_fn: self.loss_fn_async}.get(fn)

            def flush():
                if pend:
                    for row in torch.stack(pend).cpu().numpy():
                        for i in range(len(self.train_keys)):
                            rec[i].append(np.float32(row[i]))
                    del pend[:]
            for b in range(itr.N // batch_size):
                X_batch, Y_batch = it_train.next()      # sic: the reference reads it_train here too (:204)
                Z_batch = floatX(self.sampler(X_batch.shape[0], self.latent_dim))
                if loss_sync_every > 1 and fn_async is not None:
                    pend.append(fn_async(Z_batch, X_batch, Y_batch))
                    if len(pend) >= loss_sync_every:
                        flush()
                else:
                    results = fn(Z_batch, X_batch, Y_batch)
                    for i in range(len(results)):
                        rec[i].append(results[i])
                if quick_run:
                    break
            flush()
            means = tuple([np.mean(elem) for elem in rec])
            if self.rt.precision == "fast" and not np.all(np.isfinite(means)):
                # the reference (float32) would carry on; here a non-finite loss is almost always an fp16 overflow of the
                # statically scaled gradients (e.g. a degenerate BatchNorm batch) and must not pass unnoticed
                import warnings
                warnings.warn("non-finite losses in precision='fast' (fp16 storage, static loss scale %g): the fp16 "
                              "gradients overflowed; rerun with precision='tc32' / 'parity' or a smaller loss_scale"
                              % self.rt.loss_scale, RuntimeWarning)
            return means
        header = ["epoch"] + ["train_%s" % k for k in self.train_keys] + ["valid_%s" % k for k in self.train_keys]
        header += ["lr", "time", "mode"]
        if not os.path.exists(out_dir):
            os.makedirs(out_dir)
        if model_dir is not None and not os.path.exists(model_dir):
            os.makedirs(model_dir)
        f = open("%s/results.txt" % out_dir, "w" if not resume else "a")
        if not resume:
            f.write(",".join(header) + "\n")
            f.flush()
            print(",".join(header))
        else:
            if self.verbose:
                print("loading weights from: %s" % resume)
            self.load_model(resume)
        for e in range(num_epochs):
            out_str = [str(e + 1)]
            t0 = time()
            out_str += [str(r) for r in _loop(self.train_fn, it_train)]
            out_str += [str(r) for r in _loop(self.loss_fn, it_val)]
            out_str.append(str(self.lr.get_value()))
            out_str.append(str(time() - t0))
            out_str.append(self.train_mode)
            out_str = ",".join(out_str)
            print(out_str)
            f.write("%s\n" % out_str)
            f.flush()
            if self.train_mode in ['both', 'p2p'] and self.have_p2p:
                plot_grid("%s/out_%i.png" % (out_dir, e + 1), it_val, self.gen_fn,
                          is_a_grayscale=self.is_a_grayscale, is_b_grayscale=self.is_b_grayscale)
                self.generate_atob(it_train, 1, "%s/dump_train" % out_dir, deterministic=False)
                self.generate_atob(it_val, 1, "%s/dump_valid" % out_dir, deterministic=False)
            if self.train_mode in ['both', 'dcgan'] and self.have_dcgan:
                self.generate_gz(num_examples=20, batch_size=batch_size, out_dir="%s/dump_a" % out_dir,
                                 deterministic=False)
            if model_dir is not None and (e + 1) % save_every == 0:
                self.save_model("%s/%i.model" % (model_dir, e + 1))
        f.close()

    # ------------------------------------------------------------------ #
    # sampling (reference pix2pix.py:276-425)
    # ------------------------------------------------------------------ #
    def generate_atob(self, itr, num_batches, out_dir, dont_predict=False, deterministic=True):
        fn = self.gen_fn if not deterministic else self.gen_fn_det
        if not os.path.exists(out_dir):
            os.makedirs(out_dir)
        ctr = 0
        for n in range(num_batches):
            this_x, this_y = itr.next()
            pred_y = as_float_nchw(this_y, self.is_b_grayscale) if dont_predict else fn(this_x)
            this_x = as_float_nchw(this_x, self.is_a_grayscale)
            for i in range(pred_y.shape[0]):
                imsave("%s/%i.a.png" % (out_dir, ctr), convert_to_rgb(this_x[i], is_grayscale=self.is_a_grayscale))
                imsave("%s/%i.b.png" % (out_dir, ctr), convert_to_rgb(pred_y[i], is_grayscale=self.is_b_grayscale))
                ctr += 1

    def generate_gz(self, num_examples, batch_size, out_dir, deterministic=True):
        if not os.path.exists(out_dir):
            os.makedirs(out_dir)
        fn = self.z_fn if not deterministic else self.z_fn_det
        z = floatX(self.sampler(num_examples, self.latent_dim))
        ctr = 0
        for b in range(num_examples // batch_size):
            out = fn(z[b * batch_size:(b + 1) * batch_size])
            for i in range(out.shape[0]):
                imsave("%s/%i.png" % (out_dir, ctr), convert_to_rgb(out[i], is_grayscale=self.is_a_grayscale))
                ctr += 1

    def generate_interpolation(self, out_name, zsample1=None, zsample2=None, deterministic=True, mode='row',
                               figsize=(10, 10), cmap='gray'):
        assert mode in ['row', 'matrix']
        fn = self.z_fn if not deterministic else self.z_fn_det
        if zsample1 is None:
            zsample1 = floatX(self.sampler(1, self.latent_dim)[0])
        if zsample2 is None:
            zsample2 = floatX(self.sampler(1, self.latent_dim)[0])
        coefs = [0.0, 0.1, 0.3, 0.6, 0.9, 1.0] if mode == 'row' else list(np.linspace(0, 1, 25))
        rows, cols = (1, 6) if mode == 'row' else (5, 5)
        S = self.in_shp
        canvas = np.zeros((rows * S, cols * S, 3), np.float32)
        for k, a in enumerate(coefs):
            tmp = fn(floatX((1 - a) * zsample1[np.newaxis] + a * zsample2[np.newaxis]))
            y, x = divmod(k, cols)
            canvas[y * S:(y + 1) * S, x * S:(x + 1) * S] = convert_to_rgb(tmp[0], is_grayscale=self.is_a_grayscale)
        imsave(out_name, canvas)

    def generate_interpolation_clip(self, num_samples, batch_size, out_dir, deterministic=True, min_max_norm=False,
                                    concat=False):
        if not os.path.exists(out_dir):
            os.makedirs(out_dir)
        fn = self.z_fn if not deterministic else self.z_fn_det
        fn_atob = self.gen_fn if not deterministic else self.gen_fn_det
        zs = floatX(self.sampler(num_samples, self.latent_dim))
        coefs = np.linspace(0, 1, 25).astype(zs.dtype)
        all_tps = []
        for i in range(zs.shape[0] - 1):
            for a in coefs:
                all_tps.append((1 - a) * zs[i] + a * zs[i + 1])
        all_tps = np.asarray(all_tps, dtype=zs.dtype)
        ctr = 0
        S = self.in_shp
        for b in range(all_tps.shape[0] // batch_size):
            z_out = fn(all_tps[b * batch_size:(b + 1) * batch_size])
            p2p_out = fn_atob(z_out)
            for i in range(z_out.shape[0]):
                a_img, b_img = z_out[i], p2p_out[i]
                if min_max_norm:
                    a_img = (a_img - np.min(a_img)) / (np.max(a_img) - np.min(a_img))
                a_img = convert_to_rgb(a_img, is_grayscale=self.is_a_grayscale)
                b_img = convert_to_rgb(b_img, is_grayscale=self.is_b_grayscale)
                d = '%04d' % ctr
                if concat:
                    full = np.zeros((S, S * 2, 3), dtype=zs.dtype)
                    full[:, :S] = a_img
                    full[:, S:] = b_img
                    imsave("%s/concat_%s.png" % (out_dir, d), full)
                else:
                    imsave("%s/a_%s.png" % (out_dir, d), a_img)
                    imsave("%s/b_%s.png" % (out_dir, d), b_img)
                ctr += 1
