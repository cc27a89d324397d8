#!/usr/bin/env python
"""Benchmark of the hot path: one full training step (forward of G and D on real+fake,
all backward passes, simultaneous RMSprop update) of the 512x512 DCGAN heightmap model
of test1_nobn_bilin_both (reference experiments.py:98-119, pix2pix.py:142) on synthetic
crops -- BASELINE.json configs[1] (batch 32 per GPU, fp16 storage / fp32 accumulate).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload dcgan|both]

Prints ONE JSON line (rank 0).  `value` = images/s of the whole job with inputs resident
in HBM; `e2e` = the same through Pix2Pix.train_fn with pinned HOST inputs (H2D of Z,X,Y and
D2H of the five losses inside the timed region).  `--impl reference` times the CPU
restatement of the reference's step (oracle/, torch-CPU float32: the reference itself is
Python-2/Theano and cannot run here) on a bounded sample of the workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "gan-heightmaps_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np
import torch

GFLOP_PER_IMG = {"dcgan": 391.3, "both": 683.2, "p2p": 292.0}   # SURVEY.md §8d (fwd + wgrad + needed dgrad, 2 FLOP/MAC)
BATCH = {"dcgan": 32, "both": 16, "p2p": 16}            # BASELINE.json configs[1], configs[2], configs[4] (128 / 8 GPUs)
WORKLOAD_NAME = {
    "dcgan": "DCGAN 512x512 heightmap train step, batch %d/GPU, z=1000 (BASELINE configs[1])",
    "both": "DCGAN+pix2pix joint train step, 512x512, batch %d/GPU, L1+LSGAN (BASELINE configs[2])",
    "p2p": "pix2pix U-Net 512x512 -> 512x512 RGB + PatchGAN train step, batch %d/GPU (BASELINE configs[4])",
}


def _peak_value(p, key):
    """A number stored under `key`, plainly or as {"value": ...}; values quoted in PFLOP/s or TB/s are rescaled."""
    v = p[key]
    if isinstance(v, dict):
        v = v.get("value", v.get("median", v.get("max")))
    v = float(v)
    if v <= 0:
        raise ValueError(key)
    return v * 1000.0 if v < 20.0 else v


def peaks():
    """Roofline denominators: MEASURED_PEAKS.json (driver-written; B200_PROFILING.md) or the recipe's stated fallback."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        burst = _peak_value(p, "bf16_tflops")
        try:
            sustained = _peak_value(p, "bf16_tflops_sustained")
        except Exception:
            sustained = burst
        try:
            hbm = _peak_value(p, "hbm_gbs")
        except Exception:
            hbm = 6650.0
        return dict(burst=burst, sustained=sustained, hbm=hbm, src="measured")
    except Exception:
        return dict(burst=1590.0, sustained=1400.0, hbm=6650.0, src="fallback")


class ClockSampler(object):
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.th.join(timeout=2)
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        hi = sorted(sm)[len(sm) // 2:] if sm else []          # upper half = samples under load
        return {"sm_mhz": float(np.median(hi)) if hi else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def build_model(workload, device, precision, pg=None, verbose=False, lr=1e-4):
    from architectures import p2p, dcgan
    from lasagne_compat import linear, tanh, rmsprop, shared, floatX
    from pix2pix import Pix2Pix
    with_p2p, with_dcgan = workload in ("both", "p2p"), workload in ("both", "dcgan")
    return Pix2Pix(
        gen_fn_dcgan=dcgan.default_generator if with_dcgan else None,
        disc_fn_dcgan=dcgan.default_discriminator if with_dcgan else None,
        gen_params_dcgan={'num_repeats': 0, 'div': [2, 2, 4, 4, 8, 8, 8]},
        disc_params_dcgan={'num_repeats': 0, 'bn': False, 'nonlinearity': linear, 'div': [8, 4, 4, 4, 2, 2, 2]},
        gen_fn_p2p=p2p.g_unet if with_p2p else None, disc_fn_p2p=p2p.discriminator if with_p2p else None,
        gen_params_p2p={'nf': 64, 'act': tanh, 'num_repeats': 0, 'bilinear_upsample': True},
        disc_params_p2p={'nf': 64, 'bn': False, 'num_repeats': 0, 'act': linear, 'mul_factor': [1, 2, 4, 8]},
        in_shp=512, latent_dim=1000, is_a_grayscale=True, is_b_grayscale=False, lsgan=True, opt=rmsprop,
        opt_args={'learning_rate': shared(floatX(lr))}, train_mode=workload, verbose=verbose,
        device=device, precision=precision, seed=2, process_group=pg)


def liven_head(m, value):
    """At the Glorot initialisation the ReLU head of the DCGAN discriminator (reference architectures/dcgan.py:50) is
    dead on synthetic inputs: D(.) = 0 for every sample, both losses are exactly 1 and EVERY gradient tensor of the step
    is exactly zero.  The kernels run the same launches either way, but multiplying zeros draws less power, so the
    clocks -- and the measured rate -- would be those of an idle-data step, not of training.  A head bias of `value`
    puts D(.) near it, as it is after the first few real updates; the step then moves dense gradients."""
    if m.D is None or not value:
        return
    vals = m.D.get_all_param_values()
    vals[-1][:] = value
    m.D.set_all_param_values(vals)


def time_dominant_kernel(m, reps=5):
    """The convolution with the most FLOPs per launch (D's 64->128 5x5 @256^2 over the 2B real+fake batch),
    timed alone with CUDA events on the launching stream after the step has left its inputs in place."""
    import engine
    best, bf = None, 0
    for net in (m.G, m.D, m.P, m.Dp):
        if net is None:
            continue
        for op in net.ops:
            if isinstance(op, engine.ConvOp) and op.kind == "conv":
                fl = 2.0 * net.B * op.out.shape[0] * op.out.shape[1] * op.K * op.Cout
                if fl > bf:
                    best, bf, bnet = op, fl, net
    op, net = best, bnet
    op.fwd(m.rt, 0, net.B, False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        op.fwd(m.rt, 0, net.B, False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    name = "%s %dx%d %d->%d @%dx%d x%d" % (net.name, op.kh, op.kw, op.Cin, op.Cout, op.out.shape[0],
                                            op.out.shape[1], net.B)
    return name, bf, ms, getattr(op, "path", "simt")


def cpu_baseline(workload, sample_b, steps=1, warmup=1, lr=1e-4):
    """The oracle's train step (torch-CPU float32, all host threads) on `sample_b` images of the workload."""
    from oracle import step as S
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = S.experiment_kwargs('test1_nobn_bilin_both')
    which = {"both": ('G', 'D', 'P', 'Dp'), "dcgan": ('G', 'D'), "p2p": ('P', 'Dp')}[workload]
    om = S.OracleModel(S.build_nets(cfg, seed=2, which=which), opt='rmsprop', lr=lr, train_mode=workload, lsgan=True)
    Z, X, Y = S.synthetic_batch(sample_b, cfg['latent_dim'], 512, seed=0)
    for _ in range(warmup):
        om.train_fn(Z, X, Y)
    t0 = time.time()
    for _ in range(steps):
        om.train_fn(Z, X, Y)
    dt = (time.time() - t0) / steps
    return sample_b / dt, dt


def measure(workload, B, steps, warmup, precision, local, rank, world, pg, sample_clocks, head_bias=0.6, lr=1e-4):
    """Build the model of `workload`, run W warm-up + K timed steps device-resident and again end to end from pinned
    host memory.  Returns (result dict, model)."""
    import torch.distributed as dist
    m = build_model(workload, "cuda:%d" % local, precision, pg, lr=lr)
    liven_head(m, head_bias)
    from util import synthetic_batch
    Z, X, Y = synthetic_batch(B, 1000, 512, seed=100 + rank)
    Zd, Xd, Yd = (torch.from_numpy(t).cuda() for t in (Z, X, Y))
    Zp, Xp, Yp = (torch.from_numpy(t).pin_memory() for t in (Z, X, Y))

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        sync()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # W untimed warm-up steps; never fewer than 3, so that the two buffer-sizing eager steps and the CUDA-graph capture of
    # the third call are outside the timed region whatever --warmup says
    for _ in range(max(warmup, 3)):
        m.step_device(Zd, Xd, Yd, True)
    clocks = ClockSampler(local) if (sample_clocks and rank == 0) else None
    if clocks:
        clocks.start()
    l0 = m.rt.launches
    ms = timed(lambda: m.step_device(Zd, Xd, Yd, True), steps)
    launches = m.rt.launches - l0
    clk = clocks.stop() if clocks else None
    losses = m.losses.cpu().numpy()
    # share of the 2B discriminator outputs of the last timed step that are on the live side of the ReLU head: training on
    # random data can kill the head again (D(.) -> 0), and a step that multiplies zeros is flattered
    alive = float((m.D.out.buf[:2 * B].float() > 0).float().mean().item()) if m.D is not None else None
    # end to end through the public API: pinned host inputs, losses read back every step
    for _ in range(3):                    # the host path captures its own graphs on its third call
        m.train_fn(Zp, Xp, Yp)
    ms_e2e = timed(lambda: m.train_fn(Zp, Xp, Yp), steps)
    # data-parallel replicas must hold identical parameters after the same all-reduced updates
    identical = True
    if world > 1:
        for net in m._nets():
            cs = net.pflat.double().sum().reshape(1)
            lo, hi = cs.clone(), cs.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            identical = identical and bool((lo == hi).item())
    with_y = workload in ("both", "p2p")
    with_z = workload in ("both", "dcgan")
    pk = peaks()
    value = B * world * steps / (ms * 1e-3)
    step_tflops = value / world * GFLOP_PER_IMG[workload] / 1e3
    res = dict(value=value, ms_per_step=ms / steps,
               e2e={"value": B * world * steps / (ms_e2e * 1e-3), "unit": "images/s",
                    # only the tensors a model reads are copied (a DCGAN-only model never reads Y, a pix2pix-only one Z)
                    "h2d_bytes_per_step": int((Z.nbytes if with_z else 0) + X.nbytes + (Y.nbytes if with_y else 0)),
                    "d2h_bytes_per_step": 20, "ms_per_step": ms_e2e / steps},
               gpu_launches=int(launches), clocks=clk, losses=[float(v) for v in losses], head_alive_frac=alive,
               step_achieved_tflops=step_tflops, step_frac_of_burst=step_tflops / pk["burst"],
               step_frac_of_sustained=step_tflops / pk["sustained"], replicas_identical=identical)
    return res, m


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="dcgan", choices=["dcgan", "both", "p2p"])
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: BASELINE.json's)")
    ap.add_argument("--precision", default="fast", choices=["fast", "parity", "tc32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the configs[2] / configs[4] lines")
    ap.add_argument("--cpu-sample", type=int, default=4)
    ap.add_argument("--lr", type=float, default=1e-8,
                    help="RMSprop learning rate of the timed steps.  The default keeps the model at its initialisation: at the "
                         "experiment's 1e-4 the DCGAN discriminator's ReLU head dies within ~20 steps on random data (D(.) = 0, "
                         "all-zero gradients) and the step is then timed on zeros.  The optimiser kernels do the same work.")
    ap.add_argument("--head-bias", type=float, default=0.6,
                    help="bias of the DCGAN discriminator's head (0 = leave the Glorot init: dead head, all-zero gradients)")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    B = a.batch or BATCH[a.workload]
    config = {"workload": WORKLOAD_NAME[a.workload] % B, "per_gpu_batch": B, "global_batch": B * world,
              "parallelism": "dp%d" % world,
              "l2": "per-step working set (GBs of activations) exceeds the 126 MB L2; no explicit flush",
              "precision": a.precision,
              "init": "Glorot-uniform (seed 2); DCGAN discriminator head bias %g so that its ReLU head is alive and the "
                      "step moves dense gradients" % a.head_bias if a.head_bias else "Glorot-uniform (seed 2)",
              "lr": a.lr}
    base = {"metric": "512px heightmap+texture images/sec/GPU at 1/2/4/8 B200; tensor-pipe %",
            "unit": "images/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "data": "synthetic", "config": config}

    if a.impl == "reference":
        # The reference's own CPU path cannot run here (Python 2 + Theano/Lasagne): its restatement
        # (oracle/, torch-CPU float32, every host thread) is timed on a BOUNDED sample of the same workload -- a few
        # full train steps on `--cpu-sample` images -- and the line says what it ran: steps, warm-up and batch are the
        # ones executed, not the GPU arm's.
        if rank != 0:
            return
        k, w = max(1, min(a.steps, 4)), min(a.warmup, 1)
        v, dt = cpu_baseline(a.workload, a.cpu_sample, steps=k, warmup=w, lr=a.lr)
        cores = os.cpu_count() or 1
        sample = ("%d timed + %d warm-up full train steps (fwd+bwd+update) on %d images of the same 512x512 workload; "
                  "the GPU arm's step has %d images" % (k, w, a.cpu_sample, B))
        cfg_ref = dict(config, per_gpu_batch=a.cpu_sample, global_batch=a.cpu_sample, parallelism="cpu",
                       precision="float32 (oracle port)", gpu_arm_per_gpu_batch=B)
        out = dict(base, config=cfg_ref, steps=k, warmup=w, n_gpus=world, impl="reference", value=v,
                   ms_per_step=dt * 1e3, dtype="f32",
                   cpu_baseline={"value": v, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
                   e2e={"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                   gpu_launches=0)
        print(json.dumps(out))
        return

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local)
    pg = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        pg = dist.group.WORLD
    def alive_frac(r):               # the least alive replica decides (every rank must take the same branch)
        v = r["head_alive_frac"]
        if v is None:
            return 1.0
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([v], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            v = float(t.item())
        return v

    res, m = measure(a.workload, B, a.steps, a.warmup, a.precision, local, rank, world, pg, True, a.head_bias, a.lr)
    remeasured = None
    first = alive_frac(res)
    if a.head_bias and first < 0.5:
        # the discriminator's ReLU head died while training on random data: most gradients of that run were zeros and its
        # time is that of an idle-data step.  Measure once more from a fresh model; the line reports the second run.
        remeasured = {"why": "discriminator head died during the first run (share of live outputs %.2f)" % first,
                      "first_run_ms_per_step": res["ms_per_step"]}
        del m
        import gc
        gc.collect()
        torch.cuda.empty_cache()
        res, m = measure(a.workload, B, a.steps, a.warmup, a.precision, local, rank, world, pg, True, a.head_bias, a.lr)
        alive_frac(res)
    out = None
    if rank == 0:
        pk = peaks()
        kname, kflop, kms, kpath = time_dominant_kernel(m)
        traffic, ncu_tensor = None, None
        try:       # DRAM bytes per launch / tensor-pipe activity of that kernel from the committed `ncu --set full` capture
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                rec = json.load(f).get(kname, {})
            traffic, ncu_tensor = rec.get("dram_bytes_per_launch"), rec.get("tensor_pipe_active_pct")
        except Exception:
            pass
        k_tflops = kflop / (kms * 1e-3) / 1e12
        out = dict(base, value=res["value"], ms_per_step=res["ms_per_step"],
                   dtype={"fast": "f16", "tc32": "bf16x3"}.get(a.precision, "f32"), e2e=res["e2e"],
                   gpu_launches=res["gpu_launches"], clocks=res["clocks"],
                   roofline={"bound": "tensor", "achieved": k_tflops, "peak": pk["burst"], "unit": "TFLOP/s",
                             "frac": k_tflops / pk["burst"], "traffic": traffic, "kernel": kname, "kernel_path": kpath,
                             "kernel_ms": kms, "peak_source": pk["src"] + " (burst: kernel timed alone)",
                             "ncu_tensor_pipe_active_pct": ncu_tensor,
                             # the timed region is a fraction of a second at full clocks: the burst peak is the honest
                             # denominator of the whole step, the sustained one is quoted beside it
                             "step_achieved": res["step_achieved_tflops"],
                             "step_frac_of_burst": res["step_frac_of_burst"],
                             "step_frac_of_sustained": res["step_frac_of_sustained"]},
                   losses=res["losses"], head_alive_frac=res["head_alive_frac"])
        if remeasured:
            out["remeasured"] = remeasured
        if world > 1:
            out["replicas_identical"] = res["replicas_identical"]
    del m
    torch.cuda.empty_cache()
    # the other single-GPU configurations BASELINE.json names, each as its own full measurement (own model, own
    # warm-up, own timed steps): configs[2] joint step and configs[4] pix2pix-only step at its per-GPU batch
    ok = res["replicas_identical"]
    if not a.no_secondary and a.workload == "dcgan" and not a.batch:
        sec = {}
        for wl in ("both", "p2p"):
            r2, m2 = measure(wl, BATCH[wl], min(a.steps, 20), a.warmup, a.precision, local, rank, world, pg, False, a.head_bias, a.lr)
            del m2
            torch.cuda.empty_cache()
            ok = ok and r2["replicas_identical"]
            r2.pop("clocks")
            sec[wl] = dict(r2, workload=WORKLOAD_NAME[wl] % BATCH[wl], per_gpu_batch=BATCH[wl], steps=min(a.steps, 20),
                           gflop_per_image=GFLOP_PER_IMG[wl])
        if out is not None:
            out["secondary"] = sec
    if rank == 0:
        if not a.no_cpu_baseline and world == 1:        # the CPU baseline is reported by the single-GPU run only
            v, dt = cpu_baseline(a.workload, a.cpu_sample, lr=a.lr)
            out["cpu_baseline"] = {"value": v, "unit": "images/s", "cores": os.cpu_count() or 1, "kind": "port",
                                   "sample": "%d images of the same workload, 1 warm-up + 1 timed full step "
                                             "(%.1f s/step)" % (a.cpu_sample, dt)}
        print(json.dumps(out))
    _finish(world, ok)


def _finish(world, ok=True):
    """Multi-rank runs leave WITHOUT tearing the NCCL communicator down: dist.destroy_process_group() hangs in this image
    once CUDA graphs holding collectives of the communicator have been replayed (observed again in round 2 on 2 x B200
    with the models and graphs deleted and the device synchronised first: the JSON line was out, the call never
    returned, the launcher had to be killed).  Everything this process had to say is flushed; the exit status carries
    the replica check.  HMGAN_BENCH_DESTROY_PG=1 tries the clean teardown."""
    sys.stdout.flush()
    sys.stderr.flush()
    if not ok:
        sys.stderr.write("bench.py: data-parallel replicas diverged (parameter checksums differ across ranks)\n")
        sys.stderr.flush()
    if world > 1:
        torch.cuda.synchronize()
        if os.environ.get("HMGAN_BENCH_DESTROY_PG", "0") == "1":
            import torch.distributed as dist
            dist.destroy_process_group()
            sys.exit(0 if ok else 1)
        os._exit(0 if ok else 1)
    if not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
