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

GFLOP_PER_IMG = {"dcgan": 391.3, "both": 683.2}        # SURVEY.md §8d (fwd + wgrad + needed dgrad, 2 FLOP/MAC)
BATCH = {"dcgan": 32, "both": 16}                       # BASELINE.json configs[1], configs[2]


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


def build_model(workload, device, precision, pg=None, verbose=False):
    from architectures import p2p, dcgan
    from lasagne_compat import linear, tanh, rmsprop, shared, floatX
    from pix2pix import Pix2Pix
    both = workload == "both"
    return Pix2Pix(
        gen_fn_dcgan=dcgan.default_generator, disc_fn_dcgan=dcgan.default_discriminator,
        gen_params_dcgan={'num_repeats': 0, 'div': [2, 2, 4, 4, 8, 8, 8]},
        disc_params_dcgan={'num_repeats': 0, 'bn': False, 'nonlinearity': linear, 'div': [8, 4, 4, 4, 2, 2, 2]},
        gen_fn_p2p=p2p.g_unet if both else None, disc_fn_p2p=p2p.discriminator if both else None,
        gen_params_p2p={'nf': 64, 'act': tanh, 'num_repeats': 0, 'bilinear_upsample': True},
        disc_params_p2p={'nf': 64, 'bn': False, 'num_repeats': 0, 'act': linear, 'mul_factor': [1, 2, 4, 8]},
        in_shp=512, latent_dim=1000, is_a_grayscale=True, is_b_grayscale=False, lsgan=True, opt=rmsprop,
        opt_args={'learning_rate': shared(floatX(1e-4))}, train_mode='both' if both else 'dcgan', verbose=verbose,
        device=device, precision=precision, seed=2, process_group=pg)


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


def cpu_baseline(workload, sample_b, steps=1, warmup=1):
    """The oracle's train step (torch-CPU float32, all host threads) on `sample_b` images of the workload."""
    from oracle import step as S
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = S.experiment_kwargs('test1_nobn_bilin_both')
    which = ('G', 'D', 'P', 'Dp') if workload == "both" else ('G', 'D')
    om = S.OracleModel(S.build_nets(cfg, seed=2, which=which), opt='rmsprop', lr=1e-4,
                       train_mode='both' if workload == "both" else 'dcgan', lsgan=True)
    Z, X, Y = S.synthetic_batch(sample_b, cfg['latent_dim'], 512, seed=0)
    for _ in range(warmup):
        om.train_fn(Z, X, Y)
    t0 = time.time()
    for _ in range(steps):
        om.train_fn(Z, X, Y)
    dt = (time.time() - t0) / steps
    return sample_b / dt, dt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="dcgan", choices=["dcgan", "both"])
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: BASELINE.json's)")
    ap.add_argument("--precision", default="fast", choices=["fast", "parity"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=2)
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    B = a.batch or BATCH[a.workload]
    wl_name = ("DCGAN 512x512 heightmap train step, batch %d/GPU, z=1000 (BASELINE configs[1])" % B
               if a.workload == "dcgan" else
               "DCGAN+pix2pix joint train step, 512x512, batch %d/GPU, L1+LSGAN (BASELINE configs[2])" % B)
    config = {"workload": wl_name, "per_gpu_batch": B, "global_batch": B * world, "parallelism": "dp%d" % world,
              "l2": "per-step working set (GBs of activations) exceeds the 126 MB L2; no explicit flush",
              "precision": a.precision}
    base = {"metric": "512px heightmap+texture images/sec/GPU at 1/2/4/8 B200; tensor-pipe %",
            "unit": "images/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "data": "synthetic", "config": config}

    if a.impl == "reference":
        if rank != 0:
            return
        v, dt = cpu_baseline(a.workload, a.cpu_sample, steps=max(1, min(a.steps, 3)), warmup=min(a.warmup, 1))
        cores = os.cpu_count() or 1
        sample = "%d images/step of the same 512x512 workload (full step: fwd+bwd+update)" % a.cpu_sample
        out = dict(base, impl="reference", value=v, ms_per_step=dt * 1e3, dtype="f32",
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
    m = build_model(a.workload, "cuda:%d" % local, a.precision, pg)
    from util import synthetic_batch
    Z, X, Y = synthetic_batch(B, 1000, 512, seed=100 + rank)
    Zd, Xd, Yd = (torch.from_numpy(t).cuda() for t in (Z, X, Y))
    Zp, Xp, Yp = (torch.from_numpy(t).pin_memory() for t in (Z, X, Y))

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        sync()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # W untimed warm-up steps; never fewer than 3, so that the two buffer-sizing eager steps and the CUDA-graph capture of
    # the third call are outside the timed region whatever --warmup says
    for _ in range(max(a.warmup, 3)):
        m.step_device(Zd, Xd, Yd, True)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    l0 = m.rt.launches
    ms = timed(lambda: m.step_device(Zd, Xd, Yd, True), a.steps)
    launches = m.rt.launches - l0
    clk = clocks.stop() if rank == 0 else None
    losses = m.losses.cpu().numpy()
    value = B * world * a.steps / (ms * 1e-3)
    # end to end through the public API: pinned host inputs, losses read back every step
    for _ in range(3):                    # the host path captures its own pair of graphs on its third call
        m.train_fn(Zp, Xp, Yp)
    ms_e2e = timed(lambda: m.train_fn(Zp, Xp, Yp), a.steps)
    e2e = B * world * a.steps / (ms_e2e * 1e-3)
    if rank != 0:
        _finish(world)
        return
    pk = peaks()
    kname, kflop, kms, kpath = time_dominant_kernel(m)
    traffic, ncu_tensor = None, None
    try:       # DRAM bytes per launch / tensor-pipe activity of that kernel from the committed `ncu --set full` capture
        with open(os.path.join(ROOT, "profiles", "r1_ncu_traffic.json")) as f:
            rec = json.load(f).get(kname, {})
        traffic, ncu_tensor = rec.get("dram_bytes_per_launch"), rec.get("tensor_pipe_active_pct")
    except Exception:
        pass
    k_tflops = kflop / (kms * 1e-3) / 1e12
    step_tflops = value / world * GFLOP_PER_IMG[a.workload] / 1e3
    out = dict(base, value=value, ms_per_step=ms / a.steps, dtype="f16" if a.precision == "fast" else "f32",
               e2e={"value": e2e, "unit": "images/s",
                    # a DCGAN-only model never reads the texture batch Y, so train_fn does not copy it
                    "h2d_bytes_per_step": int(Z.nbytes + X.nbytes + (Y.nbytes if a.workload == "both" else 0)),
                    "d2h_bytes_per_step": 20, "ms_per_step": ms_e2e / a.steps},
               gpu_launches=int(launches), clocks=clk,
               roofline={"bound": "tensor", "achieved": k_tflops, "peak": pk["burst"], "unit": "TFLOP/s",
                         "frac": k_tflops / pk["burst"], "traffic": traffic, "kernel": kname, "kernel_path": kpath,
                         "kernel_ms": kms, "peak_source": pk["src"] + " (burst: kernel timed alone)",
                         "ncu_tensor_pipe_active_pct": ncu_tensor,
                         "step_achieved": step_tflops, "step_frac_of_sustained": step_tflops / pk["sustained"]},
               losses=[float(v) for v in losses])
    if not a.no_cpu_baseline and world == 1:        # the CPU baseline is reported by the single-GPU run only
        v, dt = cpu_baseline(a.workload, a.cpu_sample)
        out["cpu_baseline"] = {"value": v, "unit": "images/s", "cores": os.cpu_count() or 1, "kind": "port",
                               "sample": "%d images of the same workload, 1 warm-up + 1 timed full step "
                                         "(%.1f s/step)" % (a.cpu_sample, dt)}
    print(json.dumps(out))
    _finish(world)


def _finish(world):
    """Multi-rank runs leave without tearing the NCCL communicator down: destroy_process_group() was observed to hang
    here (the captured CUDA graphs hold collectives of that communicator), and a rank that lingers would stall the
    launcher.  Everything this process had to say is flushed first."""
    if world > 1:
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
