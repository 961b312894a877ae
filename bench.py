#!/usr/bin/env python
"""bench.py - throughput of the B200 audio front-end hot path (contract: see DESIGN.md section 7).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c1|c2|c3] [--impl ours|reference]

One "step" = one pass of the hot path over one batch of synthetic PCG-like clips.
Prints ONE JSON line (rank 0).  `value` is device-resident throughput, `e2e` goes through
the public host-buffer API (pinned host -> device -> features -> host, copies timed),
`roofline` is the dominant kernel's algorithmic bytes / its CUDA-event time, `cpu_baseline`
is the oracle port on the host cores over a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SR = 16000

WORKLOADS = {
    # name -> (BASELINE.json config, description)
    "c1": "CirCor-shaped synthetic PCG: 1000 clips x 8 s @16 kHz -> librosa-style 64-mel log-spectrogram [251,64] "
          "(src/util.py:481-501)",
}


# ----------------------------------------------------------------------------- CPU reference arm


_BLAS_LIMIT = None


def _cpu_worker_init():
    """One BLAS thread per worker process: the pool already uses every core."""
    global _BLAS_LIMIT
    try:
        from threadpoolctl import threadpool_limits

        _BLAS_LIMIT = threadpool_limits(1)
    except Exception:  # pragma: no cover
        pass
    import torch

    torch.set_num_threads(1)


def _cpu_logmel_one(x):
    from oracle import frontend as F

    return F.log_mel(x, f_max=8000).shape[0]


def cpu_reference_c1(n_clips: int, repeats: int = 1):
    """Oracle port (numpy restatement of the reference's librosa path), one clip per task over
    all host cores - mirrors the reference's one-file-at-a-time loop (model_util.py:138)."""
    import multiprocessing as mp

    import torch

    from heart_murmur_detection_b200 import synth

    cores = os.cpu_count() or 1
    clips = [synth.make_clip(8 * SR, 1000 + i).numpy() for i in range(n_clips)]
    ctx = mp.get_context("fork")
    torch.set_num_threads(1)
    best = None
    with ctx.Pool(cores, initializer=_cpu_worker_init) as pool:
        pool.map(_cpu_logmel_one, clips[: 2 * cores])  # warm-up (imports, page-in)
        for _ in range(repeats):
            t0 = time.perf_counter()
            frames = sum(pool.map(_cpu_logmel_one, clips, chunksize=1))
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    return {"clips_per_s": n_clips / best, "frames_per_s": frames / best, "cores": cores, "seconds": best,
            "sample": f"{n_clips} clips x 8 s, one clip per task, multiprocessing.Pool({cores})"}


# ----------------------------------------------------------------------------- clocks


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------- main


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def run_reference(args, rank):
    if rank != 0:
        return
    n = 256
    r = cpu_reference_c1(n, repeats=max(1, args.steps))
    line = {
        "impl": "reference", "metric": "log-mel clips/s", "value": r["clips_per_s"], "unit": "clips/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["seconds"] * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64 FFT / f32 mel (numpy)",
        "data": "synthetic", "config": {"workload": WORKLOADS["c1"], "sample_clips": n},
        "frames_per_s": r["frames_per_s"],
        "cpu_baseline": {"value": r["clips_per_s"], "unit": "clips/s", "cores": r["cores"], "kind": "port",
                         "sample": r["sample"]},
        "e2e": {"value": r["clips_per_s"], "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="c1", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--variant", default="auto", choices=["auto", "scalar", "packed"])
    ap.add_argument("--clips", type=int, default=1000, help="clips per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    args.warmup = max(args.warmup, 3)

    if args.impl == "reference":
        run_reference(args, rank)
        return

    # CPU baseline first (rank 0, N=1 only), before CUDA is initialised in this process (fork safety)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_c1(256, repeats=2)
        cpu = {"value": r["clips_per_s"], "unit": "clips/s", "cores": r["cores"], "kind": "port", "sample": r["sample"],
               "frames_per_s": r["frames_per_s"]}

    import torch
    import torch.distributed as dist

    from heart_murmur_detection_b200 import frontend, synth

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    n_clips = args.clips
    lens = synth.clip_lengths("c1", n_clips)
    wav, off = synth.make_batch(lens, base_seed=10_000 * rank, device=dev)
    plan = frontend.LogMelPlan(f_max=8000, variant=args.variant)
    fo = plan.frame_offsets(off)
    n_frames = int(fo[-1])
    out = torch.empty((n_frames, 64), dtype=torch.float32, device=dev)
    gathered = torch.empty((world * n_frames, 64), dtype=torch.float32, device=dev) if world > 1 else None

    def step():
        plan(wav, off, out=out)
        if world > 1:  # the path's single collective: all-gather of the feature tensors
            dist.all_gather_into_tensor(gathered, out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    plan.set_profile(True)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    power_ms, fin_ms, n_calls = plan.profile_ms()
    plan.set_profile(False)
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps
    clips_per_s = world * n_clips / (ms_step * 1e-3)

    # ---- end to end through the host-buffer API (pinned host in, pinned host out, copies timed)
    h_wav = torch.empty(wav.numel(), dtype=torch.float32, pin_memory=True)
    h_wav.copy_(wav)
    h_out = torch.empty((n_frames, 64), dtype=torch.float32, pin_memory=True)
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        frontend.logmel_from_host(plan, h_wav, off, h_out)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        frontend.logmel_from_host(plan, h_wav, off, h_out)
        if world > 1:
            dist.all_gather_into_tensor(gathered, out)
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())

    if rank == 0:
        hbm_peak, peak_src = peaks()
        bytes_per_clip = 128000 * 4 + 251 * 64 * 4  # algorithmic: f32 samples in + f32 [251,64] out
        kern_s = power_ms * 1e-3 / max(1, n_calls)
        achieved = n_clips * bytes_per_clip / kern_s / 1e9
        line = {
            "metric": "log-mel clips/s", "value": clips_per_s, "unit": "clips/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOADS[args.workload], "clips_per_gpu_per_step": n_clips,
                       "variant": args.variant, "l2_policy": "inputs larger than L2 (512 MB of samples per step)",
                       "collective": "all_gather_into_tensor(features)" if world > 1 else "none"},
            "frames_per_s": clips_per_s * 251,
            "e2e": {"value": world * n_clips / e2e_s, "unit": "clips/s", "h2d_bytes_per_step": int(wav.numel() * 4),
                    "d2h_bytes_per_step": int(n_frames * 64 * 4), "ms_per_step": e2e_s * 1e3},
            "gpu_launches": plan.last_launches * args.steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": None, "peak_source": peak_src,
                         "kernel": "logmel_power_kernel (framing+window+rFFT-1024+power+mel)",
                         "kernel_ms": kern_s * 1e3, "finalize_ms": fin_ms / max(1, n_calls),
                         "fp32_tflops_algorithmic": n_clips * 251 * 28.8e3 / kern_s / 1e12},
            "cpu_baseline": cpu,
            "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
