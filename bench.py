#!/usr/bin/env python
"""bench.py - throughput of the B200 audio front-end hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c1|c3] [--impl ours|reference]

One "step" = one pass of the hot path over one batch of synthetic PCG-like clips (per GPU).
Workloads (BASELINE.json configs):
  c2 (default)  OPERA-CT linear-probe front-end: order-5 band-pass 200-1800 Hz + silence trim +
                zero/tile pad to >= 8 s + cut at 32 s + 64-mel log-spectrogram, 5 272 ragged clips
                of 2-80 s (model_util.py:161-163 with the band-pass enabled).
  c1            CirCor-shaped log-mel only: 1000 clips x 8 s -> [251, 64] (src/util.py:481-501).
  c3            Audio-MAE front-end: kaldi fbank 128 mel, 1000 clips x 10.24 s -> [1024, 128].
Prints ONE JSON line (rank 0).  `value` is device-resident throughput (CUDA events, max over
ranks); `e2e` goes through the host-buffer API (pinned host in, features out, copies timed);
`roofline` is the dominant kernel's algorithmic bytes / its mean CUDA-event duration inside
the timed region; `cpu_baseline` is the oracle port on all host cores over a bounded sample.
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
    "c1": "CirCor-shaped synthetic PCG: 1000 clips x 8 s @16 kHz -> 64-mel log-spectrogram [251,64] "
          "(src/util.py:481-501)",
    "c2": "OPERA-CT linear-probe front-end: order-5 Butterworth band-pass 200-1800 Hz + silence trim + zero/tile pad "
          "to >=8 s + cut at 32 s + 64-mel log-spectrogram over 5272 ragged clips, log-uniform 2-80 s "
          "(model_util.py:161-163, src/util.py:205-267 with butterworth_filter=5)",
    "c3": "Audio-MAE front-end: kaldi fbank 128 mel, 25 ms / 10 ms, 1000 clips x 10.24 s padded to [1024,128] "
          "(src/util.py:845-856, audioMAE/models_mae.py:1178-1181)",
}
WORKLOADS["c2nf"] = ("OPERA-CT linear-probe front-end exactly as the reference's callers invoke it (no band-pass, SURVEY F4): "
                     "silence trim + zero/tile pad to >=8 s + cut at 32 s + 64-mel log-spectrogram over 5272 ragged clips "
                     "(model_util.py:161-163)")
WORKLOADS["c2raw"] = ("c2 from the file payload: 16-bit PCM at the recordings' native 4 kHz in host memory -> PCM16 decode + "
                      "rate conversion to 16 kHz (the librosa.load(path, sr=16000) of src/util.py:222; torchaudio.transforms."
                      "Resample algorithm, the pinned oracle) -> the c2 path")
WORKLOADS["c4"] = ("COLA continued-pretraining input: 100000 synthetic multi-site recordings of 8-60 s -> silence trim + "
                   "whole-recording 64-mel log-spectrogram (heart_pressl.py:58-99) -> AudioDataset items: random_mask, two "
                   "random_crops of 251 frames, two random_multiplies (cola_training.py:56-80), batch 64")
WORKLOADS["c5"] = ("Throughput sweep: 1000000 CirCor-shaped clips (8 s @16 kHz -> [251,64]) sharded over the GPUs in chunks "
                   "dealt round-robin, one in-place NCCL all-gather of the features per round, every rank ends with the "
                   "[1000000,251,64] tensor in global clip order")
DEFAULT_CLIPS = {"c1": 1000, "c2": 5272, "c3": 1000, "c2nf": 5272, "c4": 100000, "c5": 1000000}
CPU_SAMPLE = {"c1": 1000, "c2": 768, "c3": 1000, "c2nf": 1024, "c2raw": 768}
LENS_SEED = 1234          # clip lengths of rank r: synth.clip_lengths(seed=LENS_SEED + r)
SEED_STRIDE = 10_000_000  # clip i of rank r: synth.make_clip(seed = SEED_STRIDE * r + i)
C2_KW = dict(input_sec=8, butterworth_filter=5, pad=True, types="zero", max_sec=32)
C2NF_KW = dict(input_sec=8, butterworth_filter=None, pad=True, types="zero", max_sec=32)

# ----------------------------------------------------------------------------- CPU reference arm

NATIVE_SR = 4000  # c2raw: CirCor's native rate; host payload = 16-bit PCM at this rate


def _to_native_pcm(x16k: np.ndarray) -> np.ndarray:
    """The c2raw host payload of a synthetic 16 kHz clip: every 4th sample, quantised to 16-bit PCM."""
    return np.clip(np.round(x16k[:: SR // NATIVE_SR] * 32768.0), -32768, 32767).astype(np.int16)


def _cpu_process(workload, x, F):
    if workload == "c2raw":  # x: int16 PCM at 4 kHz -> librosa.load's decode + rate conversion -> the c2 path
        x16 = F.resample_torchaudio(x.astype(np.float32) / np.float32(32768.0), NATIVE_SR, SR)
        out = F.entire_signal(x16, spectrogram=True, **C2_KW)
        return 0 if out is None else out.shape[0]
    if workload == "c1":
        return F.log_mel(x, f_max=8000).shape[0]
    if workload in ("c2", "c2nf"):
        out = F.entire_signal(x, spectrogram=True, **(C2_KW if workload == "c2" else C2NF_KW))
        return 0 if out is None else out.shape[0]
    fb = F.kaldi_fbank_chunk(x)
    return 0 if fb is None else int(F.pad_to_model(fb.numpy()).shape[0])


def _cpu_worker(rank, cores, workload, lens, seed_base, repeats, barrier, queue, shared, offs, pregenerated=False):
    """One host core: generate its share of the sample (untimed), then run the oracle port clip by clip."""
    try:
        from threadpoolctl import threadpool_limits

        _limit = threadpool_limits(1)  # noqa: F841  one BLAS thread per worker: the workers already use every core
    except Exception:  # pragma: no cover
        pass
    import torch

    torch.set_num_threads(1)
    from heart_murmur_detection_b200 import synth
    from oracle import frontend as F

    mine = []
    for i in range(rank, len(lens), cores):
        if pregenerated:  # clips made by an earlier call (inherited shared memory)
            x = np.frombuffer(shared, dtype=np.float32)[offs[i] : offs[i + 1]].copy()
        else:
            x = synth.make_clip(int(lens[i]), seed_base + i).numpy()
            if shared is not None:  # hand the clip to the parent: the GPU arm times the very same samples
                np.frombuffer(shared, dtype=np.float32)[offs[i] : offs[i + 1]] = x
        mine.append(_to_native_pcm(x) if workload == "c2raw" else x)
    if mine:
        _cpu_process(workload, mine[0], F)  # warm-up: imports, table construction, page-in
    spans = []
    for _ in range(repeats):
        barrier.wait()
        t0 = time.perf_counter()
        frames = sum(_cpu_process(workload, x, F) for x in mine)
        spans.append((t0, time.perf_counter(), frames))
    queue.put((rank, spans))


def sample_spec(workload, n_batch, rank=0):
    """(lengths, seed base) of the first K clips of rank `rank`'s batch: the sample both arms share."""
    from heart_murmur_detection_b200 import synth

    lens = synth.clip_lengths("c2" if workload == "c2nf" else workload, n_batch, seed=LENS_SEED + rank)
    k = min(n_batch, CPU_SAMPLE[workload])
    return lens[:k], SEED_STRIDE * rank


def cpu_reference(workload: str, lens, seed_base: int, repeats: int = 1, keep_clips: bool = False, clips=None):
    """Oracle port (numpy restatement of the librosa path + live scipy / torchaudio), one clip at a
    time on every host core - mirrors the reference's one-file-at-a-time loop (model_util.py:138)
    run as `cores` independent processes, clip i on core i mod cores.  Time = first start to last
    finish (CLOCK_MONOTONIC is shared by the processes).  The clips are the FIRST K CLIPS OF THE GPU ARM'S BATCH
    (same lengths, same per-clip generator seeds); with ``keep_clips`` they are returned (shared memory) so that
    the GPU arm can place exactly these samples in its batch."""
    import multiprocessing as mp

    cores = os.cpu_count() or 1
    n_clips = len(lens)
    ctx = mp.get_context("fork")
    offs = np.zeros(n_clips + 1, dtype=np.int64)
    np.cumsum(lens, out=offs[1:])
    shared = clips if clips is not None else (ctx.RawArray("f", int(offs[-1])) if keep_clips else None)
    barrier, queue = ctx.Barrier(cores), ctx.Queue()
    procs = [ctx.Process(target=_cpu_worker, args=(r, cores, workload, lens, seed_base, repeats, barrier, queue, shared, offs,
                                                    clips is not None)) for r in range(cores)]
    for p in procs:
        p.start()
    results = [queue.get() for _ in procs]
    for p in procs:
        p.join()
    best, frames = None, 0
    for k in range(repeats):
        t0 = min(sp[k][0] for _, sp in results)
        t1 = max(sp[k][1] for _, sp in results)
        if best is None or t1 - t0 < best:
            best, frames = t1 - t0, sum(sp[k][2] for _, sp in results)
    return {"clips_per_s": n_clips / best, "frames_per_s": frames / best, "cores": cores, "seconds": best, "repeats": repeats,
            "clips": np.frombuffer(shared, dtype=np.float32) if (keep_clips or clips is not None) else None, "offsets": offs,
            "shared": shared,
            "sample": f"the first {n_clips} clips of the GPU arm's own batch of workload {workload} (same lengths, same "
                      f"per-clip seeds; {float(np.sum(lens)) / SR:.0f} s of audio, {best * cores:.0f} core-seconds per "
                      f"repeat, best of {repeats}), clip i on core i mod {cores}, one process per core, 1 BLAS thread each"}


def shared_config(wl, n_clips):
    """`config` of the JSON line: identical in the GPU arm and the reference arm (what both arms were asked to run)."""
    return {"workload": WORKLOADS[wl], "workload_id": wl, "clips_per_gpu_per_step": n_clips,
            "clip_lengths": f"synth.clip_lengths(seed={LENS_SEED} + rank)", "clip_seeds": f"{SEED_STRIDE} * rank + clip index",
            "cpu_sample": f"clips 0..{min(n_clips, CPU_SAMPLE.get(wl, 0)) - 1} of rank 0's batch, generated by the host "
                          "generator in both arms"}


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
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i",
                 str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t_begin=None, t_end=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], None, set(), []
        for ts, ln in self.lines:
            if t_begin is not None and not (t_begin <= ts <= t_end):
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power) if power else None,
                "window": "timed region + end-to-end loop + sub-workloads (the timed region alone is shorter than nvidia-smi's period)"}



# ----------------------------------------------------------------------------- sub-workloads (embedded records)


def _event_ms(fn, steps, warmup):
    import torch

    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / max(steps, 1)


def bench_fixed(wl, dev, steps, warmup, variant="auto"):
    """c1 (CirCor log-mel) / c3 (Audio-MAE fbank): one kernel family on a fixed-length batch, device resident."""
    import torch

    from heart_murmur_detection_b200 import frontend, synth

    hbm_peak, _ = peaks()
    n = DEFAULT_CLIPS[wl]
    lens = synth.clip_lengths(wl, n, seed=LENS_SEED)
    wav, off = synth.make_batch(lens, base_seed=0, device=dev)
    total = int(off[-1])
    if wl == "c1":
        plan = frontend.logmel_plan(16000, 64, 50, 8000, 1024, 512, variant)
        rows, cols = int(plan.frame_offsets(off)[-1]), 64
        out = torch.empty((rows, cols), dtype=torch.float32, device=dev)
        fn = lambda: plan(wav, off, out=out)  # noqa: E731
        frames = rows
    else:
        plan = frontend.fbank_plan(sample_rate=16000)
        rows, cols = n * 1024, 128
        out = torch.empty((rows, cols), dtype=torch.float32, device=dev)
        fn = lambda: plan(wav, off, rows_per_clip=1024, out=out)  # noqa: E731
        frames = int(plan.num_frames(np.diff(off)).sum())
    _event_ms(fn, 0, warmup)
    plan.set_profile(True)
    ms = _event_ms(fn, steps, 0)
    if wl == "c1":
        p_ms, f_ms, calls = plan.profile_ms()
        kern = {"logmel_power": p_ms / calls, "logmel_finalize": f_ms / calls}
        alg = {"logmel_power": 4 * total + rows * cols * 4, "logmel_finalize": 2 * rows * cols * 4}
        dom = "logmel_power"
    else:
        k_ms, calls = plan.profile_ms()
        kern = {"fbank": k_ms / calls}
        alg = {"fbank": 4 * total + rows * cols * 4}
        dom = "fbank"
    plan.set_profile(False)
    ach = alg[dom] / (kern[dom] * 1e-3) / 1e9
    return {"workload": WORKLOADS[wl], "value": n / (ms * 1e-3), "unit": "clips/s", "frames_per_s": frames / (ms * 1e-3),
            "ms_per_step": ms, "steps": steps, "clips_per_step": n, "launches_per_step": int(plan.last_launches),
            "kernels_ms_per_launch": {k: round(v, 4) for k, v in kern.items()},
            "roofline": {"bound": "hbm", "kernel": dom, "kernel_ms": kern[dom], "algorithmic_bytes_per_launch": alg[dom],
                         "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak},
            "l2_policy": f"inputs larger than L2 ({total * 4 / 1e6:.0f} MB of samples per step)"}


def bench_c2raw(dev, wav, off, steps, warmup):
    """c2 from native-rate PCM16: device-resident (int16 batch in HBM -> fused decode + resample -> c2 path) and end
    to end through pipeline.entire_signal_from_host(sr_in=4000) from pinned host memory."""
    import torch

    from heart_murmur_detection_b200 import frontend, pipeline

    n = off.size - 1
    n4 = np.diff(off) // (SR // NATIVE_SR)
    o4 = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(n4, out=o4[1:])
    d_pcm = torch.empty(int(o4[-1]), dtype=torch.int16, device=dev)
    step4 = SR // NATIVE_SR
    for i in range(n):  # every 4th sample of each clip, quantised like a 16-bit WAV (same rule as the CPU arm)
        seg = wav[int(off[i]) : int(off[i]) + step4 * int(n4[i]) : step4]
        d_pcm[int(o4[i]) : int(o4[i + 1])] = torch.clamp(torch.round(seg * 32768.0), -32768, 32767).to(torch.int16)
    rplan = frontend.resample_plan(NATIVE_SR, SR, **frontend.RESAMPLE_PRESETS["torchaudio"])
    o16 = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(rplan.out_lengths(n4), out=o16[1:])
    wav16 = torch.empty(int(o16[-1]), dtype=torch.float32, device=dev)
    probe = pipeline.entire_signal_batch(rplan(d_pcm, o4, out=wav16)[0], o16, spectrogram=True, **C2_KW)
    rows = int(probe.row_offsets[-1])
    del probe
    ub_rows = int((1 + np.maximum(np.diff(o16), 8 * SR) // 512).sum())
    work_buf = torch.empty(int(o16[-1]) + 8 * SR * n, dtype=torch.float32, device=dev)
    out = torch.empty((ub_rows, 64), dtype=torch.float32, device=dev)

    def step():
        rplan(d_pcm, o4, out=wav16)
        pipeline.entire_signal_device(wav16, o16, work=work_buf, out=out, **C2_KW)

    ms = _event_ms(step, steps, warmup)
    ms_rs = _event_ms(lambda: rplan(d_pcm, o4, out=wav16), steps, 1)
    h_pcm = torch.empty(d_pcm.numel(), dtype=torch.int16, pin_memory=True)
    h_pcm.copy_(d_pcm)
    ub = int((1 + np.maximum(np.diff(o16), 8 * SR) // 512).sum())
    h_out = torch.empty((ub, 64), dtype=torch.float32, pin_memory=True)
    del wav16, work_buf, out

    def e2e_step():
        pipeline.entire_signal_from_host(h_pcm, o4, h_out, sr_in=NATIVE_SR, **C2_KW)

    for _ in range(2):
        e2e_step()
    torch.cuda.synchronize()
    reps = 5
    t0 = time.perf_counter()
    for _ in range(reps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / reps
    hbm_peak, _ = peaks()
    rs_bytes = 2 * int(o4[-1]) + 4 * int(o16[-1])
    return {"workload": WORKLOADS["c2raw"], "value": n / (ms * 1e-3), "unit": "clips/s", "ms_per_step": ms, "clips_per_step": n,
            "native_rate_hz": NATIVE_SR, "resample_kernel_ms": ms_rs,
            "resample_roofline": {"bound": "hbm", "algorithmic_bytes_per_launch": rs_bytes, "achieved": rs_bytes / (ms_rs * 1e-3) / 1e9,
                                  "peak": hbm_peak, "unit": "GB/s", "frac": rs_bytes / (ms_rs * 1e-3) / 1e9 / hbm_peak},
            "e2e": {"value": n / e2e_s, "unit": "clips/s", "ms_per_step": e2e_s * 1e3, "h2d_bytes_per_step": 2 * int(o4[-1]),
                    "d2h_bytes_per_step": rows * 64 * 4, "steps": reps,
                    "api": "pipeline.entire_signal_from_host(sr_in=4000)"}}


def bench_c4(dev, rank, n_total=100000, pool=2048, max_len=251, batch=64):
    """BASELINE config 4: producer (heart_pressl.py:58-99: trim + whole-recording log-mel, recordings shorter than
    8 s dropped) + COLA Dataset items (cola_training.py:56-80) on the fly.  The 100 000 recordings are streamed as
    ceil(100000 / pool) passes over a resident pool of `pool` distinct recordings (100 000 x 34 s would be 218 GB);
    every pass re-runs the full path: kernels, host planning and the Python-RNG-exact draws."""
    import random

    import torch

    from heart_murmur_detection_b200 import datasets, pipeline, synth

    lens = synth.clip_lengths("c4", pool, seed=4000 + rank)
    wav, off = synth.make_batch(lens, base_seed=SEED_STRIDE * rank + 5_000_000, device=dev)
    total = int(off[-1])
    rows_ub = int((1 + np.diff(off) // 512).sum())
    out = torch.empty((rows_ub, 64), dtype=torch.float32, device=dev)
    passes = -(-n_total // pool)
    t_prod = t_data = 0.0
    items = pairs_checksum = 0

    def one_pass(timed):
        nonlocal t_prod, t_data, items, pairs_checksum
        t0 = time.perf_counter()
        fb = pipeline.entire_signal_batch(wav, off, input_sec=8, spectrogram=True, out=out)
        store = datasets.SpecStore.from_features(fb.features, fb.row_offsets)
        if timed:
            torch.cuda.synchronize()
        t1 = time.perf_counter()
        n_items = len(fb.row_offsets) - 1
        x1, x2 = datasets.cola_batch(store, np.arange(n_items), max_len=max_len, augment=True)
        if timed:
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            t_prod += t1 - t0
            t_data += t2 - t1
            items += n_items
        return x1, x2, fb

    random.seed(2024 + rank)
    one_pass(False)
    torch.cuda.synchronize()
    t_begin = time.perf_counter()
    for _ in range(passes):
        x1, x2, fb = one_pass(True)
    wall = time.perf_counter() - t_begin
    n_rec = passes * pool
    frames = int(fb.row_offsets[-1]) * passes
    return {"workload": WORKLOADS["c4"], "value": n_rec / wall, "unit": "recordings/s", "recordings": n_rec,
            "seconds": wall, "passes": passes, "pool_recordings": pool, "audio_seconds_per_s": passes * total / SR / wall,
            "frames_per_s": frames / wall, "items_per_s": items / wall, "pairs_shape": [int(v) for v in x1.shape],
            "batches_of_64_per_pass": -(-int(x1.shape[0]) // batch),
            "producer_ms_per_pass": 1e3 * t_prod / passes, "dataset_ms_per_pass": 1e3 * t_data / passes,
            "note": "wall clock with a device synchronise after each stage (host planning and the Python-RNG-exact "
                    "draws are part of the path); one cola_batch call per pass = the pass's items in DataLoader order, "
                    "i.e. consecutive batches of 64"}


def bench_c5(dev, rank, world, variant="auto", n_total=1_000_000, chunk=5000, pool_chunks=2, passes=2):
    """BASELINE config 5, STRONG scaling: the job is fixed (1 M clips); chunks of `chunk` clips are dealt round-robin
    (dist.chunk_rounds), every rank writes its features in place into the global-order tensor and each round is
    completed by one in-place NCCL all-gather on a side stream, overlapped with the next round's kernels.  Input
    clips come from a resident pool (1 M x 512 KB = 512 GB does not fit): global clip g is pool clip g mod pool."""
    import torch
    import torch.distributed as dist

    from heart_murmur_detection_b200 import dist as hd
    from heart_murmur_detection_b200 import frontend, synth

    chunk = int(os.environ.get("HMFE_C5_CHUNK", chunk))  # clips per rank per round (A/B runs)
    n_samp, T = 8 * SR, 251
    pool = chunk * pool_chunks
    group = None
    # The sweep is bound by the all-gather, not by the kernels (N = 8: 36 ms of compute per job), so its collective gets a
    # communicator of its own with more CTAs than NCCL's default, and 320 MB per rank and round.  Measured at N = 8:
    # 1 250 clips per round, default CTAs 9.06 M clips/s (509 GB/s into every GPU); 5 000 clips 9.75 M; + 32 CTAs 9.88 M;
    # + 64 CTAs 10.18 M (572 GB/s).  HMFE_C5_NCCL_CTAS=0 keeps the default communicator.
    ctas = int(os.environ.get("HMFE_C5_NCCL_CTAS", 64))
    if world > 1 and ctas > 0:
        opts = dist.ProcessGroupNCCL.Options()
        opts.config.min_ctas = opts.config.max_ctas = ctas
        group = dist.new_group(ranks=list(range(world)), pg_options=opts)
    wav, _ = synth.make_batch(np.full(pool, n_samp, dtype=np.int64), base_seed=77_000_000, device=dev)  # same on every rank
    off = np.arange(chunk + 1, dtype=np.int64) * n_samp
    plan = frontend.logmel_plan(16000, 64, 50, 8000, 1024, 512, variant)
    rounds = hd.chunk_rounds(n_total, chunk, world)
    final = torch.empty((n_total * T, 64), dtype=torch.float32, device=dev)
    rows_chunk = chunk * T
    comm = torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream()

    def job(gather):
        for j in range(rounds):
            c = j * world + rank
            pc = c % pool_chunks
            plan(wav[pc * chunk * n_samp : (pc + 1) * chunk * n_samp], off, out=final[c * rows_chunk : (c + 1) * rows_chunk])
            if gather and world > 1:
                ev = torch.cuda.Event()
                ev.record(main)
                with torch.cuda.stream(comm):
                    comm.wait_event(ev)
                    hd.all_gather_round_inplace(final, j, rows_chunk, group=group)
        main.wait_stream(comm)

    def timed(gather):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        job(gather)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    final.zero_()
    timed(True)  # warm-up pass (also the pass the identity check reads)
    # in-run identity check: sampled clips of the gathered tensor == the same clip computed alone on this rank (P = 1)
    rng = np.random.default_rng(5)
    sample = np.unique(np.concatenate([rng.integers(0, n_total, size=24), [0, n_total - 1, chunk * world - 1, chunk * world]]))
    one = np.array([0, n_samp], dtype=np.int64)
    same = True
    for g in sample:
        pc, i = (int(g) // chunk) % pool_chunks, int(g) % chunk
        ref, _ = plan(wav[(pc * chunk + i) * n_samp : (pc * chunk + i + 1) * n_samp], one)
        same = same and bool(torch.equal(final[int(g) * T : (int(g) + 1) * T], ref))
    flag = torch.tensor([1 if same else 0], device=dev)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    ms = min(timed(True) for _ in range(passes))
    ms_compute = timed(False) if world > 1 else ms
    del final
    torch.cuda.empty_cache()
    return {"workload": WORKLOADS["c5"], "value": n_total / (ms * 1e-3), "unit": "clips/s", "scaling": "strong",
            "n_gpus": world, "clips_total": n_total, "frames_per_s": n_total * T / (ms * 1e-3), "ms_per_job": ms,
            "ms_per_job_compute_only": ms_compute, "value_compute_only": n_total / (ms_compute * 1e-3),
            "rounds": rounds, "chunk_clips": chunk, "pool_clips": pool,
            "gathered_bytes_per_rank": (world - 1) * (n_total // world) * T * 64 * 4,
            "nvlink_rx_gbs_per_gpu": ((world - 1) * (n_total // world) * T * 64 * 4) / (ms * 1e-3) / 1e9 if world > 1 else 0.0,
            "collective": "none" if world == 1 else f"in-place NCCL all_gather_into_tensor per round on a side stream, own communicator ({ctas or 'default'} CTAs)",
            "identity_check": {"sampled_clips": int(sample.size), "bit_identical_to_p1_on_every_rank": bool(flag.item()),
                               "how": "gathered rows of sampled global clips == the same clip run alone through the same plan"},
            "passes_timed": passes, "timing": "CUDA events around one whole job, barrier + synchronise both sides, max over ranks, best pass"}


def check_all_gather_features(dev, rank, world):
    """The ragged product path on NCCL, outside the timed region: one global ragged batch (identical on every rank),
    dist.shard_by_length -> local pipeline -> dist.all_gather_features (metadata gathers + order-restoring index_select)
    against the single-GPU result computed locally.  Trim, padding and log-mel are per clip: without the band-pass the
    gathered tensor must equal the single-GPU one bit for bit.  The one-pass band-pass picks its chunk length per call
    (wave quantisation over the batch it is given), and a chunk that starts elsewhere moves the float64 recurrence at the
    1e-13 level, which flips a float32 rounding now and then: with the band-pass the check reports the largest difference
    (tools/check_shard_invariance.py measures the same on one GPU: 69 of 9.07 M elements, 2.5e-7 of the [0, 1] range)."""
    import torch
    import torch.distributed as dist

    from heart_murmur_detection_b200 import dist as hd
    from heart_murmur_detection_b200 import pipeline, synth

    n = 256
    lens = synth.clip_lengths("c2", n, seed=99)
    wav, off = synth.make_batch(lens, base_seed=123_000_000, device=dev)
    shard = hd.shard_by_length(lens, world)[rank]
    lo = np.zeros(len(shard) + 1, dtype=np.int64)
    np.cumsum(lens[shard], out=lo[1:])
    lw = torch.cat([wav[int(off[i]) : int(off[i + 1])] for i in shard])
    rec = {"api": "dist.shard_by_length + pipeline.entire_signal_batch + dist.all_gather_features", "clips": n}
    for key, bw in (("no_bandpass", None), ("bandpass", 5)):
        kw = dict(input_sec=8, butterworth_filter=bw, pad=True, types="zero", max_sec=32, spectrogram=True)
        full = pipeline.entire_signal_batch(wav, off, **kw)
        res = pipeline.entire_signal_batch(lw, lo, **kw)
        rows = np.zeros(len(shard), dtype=np.int64)
        rows[res.chunks.clip_ids] = np.diff(res.row_offsets)
        out, ro = hd.all_gather_features(res.features[: int(res.row_offsets[-1])], rows, shard, n)
        ref = full.features[: int(full.row_offsets[-1])]
        same_shape = out.shape == ref.shape
        diff = (out - ref).abs() if same_shape else None
        stats = torch.tensor([1.0 if same_shape and bool(torch.equal(out, ref)) else 0.0,
                              float(diff.max()) if same_shape else float("inf"),
                              float((diff > 0).sum()) if same_shape else float("inf")], device=dev, dtype=torch.float64)
        ident = stats[:1].clone()
        dist.all_reduce(ident, op=dist.ReduceOp.MIN)
        worst = stats[1:].clone()
        dist.all_reduce(worst, op=dist.ReduceOp.MAX)
        rec[key] = {"rows": int(ref.shape[0]), "bit_identical_to_p1_on_every_rank": bool(ident.item()),
                    "max_abs_diff": float(worst[0].item()), "differing_elements": int(worst[1].item()),
                    "elements": int(ref.numel())}
    rec["bit_identical_to_p1_on_every_rank"] = rec["no_bandpass"]["bit_identical_to_p1_on_every_rank"]
    rec["within_tolerance_with_bandpass"] = rec["bandpass"]["max_abs_diff"] <= 2e-4  # the tolerance of the normalised log-mel
    return rec


def run_sub_as_main(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    wl = args.workload
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    t_begin = time.perf_counter()
    if wl == "c5":
        rec = bench_c5(dev, rank, world, args.variant, n_total=args.clips or DEFAULT_CLIPS["c5"])
        scaling = "strong"
    else:
        rec = bench_c4(dev, rank, n_total=args.clips or DEFAULT_CLIPS["c4"])
        if world > 1:
            t = torch.tensor([rec["value"]], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            rec["value"] = float(t.item())
        scaling = "weak"
    clocks = sampler.stop(t_begin, time.perf_counter()) if rank == 0 else None
    if rank == 0:
        line = {"metric": "front-end clips/s", "value": rec["value"], "unit": rec["unit"], "n_gpus": world,
                "steps": rec.get("passes_timed", rec.get("passes", 1)), "warmup": 1,
                "ms_per_step": rec.get("ms_per_job", 1e3 * rec.get("seconds", 0.0)), "higher_is_better": True,
                "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": shared_config(wl, args.clips or DEFAULT_CLIPS[wl]), "e2e": None,
                "e2e_note": "inputs are generated on the device (c5: 512 GB of samples would not fit the host); the "
                            "end-to-end figure of the path is the default workload's",
                "workloads": {wl: rec}, "clocks": clocks}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------- main


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json, copy kernel)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic(workload, kernel, n_clips):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel` from the committed ncu --set full
    capture (profiles/*_traffic.json, written by tools/make_profiles.py); None when the capture was taken at
    another batch size."""
    import glob

    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json")), reverse=True):
        try:
            rec = json.load(open(path)).get(workload, {}).get(kernel)
        except Exception:
            continue
        if rec and int(rec.get("clips_per_launch", -1)) == int(n_clips):
            return float(rec["dram_bytes_read"]) + float(rec["dram_bytes_write"]), os.path.basename(path)
    return None, None


def run_reference(args, rank):
    if rank != 0:
        return
    wl = args.workload
    n_batch = args.clips or DEFAULT_CLIPS[wl]
    lens, seed_base = sample_spec(wl, n_batch)
    # every "step" of this arm is one pass over the bounded sample; the run is sized to end within minutes
    repeats = max(1, min(args.steps, 3))
    r = cpu_reference(wl, lens, seed_base, repeats=repeats)
    line = {
        "impl": "reference", "metric": "front-end clips/s", "value": r["clips_per_s"], "unit": "clips/s",
        "n_gpus": args.gpus, "steps": repeats, "steps_requested": args.steps, "warmup": 1, "ms_per_step": r["seconds"] * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64 FFT + f64 IIR / f32 mel (numpy, scipy)",
        "data": "synthetic", "config": shared_config(wl, n_batch),
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
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--variant", default="auto", choices=["auto", "scalar", "packed", "pair", "tc"])
    ap.add_argument("--clips", type=int, default=0, help="clips per GPU per step (0 = the workload's size)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-sub", action="store_true", help="skip the embedded c1 / c3 / c4 / c5 sub-records")
    ap.add_argument("--gather", default="auto", choices=["auto", "copy", "multicast", "nccl"],
                    help="N > 1: all-gather through peer memory (dist.PeerAllGather: copy engines or NVLink multicast "
                         "push) or NCCL.  auto = copy engines up to 4 GPUs, NCCL beyond (measured, DESIGN.md section 8)")
    ap.add_argument("--push-ctas", type=int, default=16)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    args.warmup = max(args.warmup, 3)
    wl = args.workload

    if args.impl == "reference":
        run_reference(args, rank)
        return

    # CPU baseline first (rank 0, N=1 only), before CUDA is initialised in this process (fork safety)
    if wl in ("c4", "c5"):
        return run_sub_as_main(args, rank, world, local_rank)
    cpu, host_clips, cpu_raw = None, None, None
    n_clips = args.clips or DEFAULT_CLIPS[wl]
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        k_lens, k_seed = sample_spec(wl, n_clips)
        r = cpu_reference(wl, k_lens, k_seed, repeats=2, keep_clips=True)
        host_clips = (r["clips"], r["offsets"])
        cpu = {"value": r["clips_per_s"], "unit": "clips/s", "cores": r["cores"], "kind": "port", "sample": r["sample"],
               "frames_per_s": r["frames_per_s"]}
        if wl == "c2" and not args.no_sub:  # the same clips as 4 kHz PCM16 through decode + resample + c2 on the CPU
            r2 = cpu_reference("c2raw", k_lens, k_seed, repeats=2, clips=r["shared"])
            cpu_raw = {"value": r2["clips_per_s"], "unit": "clips/s", "cores": r2["cores"], "kind": "port",
                       "sample": r2["sample"] + "; every 4th sample as 16-bit PCM, torchaudio.functional.resample 4 kHz -> 16 kHz "
                                 "(the oracle of the GPU resampler) inside the timed region"}

    import torch
    import torch.distributed as dist

    from heart_murmur_detection_b200 import frontend, pipeline, synth

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    lens = synth.clip_lengths("c2" if wl == "c2nf" else wl, n_clips, seed=LENS_SEED + rank)
    kw2 = C2NF_KW if wl == "c2nf" else C2_KW
    wav, off = synth.make_batch(lens, base_seed=SEED_STRIDE * rank, device=dev)
    if host_clips is not None:  # the clips the CPU arm timed ARE the first K clips of this batch, bit for bit
        wav[: host_clips[0].size].copy_(torch.from_numpy(host_clips[0]))
        assert np.array_equal(host_clips[1], off[: host_clips[1].size])
        host_clips = None
    total_samples = int(off[-1])
    ctx = frontend.default_ctx()
    lm_plan = frontend.logmel_plan(16000, 64, 50, 8000, 1024, 512, args.variant)
    fb_plan = frontend.fbank_plan(sample_rate=16000)
    state = {}

    # Every workload writes its features into the buffer it is given, so that with N > 1 the kernels
    # write straight into the all-gather send buffer (no staging copy).
    work_buf = None
    if wl == "c1":
        fo = lm_plan.frame_offsets(off)
        out_rows = int(fo[-1])

        def step(out):
            lm_plan(wav, off, out=out)
            state.update(features=out, rows=out_rows, launches=lm_plan.last_launches)
    elif wl in ("c2", "c2nf"):
        probe = pipeline.entire_signal_batch(wav, off, spectrogram=True, **kw2)
        out_rows = int(probe.row_offsets[-1])
        chunk_samples_probe = int(probe.chunks.lengths.sum())
        del probe
        # the step is enqueued without any host round trip (device-side planner); buffers hold the upper bounds
        L8 = 8 * SR
        work_buf = torch.empty(total_samples + L8 * n_clips, dtype=torch.float32, device=dev)
        ub_rows = int((1 + np.maximum(np.diff(off), L8) // 512).sum())

        def step(out):
            res = pipeline.entire_signal_device(wav, off, work=work_buf, out=out, **kw2)
            state.update(features=res.features, rows=out_rows, launches=res.launches, res=res)
    else:
        out_rows = n_clips * 1024

        def step(out):
            fb_plan(wav, off, rows_per_clip=1024, out=out)
            state.update(features=out, rows=out_rows, launches=fb_plan.last_launches)

    if wl in ("c2", "c2nf") and args.variant != "auto":  # make the pipeline pick the requested log-mel variant
        frontend._plans[("logmel", torch.cuda.current_device(), 16000, 64, 50.0, 8000.0, 1024, 512, "auto", "constant")] = lm_plan

    n_cols = 128 if wl == "c3" else 64
    rows = out_rows
    max_rows = ub_rows if wl in ("c2", "c2nf") else rows  # send / gather buffers hold the upper bound of rows
    if world > 1:
        t = torch.tensor([rows], device=dev, dtype=torch.int64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        max_rows = int(t.item())
    # two send buffers (and two gather targets): the all-gather of step i runs on its own stream(s) while
    # step i + 1 computes into the other buffer; the timed region ends after the last gather.
    peer_ag = None
    if args.gather == "auto":
        args.gather = "copy" if world <= 4 else "nccl"
    if world > 1 and args.gather != "nccl":
        from heart_murmur_detection_b200.dist import PeerAllGather

        # kernels write straight into this rank's slot of the gathered buffer
        try:
            peer_ag = PeerAllGather(max_rows, n_cols, mode=args.gather, push_ctas=args.push_ctas)
        except Exception as e:  # no symmetric memory on this box: the NCCL collective is always available
            if rank == 0:
                print(f"bench: peer-memory all-gather unavailable ({type(e).__name__}: {e}); using NCCL", file=sys.stderr)
            peer_ag, args.gather = None, "nccl"
    if peer_ag is not None:
        send = [peer_ag.slot(0), peer_ag.slot(1)]
        comm = peer_ag.comm
    else:
        send = [torch.zeros((max_rows, n_cols), dtype=torch.float32, device=dev) for _ in range(2 if world > 1 else 1)]
    out = send[0]
    gather_group = None
    if world > 1 and peer_ag is None:
        gathered = [torch.empty((world * max_rows, n_cols), dtype=torch.float32, device=dev) for _ in range(2)]
        comm = torch.cuda.Stream(device=dev)
        if int(os.environ.get("HMFE_GATHER_NCCL_CTAS", 0)) > 0:  # A/B: a communicator with a fixed CTA count for the step's gather
            opts = dist.ProcessGroupNCCL.Options()
            opts.config.min_ctas = opts.config.max_ctas = int(os.environ["HMFE_GATHER_NCCL_CTAS"])
            gather_group = dist.new_group(ranks=list(range(world)), pg_options=opts)
        ev_ready = [torch.cuda.Event() for _ in range(2)]
        ev_sent = [torch.cuda.Event() for _ in range(2)]
    step_no = [0]

    def full_step():
        i = step_no[0] % len(send)
        step_no[0] += 1
        main = torch.cuda.current_stream()
        if peer_ag is not None:
            peer_ag.wait_reusable(i, main)
            step(send[i])
            peer_ag.gather(i, main)
            return
        if world > 1 and step_no[0] > 2:
            main.wait_event(ev_sent[i])  # the gather that read send[i] two steps ago
        step(send[i])
        if world > 1:  # the path's single collective: all-gather of the (row-padded) feature blocks
            ev_ready[i].record(main)
            with torch.cuda.stream(comm):
                comm.wait_event(ev_ready[i])
                dist.all_gather_into_tensor(gathered[i], send[i], group=gather_group)
                ev_sent[i].record(comm)

    def barrier():
        if world > 1:
            torch.cuda.current_stream().wait_stream(comm)
            dist.barrier()
        torch.cuda.synchronize()

    step(send[0])
    torch.cuda.synchronize()
    for _ in range(args.warmup):
        full_step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    t_begin = time.perf_counter()
    ctx.set_profile(True)
    lm_plan.set_profile(True)
    fb_plan.set_profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        full_step()
    if world > 1:
        torch.cuda.current_stream().wait_stream(comm)  # the last all-gather is inside the timed region
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    iir_plan = ctx.last_iir_plan() if wl == "c2" else None  # what the timed steps used (later calls overwrite it)
    kern = {k: v for k, v in ctx.profile_ms().items() if v[1]}
    p_ms, f_ms, p_n = lm_plan.profile_ms()
    if p_n:
        kern["logmel_power"] = (p_ms, p_n)
        kern["logmel_finalize"] = (f_ms, p_n)
    k_ms, k_n = fb_plan.profile_ms()
    if k_n:
        kern["fbank"] = (k_ms, k_n)
    ctx.set_profile(False)
    lm_plan.set_profile(False)
    fb_plan.set_profile(False)
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    clips_per_s = world * n_clips / (ms_step * 1e-3)
    n_frames_valid = rows if wl != "c3" else int(fb_plan.num_frames(np.diff(off)).sum())

    # ---- end to end through the host-buffer API (pinned host in, pinned host out, copies timed)
    e2e = None
    if not args.no_e2e:
        h_wav = torch.empty(total_samples, dtype=torch.float32, pin_memory=True)
        h_wav.copy_(wav)
        h_out = torch.empty((ub_rows if wl in ("c2", "c2nf") else rows, n_cols), dtype=torch.float32, pin_memory=True)

        def e2e_step():
            if wl == "c1":
                frontend.logmel_from_host(lm_plan, h_wav, off, h_out)
            elif wl in ("c2", "c2nf"):
                pipeline.entire_signal_from_host(h_wav, off, h_out, **kw2)
            else:
                frontend.fbank_from_host(fb_plan, h_wav, off, h_out, rows_per_clip=1024)
            if world > 1 and peer_ag is not None:
                peer_ag.wait_reusable(0)
                peer_ag.gather(0)
                peer_ag.finish()
            elif world > 1:
                dist.all_gather_into_tensor(gathered[0], send[0])

        e2e_steps = max(3, min(args.steps, 5))
        for _ in range(2):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        barrier()
        e2e_s = (time.perf_counter() - t0) / e2e_steps
        t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
        e2e = {"value": world * n_clips / e2e_s, "unit": "clips/s", "h2d_bytes_per_step": total_samples * 4,
               "d2h_bytes_per_step": rows * n_cols * 4, "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
               "api": {"c1": "frontend.logmel_from_host", "c2": "pipeline.entire_signal_from_host",
                       "c2nf": "pipeline.entire_signal_from_host", "c3": "frontend.fbank_from_host"}[wl]}
        if True:  # same call with the 16-bit WAV payload as the host buffer (decoded on the device)
            h_pcm = torch.clamp(torch.round(h_wav * 32768.0), -32768, 32767).to(torch.int16).pin_memory()
            ub = int((1 + np.maximum(np.diff(off), 8 * SR) // 512).sum()) if wl in ("c2", "c2nf") else rows
            h_out16 = torch.empty((ub, n_cols), dtype=torch.float32, pin_memory=True)

            def pcm_step():
                if wl == "c1":
                    frontend.logmel_from_host(lm_plan, h_pcm, off, h_out16)
                elif wl in ("c2", "c2nf"):
                    pipeline.entire_signal_from_host(h_pcm, off, h_out16, **kw2)
                else:
                    frontend.fbank_from_host(fb_plan, h_pcm, off, h_out16, rows_per_clip=1024)

            for _ in range(2):
                pcm_step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                pcm_step()
            barrier()
            s16 = (time.perf_counter() - t0) / e2e_steps
            t = torch.tensor([s16], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e["pcm16_host_input"] = {"value": world * n_clips / float(t.item()), "unit": "clips/s",
                                       "h2d_bytes_per_step": total_samples * 2, "ms_per_step": float(t.item()) * 1e3,
                                       "note": "same API, host buffer = 16-bit PCM WAV payload (quantised copy of the "
                                               "synthetic clips), int16 -> float32 / 32768 on the device"}
    # ---- the other BASELINE configs as sub-records of the same line (c1, c3, c4 at N = 1; the c5 sweep at every N)
    chunk_samples = chunk_samples_probe if wl in ("c2", "c2nf") else total_samples
    launches_per_step = int(state["launches"])
    subs = {}
    if not args.no_sub and wl == "c2" and world == 1:
        state.clear()
        pipeline._host_pipes.clear()
        torch.cuda.empty_cache()
        subs["c2raw"] = bench_c2raw(dev, wav, off, args.steps, args.warmup)
        subs["c2raw"]["cpu_baseline"] = cpu_raw
    if not args.no_sub and wl == "c2":
        state.clear()
        h_wav = h_out = h_pcm = h_out16 = None  # noqa: F841
        del wav, work_buf, send, out
        if world > 1 and peer_ag is None:
            del gathered
        peer_ag = None if peer_ag is None else peer_ag  # symmetric buffers stay (freed at exit)
        pipeline._host_pipes.clear()
        torch.cuda.empty_cache()
        if world == 1:
            subs["c1"] = bench_fixed("c1", dev, args.steps, args.warmup, args.variant)
            subs["c3"] = bench_fixed("c3", dev, args.steps, args.warmup)
            subs["c4"] = bench_c4(dev, rank)
        else:
            subs["gather_check"] = check_all_gather_features(dev, rank, world)
        subs["c5"] = bench_c5(dev, rank, world, args.variant)
    clocks = sampler.stop(t_begin, time.perf_counter()) if rank == 0 else None

    if rank == 0:
        hbm_peak, peak_src = peaks()
        # algorithmic bytes per launch of each kernel
        out_bytes = rows * n_cols * 4
        alg = {
            "logmel_power": 4 * chunk_samples + out_bytes,
            "logmel_finalize": 2 * out_bytes,
            "fbank": 4 * total_samples + out_bytes,
            "iir_zero_state": 4 * total_samples,
            "iir_final": 8 * total_samples,
            "iir_overlap": 8 * total_samples,
            "trim_power": 4 * total_samples,
        }
        per_launch = {k: v[0] / v[1] for k, v in kern.items()}
        # the dominant kernel: the longest one; kernels within 5 % of it count as tied (c2: band-pass 3.3 ms, log-mel 3.2-3.3 ms,
        # the order flips from run to run) and the tie goes to the one that moves more bytes, the one an HBM roofline is about
        t_max = max(per_launch.values())
        dom = max((k for k in per_launch if per_launch[k] >= 0.95 * t_max), key=lambda k: (alg.get(k, 0), per_launch[k]))
        dom_s = per_launch[dom] * 1e-3
        achieved = alg.get(dom, 0) / dom_s / 1e9
        step_kernel_ms = sum(v[0] for v in kern.values()) / args.steps
        traffic, traffic_src = measured_traffic(wl, dom, n_clips)
        roof = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "kernel": dom,
                "kernel_ms": per_launch[dom],
                "algorithmic_bytes_per_launch": alg.get(dom),
                "kernels_ms_per_launch": {k: round(v, 4) for k, v in sorted(per_launch.items())},
                "kernels_gbs": {k: round(alg[k] / (per_launch[k] * 1e-3) / 1e9, 1) for k in per_launch if k in alg},
                "kernels_frac": {k: round(alg[k] / (per_launch[k] * 1e-3) / 1e9 / hbm_peak, 3) for k in per_launch if k in alg},
                "sum_kernel_ms_per_step": step_kernel_ms,
                "note": "FFT kernels are FP32-issue bound, not HBM bound (DESIGN.md section 5): fp32 pipe fraction "
                        "is reported in profiles/"}
        line = {
            "metric": "front-end clips/s", "value": clips_per_s, "unit": "clips/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 (FFT, mel) + f64 (IIR)", "data": "synthetic",
            "config": shared_config(wl, n_clips),
            "run": {"audio_seconds_per_gpu_per_step": total_samples / SR, "variant": args.variant,
                    "l2_policy": f"inputs larger than L2 ({total_samples * 4 / 1e6:.0f} MB of samples per step)",
                    "collective": ("none" if world == 1 else f"peer-memory all-gather of the row-padded features (symmetric memory, {peer_ag.mode}, device barrier), overlapped with the next step's kernels" if peer_ag is not None else "NCCL all_gather_into_tensor(row-padded features), overlapped with the next step's kernels"),
                    **({"iir_plan": iir_plan} if iir_plan else {})},
            "frames_per_s": world * n_frames_valid / (ms_step * 1e-3),
            "audio_seconds_per_s": world * total_samples / SR / (ms_step * 1e-3),
            "e2e": e2e,
            "gpu_launches": launches_per_step * args.steps,
            "roofline": roof,
            "cpu_baseline": cpu,
            "clocks": clocks,
            "workloads": subs,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
