"""Throughput sample of the sibling front-ends (CLAP log-mel, HeAR mel-PCEN) on synthetic clips, CUDA-event timed.
``--small`` runs one tiny call of each (used under compute-sanitizer)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

from heart_murmur_detection_b200 import clap_input as ci  # noqa: E402
from heart_murmur_detection_b200 import hear_input as hi  # noqa: E402
from signals import golden_signal  # noqa: E402


def timed(fn, steps=20, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


def main():
    small = "--small" in sys.argv
    n_hear, n_clap = (3, 2) if small else (4096, 512)
    base = np.stack([golden_signal(32000, seed=500 + i) for i in range(min(n_hear, 16))])
    hear = torch.from_numpy(np.tile(base, (n_hear // len(base) + 1, 1))[:n_hear]).cuda().contiguous()
    basec = np.stack([golden_signal(5 * 44100, seed=600 + i, sr=44100) for i in range(min(n_clap, 4))])
    clap = torch.from_numpy(np.tile(basec, (n_clap // len(basec) + 1, 1))[:n_clap]).cuda().contiguous()
    plan = hi.hear_plan()
    if small:
        plan(hear)
        plan(hear[:, :20000].contiguous())
        ci.logmel_batch(clap)
        torch.cuda.synchronize()
        print("ok")
        return
    out = {}
    ms = timed(lambda: plan(hear))
    out["hear_mel_pcen"] = {"clips": n_hear, "ms": ms, "clips_per_s": n_hear / ms * 1e3,
                            "algorithmic_GBps": n_hear * (32000 * 4 + 192 * 128 * 4) / ms / 1e6}
    ms = timed(lambda: plan.mel_power(hear))
    out["hear_mel_only"] = {"clips": n_hear, "ms": ms}
    ms = timed(lambda: ci.logmel_batch(clap))
    out["clap_logmel"] = {"clips": n_clap, "ms": ms, "clips_per_s": n_clap / ms * 1e3,
                          "algorithmic_GBps": n_clap * (220500 * 4 + 690 * 64 * 4) / ms / 1e6}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
