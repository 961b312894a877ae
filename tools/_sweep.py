import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from heart_murmur_detection_b200 import synth
from heart_murmur_detection_b200.frontend import LogMelPlan
lens = synth.clip_lengths("c1", 1000)
wav, off = synth.make_batch(lens, base_seed=11, device="cuda")
out = torch.empty((251 * 1000, 64), device="cuda")
for variant, warps in (("packed", "11"), ("tc", "11"), ("tc", "8")):
    for st in ("0", "250", "500", "1000", "2000", "4000"):
        os.environ["HMFE_TC_FFT_WARPS"] = warps; os.environ["HMFE_LOGMEL_STAGGER_NS"] = st
        plan = LogMelPlan(f_max=8000, variant=variant)
        for _ in range(5): plan(wav, off, out=out)
        torch.cuda.synchronize(); plan.set_profile(True)
        for _ in range(20): plan(wav, off, out=out)
        p, f, n = plan.profile_ms()
        print(f"{variant} warps={warps} stagger={st}: {p/n:.4f} ms", flush=True)
