"""Debug / first timing of the tensor-core log-mel variant against the packed FP32 variant (run on the GPU box)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
import numpy as np, torch
from signals import golden_signal
from heart_murmur_detection_b200 import synth
from heart_murmur_detection_b200.frontend import LogMelPlan

def run(plan, clips, mode="power"):
    off = np.zeros(len(clips) + 1, dtype=np.int64); np.cumsum([len(c) for c in clips], out=off[1:])
    wav = torch.from_numpy(np.concatenate(clips).astype(np.float32)).cuda()
    out, fo = plan(wav, off, mode=mode); torch.cuda.synchronize()
    return [out[fo[i]:fo[i+1]].cpu().numpy() for i in range(len(clips))]

if "--shapes" in sys.argv:
    clips = [golden_signal(n, seed=3 + n % 11) for n in (48000, 7777, 128000, 1, 600, 1500)]
    for kw in (dict(f_max=2000), dict(f_max=8000, hop=256), dict(f_max=8000, hop=320, n_mels=32), dict(f_max=4000, hop=500), dict(f_max=8000)):
        for warps in ("11", "8"):
            os.environ["HMFE_TC_FFT_WARPS"] = warps
            tc, ref = LogMelPlan(variant="tc", **kw), LogMelPlan(variant="packed", **kw)
            a, b = run(tc, clips), run(ref, clips)
            for i, (x, y) in enumerate(zip(a, b)):
                err = np.abs(x - y).max() / max(np.abs(y).max(), 1e-30)
                bad = np.argwhere(np.abs(x - y) > 1e-4 * np.abs(y).max())
                print(kw, "warps", warps, "clip", i, len(clips[i]), "shape", x.shape, "max|d|/max|S| = %.3g" % err, "status", tc.tc_status(),
                      "first bad (frame, mel):", bad[:4].tolist() if len(bad) else "-", "nan" if not np.isfinite(x).all() else "")

if "--time" in sys.argv:
    lens = synth.clip_lengths("c1", 1000)
    wav, off = synth.make_batch(lens, base_seed=11, device="cuda")
    for variant, warps in (("packed", "11"), ("tc", "11"), ("tc", "8")):
        os.environ["HMFE_TC_FFT_WARPS"] = warps
        plan = LogMelPlan(f_max=8000, variant=variant)
        out = torch.empty((251 * 1000, 64), device="cuda")
        for _ in range(5): plan(wav, off, out=out)
        torch.cuda.synchronize()
        plan.set_profile(True)
        for _ in range(20): plan(wav, off, out=out)
        p, f, n = plan.profile_ms()
        print(f"c1 {variant} fft_warps={warps}: power kernel {p/n:.4f} ms, finalize {f/n:.4f} ms, status {plan.tc_status() if variant=='tc' else 0}")

if "--one" in sys.argv:  # a single launch of one variant (for ncu): --one tc|packed [fft_warps]
    i = sys.argv.index("--one")
    variant = sys.argv[i + 1]
    if len(sys.argv) > i + 2: os.environ["HMFE_TC_FFT_WARPS"] = sys.argv[i + 2]
    lens = synth.clip_lengths("c1", 1000)
    wav, off = synth.make_batch(lens, base_seed=11, device="cuda")
    plan = LogMelPlan(f_max=8000, variant=variant)
    out = torch.empty((251 * 1000, 64), device="cuda")
    for _ in range(3): plan(wav, off, out=out)
    torch.cuda.synchronize()
    print("done", variant, plan.tc_status() if variant == "tc" else 0)
