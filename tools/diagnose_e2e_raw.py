"""End-to-end path from native-rate PCM16 host buffers (c2raw): time against the sub-batch size, and where the host spends it."""
import cProfile, io, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from heart_murmur_detection_b200 import pipeline, synth
C2_KW = dict(input_sec=8, butterworth_filter=5, pad=True, types="zero", max_sec=32)
dev = torch.device("cuda", 0)
lens = synth.clip_lengths("c2", 5272, seed=1234)
n4 = lens // 4
o4 = np.zeros(lens.size + 1, dtype=np.int64); np.cumsum(n4, out=o4[1:])
h_pcm = torch.randint(-2000, 2000, (int(o4[-1]),), dtype=torch.int16).pin_memory()
ub = int((1 + np.maximum(n4 * 4, 128000) // 512).sum())
h_out = torch.empty((ub, 64), dtype=torch.float32, pin_memory=True)
for cb in (1 << 30, 2 << 30, 3 << 30, 4 << 30, 8 << 30):
    for _ in range(2):
        pipeline.entire_signal_from_host(h_pcm, o4, h_out, sr_in=4000, chunk_bytes=cb, **C2_KW)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(4):
        pipeline.entire_signal_from_host(h_pcm, o4, h_out, sr_in=4000, chunk_bytes=cb, **C2_KW)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 4
    print(f"chunk_bytes={cb >> 20} MB: {dt * 1e3:.1f} ms -> {5272 / dt:.0f} clips/s", flush=True)
pr = cProfile.Profile(); pr.enable()
pipeline.entire_signal_from_host(h_pcm, o4, h_out, sr_in=4000, **C2_KW)
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(14); print(s.getvalue()[:3500])
