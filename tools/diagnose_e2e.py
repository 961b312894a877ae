"""Diagnose the end-to-end (host buffer) path: raw PCIe bandwidth vs pipeline time vs host overhead."""
import cProfile, io, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from heart_murmur_detection_b200 import frontend, pipeline, synth
C2_KW = dict(input_sec=8, butterworth_filter=5, pad=True, types="zero", max_sec=32)
dev = torch.device("cuda", 0)
lens = synth.clip_lengths("c2", 5272, seed=1234)
wav, off = synth.make_batch(lens, base_seed=0, device=dev)
total = int(off[-1])
h_wav = torch.empty(total, dtype=torch.float32, pin_memory=True); h_wav.copy_(wav)
d = torch.empty_like(wav)
for name, n in (("1GB", 1 << 28), ("all", total)):
    for _ in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        d[:n].copy_(h_wav[:n], non_blocking=True); torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    print(f"H2D {name}: {n*4/dt/1e9:.1f} GB/s ({dt*1e3:.1f} ms)")
res = pipeline.entire_signal_batch(wav, off, spectrogram=True, **C2_KW)
rows = int(res.row_offsets[-1])
h_out = torch.empty((rows, 64), dtype=torch.float32, pin_memory=True)
for _ in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    h_out.copy_(res.features[:rows], non_blocking=True); torch.cuda.synchronize()
    dt = time.perf_counter() - t0
print(f"D2H features: {rows*256/dt/1e9:.1f} GB/s ({dt*1e3:.1f} ms)")
for cb in (256 << 20, 1 << 30):
    for _ in range(2):
        pipeline.entire_signal_from_host(h_wav, off, h_out, chunk_bytes=cb, **C2_KW)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3):
        pipeline.entire_signal_from_host(h_wav, off, h_out, chunk_bytes=cb, **C2_KW)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
    print(f"from_host chunk_bytes={cb>>20}MB: {dt*1e3:.1f} ms  -> {5272/dt:.0f} clips/s")
pr = cProfile.Profile(); pr.enable()
pipeline.entire_signal_from_host(h_wav, off, h_out, **C2_KW)
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(25); print(s.getvalue()[:5000])
os.system("nvidia-smi topo -m | head -20; lscpu | grep -i 'numa\\|model name\\|^CPU(s)'")
