"""Does the result of a clip depend on which other clips share its batch?  (What the multi-GPU identity check of bench.py
sees when the ranks hold 1/8 of the batch.)  Log-mel, trim and padding are per clip; the one-pass band-pass picks its chunk
length per CALL (wave quantisation), and a chunk that starts elsewhere changes the float64 recurrence at the 1e-13 level,
which can flip a float32 rounding now and then.  Prints the differences with and without the band-pass, and with the exact
scan."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from heart_murmur_detection_b200 import dist as hd, frontend, pipeline, synth

n, world = 256, 8
lens = synth.clip_lengths("c2", n, seed=99)
wav, off = synth.make_batch(lens, base_seed=123_000_000, device="cuda")
for name, kw, algo in (("no band-pass", dict(butterworth_filter=None), "auto"), ("band-pass, auto", dict(butterworth_filter=5), "auto"),
                       ("band-pass, exact scan", dict(butterworth_filter=5), "scan")):
    frontend.default_ctx().set_iir_algo(algo)
    kw = dict(input_sec=8, pad=True, types="zero", max_sec=32, spectrogram=True, **kw)
    full = pipeline.entire_signal_batch(wav, off, **kw)
    ref = full.features[: int(full.row_offsets[-1])]
    worst, differing, plans = 0.0, 0, set()
    for rank in range(world):
        shard = hd.shard_by_length(lens, world)[rank]
        lo = np.zeros(len(shard) + 1, dtype=np.int64)
        np.cumsum(lens[shard], out=lo[1:])
        lw = torch.cat([wav[int(off[i]) : int(off[i + 1])] for i in shard])
        res = pipeline.entire_signal_batch(lw, lo, **kw)
        plans.add(str(frontend.default_ctx().last_iir_plan()))
        for k, cid in enumerate(res.chunks.clip_ids):
            g = int(shard[cid])
            a = res.features[int(res.row_offsets[k]) : int(res.row_offsets[k + 1])]
            # rows of global clip g in the full result
            kk = int(np.where(full.chunks.clip_ids == g)[0][0])
            b = ref[int(full.row_offsets[kk]) : int(full.row_offsets[kk + 1])]
            d = (a - b).abs()
            worst = max(worst, float(d.max()))
            differing += int((d > 0).sum())
    print(f"{name:24s}: max |diff| {worst:.3g}, differing elements {differing} of {ref.numel()}, plans {plans}")
frontend.default_ctx().set_iir_algo("auto")
