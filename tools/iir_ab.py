"""A/B of the one-pass band-pass kernel (run on the GPU box): HMFE_IIR_CONV=0|1|3 python tools/iir_ab.py  (the integer-conversion variants 1 and 3 are compiled only with -DHMFE_IIR_BUILD_CONV_VARIANTS; HMFE_IIR_PIPE=1 selects the pipelined cascade)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from heart_murmur_detection_b200 import frontend as fe, synth

lens = synth.clip_lengths("c2", 5272, seed=1234)
wav, off = synth.make_batch(lens, base_seed=0, device="cuda")
sos = fe.butter_bandpass_sos(200, 1800, 16000, order=5)
ctx = fe.default_ctx()
out = torch.empty_like(wav)
for _ in range(3):
    fe.iir_sos_trim(wav, off, sos, out=out, ctx=ctx)
torch.cuda.synchronize()
ctx.set_profile(True)
for _ in range(10):
    _, se = fe.iir_sos_trim(wav, off, sos, out=out, ctx=ctx)
ms = {k: v[0] / v[1] for k, v in ctx.profile_ms().items() if v[1]}
print("HMFE_IIR_CONV", os.environ.get("HMFE_IIR_CONV", "default"), "kernels ms:", {k: round(v, 4) for k, v in ms.items()},
      "checksum", float(out.double().abs().sum()), "se sum", int(se.sum()))
path = "gpurun_out/iir_ab_ref.pt"
if os.environ.get("HMFE_IIR_CONV", "0") == "0":
    torch.save(out[: 50_000_000].cpu(), path)
elif os.path.exists(path):
    ref = torch.load(path)
    d = (out[: 50_000_000].cpu() - ref).abs()
    print("   vs conversion instructions: max |diff| =", float(d.max()), "differing samples:", int((d > 0).sum()), "of", d.numel())
