import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
import numpy as np, torch
from heart_murmur_detection_b200 import frontend as fe
from signals import golden_signal
lens=[1,2,31,32,33,799,800,801,5000, 20000]
clips=[golden_signal(n, seed=3+i) for i,n in enumerate(lens)]
off=np.zeros(len(clips)+1,dtype=np.int64); np.cumsum(lens,out=off[1:])
wav=torch.from_numpy(np.concatenate(clips)).cuda()
sos=fe.butter_bandpass_sos(200,1800,16000,5)
ctx=fe.Context(); ctx.set_iir_algo("overlap")
y=fe.iir_sos(wav,off,sos,ctx=ctx); torch.cuda.synchronize(); print("aligned ok", ctx.last_iir_plan())
sh=torch.empty(wav.numel()+1,dtype=torch.float32,device="cuda"); sh[1:].copy_(wav)
outs=torch.zeros(wav.numel()+1,dtype=torch.float32,device="cuda")
fe.iir_sos(sh[1:],off,sos,out=outs[1:],ctx=ctx); torch.cuda.synchronize(); print("phase1 ok", ctx.last_iir_plan())
print((outs[1:]-y).abs().max().item())
