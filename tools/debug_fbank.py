import sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests/golden')
from signals import golden_signal
from heart_murmur_detection_b200.frontend import FbankPlan
import torchaudio
def ref(x):
    w = torch.tensor(x - x.mean()).reshape(1, -1)
    return torchaudio.compliance.kaldi.fbank(w, channel=0, frame_length=25, htk_compat=True, sample_frequency=16000, use_energy=False, window_type="hanning", num_mel_bins=128, dither=0.0, frame_shift=10).numpy()
plan = FbankPlan()
lens = [401, 560, 32000, 128000, 160000, 163840, 400, 719, 720, 721]
clips = [golden_signal(n, seed=21 + i, lead=0, tail=0) for i, n in enumerate(lens)]
clips.append((golden_signal(50000, 40, lead=0, tail=0) + 0.3).astype(np.float32))
off = np.zeros(len(clips)+1, np.int64); np.cumsum([len(c) for c in clips], out=off[1:])
wav = torch.from_numpy(np.concatenate(clips)).cuda()
for trial in range(2):
    out, ro = plan(wav, off)
    o = out.cpu().numpy()
    for i, x in enumerate(clips):
        r = ref(x); g = o[ro[i]:ro[i+1]]; d = np.abs(g - r); idx = np.argwhere(d > 2.3e-3)
        if len(idx): print('trial', trial, 'clip', i, 'nbad', len(idx), 'rows', sorted(set(idx[:,0].tolist()))[:20], 'cols', sorted(set(idx[:,1].tolist()))[:40], [(float(g[a,b]), float(r[a,b])) for a,b in idx[:4]])
        else: print('trial', trial, 'clip', i, 'ok', d.max())
