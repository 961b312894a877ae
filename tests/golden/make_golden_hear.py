"""Golden fixtures for the HeAR mel-PCEN front-end, produced by EXECUTING the reference's own
``src/benchmark/baseline/hear/python/data_processing/audio_utils.py`` (pure torch / numpy / scipy, importable here).

    python tests/golden/make_golden_hear.py        # needs /root/reference; writes tests/golden/ref_hear.npz

The fixtures travel to the GPU box; the reference does not.
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from cases import HEAR_CASES  # noqa: E402
from signals import golden_signal  # noqa: E402

REF = "/root/reference/src/benchmark/baseline/hear/python/data_processing/audio_utils.py"


def main():
    spec = importlib.util.spec_from_file_location("ref_hear_audio_utils", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    torch.set_num_threads(1)
    out = {"mel_matrix": ref._linear_to_mel_weight_matrix().numpy()}
    for name, (n, clips) in HEAR_CASES.items():
        x = torch.from_numpy(np.stack([golden_signal(n, seed=s) for s in clips]))
        y = ref.preprocess_audio(x.clone())
        assert y.shape == (len(clips), 1, 192, 128) and y.dtype == torch.float32
        out[f"out/{name}"] = y.numpy()
    # intermediates of _mel_pcen for one case (the same statements, :357-383)
    n, clips = HEAR_CASES["b3"]
    x = torch.from_numpy(np.stack([golden_signal(n, seed=s) for s in clips])).float()
    x -= torch.min(x)
    x = x / (torch.max(x) + 1e-8)
    x = (x * 2) - 1
    stft = ref._compute_stft(x, frame_length=400, fft_length=400, frame_step=160, window_fn=torch.hann_window, pad_end=True)
    mel = torch.matmul(torch.square(torch.abs(stft)), ref._linear_to_mel_weight_matrix())
    out["mel/b3"] = mel.numpy()
    out["pcen/b3"] = ref._pcen_function(mel).numpy()
    np.savez_compressed(os.path.join(HERE, "ref_hear.npz"), **{k.replace("/", "|"): v for k, v in out.items()})
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
