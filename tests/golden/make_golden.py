#!/usr/bin/env python
"""Generate the golden fixtures by EXECUTING THE REAL REFERENCE CODE.

Runs only in the build container (needs ``/root/reference``); the GPU box and the
test-suite only read the committed outputs ``ref_util.json`` / ``ref_util.npz``.

``/root/reference/src/util.py`` and ``src/benchmark/baseline/extract_feature.py``
are imported unmodified.  Their third-party imports that are not installable
here are shimmed:

* ``librosa``  -> ``oracle.librosa_restated.as_librosa_module`` (numpy restatement of
  librosa 0.10.1; ``librosa.load`` serves in-memory arrays keyed by path),
* ``matplotlib`` / ``seaborn`` / ``opensmile`` -> empty stubs (plotting / unrelated).

So the fixtures pin the reference's *own* Python (pad / split / crop index work,
RNG side effects, composition order, scipy + torchaudio calls) exactly as executed;
the librosa arithmetic underneath is the restatement (parity unpinned vs librosa).

Usage:  python tests/golden/make_golden.py
"""
from __future__ import annotations

import json
import os
import random
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from cases import (CYCLES, HTSAT_T, PAD_SPLIT_LENGTHS, PAD_SPLIT_SECS, RECORDINGS, SR, digest, hash_spec,  # noqa: E402
                   htsat_bn_params, sha, sha_list)
from signals import golden_signal  # noqa: E402

from oracle import librosa_restated as lr  # noqa: E402

REFERENCE = "/root/reference"


def install_shims(store):
    def load(path, sr):
        key = os.path.basename(path)[: -len(".wav")]
        return store[key].copy(), sr

    sys.modules["librosa"] = lr.as_librosa_module(load)
    for name in ("matplotlib", "matplotlib.pyplot", "seaborn", "opensmile", "torchlibrosa", "torchlibrosa.augmentation",
                 "torchlibrosa.stft"):
        sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    # src/model/htsat/htsat.py imports these at module level; reshape_wav2img / bn0 do not use them
    sys.modules["torchlibrosa.augmentation"].SpecAugmentation = object
    sys.modules["torchlibrosa.stft"].LogmelFilterBank = object
    sys.modules["torchlibrosa.stft"].Spectrogram = object
    sys.path.insert(0, REFERENCE)


def rng_fingerprint():
    return sha(np.array(random.getstate()[1], dtype=np.uint64))


def main():
    store = {name: golden_signal(n, seed, SR, lead, tail) for name, n, seed, lead, tail in RECORDINGS}
    install_shims(store)
    import src.util as ref  # noqa: the real reference
    import src.benchmark.baseline.extract_feature as ref_ef  # noqa

    meta = {"reference_files": ["src/util.py", "src/benchmark/baseline/extract_feature.py"], "cases": {}}
    arrays = {}
    C = meta["cases"]

    # ---- pad / split index work (src/util.py:504-620, extract_feature.py:250-259) ----
    for sec in PAD_SPLIT_SECS:
        for n in PAD_SPLIT_LENGTHS:
            x = golden_signal(n, seed=n % 97, sr=SR, lead=0, tail=0)
            for types_ in ("repeat", "zero"):
                random.seed(99)
                out = ref.split_pad_sample([x, 0, 0], sec, SR, types=types_)
                C[f"split_pad/{sec}/{n}/{types_}"] = {
                    "n_chunks": len(out),
                    "sha": sha_list([o[0] for o in out]),
                    "dtypes": sorted({str(o[0].dtype) for o in out}),
                    "rng_after": rng_fingerprint(),
                }
            out = ref_ef.split_sample(x, sec, SR)
            C[f"split_sample/{sec}/{n}"] = {"n_chunks": len(out), "lens": [len(o) for o in out], "sha": sha_list(out)}
            C[f"droplast/{sec}/{n}"] = bool(ref.decide_droplast(x, SR, sec))

    # ---- spectrogram-domain ops (src/util.py:26-51) ----
    for T, F, crop in [(251, 64, 251), (400, 64, 251), (1022, 128, 1024 // 2), (63, 64, 32), (3750, 64, 251)]:
        spec = hash_spec(T, F, seed=T)
        random.seed(1000 + T)
        draws = {}
        m = ref.random_mask(spec)
        draws["mask_rows"] = [int(i) for i in np.flatnonzero(np.all(m == m[:, :1], axis=1) & (m[:, 0] != spec[:, 0]))]
        draws["mask_mean"] = float(spec.mean())
        draws["mask_sha"] = sha(m)
        c1 = ref.random_crop(m, crop_size=crop)
        c2 = ref.random_crop(m, crop_size=crop)
        draws["crop_sha"] = [sha(c1), sha(c2)]
        g1 = ref.random_multiply(c1)
        g2 = ref.random_multiply(c2)
        draws["mul_sha"] = [sha(g1), sha(g2)]
        draws["first_sha"] = sha(ref.crop_first(spec, crop_size=crop))
        draws["rng_after"] = rng_fingerprint()
        C[f"specops/{T}x{F}/{crop}"] = draws

    # ---- band-pass (src/util.py:113-126) ----
    for name in ("r_mid", "r_8s"):
        y = ref._butter_bandpass_filter(store[name], 200, 1800, SR, order=5)
        assert y.dtype == np.float64
        C[f"bandpass/{name}"] = digest(y)
        arrays[f"bandpass/{name}/head"] = y[:4096]
    b, a = ref._butter_bandpass(200, 1800, SR, order=5)
    arrays["bandpass/b"], arrays["bandpass/a"] = b, a

    # ---- log-mel (src/util.py:481-501) ----
    for name, fmax in (("r_mid", 8000), ("r_8s", 8000), ("r_8s", 2000), ("r_short", 8000)):
        out = ref.pre_process_audio_mel_t(store[name], f_max=fmax)
        C[f"logmel/{name}/{fmax}"] = digest(out)
        if name in ("r_8s", "r_short"):
            arrays[f"logmel/{name}/{fmax}"] = out
    arrays["logmel/zeros"] = ref.pre_process_audio_mel_t(np.zeros(4000, dtype=np.float32), f_max=8000)

    # ---- composite entry points (src/util.py:141-267,309-364,794-860; extract_feature.py:213-247) ----
    def record(key, out):
        if out is None:
            C[key] = {"none": True}
        elif isinstance(out, np.ndarray):
            C[key] = {"shape": list(out.shape), "dtype": str(out.dtype), "sha": sha(out)}
            if out.ndim == 2:
                C[key]["digest"] = digest(out)
        else:
            outs = [o.numpy() if hasattr(o, "numpy") else np.asarray(o) for o in out]
            C[key] = {"n": len(outs), "shapes": [list(o.shape) for o in outs], "sha": sha_list(outs)}
            C[key]["digests"] = [digest(o) for o in outs if o.ndim == 2]

    for name, *_ in RECORDINGS:
        for kw in (
            dict(input_sec=8, spectrogram=True, pad=True, types="zero", max_sec=32),      # model_util.py:161-163
            dict(input_sec=8, spectrogram=True, pad=True),                                # model_util.py:165 (repeat pad)
            dict(input_sec=8, spectrogram=True),                                          # heart_pressl.py:79 (pad=False)
            dict(input_sec=2, spectrogram=False, pad=True),                               # pascal_processing.py:169-175 shape
            dict(input_sec=8, spectrogram=True, pad=True, types="zero", max_sec=32, butterworth_filter=5),
        ):
            tag = ",".join(f"{k}={v}" for k, v in kw.items())
            record(f"entire/{name}/{tag}", ref.get_entire_signal_librosa("mem", name, **kw))
        for kw in (
            dict(input_sec=8.18, spectrogram=True),                                       # finetuning.py:1126
            dict(input_sec=4.09, spectrogram=True, trim_tail=True),
            dict(input_sec=2, spectrogram=False),
        ):
            tag = ",".join(f"{k}={v}" for k, v in kw.items())
            record(f"split/{name}/{tag}", ref.get_split_signal_librosa("mem", name, **kw))
        record(f"fbank_pad/{name}/10", ref.get_split_signal_fbank_pad("mem", name, input_sec=10, spectrogram=True))
        record(f"fbank_pad/{name}/2", ref.get_split_signal_fbank_pad("mem", name, input_sec=2, spectrogram=True))
        record(f"fbank/{name}/10", ref_ef.get_split_signal_fbank("mem", name, input_sec=10))
        record(f"segments/{name}/8", ref.get_individual_segments_librosa("mem", name, input_sec=8, spectrogram=True))
        record(f"segments_audio/{name}/4", ref.get_individual_segments_librosa("mem", name, input_sec=4))

    # ---- ICBHI cycle slicing (src/util.py:374-422) ----
    import pandas as pd

    ann = pd.DataFrame(CYCLES, columns=["Start", "End", "Crackles", "Wheezes", "Disease"])
    for split, n_cls in (("cycle", 4), ("cycle", 2), ("diagnosis", 3), ("diagnosis", 2)):
        for bw in (None, 5):
            out = ref.get_individual_cycles_librosa(split, ann, "mem", "r_long", SR, n_cls, butterworth_filter=bw)
            C[f"cycles/{split}/{n_cls}/{bw}"] = {
                "labels": [lab for _, lab in out],
                "lens": [len(a) for a, _ in out],
                "dtypes": sorted({str(a.dtype) for a, _ in out}),
                "sha": sha_list([a for a, _ in out]) if bw is None else None,
                "digests": [digest(a) for a, _ in out],
            }

    # ---- HTS-AT input stage (src/model/htsat/htsat.py:889-891 bn0 + :829-858 reshape_wav2img, executed) ----
    import torch

    import src.model.htsat.htsat as ref_htsat

    weight, bias, mean, var = htsat_bn_params(64)
    bn0 = torch.nn.BatchNorm2d(64)
    with torch.no_grad():
        bn0.weight.copy_(torch.from_numpy(weight))
        bn0.bias.copy_(torch.from_numpy(bias))
        bn0.running_mean.copy_(torch.from_numpy(mean))
        bn0.running_var.copy_(torch.from_numpy(var))
    bn0.eval()
    self_like = types.SimpleNamespace(spec_size=256, freq_ratio=4)
    for T in HTSAT_T:
        x = torch.from_numpy(hash_spec(T, 64, seed=700 + T))[None, None]
        with torch.no_grad():
            y = bn0(x.transpose(1, 3)).transpose(1, 3)  # htsat.py:889-891
            img = ref_htsat.HTSAT_Swin_Transformer.reshape_wav2img(self_like, y)
        C[f"htsat_input/{T}"] = digest(img[0, 0].numpy(), n_probe=96)

    # ---- VGGish input front-end (src/benchmark/baseline/vggish/vggish_input.py:52-125, executed) ----
    sys.path.insert(0, os.path.join(REFERENCE, "src", "benchmark", "baseline", "vggish"))
    import vggish_input as ref_vgg  # imports the reference's mel_features / vggish_params

    for name in ("r_short", "r_mid", "r_long"):
        ex = ref_vgg.waveform_to_examples(store[name], 16000)
        C[f"vggish/{name}"] = {"shape": list(ex.shape), "dtype": str(ex.dtype), "digest": digest(ex, n_probe=96)}
    arrays["vggish/r_short"] = ref_vgg.waveform_to_examples(store["r_short"], 16000).astype(np.float32)
    arrays["vggish/mel_matrix"] = ref_vgg.mel_features.spectrogram_to_mel_matrix(
        num_mel_bins=64, num_spectrogram_bins=257, audio_sample_rate=16000, lower_edge_hertz=125, upper_edge_hertz=7500)

    # trim indices straight from the shimmed call the reference makes
    for name, *_ in RECORDINGS:
        _, idx = lr.trim(store[name], frame_length=1600, hop_length=800)
        C[f"trim/{name}"] = [int(idx[0]), int(idx[1])]

    import scipy
    import torch
    import torchaudio

    meta["versions"] = {
        "numpy": np.__version__,
        "scipy": scipy.__version__,
        "torch": torch.__version__,
        "torchaudio": torchaudio.__version__,
        "librosa": "restated 0.10.1 (oracle/librosa_restated.py)",
    }
    with open(os.path.join(HERE, "ref_util.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(HERE, "ref_util.npz"), **{k.replace("/", "|"): v for k, v in arrays.items()})
    print("cases:", len(C), "arrays:", len(arrays))


if __name__ == "__main__":
    main()
