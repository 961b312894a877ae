"""Drop-in mirror of the audio front-end of the CLAP baseline (sibling front-ends):

* ``src/benchmark/baseline/msclap/CLAPWrapper.py:263-299`` (``read_audio`` + ``load_audio_into_tensor``):
  ``T.Resample(sr, 44100)`` -> flatten -> clips not longer than ``duration`` s are repeated and cut to
  ``duration * 44100`` samples, longer ones are cropped at ``random.randrange(n - L)``;
* ``src/benchmark/baseline/msclap/models/audio.py:146-175,190-196`` (Cnn14 input stage) with the constants of
  ``configs/config_2022.yml:10-17``: torchlibrosa ``Spectrogram(n_fft=1024, hop_length=320, window="hann",
  center=True, pad_mode="reflect")`` (power 2) -> ``LogmelFilterBank(sr=44100, n_mels=64, fmin=50, fmax=14000,
  ref=1.0, amin=1e-10, top_db=None)`` = ``10 log10(max(1e-10, S @ librosa.filters.mel(...).T))``.

Everything runs in the kernels of the main path: the polyphase resampler, the gather kernel (repeat / cut) and
the fused STFT-1024 + mel kernel with hop 320, reflect centre padding and the absolute-dB epilogue
(``HMFE_PAD_REFLECT``, ``HMFE_LOGMEL_OUT_DB_ABS``).  The crop start is drawn on the host from Python's global
``random`` exactly where the reference draws it, so a seeded run selects the same samples.
"""
from __future__ import annotations

import random

import numpy as np
import torch

from . import frontend as fe

# configs/config_2022.yml
SAMPLING_RATE = 44100
DURATION = 5
FMIN = 50
FMAX = 14000
HOP_SIZE = 320
MEL_BINS = 64
WINDOW_SIZE = 1024


def plan_fixed_duration(n: int, audio_duration: int, sample_rate: int):
    """Index plan of load_audio_into_tensor (CLAPWrapper.py:279-299) for a clip of n samples: a gather record
    (repeat, cut at L) or a view (crop at a start drawn from ``random.randrange``)."""
    L = int(audio_duration * sample_rate)
    if n <= 0:
        raise ValueError("empty audio (the reference divides by zero here)")
    if L >= n:
        if n == L:
            return fe._view(0, L)
        return fe._gather(L, src_start=0, period=n, a_end=L, a_phase=0, b_end=L)
    return fe._view(random.randrange(n - L), L)


def load_audio_batch(wav: torch.Tensor, offsets, sample_rate: int, audio_duration: int = DURATION, resample: bool = True,
                     target_rate: int = SAMPLING_RATE):
    """Batched load_audio_into_tensor: ragged clips at ``sample_rate`` -> [n_clips, duration * rate] float32 (CUDA).
    As in the reference the output rate is ``target_rate`` whether or not ``resample`` is set (read_audio returns
    ``resample_rate`` in both cases, CLAPWrapper.py:268-272)."""
    o = fe._as_offsets(offsets)
    if resample and target_rate != sample_rate:
        wav, o = fe.resample_plan(sample_rate, target_rate)(wav, o)
    L = int(audio_duration * target_rate)
    n_clips = o.size - 1
    out = torch.empty((n_clips, L), dtype=torch.float32, device=wav.device)
    descs = np.zeros(n_clips, dtype=fe.GATHER_DTYPE)
    for i in range(n_clips):
        ch = plan_fixed_duration(int(o[i + 1] - o[i]), audio_duration, target_rate)
        if ch[0] == "view":  # a copy of L samples is a gather with an empty repeat part
            descs[i] = (int(o[i]), i * L, L, 1, 0, 0, L, ch[1])
        else:
            _, length, src_start, period, a_end, a_phase, b_end, b_start = ch
            descs[i] = (int(o[i]) + src_start, i * L, length, period, a_end, a_phase, b_end, b_start)
    if n_clips:
        fe.gather(wav, out, descs)
    return out


def logmel_batch(audio: torch.Tensor) -> torch.Tensor:
    """Cnn14 input stage (audio.py:190-192): [batch, data_length] float32 CUDA -> [batch, 1, time_steps, 64] log-mel
    in dB, time_steps = 1 + data_length // 320."""
    if audio.dim() != 2:
        raise ValueError("audio must be [batch_size, data_length]")
    audio = audio.contiguous()
    B, n = audio.shape
    plan = fe.logmel_plan(SAMPLING_RATE, MEL_BINS, FMIN, FMAX, WINDOW_SIZE, HOP_SIZE, pad_mode="reflect")
    out, fo = plan(audio.reshape(-1), np.arange(B + 1, dtype=np.int64) * n, mode="db_abs")
    return out.view(B, 1, 1 + n // HOP_SIZE, MEL_BINS)


def preprocess_audio(clips, sample_rates, resample: bool = True) -> torch.Tensor:
    """CLAPWrapper.preprocess_audio (:301-314) for decoded clips instead of paths: list of 1-D (or [channels, n],
    flattened like the reference's reshape(-1)) float arrays with their sample rates -> [n, 1, 220500] CUDA."""
    outs = []
    for x, sr in zip(clips, sample_rates):
        x = torch.as_tensor(np.ascontiguousarray(np.asarray(x, dtype=np.float32))).cuda()
        outs.append(_one(x, int(sr), resample))
    return torch.stack(outs).unsqueeze(1) if outs else torch.empty((0, 1, DURATION * SAMPLING_RATE), device="cuda")


def _one(x: torch.Tensor, sr: int, resample: bool) -> torch.Tensor:
    if x.dim() == 2:  # the reference resamples every channel, then flattens channel after channel (:270-271,278)
        if resample and sr != SAMPLING_RATE:
            n = x.shape[1]
            y, _ = fe.resample_plan(sr, SAMPLING_RATE)(x.reshape(-1), np.arange(x.shape[0] + 1, dtype=np.int64) * n)
            x, sr = y, SAMPLING_RATE
        else:
            x = x.reshape(-1)
        resample = False
    return load_audio_batch(x, [0, x.numel()], sr, DURATION, resample)[0]
