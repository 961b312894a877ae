"""Drop-in mirror of the VGGish input front-end of the reference (sibling front-ends):
``src/benchmark/baseline/vggish/vggish_input.py:52-125`` (waveform_to_examples) on top of
``mel_features.py`` (:35-83 frame, :85-122 periodic_hann, :125-170 stft_magnitude, :172-195 hertz_to_mel,
:196-340 spectrogram_to_mel_matrix, :342-400 log_mel_spectrogram) with the constants of ``vggish_params.py``.

25 ms / 10 ms frames at 16 kHz, periodic Hann, 512-point FFT, MAGNITUDE spectrum, 64 HTK-mel bands
125-7500 Hz, log(mel + 0.01), examples of 96 frames without overlap.  The framing, FFT, mel and log
run in the Kaldi-fbank kernel with these constants (``frontend.FbankPlan.custom``); window and mel
matrix are built on the host in float64 exactly as the reference does and rounded once to float32.
The reference computes in float64 and returns float64: values agree to <= 1e-4 in the log domain (the
offset 0.01 amplifies the float32 FFT's absolute error near the floor), dtype is kept.
"""
from __future__ import annotations

import threading

import numpy as np
import torch

from . import frontend as fe

# vggish_params.py
NUM_FRAMES = 96
NUM_BANDS = 64
SAMPLE_RATE = 16000
STFT_WINDOW_LENGTH_SECONDS = 0.025
STFT_HOP_LENGTH_SECONDS = 0.010
NUM_MEL_BINS = NUM_BANDS
MEL_MIN_HZ = 125
MEL_MAX_HZ = 7500
LOG_OFFSET = 0.01
EXAMPLE_WINDOW_SECONDS = 0.96
EXAMPLE_HOP_SECONDS = 0.96


def periodic_hann(window_length):
    return 0.5 - 0.5 * np.cos(2 * np.pi / window_length * np.arange(window_length))


def hertz_to_mel(frequencies_hertz):
    return 1127.0 * np.log(1.0 + (np.asarray(frequencies_hertz, dtype=np.float64) / 700.0))


def spectrogram_to_mel_matrix(num_mel_bins=20, num_spectrogram_bins=129, audio_sample_rate=8000, lower_edge_hertz=125.0,
                              upper_edge_hertz=3800.0):
    """[num_spectrogram_bins, num_mel_bins] triangles in the mel domain, DC row zero (mel_features.py:196-340)."""
    nyquist = audio_sample_rate / 2.0
    if lower_edge_hertz < 0.0:
        raise ValueError("lower_edge_hertz %.1f must be >= 0" % lower_edge_hertz)
    if lower_edge_hertz >= upper_edge_hertz:
        raise ValueError("lower_edge_hertz %.1f >= upper_edge_hertz %.1f" % (lower_edge_hertz, upper_edge_hertz))
    if upper_edge_hertz > nyquist:
        raise ValueError("upper_edge_hertz %.1f is greater than Nyquist %.1f" % (upper_edge_hertz, nyquist))
    bins_mel = hertz_to_mel(np.linspace(0.0, nyquist, num_spectrogram_bins))
    edges = np.linspace(hertz_to_mel(lower_edge_hertz), hertz_to_mel(upper_edge_hertz), num_mel_bins + 2)
    w = np.empty((num_spectrogram_bins, num_mel_bins))
    for i in range(num_mel_bins):
        lo, mid, hi = edges[i : i + 3]
        w[:, i] = np.maximum(0.0, np.minimum((bins_mel - lo) / (mid - lo), (hi - bins_mel) / (hi - mid)))
    w[0, :] = 0.0
    return w


_plan_lock = threading.Lock()
_plans: dict = {}


def _plan():
    key = (torch.cuda.current_device(),)
    with _plan_lock:
        p = _plans.get(key)
        if p is None:
            win = int(round(SAMPLE_RATE * STFT_WINDOW_LENGTH_SECONDS))
            hop = int(round(SAMPLE_RATE * STFT_HOP_LENGTH_SECONDS))
            fft_length = 2 ** int(np.ceil(np.log(win) / np.log(2.0)))
            assert (win, hop, fft_length) == (400, 160, 512)
            mel = spectrogram_to_mel_matrix(num_mel_bins=NUM_MEL_BINS, num_spectrogram_bins=fft_length // 2 + 1,
                                            audio_sample_rate=SAMPLE_RATE, lower_edge_hertz=MEL_MIN_HZ,
                                            upper_edge_hertz=MEL_MAX_HZ)
            p = _plans[key] = fe.FbankPlan.custom(periodic_hann(win), mel.T, sample_rate=SAMPLE_RATE, shift=hop,
                                                  magnitude=True, log_offset=LOG_OFFSET)
        return p


def log_mel_spectrogram_batch(wav: torch.Tensor, offsets):
    """log_mel_spectrogram (mel_features.py:342-400) with the VGGish constants over a ragged batch on the GPU.
    Returns (float32 CUDA tensor [sum frames, 64], row offsets)."""
    return _plan()(wav, offsets)


def waveform_to_examples(data, sample_rate):
    """[num_examples, 96, 64] float64 (vggish_input.py:52-125).  As in the reference the resampling branch is
    commented out: ``data`` is taken to be at 16 kHz whatever ``sample_rate`` says."""
    data = np.asarray(data)
    if len(data.shape) > 1:
        data = np.mean(data, axis=1)
    wav = torch.from_numpy(np.ascontiguousarray(data, dtype=np.float32)).cuda()
    log_mel, _ = log_mel_spectrogram_batch(wav, np.array([0, wav.numel()], dtype=np.int64))
    log_mel = log_mel.cpu().numpy().astype(np.float64)
    features_sample_rate = 1.0 / STFT_HOP_LENGTH_SECONDS
    window = int(round(EXAMPLE_WINDOW_SECONDS * features_sample_rate))
    hop = int(round(EXAMPLE_HOP_SECONDS * features_sample_rate))
    n = 1 + int(np.floor((log_mel.shape[0] - window) / hop))
    n = max(n, 0)
    return np.stack([log_mel[i * hop : i * hop + window] for i in range(n)]) if n else np.zeros((0, window, NUM_BANDS))
