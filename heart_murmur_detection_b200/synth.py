"""Synthetic PCG-like (phonocardiogram) recordings for tests and benchmarks.

Shapes follow BASELINE.json's configs.  Every clip is a
periodic train of S1/S2 heart sounds (Gaussian-windowed tones), an optional
band-limited systolic murmur (sum of random 100-400 Hz tones), white noise at
-60 dBFS and U(0,0.5) s of -90 dBFS "silence" either side so that the silence
trim is non-trivial.  Per-clip scalars come from ``numpy.random.default_rng(seed)``
(``seed = base_seed + clip_index``), so a clip is reproducible on any device;
the noise comes from a torch generator on the target device.

This is plumbing (input generation), not part of the measured path.
"""
from __future__ import annotations

import math

import numpy as np
import torch

SR = 16000


def clip_lengths(config: str, n_clips: int, seed: int = 1234) -> np.ndarray:
    """Sample counts per clip for a BASELINE.json config.

    c1: fixed 8 s.  c2: log-uniform 2-80 s (CirCor/PhysioNet16-shaped ragged).
    c3: fixed 10.24 s.  c4: uniform 8-60 s recordings.
    """
    rng = np.random.default_rng(seed)
    if config == "c1":
        return np.full(n_clips, 8 * SR, dtype=np.int64)
    if config == "c2":
        sec = np.exp(rng.uniform(math.log(2.0), math.log(80.0), size=n_clips))
        return np.maximum(1, np.round(sec * SR)).astype(np.int64)
    if config == "c3":
        return np.full(n_clips, 163840, dtype=np.int64)
    if config == "c4":
        return np.round(rng.uniform(8.0, 60.0, size=n_clips) * SR).astype(np.int64)
    raise ValueError(config)


def _beat_template(rng: np.random.Generator, period: int) -> np.ndarray:
    """One cardiac cycle of ``period`` samples (float64)."""
    t = np.arange(period, dtype=np.float64) / SR
    per = period / SR

    def burst(center, f, sigma, amp):
        # wrap-around Gaussian so the template is exactly periodic
        d = (t - center + per / 2) % per - per / 2
        return amp * np.exp(-0.5 * (d / sigma) ** 2) * np.sin(2 * np.pi * f * d)

    x = burst(0.08 * per + 0.03, rng.uniform(30, 60), 0.030, 1.0)
    x += burst(0.08 * per + 0.03 + 0.30 * per, rng.uniform(50, 90), 0.020, 0.6)
    if rng.uniform() < 0.3:  # systolic murmur between S1 and S2
        c = 0.08 * per + 0.03 + 0.15 * per
        d = (t - c + per / 2) % per - per / 2
        env = np.exp(-0.5 * (d / (0.08 * per)) ** 2)
        f = rng.uniform(100, 400, size=8)
        ph = rng.uniform(0, 2 * np.pi, size=8)
        m = np.sin(2 * np.pi * f[:, None] * t[None, :] + ph[:, None]).sum(0) / math.sqrt(8)
        x += 0.05 * env * m
    return x / np.max(np.abs(x)) * 0.5


def make_clip(n_samples: int, seed: int, device="cpu", generator: torch.Generator | None = None) -> torch.Tensor:
    """One float32 clip of ``n_samples`` samples on ``device``."""
    rng = np.random.default_rng(seed)
    bpm = rng.uniform(60, 120)
    period = int(round(SR * 60.0 / bpm))
    tmpl = torch.from_numpy(_beat_template(rng, period).astype(np.float32)).to(device)
    lead = int(rng.uniform(0, 0.5) * SR)
    tail = int(rng.uniform(0, 0.5) * SR)
    if lead + tail >= n_samples:  # very short clip: keep at least half active
        lead = tail = n_samples // 4
    phase = int(rng.integers(0, period))
    idx = (torch.arange(n_samples, device=device) + phase) % period
    x = tmpl[idx]
    if generator is None:
        generator = torch.Generator(device=device)
        generator.manual_seed(seed)
    noise = torch.randn(n_samples, device=device, generator=generator)
    active = torch.zeros(n_samples, device=device, dtype=torch.bool)
    active[lead : n_samples - tail] = True
    x = torch.where(active, x + 1e-3 * noise, (10 ** (-90 / 20)) * noise)
    return x.to(torch.float32)


def make_batch(lengths, base_seed: int = 0, device="cpu"):
    """Ragged batch: (concatenated float32 samples, int64 offsets[n+1] as numpy)."""
    lengths = np.asarray(lengths, dtype=np.int64)
    offsets = np.zeros(len(lengths) + 1, dtype=np.int64)
    np.cumsum(lengths, out=offsets[1:])
    wav = torch.empty(int(offsets[-1]), dtype=torch.float32, device=device)
    gen = torch.Generator(device=device)
    gen.manual_seed(base_seed)
    for i, n in enumerate(lengths):
        wav[offsets[i] : offsets[i + 1]] = make_clip(int(n), base_seed + i, device=device, generator=gen)
    return wav, offsets
