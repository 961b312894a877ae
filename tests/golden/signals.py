"""Bit-reproducible test signals built from integer arithmetic only.

Golden fixtures store hashes of reference outputs, so the *inputs* must be
reproducible to the bit on any machine and numpy version.  Everything here is
int64/uint64 arithmetic followed by an exact power-of-two scaling to float32.
The signal is PCG-like in shape: periodic two-burst beats (table-sine tones under
triangular envelopes), hash noise, and near-silent lead-in / lead-out.
"""
from __future__ import annotations

import math

import numpy as np

_TABLE_BITS = 10
_SIN = np.array(
    [int(round(32767 * math.sin(2 * math.pi * (k + 0.5) / (1 << _TABLE_BITS)))) for k in range(1 << _TABLE_BITS)],
    dtype=np.int64,
)


def _hash16(n: np.ndarray, seed: int) -> np.ndarray:
    """Counter-based hash -> int64 in [-32768, 32767] (splitmix64 finaliser)."""
    with np.errstate(over="ignore"):
        h = (n.astype(np.uint64) + np.uint64((seed * 0x632BE59BD9B4E019) & 0xFFFFFFFFFFFFFFFF)) * np.uint64(
            0x9E3779B97F4A7C15
        )
        h ^= h >> np.uint64(30)
        h *= np.uint64(0xBF58476D1CE4E5B9)
        h ^= h >> np.uint64(27)
        h *= np.uint64(0x94D049BB133111EB)
        h ^= h >> np.uint64(31)
    return (h >> np.uint64(48)).astype(np.int64) - 32768


def _tone(n: np.ndarray, freq_hz: float, sr: int) -> np.ndarray:
    step = int(round(freq_hz * (1 << 32) / sr))  # 32-bit phase accumulator
    phase = (n * step) & 0xFFFFFFFF
    return _SIN[phase >> (32 - _TABLE_BITS)]


def _tri(p: np.ndarray, center: int, half: int) -> np.ndarray:
    return np.maximum(0, half - np.abs(p - center))


def golden_signal(n: int, seed: int, sr: int = 16000, lead: int | None = None, tail: int | None = None) -> np.ndarray:
    """float32 [n]; peak about 0.5; noise floor about -60 dBFS; silent edges about -96 dBFS."""
    idx = np.arange(n, dtype=np.int64)
    bpm = 60 + (seed * 37) % 61
    period = (sr * 60) // bpm
    p = idx % period
    f1 = 30 + (seed * 13) % 31
    f2 = 50 + (seed * 7) % 41
    h1 = (30 * sr) // 1000 * 2
    h2 = (20 * sr) // 1000 * 2
    c1 = h1 + period // 20
    c2 = c1 + (3 * period) // 10
    x = (_tri(p, c1, h1) * _tone(idx, f1, sr)) // h1  # |.| <= 32767
    x += (6 * _tri(p, c2, h2) * _tone(idx, f2, sr)) // (10 * h2)
    if seed % 3 == 0:  # murmur-ish band noise between the bursts
        cm = (c1 + c2) // 2
        hm = max(1, (c2 - c1) // 2)
        x += (_tri(p, cm, hm) * ((_tone(idx, 150 + seed % 200, sr) + _tone(idx, 310 + seed % 90, sr)) // 40)) // hm
    x = x // 2  # peak ~0.5 of 32768
    x += _hash16(idx, seed) // 1024  # about +-32 counts ~ -60 dBFS
    if lead is None:
        lead = (seed * 811) % (sr // 2)
    if tail is None:
        tail = (seed * 467) % (sr // 2)
    if lead + tail >= n:
        lead = tail = n // 4
    quiet = _hash16(idx, seed + 1000) // 16384  # +-2 counts / 2^17 scale below
    active = (idx >= lead) & (idx < n - tail)
    # active part on a 2^-15 grid, silent part on a 2^-17 grid: both exact in float32
    out = np.where(active, x.astype(np.float64) / 32768.0, quiet.astype(np.float64) / 131072.0)
    return out.astype(np.float32)
