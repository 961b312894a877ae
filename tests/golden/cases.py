"""Case tables shared by ``make_golden.py`` (runs the real reference) and the tests."""
from __future__ import annotations

import hashlib

import numpy as np

SR = 16000

# (n_samples, seed) used for the pad / split index work.  Lengths straddle every
# branch: shorter than half the target, between half and full, exactly the target,
# just over, multiples of the half-hop, and long enough for several 50 % overlap chunks.
PAD_SPLIT_LENGTHS = [1, 7, 401, 3999, 15999, 16000, 16001, 31999, 32000, 32001, 40000, 47999, 48000, 48001,
                     63999, 64000, 64001, 80000, 100000, 131071, 160000, 200001]
# desired_length in seconds -> L = int(sec * SR): 32000, 65440, 16000
PAD_SPLIT_SECS = [2, 4.09, 1]

# whole-recording cases for the composite entry points: (name, n_samples, seed, lead, tail)
RECORDINGS = [
    ("r_short", 20000, 3, 2400, 1700),       # 1.25 s: shorter than input_sec -> None / padded
    ("r_mid", 90000, 4, 0, 5200),            # 5.6 s
    ("r_8s", 128000 + 1600 + 800, 5, 1600, 800),  # trims to ~8 s
    ("r_long", 300000, 6, 7000, 9000),       # 18.75 s: several chunks
    ("r_vlong", 600000, 9, 100, 100),        # 37.5 s: exercises max_sec=32
]


# annotated cycles for get_individual_cycles_librosa on r_long: (Start s, End s, Crackles, Wheezes, Disease)
CYCLES = [(0.5, 2.1, 0, 1, "COPD"), (2.1, 5.037, 1, 0, "Healthy"), (5.037, 9.9, 1, 1, "URTI"), (17.0, 30.0, 0, 0, "Asthma"),
          (25.0, 26.0, 0, 1, "LRTI")]


# HTS-AT input stage: spectrogram lengths (frames) and the deterministic BatchNorm parameters used for the fixtures
HTSAT_T = [251, 256, 63, 1001, 1024, 1, 4]


def htsat_bn_params(F=64):
    k = np.arange(F, dtype=np.float64)
    weight = (0.75 + 0.5 * ((k * 37) % 64) / 64).astype(np.float32)
    bias = (-0.25 + 0.5 * ((k * 11) % 64) / 64).astype(np.float32)
    mean = (0.2 + 0.4 * ((k * 23) % 64) / 64).astype(np.float32)
    var = (0.02 + 0.1 * ((k * 5) % 64) / 64).astype(np.float32)
    return weight, bias, mean, var


def sha(arr) -> str:
    a = np.ascontiguousarray(arr)
    h = hashlib.sha256()
    h.update(str(a.dtype).encode())
    h.update(str(a.shape).encode())
    h.update(a.tobytes())
    return h.hexdigest()


def sha_list(arrs) -> str:
    h = hashlib.sha256()
    for a in arrs:
        h.update(sha(a).encode())
    return h.hexdigest()


def digest(arr, n_probe=48) -> dict:
    """Compact, tolerance-comparable summary of a float array (full arrays are kept
    only for a few cases to bound the fixture size)."""
    a = np.ascontiguousarray(arr)
    flat = a.reshape(-1).astype(np.float64)
    pos = (np.arange(n_probe, dtype=np.int64) * 2654435761 % max(1, flat.size)) if flat.size else np.zeros(0, np.int64)
    return {
        "shape": list(a.shape),
        "dtype": str(a.dtype),
        "sum": float(flat.sum()),
        "sumsq": float((flat * flat).sum()),
        "probe_pos": [int(p) for p in pos],
        "probe": [float(flat[p]) for p in pos],
    }


def check_digest(arr, d, rtol=1e-6, atol=1e-6):
    a = np.ascontiguousarray(arr)
    assert list(a.shape) == d["shape"], (a.shape, d["shape"])
    assert str(a.dtype) == d["dtype"], (a.dtype, d["dtype"])
    flat = a.reshape(-1).astype(np.float64)
    scale = max(1.0, float(np.abs(flat).sum()))
    assert abs(float(flat.sum()) - d["sum"]) <= rtol * scale + atol
    assert abs(float((flat * flat).sum()) - d["sumsq"]) <= rtol * max(1.0, d["sumsq"]) + atol
    if flat.size:
        np.testing.assert_allclose(flat[np.array(d["probe_pos"])], np.array(d["probe"]), rtol=rtol, atol=atol)


def hash_spec(T: int, F: int, seed: int) -> np.ndarray:
    """Bit-reproducible pseudo-spectrogram in [0, 1) on a 2^-20 grid."""
    idx = np.arange(T * F, dtype=np.uint64)
    with np.errstate(over="ignore"):
        h = (idx + np.uint64(seed * 1000003)) * np.uint64(0x9E3779B97F4A7C15)
        h ^= h >> np.uint64(29)
        h *= np.uint64(0xBF58476D1CE4E5B9)
        h ^= h >> np.uint64(32)
    v = (h >> np.uint64(44)).astype(np.float64) / float(1 << 20)
    return v.astype(np.float32).reshape(T, F)


# HeAR front-end cases: name -> (samples per clip, golden_signal seeds); clips shorter than 32 000 are zero padded
HEAR_CASES = {
    "b3": (32000, (101, 102, 103)),
    "single": (32000, (104,)),
    "short": (20000, (105, 106)),
}


# Dataset __getitem__ fixtures (make_golden_datasets.py): pseudo-spectrogram lengths and the item order of a batch
DATASET_SPECS = {
    "rows64": [251, 400, 1001, 2500, 760, 300],     # >= max_len = 251; 1001 and 2500 exceed the 3 * max_len window
    "rows128": [998, 1022, 1024, 1500, 2048, 16],   # audiomae: padded and cropped to 1024
    "order": [3, 0, 4, 1, 2, 0, 5],
}
