"""Batched, device-resident API over the hmfe C ABI.

A ragged batch is (``wav``: 1-D float32 CUDA tensor holding all clips back to back,
``offsets``: int64 numpy array of n_clips+1 sample offsets on the host).  torch is used
only for device memory and streams.
"""
from __future__ import annotations

import ctypes as C
import threading

import numpy as np
import torch

from . import _lib
from ._lib import check

OUT_MODES = {"normalised": 0, "db": 1, "power": 2}
VARIANTS = {"auto": 0, "scalar": 1, "packed": 2}


def _stream_ptr(stream=None):
    s = stream if stream is not None else torch.cuda.current_stream()
    return C.c_void_p(s.cuda_stream)


def _as_offsets(offsets) -> np.ndarray:
    o = np.ascontiguousarray(np.asarray(offsets, dtype=np.int64))
    if o.ndim != 1 or o.size < 1:
        raise ValueError("offsets must be a 1-D int64 array of n_clips+1 entries")
    return o


def _require_cuda_f32(t: torch.Tensor, name: str):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
        raise TypeError(f"{name} must be a contiguous float32 CUDA tensor (no CPU fallback)")


class LogMelPlan:
    """Fused STFT-power + mel + dB/min-max for ragged batches (src/util.py:481-501)."""

    def __init__(self, sample_rate=16000, n_mels=64, f_min=50, f_max=2000, nfft=1024, hop=512, variant="auto",
                 device=None):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.sample_rate, self.n_mels, self.nfft, self.hop = int(sample_rate), int(n_mels), int(nfft), int(hop)
        self.f_min, self.f_max = float(f_min), float(f_max)
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            check(
                _lib.hmfe_logmel_plan_create(C.byref(self._h), self.sample_rate, self.nfft, self.hop, self.n_mels,
                                             self.f_min, self.f_max, VARIANTS[variant]),
                "hmfe_logmel_plan_create",
            )

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            _lib.hmfe_logmel_plan_destroy(h)
            self._h = None

    def frame_offsets(self, offsets) -> np.ndarray:
        o = _as_offsets(offsets)
        T = 1 + np.diff(o) // self.hop
        fo = np.zeros(o.size, dtype=np.int64)
        np.cumsum(T, out=fo[1:])
        return fo

    def mel_basis(self) -> np.ndarray:
        out = np.empty((self.n_mels, self.nfft // 2 + 1), dtype=np.float32)
        check(_lib.hmfe_logmel_mel_basis(self._h, out.ctypes.data_as(C.c_void_p)))
        return out

    @property
    def last_launches(self) -> int:
        return int(_lib.hmfe_logmel_last_launches(self._h))

    def set_profile(self, enable: bool):
        check(_lib.hmfe_logmel_set_profile(self._h, int(bool(enable))))

    def profile_ms(self):
        """(stft+mel kernel ms, dB/min-max kernel ms, calls) since the last query."""
        a, f, n = C.c_double(), C.c_double(), C.c_int()
        check(_lib.hmfe_logmel_profile_ms(self._h, C.byref(a), C.byref(f), C.byref(n)))
        return a.value, f.value, n.value

    def __call__(self, wav: torch.Tensor, offsets, out: torch.Tensor | None = None, mode="normalised", stream=None):
        """Returns (out [sum T_i, n_mels] float32 CUDA, frame_offsets int64 numpy)."""
        _require_cuda_f32(wav, "wav")
        o = _as_offsets(offsets)
        if int(o[-1]) > wav.numel() or int(o[0]) < 0:
            raise ValueError("offsets exceed the wav buffer")
        fo = self.frame_offsets(o)
        if out is None:
            out = torch.empty((int(fo[-1]), self.n_mels), dtype=torch.float32, device=wav.device)
        else:
            _require_cuda_f32(out, "out")
            if out.numel() < int(fo[-1]) * self.n_mels:
                raise ValueError("out is too small")
        with torch.cuda.device(wav.device):
            check(
                _lib.hmfe_logmel_batch(self._h, C.c_void_p(wav.data_ptr()), o.ctypes.data_as(C.c_void_p), o.size - 1,
                                       C.c_void_p(out.data_ptr()), OUT_MODES[mode], _stream_ptr(stream)),
                "hmfe_logmel_batch",
            )
        return out, fo


_plans: dict = {}
_plans_lock = threading.Lock()


def logmel_plan(sample_rate=16000, n_mels=64, f_min=50, f_max=2000, nfft=1024, hop=512, variant="auto") -> LogMelPlan:
    key = ("logmel", torch.cuda.current_device(), int(sample_rate), int(n_mels), float(f_min), float(f_max), int(nfft),
           int(hop), variant)
    with _plans_lock:
        p = _plans.get(key)
        if p is None:
            p = _plans[key] = LogMelPlan(sample_rate, n_mels, f_min, f_max, nfft, hop, variant)
        return p


def logmel_from_host(plan: LogMelPlan, h_wav: torch.Tensor, offsets, h_out: torch.Tensor | None = None,
                     mode="normalised", chunk_bytes: int = 32 << 20):
    """Host-buffer entry: (pinned) host samples in, (pinned) host features out.

    The batch is cut into chunks of about ``chunk_bytes`` of samples; chunk i runs on stream
    i % 2 (H2D copy -> kernels -> D2H copy), so the copies of one chunk overlap the kernels of
    the other.  Returns (h_out [sum T_i, n_mels], frame_offsets).  Synchronous on return.
    """
    if h_wav.is_cuda or h_wav.dtype != torch.float32:
        raise TypeError("h_wav must be a float32 host tensor")
    o = _as_offsets(offsets)
    n = o.size - 1
    fo = plan.frame_offsets(o)
    if h_out is None:
        h_out = torch.empty((int(fo[-1]), plan.n_mels), dtype=torch.float32, pin_memory=True)
    dev = plan.device
    bounds = [0]
    while bounds[-1] < n:
        c0 = bounds[-1]
        c1 = int(np.searchsorted(o, o[c0] + chunk_bytes // 4, side="right")) - 1
        bounds.append(min(n, max(c0 + 1, c1)))
    max_samples = max(int(o[b] - o[a]) for a, b in zip(bounds[:-1], bounds[1:]))
    max_frames = max(int(fo[b] - fo[a]) for a, b in zip(bounds[:-1], bounds[1:]))
    cache = plan.__dict__.setdefault("_host_pipe", {})
    if cache.get("cap", (0, 0)) < (max_samples, max_frames) or cache.get("cap", (0, 0))[1] < max_frames:
        cache["wav"] = [torch.empty(max_samples, dtype=torch.float32, device=dev) for _ in range(2)]
        cache["out"] = [torch.empty((max_frames, plan.n_mels), dtype=torch.float32, device=dev) for _ in range(2)]
        cache["streams"] = [torch.cuda.Stream(device=dev) for _ in range(2)]
        cache["cap"] = (max_samples, max_frames)
    cur = torch.cuda.current_stream(dev)
    for s in cache["streams"]:
        s.wait_stream(cur)
    for i, (a, b) in enumerate(zip(bounds[:-1], bounds[1:])):
        s = cache["streams"][i % 2]
        dw, do = cache["wav"][i % 2], cache["out"][i % 2]
        ns, nf = int(o[b] - o[a]), int(fo[b] - fo[a])
        with torch.cuda.stream(s):
            dw[:ns].copy_(h_wav[int(o[a]) : int(o[b])], non_blocking=True)
            plan(dw, o[a : b + 1] - o[a], out=do, mode=mode, stream=s)
            h_out[int(fo[a]) : int(fo[b])].copy_(do[:nf], non_blocking=True)
    for s in cache["streams"]:
        s.synchronize()
    return h_out, fo
