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

OUT_MODES = {"normalised": 0, "db": 1, "power": 2, "db_abs": 3}
PAD_MODES = {"constant": 0, "reflect": 1}
VARIANTS = {"auto": 0, "scalar": 1, "packed": 2, "pair": 3, "tc": 4}


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
                 device=None, pad_mode="constant"):
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
            self.pad_mode = pad_mode
            if pad_mode != "constant":
                check(_lib.hmfe_logmel_plan_set_pad_mode(self._h, PAD_MODES[pad_mode]), "hmfe_logmel_plan_set_pad_mode")

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and _lib is not None:  # module globals are gone at interpreter shutdown
            try:
                _lib.hmfe_logmel_plan_destroy(h)
            except TypeError:  # interpreter shutdown: the ctypes entry is already gone
                pass
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

    def tc_status(self) -> int:
        """Protocol-error word of the tensor-core variant's bounded waits (0 = fine); synchronises."""
        w = C.c_uint32()
        check(_lib.hmfe_logmel_tc_status(self._h, C.byref(w)), "hmfe_logmel_tc_status")
        return int(w.value)

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


def logmel_plan(sample_rate=16000, n_mels=64, f_min=50, f_max=2000, nfft=1024, hop=512, variant="auto",
                pad_mode="constant") -> LogMelPlan:
    key = ("logmel", torch.cuda.current_device(), int(sample_rate), int(n_mels), float(f_min), float(f_max), int(nfft),
           int(hop), variant, pad_mode)
    with _plans_lock:
        p = _plans.get(key)
        if p is None:
            p = _plans[key] = LogMelPlan(sample_rate, n_mels, f_min, f_max, nfft, hop, variant, pad_mode=pad_mode)
        return p


def logmel_from_host(plan: LogMelPlan, h_wav: torch.Tensor, offsets, h_out: torch.Tensor | None = None,
                     mode="normalised", chunk_bytes: int = 32 << 20):
    """Host-buffer entry: (pinned) host samples in, (pinned) host features out.

    The batch is cut into chunks of about ``chunk_bytes`` of samples; chunk i runs on stream
    i % 2 (H2D copy -> kernels -> D2H copy), so the copies of one chunk overlap the kernels of
    the other.  Returns (h_out [sum T_i, n_mels], frame_offsets).  Synchronous on return.
    """
    if h_wav.is_cuda or h_wav.dtype not in (torch.float32, torch.int16):
        raise TypeError("h_wav must be a float32 (samples) or int16 (16-bit PCM payload) host tensor")
    pcm = h_wav.dtype == torch.int16
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
        cache["pcm"] = [torch.empty(max_samples, dtype=torch.int16, device=dev) for _ in range(2)]
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
            if pcm:  # 2-byte payload over PCIe, / 32768 on the device
                dp = cache["pcm"][i % 2]
                dp[:ns].copy_(h_wav[int(o[a]) : int(o[b])], non_blocking=True)
                pcm16_to_f32(dp[:ns], out=dw, stream=s)
            else:
                dw[:ns].copy_(h_wav[int(o[a]) : int(o[b])], non_blocking=True)
            plan(dw, o[a : b + 1] - o[a], out=do, mode=mode, stream=s)
            h_out[int(fo[a]) : int(fo[b])].copy_(do[:nf], non_blocking=True)
    for s in cache["streams"]:
        s.synchronize()
    return h_out, fo


# =============================================================================================
# Context for plan-less stages
# =============================================================================================


class Context:
    """hmfe_ctx: descriptor staging + device scratch for trim / gather / IIR / spectrogram ops."""

    def __init__(self, device=None):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            check(_lib.hmfe_ctx_create(C.byref(self._h)), "hmfe_ctx_create")

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and _lib is not None:  # module globals are gone at interpreter shutdown
            try:
                _lib.hmfe_ctx_destroy(h)
            except TypeError:  # interpreter shutdown: the ctypes entry is already gone
                pass
            self._h = None

    @property
    def last_launches(self) -> int:
        return int(_lib.hmfe_ctx_last_launches(self._h))

    def set_profile(self, enable: bool):
        check(_lib.hmfe_ctx_set_profile(self._h, int(bool(enable))))

    def use_workspace(self, workspace: torch.Tensor | None, max_clips: int | None = None):
        """Caller-provided memory (include/hmfe.h): ``workspace`` (uint8 CUDA tensor, kept alive here) becomes the
        scratch of every stage run on this context; with ``max_clips`` the descriptor staging is reserved once and
        no later call allocates.  Sizes: ``trim_workspace_bytes`` / ``iir_workspace_bytes``."""
        self._ws = workspace
        ptr, n = (C.c_void_p(workspace.data_ptr()), workspace.numel()) if workspace is not None else (C.c_void_p(), 0)
        with torch.cuda.device(self.device):
            check(_lib.hmfe_ctx_set_workspace(self._h, ptr, n), "hmfe_ctx_set_workspace")
            if max_clips is not None:
                check(_lib.hmfe_ctx_reserve(self._h, int(max_clips)), "hmfe_ctx_reserve")

    def set_iir_algo(self, algo: str = "auto"):
        """"auto" | "scan" (exact chunked scan) | "overlap" (one pass with warm-up) - include/hmfe.h."""
        check(_lib.hmfe_ctx_set_iir_algo(self._h, _lib.IIR_ALGOS[algo]), "hmfe_ctx_set_iir_algo")

    def set_iir_rows(self, rows: str = "auto"):
        """"auto" | "scalar" (force the 4-byte row variant of the overlap kernel)."""
        check(_lib.hmfe_ctx_set_iir_rows(self._h, _lib.IIR_ROWS[rows]), "hmfe_ctx_set_iir_rows")

    def last_iir_plan(self) -> dict:
        a, c, w, r = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        check(_lib.hmfe_ctx_last_iir_plan(self._h, C.byref(a), C.byref(c), C.byref(w), C.byref(r)))
        return {"algo": {v: k for k, v in _lib.IIR_ALGOS.items()}.get(a.value, "none"), "chunk": c.value,
                "warmup": w.value, "rows": {v: k for k, v in _lib.IIR_ROWS.items()}.get(r.value, "none")}

    def profile_ms(self) -> dict:
        """{kernel name: (total ms, launches)} since the last query (synchronises)."""
        n = len(_lib.KERNEL_NAMES)
        ms, cnt = (C.c_double * n)(), (C.c_int * n)()
        check(_lib.hmfe_ctx_profile_ms(self._h, ms, cnt))
        return {k: (ms[i], cnt[i]) for i, k in enumerate(_lib.KERNEL_NAMES)}


def trim_workspace_bytes(offsets, frame_length=1600, hop_length=800) -> int:
    o = _as_offsets(offsets)
    return int(_lib.hmfe_trim_workspace_bytes(o.ctypes.data_as(C.c_void_p), o.size - 1, int(frame_length), int(hop_length)))


def iir_workspace_bytes(offsets, n_sections=5, hop_length=800) -> int:
    o = _as_offsets(offsets)
    return int(_lib.hmfe_iir_workspace_bytes(o.ctypes.data_as(C.c_void_p), o.size - 1, int(n_sections), int(hop_length)))


_ctxs: dict = {}


def default_ctx() -> Context:
    key = (torch.cuda.current_device(), threading.get_ident())
    with _plans_lock:
        c = _ctxs.get(key)
        if c is None:
            c = _ctxs[key] = Context()
        return c


# =============================================================================================
# Butterworth band-pass design (host, float64) - the coefficients scipy.signal.butter returns
# for src/util.py:113-119, derived from the textbook construction (analog prototype ->
# band-pass transform -> bilinear transform).
# =============================================================================================


def butter_bandpass_zpk(lowcut, highcut, fs, order=5):
    nyq = 0.5 * fs
    wn = np.array([lowcut / nyq, highcut / nyq], dtype=np.float64)
    if not (0 < wn[0] < wn[1] < 1):
        raise ValueError("band edges must satisfy 0 < low < high < fs/2")
    n = int(order)
    m = np.arange(-n + 1, n, 2)
    p = -np.exp(1j * np.pi * m / (2 * n))  # analog Butterworth prototype, unit cutoff
    fs_d = 2.0
    warped = 2 * fs_d * np.tan(np.pi * wn / fs_d)
    bw, wo = warped[1] - warped[0], np.sqrt(warped[0] * warped[1])
    p_lp = p * bw / 2
    root = np.sqrt(p_lp.astype(complex) ** 2 - wo**2)
    p_bp = np.concatenate([p_lp + root, p_lp - root])
    z_bp = np.zeros(n)
    k_bp = bw**n
    fs2 = 2 * fs_d
    z_d = np.concatenate([(fs2 + z_bp) / (fs2 - z_bp), -np.ones(len(p_bp) - len(z_bp))])
    p_d = (fs2 + p_bp) / (fs2 - p_bp)
    k_d = k_bp * np.real(np.prod(fs2 - z_bp) / np.prod(fs2 - p_bp))
    return z_d, p_d, float(k_d)


def butter_bandpass_ba(lowcut, highcut, fs, order=5):
    """(b, a) transfer-function coefficients - what ``_butter_bandpass`` returns."""
    z, p, k = butter_bandpass_zpk(lowcut, highcut, fs, order)
    b = k * np.real(np.poly(z))
    a = np.real(np.poly(p))
    return b, a


def butter_bandpass_sos(lowcut, highcut, fs, order=5) -> np.ndarray:
    """Second-order sections [order, 6]: zeros (+1, -1) and one conjugate pole pair per section."""
    z, p, k = butter_bandpass_zpk(lowcut, highcut, fs, order)
    tol = 1e-12
    upper = p[np.imag(p) > tol]
    upper = upper[np.argsort(np.abs(upper))]
    real = np.sort(np.real(p[np.abs(np.imag(p)) <= tol]))
    if 2 * len(upper) + len(real) != len(p) or len(real) % 2:
        raise ValueError("unexpected pole layout in a Butterworth band-pass design")
    dens = [[1.0, -2.0 * np.real(q), np.abs(q) ** 2] for q in upper]
    dens += [[1.0, -(real[i] + real[i + 1]), real[i] * real[i + 1]] for i in range(0, len(real), 2)]
    sos = np.zeros((len(dens), 6), dtype=np.float64)
    for i, den in enumerate(dens):
        sos[i] = [1.0, 0.0, -1.0] + den
    sos[0, :3] *= k
    return sos


# =============================================================================================
# Device stages
# =============================================================================================


def iir_sos(wav: torch.Tensor, offsets, sos: np.ndarray, out: torch.Tensor | None = None, out_dtype=torch.float32,
            ctx: Context | None = None, stream=None) -> torch.Tensor:
    """Causal zero-state SOS cascade in float64 (``lfilter`` semantics) over a ragged batch."""
    _require_cuda_f32(wav, "wav")
    o = _as_offsets(offsets)
    sos = np.ascontiguousarray(sos, dtype=np.float64)
    ctx = ctx or default_ctx()
    if out is None:
        out = torch.empty(wav.numel(), dtype=out_dtype, device=wav.device)
    y32 = C.c_void_p(out.data_ptr()) if out.dtype == torch.float32 else C.c_void_p()
    y64 = C.c_void_p(out.data_ptr()) if out.dtype == torch.float64 else C.c_void_p()
    if not (y32 or y64):
        raise TypeError("out must be float32 or float64")
    with torch.cuda.device(wav.device):
        check(
            _lib.hmfe_iir_sos_batch(ctx._h, C.c_void_p(wav.data_ptr()), o.ctypes.data_as(C.c_void_p), o.size - 1,
                                    sos.ctypes.data_as(C.c_void_p), sos.shape[0], y32, y64, _stream_ptr(stream)),
            "hmfe_iir_sos_batch",
        )
    return out


def iir_sos_trim(wav: torch.Tensor, offsets, sos: np.ndarray, out: torch.Tensor | None = None, frame_length=1600,
                 hop_length=800, top_db=60.0, out64: torch.Tensor | None = None, ctx: Context | None = None, stream=None):
    """Band-pass + silence-trim indices of the filtered signal in one call (src/util.py:226-244).

    Returns (filtered float32 signal, int64 CUDA tensor [n_clips, 2] of clip-relative (start, end))."""
    _require_cuda_f32(wav, "wav")
    o = _as_offsets(offsets)
    sos = np.ascontiguousarray(sos, dtype=np.float64)
    ctx = ctx or default_ctx()
    if out is None:
        out = torch.empty(wav.numel(), dtype=torch.float32, device=wav.device)
    if out.dtype != torch.float32:
        raise TypeError("out must be float32 (pass out64 for the float64 copy)")
    y64 = C.c_void_p(out64.data_ptr()) if out64 is not None else C.c_void_p()
    se = torch.empty((o.size - 1, 2), dtype=torch.int64, device=wav.device)
    with torch.cuda.device(wav.device):
        check(
            _lib.hmfe_iir_sos_trim_batch(ctx._h, C.c_void_p(wav.data_ptr()), o.ctypes.data_as(C.c_void_p), o.size - 1,
                                         sos.ctypes.data_as(C.c_void_p), sos.shape[0], C.c_void_p(out.data_ptr()), y64,
                                         int(frame_length), int(hop_length), float(top_db), C.c_void_p(se.data_ptr()),
                                         _stream_ptr(stream)),
            "hmfe_iir_sos_trim_batch",
        )
    return out, se


def sosfiltfilt(wav: torch.Tensor, offsets, sos: np.ndarray, out_dtype=torch.float64, padlen: int | None = None,
                ctx: Context | None = None, stream=None) -> torch.Tensor:
    """Zero-phase SOS filtering of every clip: ``scipy.signal.sosfiltfilt(sos, x)`` (odd padding).
    The reference's own band-pass is the causal ``iir_sos`` (lfilter); this is the extra mode."""
    _require_cuda_f32(wav, "wav")
    o = _as_offsets(offsets)
    sos = np.ascontiguousarray(sos, dtype=np.float64)
    ctx = ctx or default_ctx()
    pl_ = -1 if padlen is None else int(padlen)
    edge = int(_lib.hmfe_sosfiltfilt_padlen(sos.ctypes.data_as(C.c_void_p), sos.shape[0])) if pl_ < 0 else pl_
    nbytes = int(_lib.hmfe_sosfiltfilt_workspace_bytes(o.ctypes.data_as(C.c_void_p), o.size - 1, edge))
    work = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=wav.device)
    out = torch.empty(wav.numel(), dtype=out_dtype, device=wav.device)
    y32 = C.c_void_p(out.data_ptr()) if out.dtype == torch.float32 else C.c_void_p()
    y64 = C.c_void_p(out.data_ptr()) if out.dtype == torch.float64 else C.c_void_p()
    if not (y32 or y64):
        raise TypeError("out_dtype must be float32 or float64")
    with torch.cuda.device(wav.device):
        check(
            _lib.hmfe_sosfiltfilt_batch(ctx._h, C.c_void_p(wav.data_ptr()), o.ctypes.data_as(C.c_void_p), o.size - 1,
                                        sos.ctypes.data_as(C.c_void_p), sos.shape[0], pl_, C.c_void_p(work.data_ptr()), nbytes,
                                        y32, y64, _stream_ptr(stream)),
            "hmfe_sosfiltfilt_batch",
        )
    return out


def pcm16_to_f32(pcm: torch.Tensor, out: torch.Tensor | None = None, stream=None) -> torch.Tensor:
    """int16 CUDA samples -> float32 / 32768 (soundfile's PCM16 convention, exact)."""
    if not (isinstance(pcm, torch.Tensor) and pcm.is_cuda and pcm.dtype == torch.int16 and pcm.is_contiguous()):
        raise TypeError("pcm must be a contiguous int16 CUDA tensor (no CPU fallback)")
    if out is None:
        out = torch.empty(pcm.numel(), dtype=torch.float32, device=pcm.device)
    _require_cuda_f32(out, "out")
    if out.numel() < pcm.numel():
        raise ValueError("out is too small")
    with torch.cuda.device(pcm.device):
        check(_lib.hmfe_pcm16_decode(C.c_void_p(pcm.data_ptr()), pcm.numel(), C.c_void_p(out.data_ptr()),
                                     _stream_ptr(stream)), "hmfe_pcm16_decode")
    return out


def trim_indices(wav: torch.Tensor, offsets, frame_length=1600, hop_length=800, top_db=60.0, ctx: Context | None = None,
                 stream=None) -> torch.Tensor:
    """int64 CUDA tensor [n_clips, 2] of clip-relative (start, end) - ``librosa.effects.trim`` indices."""
    _require_cuda_f32(wav, "wav")
    o = _as_offsets(offsets)
    ctx = ctx or default_ctx()
    se = torch.empty((o.size - 1, 2), dtype=torch.int64, device=wav.device)
    with torch.cuda.device(wav.device):
        check(
            _lib.hmfe_trim_batch(ctx._h, C.c_void_p(wav.data_ptr()), o.ctypes.data_as(C.c_void_p), o.size - 1,
                                 int(frame_length), int(hop_length), float(top_db), C.c_void_p(se.data_ptr()),
                                 _stream_ptr(stream)),
            "hmfe_trim_batch",
        )
    return se


GATHER_DTYPE = np.dtype(
    [("src_off", "<i8"), ("dst_off", "<i8"), ("len", "<i4"), ("period", "<i4"), ("a_end", "<i4"), ("a_phase", "<i4"),
     ("b_end", "<i4"), ("b_start", "<i4")]
)
assert GATHER_DTYPE.itemsize == C.sizeof(_lib.GatherDesc)

CROP_DTYPE = np.dtype([("src_row", "<i8"), ("n_rows", "<i4"), ("spec_id", "<i4"), ("gain", "<f4"), ("mask_off", "<i4")])
assert CROP_DTYPE.itemsize == C.sizeof(_lib.CropDesc)


def gather(src: torch.Tensor, dst: torch.Tensor, descs: np.ndarray, ctx: Context | None = None, stream=None):
    """Apply host-built ``hmfe_gather_desc`` records (offsets are element offsets into src / dst)."""
    _require_cuda_f32(src, "src")
    _require_cuda_f32(dst, "dst")
    descs = np.ascontiguousarray(descs, dtype=GATHER_DTYPE)
    ctx = ctx or default_ctx()
    with torch.cuda.device(src.device):
        check(
            _lib.hmfe_gather_batch(ctx._h, C.c_void_p(src.data_ptr()), C.c_void_p(dst.data_ptr()),
                                   descs.ctypes.data_as(C.c_void_p), descs.size, _stream_ptr(stream)),
            "hmfe_gather_batch",
        )
    return dst


class FbankPlan:
    """Kaldi fbank for ragged batches (src/util.py:845-856; extract_feature.py:232-243)."""

    def __init__(self, sample_rate=16000, frame_length=25.0, frame_shift=10.0, num_mel_bins=128, low_freq=20.0,
                 high_freq=0.0, preemphasis=0.97, device=None):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.n_mels = int(num_mel_bins)
        self.win = int(sample_rate * frame_length * 0.001)
        self.shift = int(sample_rate * frame_shift * 0.001)
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            check(
                _lib.hmfe_fbank_plan_create(C.byref(self._h), int(sample_rate), float(frame_length), float(frame_shift),
                                            self.n_mels, float(low_freq), float(high_freq), float(preemphasis)),
                "hmfe_fbank_plan_create",
            )

    @classmethod
    def custom(cls, window, mel, sample_rate=16000, shift=160, remove_dc=False, preemphasis=0.0, magnitude=False,
               log_offset=None, device=None):
        """Same kernel with the caller's window [win] and dense mel matrix [n_mels, 257] (one contiguous band
        per row): sibling front-ends such as VGGish (periodic Hann, |X|, log(x + 0.01))."""
        self = cls.__new__(cls)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        window = np.ascontiguousarray(window, dtype=np.float32)
        mel = np.ascontiguousarray(mel, dtype=np.float32)
        if mel.ndim != 2 or mel.shape[1] != 257:
            raise ValueError("mel must be [n_mels, 257] (512-point FFT)")
        self.n_mels, self.win, self.shift = int(mel.shape[0]), int(window.size), int(shift)
        flags = (_lib.FB_REMOVE_DC if remove_dc else 0) | (_lib.FB_MAGNITUDE if magnitude else 0) | (
            _lib.FB_LOG_OFFSET if log_offset is not None else 0)
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            check(
                _lib.hmfe_fbank_plan_create_custom(C.byref(self._h), int(sample_rate), self.win, self.shift, self.n_mels,
                                                   window.ctypes.data_as(C.c_void_p), mel.ctypes.data_as(C.c_void_p), flags,
                                                   float(preemphasis), float(log_offset or 0.0)),
                "hmfe_fbank_plan_create_custom",
            )
        return self

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and _lib is not None:  # module globals are gone at interpreter shutdown
            try:
                _lib.hmfe_fbank_plan_destroy(h)
            except TypeError:  # interpreter shutdown: the ctypes entry is already gone
                pass
            self._h = None

    def num_frames(self, lengths) -> np.ndarray:
        n = np.asarray(lengths, dtype=np.int64)
        return np.where(n >= self.win, 1 + (n - self.win) // self.shift, 0)

    def mel_basis(self) -> np.ndarray:
        out = np.empty((self.n_mels, 257), dtype=np.float32)
        check(_lib.hmfe_fbank_mel_basis(self._h, out.ctypes.data_as(C.c_void_p)))
        return out

    @property
    def last_launches(self) -> int:
        return int(_lib.hmfe_fbank_last_launches(self._h))

    def set_profile(self, enable: bool):
        check(_lib.hmfe_fbank_set_profile(self._h, int(bool(enable))))

    def profile_ms(self):
        a, n = C.c_double(), C.c_int()
        check(_lib.hmfe_fbank_profile_ms(self._h, C.byref(a), C.byref(n)))
        return a.value, n.value

    def views(self, wav: torch.Tensor, starts, lengths, rows_per_clip=0, out: torch.Tensor | None = None, stream=None):
        """fbank of the clips wav[starts[i] : starts[i]+lengths[i]].  Returns (out, row_offsets)."""
        _require_cuda_f32(wav, "wav")
        st = np.ascontiguousarray(starts, dtype=np.int64)
        ln = np.ascontiguousarray(lengths, dtype=np.int64)
        m = self.num_frames(ln)
        rows = np.full_like(m, rows_per_clip) if rows_per_clip else m
        ro = np.zeros(len(m) + 1, dtype=np.int64)
        np.cumsum(rows, out=ro[1:])
        if out is None:
            out = torch.empty((int(ro[-1]), self.n_mels), dtype=torch.float32, device=wav.device)
        with torch.cuda.device(wav.device):
            check(
                _lib.hmfe_fbank_batch_views(self._h, C.c_void_p(wav.data_ptr()), st.ctypes.data_as(C.c_void_p),
                                            ln.ctypes.data_as(C.c_void_p), len(ln), C.c_void_p(out.data_ptr()),
                                            int(rows_per_clip), _stream_ptr(stream)),
                "hmfe_fbank_batch_views",
            )
        return out, ro

    def __call__(self, wav: torch.Tensor, offsets, rows_per_clip=0, out=None, stream=None):
        o = _as_offsets(offsets)
        return self.views(wav, o[:-1], np.diff(o), rows_per_clip, out, stream)


def fbank_from_host(plan: FbankPlan, h_wav: torch.Tensor, offsets, h_out: torch.Tensor | None = None, rows_per_clip=0,
                    chunk_bytes: int = 64 << 20):
    """Host-buffer entry of the Kaldi fbank: (pinned) host samples in, (pinned) host features out.

    Chunk i runs on stream i % 2 (H2D copy -> kernel -> D2H copy): the two PCIe directions and the
    kernel of neighbouring chunks overlap.  Returns (h_out [rows, n_mels], row_offsets)."""
    if h_wav.is_cuda or h_wav.dtype not in (torch.float32, torch.int16):
        raise TypeError("h_wav must be a float32 (samples) or int16 (16-bit PCM payload) host tensor")
    pcm = h_wav.dtype == torch.int16
    o = _as_offsets(offsets)
    n = o.size - 1
    m = plan.num_frames(np.diff(o))
    rows = np.full_like(m, rows_per_clip) if rows_per_clip else m
    ro = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(rows, out=ro[1:])
    if h_out is None:
        h_out = torch.empty((int(ro[-1]), plan.n_mels), dtype=torch.float32, pin_memory=True)
    dev = plan.device
    bounds = [0]
    while bounds[-1] < n:
        c0 = bounds[-1]
        c1 = int(np.searchsorted(o, o[c0] + chunk_bytes // 4, side="right")) - 1
        bounds.append(min(n, max(c0 + 1, c1)))
    subs = list(zip(bounds[:-1], bounds[1:]))
    max_samples = max(int(o[b] - o[a]) for a, b in subs)
    max_rows = max(int(ro[b] - ro[a]) for a, b in subs)
    cache = plan.__dict__.setdefault("_host_pipe", {})
    cap = cache.get("cap", (0, 0))
    if cap[0] < max_samples or cap[1] < max_rows:
        cap = (max(cap[0], max_samples), max(cap[1], max_rows))
        cache["wav"] = [torch.empty(cap[0], dtype=torch.float32, device=dev) for _ in range(2)]
        cache["pcm"] = [torch.empty(cap[0], dtype=torch.int16, device=dev) for _ in range(2)]
        cache["out"] = [torch.empty((cap[1], plan.n_mels), dtype=torch.float32, device=dev) for _ in range(2)]
        cache["streams"] = [torch.cuda.Stream(device=dev) for _ in range(2)]
        cache["cap"] = cap
    cur = torch.cuda.current_stream(dev)
    for s in cache["streams"]:
        s.wait_stream(cur)
    for i, (a, b) in enumerate(subs):
        s = cache["streams"][i % 2]
        dw, do = cache["wav"][i % 2], cache["out"][i % 2]
        ns, nr = int(o[b] - o[a]), int(ro[b] - ro[a])
        with torch.cuda.stream(s):
            if pcm:
                dp = cache["pcm"][i % 2]
                dp[:ns].copy_(h_wav[int(o[a]) : int(o[b])], non_blocking=True)
                pcm16_to_f32(dp[:ns], out=dw, stream=s)
            else:
                dw[:ns].copy_(h_wav[int(o[a]) : int(o[b])], non_blocking=True)
            plan(dw, o[a : b + 1] - o[a], rows_per_clip=rows_per_clip, out=do, stream=s)
            h_out[int(ro[a]) : int(ro[b])].copy_(do[:nr], non_blocking=True)
    for s in cache["streams"]:
        s.synchronize()
    return h_out, ro


class ResamplePlan:
    """Polyphase windowed-sinc resampler (torchaudio.transforms.Resample algorithm)."""

    def __init__(self, orig_freq, new_freq=16000, lowpass_filter_width=6, rolloff=0.99, method="sinc_interp_hann",
                 beta=None, device=None):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.orig, self.new = int(orig_freq), int(new_freq)
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            check(
                _lib.hmfe_resample_plan_create(C.byref(self._h), self.orig, self.new, int(lowpass_filter_width),
                                               float(rolloff), {"sinc_interp_hann": 0, "sinc_interp_kaiser": 1}[method],
                                               float(beta or 0.0)),
                "hmfe_resample_plan_create",
            )

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and _lib is not None:  # module globals are gone at interpreter shutdown
            try:
                _lib.hmfe_resample_plan_destroy(h)
            except TypeError:  # interpreter shutdown: the ctypes entry is already gone
                pass
            self._h = None

    def out_lengths(self, lengths) -> np.ndarray:
        n = np.asarray(lengths, dtype=np.int64)
        g = np.gcd(self.orig, self.new)
        u, d = self.new // g, self.orig // g
        return (n * u + d - 1) // d

    def taps(self) -> np.ndarray:
        u, w = C.c_int(), C.c_int()
        check(_lib.hmfe_resample_taps(self._h, C.byref(u), C.byref(w), None))
        out = np.empty((u.value, w.value), dtype=np.float32)
        check(_lib.hmfe_resample_taps(self._h, None, None, out.ctypes.data_as(C.c_void_p)))
        return out

    @property
    def last_launches(self) -> int:
        return int(_lib.hmfe_resample_last_launches(self._h))

    def __call__(self, wav: torch.Tensor, offsets, stream=None, out: torch.Tensor | None = None):
        """Returns (resampled wav, new offsets).  ``wav``: float32 samples or int16 PCM (decoded as x / 32768 in the
        same pass) on the device; ``out``: optional float32 buffer of at least the resampled size."""
        pcm = isinstance(wav, torch.Tensor) and wav.is_cuda and wav.dtype == torch.int16 and wav.is_contiguous()
        if not pcm:
            _require_cuda_f32(wav, "wav")
        o = _as_offsets(offsets)
        no = np.zeros(o.size, dtype=np.int64)
        np.cumsum(self.out_lengths(np.diff(o)), out=no[1:])
        if out is None:
            out = torch.empty(int(no[-1]), dtype=torch.float32, device=wav.device)
        else:
            _require_cuda_f32(out, "out")
            if out.numel() < int(no[-1]):
                raise ValueError("out is too small")
        fn = _lib.hmfe_resample_batch_pcm16 if pcm else _lib.hmfe_resample_batch
        with torch.cuda.device(wav.device):
            check(
                fn(self._h, C.c_void_p(wav.data_ptr()), o.ctypes.data_as(C.c_void_p), o.size - 1, C.c_void_p(out.data_ptr()),
                   _stream_ptr(stream)),
                "hmfe_resample_batch",
            )
        return out, no


# Named filter designs for ``resample_plan(..., **RESAMPLE_PRESETS[name])``.
#   "torchaudio"   torchaudio.transforms.Resample defaults (Hann-windowed sinc, 6 zero crossings, roll-off 0.99): the
#                  resampler the reference uses at src/model/models_eval.py:964-968 and the PINNED oracle of this kernel.
#   "soxr_hq_like" a Kaiser-windowed sinc with the published figures of libsoxr's HQ recipe, which is what
#                  ``librosa.load(path, sr=16000)`` runs (librosa 0.10.1 res_type="soxr_hq", environment.yml:117,188):
#                  pass band flat to 0.913 of the input Nyquist, stop band from 1.0, about 125 dB rejection (20 bit).
#                  Kaiser design formulas: beta = 0.1102 (A - 8.7) = 12.8, transition 0.087 Nyquist -> about 94 zero
#                  crossings each side, cut-off in the middle of the transition band (roll-off 0.9565).  libsoxr itself
#                  is not installable here: parity with it is UNPINNED; tests quantify the response of this design.
RESAMPLE_PRESETS = {
    "torchaudio": dict(lowpass_filter_width=6, rolloff=0.99, method="sinc_interp_hann"),
    "soxr_hq_like": dict(lowpass_filter_width=94, rolloff=0.9565, method="sinc_interp_kaiser", beta=12.8),
}


def fbank_plan(**kw) -> FbankPlan:
    key = ("fbank", torch.cuda.current_device(), tuple(sorted(kw.items())))
    with _plans_lock:
        p = _plans.get(key)
        if p is None:
            p = _plans[key] = FbankPlan(**kw)
        return p


def resample_plan(orig_freq, new_freq=16000, **kw) -> ResamplePlan:
    key = ("resample", torch.cuda.current_device(), int(orig_freq), int(new_freq), tuple(sorted(kw.items())))
    with _plans_lock:
        p = _plans.get(key)
        if p is None:
            p = _plans[key] = ResamplePlan(orig_freq, new_freq, **kw)
        return p


def logmel_views(plan: LogMelPlan, wav: torch.Tensor, starts, lengths, out=None, mode="normalised", stream=None, alt=None):
    """Log-mel of the clips wav[starts[i] : starts[i]+lengths[i]]; with ``alt`` given, a negative start s
    selects alt[-s-1 : -s-1+length] (padded copies kept apart from the signal).  Returns (out, frame_offsets)."""
    _require_cuda_f32(wav, "wav")
    if alt is not None:
        _require_cuda_f32(alt, "alt")
    st = np.ascontiguousarray(starts, dtype=np.int64)
    ln = np.ascontiguousarray(lengths, dtype=np.int64)
    fo = np.zeros(len(ln) + 1, dtype=np.int64)
    np.cumsum(1 + ln // plan.hop, out=fo[1:])
    if out is None:
        out = torch.empty((int(fo[-1]), plan.n_mels), dtype=torch.float32, device=wav.device)
    with torch.cuda.device(wav.device):
        check(
            _lib.hmfe_logmel_batch_views2(plan._h, C.c_void_p(wav.data_ptr()),
                                          C.c_void_p(alt.data_ptr()) if alt is not None else C.c_void_p(),
                                          st.ctypes.data_as(C.c_void_p), ln.ctypes.data_as(C.c_void_p), len(ln),
                                          C.c_void_p(out.data_ptr()), OUT_MODES[mode], _stream_ptr(stream)),
            "hmfe_logmel_batch_views2",
        )
    return out, fo


# =============================================================================================
# Spectrogram-domain dataset ops (src/util.py:26-51)
# =============================================================================================


def spec_means(spec: torch.Tensor, row_offsets, ctx: Context | None = None, stream=None) -> torch.Tensor:
    _require_cuda_f32(spec, "spec")
    ro = _as_offsets(row_offsets)
    ctx = ctx or default_ctx()
    mean = torch.empty(ro.size - 1, dtype=torch.float32, device=spec.device)
    with torch.cuda.device(spec.device):
        check(
            _lib.hmfe_spec_mean_batch(ctx._h, C.c_void_p(spec.data_ptr()), ro.ctypes.data_as(C.c_void_p), ro.size - 1,
                                      int(spec.shape[-1]), C.c_void_p(mean.data_ptr()), _stream_ptr(stream)),
            "hmfe_spec_mean_batch",
        )
    return mean


def spec_mean_ranges(spec: torch.Tensor, row_lo, row_hi, ctx: Context | None = None, stream=None) -> torch.Tensor:
    """Mean of rows [row_lo[i], row_hi[i]) of ``spec`` for every i (float64 accumulation, float32 result)."""
    _require_cuda_f32(spec, "spec")
    lo = np.ascontiguousarray(row_lo, dtype=np.int64)
    hi = np.ascontiguousarray(row_hi, dtype=np.int64)
    ctx = ctx or default_ctx()
    mean = torch.empty(lo.size, dtype=torch.float32, device=spec.device)
    with torch.cuda.device(spec.device):
        check(
            _lib.hmfe_spec_mean_ranges(ctx._h, C.c_void_p(spec.data_ptr()), lo.ctypes.data_as(C.c_void_p),
                                       hi.ctypes.data_as(C.c_void_p), lo.size, int(spec.shape[-1]), C.c_void_p(mean.data_ptr()),
                                       _stream_ptr(stream)),
            "hmfe_spec_mean_ranges",
        )
    return mean


def spec_crop(spec: torch.Tensor, descs: np.ndarray, out_rows: int, row_mask: torch.Tensor | None = None,
              means: torch.Tensor | None = None, ctx: Context | None = None, stream=None) -> torch.Tensor:
    """Apply ``hmfe_crop_desc`` records: returns [n_items, out_rows, n_cols]."""
    _require_cuda_f32(spec, "spec")
    descs = np.ascontiguousarray(descs, dtype=CROP_DTYPE)
    ctx = ctx or default_ctx()
    n_cols = int(spec.shape[-1])
    out = torch.empty((descs.size, int(out_rows), n_cols), dtype=torch.float32, device=spec.device)
    if row_mask is not None and not (row_mask.is_cuda and row_mask.dtype == torch.uint8 and row_mask.is_contiguous()):
        raise TypeError("row_mask must be a contiguous uint8 CUDA tensor")
    with torch.cuda.device(spec.device):
        check(
            _lib.hmfe_spec_crop_batch(ctx._h, C.c_void_p(spec.data_ptr()), n_cols, descs.ctypes.data_as(C.c_void_p),
                                      descs.size, C.c_void_p(row_mask.data_ptr()) if row_mask is not None else None,
                                      C.c_void_p(means.data_ptr()) if means is not None else None,
                                      C.c_void_p(out.data_ptr()), int(out_rows), _stream_ptr(stream)),
            "hmfe_spec_crop_batch",
        )
    return out


def htsat_input(spec: torch.Tensor, src_rows, n_rows, bn_weight, bn_bias, bn_mean, bn_var, eps=1e-5, spec_size=256,
                ctx: Context | None = None, stream=None) -> torch.Tensor:
    """HTS-AT input stage (htsat.py:889-891 bn0 in inference form, :829-858 reshape_wav2img) for items
    ``spec[src_rows[i] : src_rows[i] + n_rows[i]]``.  Returns ``[n_items, 1, spec_size, spec_size]``."""
    _require_cuda_f32(spec, "spec")
    rows = np.ascontiguousarray(src_rows, dtype=np.int64)
    cnt = np.ascontiguousarray(n_rows, dtype=np.int32)
    w, bb = np.asarray(bn_weight, dtype=np.float32), np.asarray(bn_bias, dtype=np.float32)
    mu, var = np.asarray(bn_mean, dtype=np.float32), np.asarray(bn_var, dtype=np.float32)
    scale = (w / np.sqrt(var + np.float32(eps))).astype(np.float32)  # float32 like torch's batch_norm
    shift = (bb - mu * scale).astype(np.float32)
    ctx = ctx or default_ctx()
    out = torch.empty((rows.size, 1, int(spec_size), int(spec_size)), dtype=torch.float32, device=spec.device)
    with torch.cuda.device(spec.device):
        check(
            _lib.hmfe_htsat_input_batch(ctx._h, C.c_void_p(spec.data_ptr()), int(spec.shape[-1]),
                                        rows.ctypes.data_as(C.c_void_p), cnt.ctypes.data_as(C.c_void_p), rows.size,
                                        scale.ctypes.data_as(C.c_void_p), shift.ctypes.data_as(C.c_void_p), int(spec_size),
                                        C.c_void_p(out.data_ptr()), _stream_ptr(stream)),
            "hmfe_htsat_input_batch",
        )
    return out


# =============================================================================================
# Host index planners - integer work, bit-exact with the reference
# (src/util.py:504-620, 257-259; extract_feature.py:250-259).  A chunk is either a straight
# VIEW of the (trimmed) clip or a GATHER record (tile / repeat / zero pad).
# =============================================================================================

import math
import random as _random


def _view(start, length):
    return ("view", int(start), int(length))


def _gather(length, src_start=0, period=1, a_end=0, a_phase=0, b_end=0, b_start=0):
    return ("gather", int(length), int(src_start), int(period), int(a_end), int(a_phase), int(b_end), int(b_start))


def plan_zero_padding(src_start, src_len, L):
    """_zero_padding (src/util.py:504-519) of clip[src_start : src_start+src_len] to L samples."""
    if src_len > L:
        raise ValueError("slice longer than the padded length (the reference would raise here too)")
    if src_len / L < 0.5:
        copies = (L - 1) // src_len if src_len > 0 else 0  # while cursor + len < L
        return _gather(L, src_start=src_start, period=max(1, src_len), a_end=copies * src_len, a_phase=0,
                       b_end=copies * src_len)
    if src_len == L:
        return _view(src_start, L)
    return _gather(L, src_start=src_start, period=max(1, src_len), a_end=0, b_end=src_len, b_start=0)


def plan_equally_slice_pad(n, desired_length, sample_rate):
    """_equally_slice_pad_sample (src/util.py:522-547)."""
    L = int(desired_length * sample_rate)
    n_slices = int(math.ceil((n / sample_rate) / desired_length))
    per = n // n_slices
    out, lo = [], 0
    for _ in range(n_slices):
        hi = min(lo + per, n)
        out.append(plan_zero_padding(lo, hi - lo, L))
        lo = hi
    return out


def plan_duplicate_padding(n, src_start, src_len, L):
    """_duplicate_padding (src/util.py:550-575): source at the end, tail of the doubled clip in
    front (the reference's seeded draw 0.0617 < 0.5 always selects this branch).  The reseed of
    Python's global RNG is reproduced by the caller (``reseed_like_reference``)."""
    left = L - src_len
    if left == 0:
        return _view(src_start, L)
    len_aug = n
    while len_aug < left:
        len_aug *= 2
    return _gather(L, src_start=0, period=n, a_end=left, a_phase=(len_aug - left) % n, b_end=L, b_start=src_start)


def reseed_like_reference():
    """Side effect of every _duplicate_padding call (src/util.py:564-565)."""
    _random.seed(7456)
    _random.random()


def plan_split_pad(n, desired_length, sample_rate, types="repeat"):
    """split_pad_sample (src/util.py:578-620) for a clip of n samples -> list of chunks."""
    if types == "zero":
        return plan_equally_slice_pad(n, desired_length, sample_rate)
    L = int(desired_length * sample_rate)
    out = []
    if n > L:
        hop = L // 2
        n_full = 1 + (n - L) // hop
        out += [_view(j * hop, L) for j in range(n_full)]
        last = n_full * hop
        out.append(plan_duplicate_padding(n, last, n - last, L))
    else:
        out.append(plan_duplicate_padding(n, 0, n, L))
    return out


def plan_split_sample(n, desired_length, sample_rate):
    """split_sample (extract_feature.py:250-259): non-overlapping views, last one short."""
    L = int(desired_length * sample_rate)
    return [_view(L * i, min(L, n - L * i)) for i in range(int(np.ceil(n / L)))]


def materialise_chunks(work: torch.Tensor, used: int, clip_starts, chunk_lists, ctx: Context | None = None, stream=None,
                       owned: bool = True):
    """Turn per-clip chunk plans into (starts, lengths, clip_ids) views into ``work``.

    ``work[:used]`` holds the signal; gather chunks are written behind it (the buffer is
    re-allocated with the signal copied if its spare capacity is too small, or if ``owned`` is False:
    ``work`` is then the caller's tensor and whatever lies behind ``used`` in it is the caller's).
    Returns (work, starts, lengths, clip_ids, is_view).
    """
    starts, lengths, clip_ids, recs, is_view = [], [], [], [], []
    tail = used
    for cid, (c0, chunks) in enumerate(zip(clip_starts, chunk_lists)):
        for ch in chunks or ():
            if ch[0] == "view":
                starts.append(c0 + ch[1])
                lengths.append(ch[2])
            else:
                _, length, src_start, period, a_end, a_phase, b_end, b_start = ch
                recs.append((c0 + src_start, tail, length, period, a_end, a_phase, b_end, b_start))
                starts.append(tail)
                lengths.append(length)
                tail += length
            clip_ids.append(cid)
            is_view.append(ch[0] == "view")
    if recs:
        if tail > work.numel() or not owned:
            bigger = torch.empty(tail, dtype=torch.float32, device=work.device)
            bigger[:used].copy_(work[:used])
            work = bigger
        descs = np.array(recs, dtype=GATHER_DTYPE)
        gather(work, work, descs, ctx=ctx, stream=stream)
    return (work, np.asarray(starts, dtype=np.int64), np.asarray(lengths, dtype=np.int64),
            np.asarray(clip_ids, dtype=np.int64), np.asarray(is_view, dtype=bool))
