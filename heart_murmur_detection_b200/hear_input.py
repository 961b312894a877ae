"""Drop-in mirror of the HeAR audio front-end of the reference (sibling front-ends):
``src/benchmark/baseline/hear/python/data_processing/audio_utils.py:448-476`` (``preprocess_audio``) with
``_mel_pcen`` (:357-383), ``_compute_stft`` (:22-115), ``_linear_to_mel_weight_matrix`` (:264-358),
``_pcen_function`` / ``_ema`` (:121-246) and ``_torch_resize_bilinear_tf_compat`` (:386-445).

2 s clips at 16 kHz -> batch-wide min / max scaling -> 400-sample periodic Hann frames every 160 samples,
400-point FFT -> 128 HTK-mel bands -> PCEN -> bilinear resize to [192, 128].  All of it runs in
``csrc/hear_pcen.cu`` (a 25 x 16 mixed-radix FFT kernel and a PCEN + resize kernel).  The window and the mel
matrix are plan constants built once on the host with the same float32 torch ops the reference uses, so they
are bit-identical to the reference's.
"""
from __future__ import annotations

import ctypes as C
import threading

import torch

from . import _lib
from ._lib import check

SAMPLE_RATE = 16000
CLIP_SAMPLES = 32000  # audio_utils.py:466-472
FRAME_LENGTH = 16 * 25
FRAME_STEP = 160
NUM_MEL_BINS = 128
OUT_SIZE = (192, 128)  # :475


def _hertz_to_mel(frequencies_hertz: torch.Tensor) -> torch.Tensor:
    return 2595.0 * torch.log10(1.0 + frequencies_hertz / 700.0)


def linear_to_mel_weight_matrix(num_mel_bins=128, num_spectrogram_bins=201, sample_rate=16000.0, lower_edge_hertz=0.0,
                                upper_edge_hertz=8000.0, dtype=torch.float32) -> torch.Tensor:
    """_linear_to_mel_weight_matrix (:264-358): [num_spectrogram_bins, num_mel_bins]; same errors for bad arguments."""
    if num_mel_bins <= 0:
        raise ValueError(f"num_mel_bins must be positive. Got: {num_mel_bins}.")
    if num_spectrogram_bins <= 0:
        raise ValueError(f"num_spectrogram_bins must be positive. Got: {num_spectrogram_bins}.")
    if sample_rate <= 0:
        raise ValueError(f"sample_rate must be positive. Got: {sample_rate}.")
    if lower_edge_hertz < 0.0:
        raise ValueError(f"lower_edge_hertz must be non-negative. Got: {lower_edge_hertz}.")
    if lower_edge_hertz >= upper_edge_hertz:
        raise ValueError("lower_edge_hertz must be smaller than upper_edge_hertz. Got: "
                         f"lower_edge_hertz={lower_edge_hertz}, upper_edge_hertz={upper_edge_hertz}.")
    if upper_edge_hertz > sample_rate / 2.0:
        raise ValueError("upper_edge_hertz must not be larger than the Nyquist frequency"
                         f"({sample_rate / 2.0}). Got: upper_edge_hertz={upper_edge_hertz}.")
    zero = torch.tensor(0.0, dtype=dtype)
    nyquist = torch.tensor(sample_rate, dtype=dtype) / 2.0
    bins_mel = _hertz_to_mel(torch.linspace(zero, nyquist, num_spectrogram_bins, dtype=dtype)[1:]).unsqueeze(1)
    edges = torch.linspace(_hertz_to_mel(torch.tensor(lower_edge_hertz, dtype=dtype)),
                           _hertz_to_mel(torch.tensor(upper_edge_hertz, dtype=dtype)), num_mel_bins + 2, dtype=dtype)
    edges = edges.unfold(0, 3, 1)
    lower, center, upper = (edges[:, i].unsqueeze(0) for i in range(3))
    w = torch.maximum(zero, torch.minimum((bins_mel - lower) / (center - lower), (upper - bins_mel) / (upper - center)))
    return torch.nn.functional.pad(w, (0, 0, 1, 0), mode="constant", value=0.0)


class HearPlan:
    """Plan of the mel-PCEN kernels: window, mel matrix and PCEN constants (defaults of _pcen_function, :193-201)."""

    def __init__(self, num_mel_bins=NUM_MEL_BINS, alpha=0.8, smooth_coef=0.04, delta=2.0, root=2.0, floor=1e-8, device=None):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.n_mels = int(num_mel_bins)
        win = torch.hann_window(FRAME_LENGTH).contiguous()
        mel = linear_to_mel_weight_matrix(num_mel_bins=self.n_mels, num_spectrogram_bins=FRAME_LENGTH // 2 + 1).contiguous()
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            check(
                _lib.hmfe_hear_plan_create(C.byref(self._h), C.c_void_p(win.data_ptr()), C.c_void_p(mel.data_ptr()),
                                           self.n_mels, float(alpha), float(smooth_coef), float(delta), float(root),
                                           float(floor)),
                "hmfe_hear_plan_create",
            )
        self._ws = None

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and _lib is not None:  # module globals are gone at interpreter shutdown
            try:
                _lib.hmfe_hear_plan_destroy(h)
            except TypeError:  # interpreter shutdown: the ctypes entry is already gone
                pass
            self._h = None

    @property
    def last_launches(self) -> int:
        return int(_lib.hmfe_hear_last_launches(self._h))

    def _workspace(self, n_clips, n_padded, device):
        need = int(_lib.hmfe_hear_workspace_bytes(self._h, n_clips, n_padded))
        if self._ws is None or self._ws.numel() < need or self._ws.device != device:
            self._ws = torch.empty(need, dtype=torch.uint8, device=device)
        return self._ws

    def _check(self, audio):
        if not (audio.is_cuda and audio.dtype == torch.float32 and audio.dim() == 2 and audio.is_contiguous()):
            raise TypeError("audio must be a contiguous float32 CUDA tensor [n_clips, n_samples]")

    def mel_power(self, audio: torch.Tensor, n_padded=CLIP_SAMPLES, stream=None) -> torch.Tensor:
        """[n_clips, ceil(n_padded / 160), n_mels] mel power (the input of the PCEN stage)."""
        self._check(audio)
        B, n = audio.shape
        T = int(_lib.hmfe_hear_num_frames(int(n_padded)))
        out = torch.empty((B, T, self.n_mels), dtype=torch.float32, device=audio.device)
        ws = self._workspace(B, n_padded, audio.device)
        with torch.cuda.device(audio.device):
            check(
                _lib.hmfe_hear_mel_batch(self._h, C.c_void_p(audio.data_ptr()), B, n, int(n_padded), C.c_void_p(out.data_ptr()),
                                         C.c_void_p(ws.data_ptr()), ws.numel(), _stream_ptr(stream)),
                "hmfe_hear_mel_batch",
            )
        return out

    def __call__(self, audio: torch.Tensor, n_padded=CLIP_SAMPLES, out_rows=OUT_SIZE[0], stream=None) -> torch.Tensor:
        """[n_clips, out_rows, n_mels] mel-PCEN image rows."""
        self._check(audio)
        B, n = audio.shape
        out = torch.empty((B, int(out_rows), self.n_mels), dtype=torch.float32, device=audio.device)
        ws = self._workspace(B, n_padded, audio.device)
        with torch.cuda.device(audio.device):
            check(
                _lib.hmfe_hear_mel_pcen_batch(self._h, C.c_void_p(audio.data_ptr()), B, n, int(n_padded), int(out_rows),
                                              C.c_void_p(out.data_ptr()), C.c_void_p(ws.data_ptr()), ws.numel(),
                                              _stream_ptr(stream)),
                "hmfe_hear_mel_pcen_batch",
            )
        return out


def _stream_ptr(stream):
    s = torch.cuda.current_stream() if stream is None else stream
    return C.c_void_p(s.cuda_stream)


_plans: dict = {}
_lock = threading.Lock()


def hear_plan() -> HearPlan:
    key = torch.cuda.current_device()
    with _lock:
        p = _plans.get(key)
        if p is None:
            p = _plans[key] = HearPlan()
        return p


def preprocess_audio(audio: torch.Tensor) -> torch.Tensor:
    """Same call as the reference (:448-476): ``[..., samples]`` rank-2 tensor of 2 s clips at 16 kHz (shorter clips
    are zero padded, longer ones rejected) -> ``[B, 1, 192, 128]`` float32 on the device of ``audio``.  As in the
    reference the min / max scaling is taken over the WHOLE batch (:361-365), so results depend on batch composition."""
    if audio.ndim != 2:
        raise ValueError(f"Input audio must have rank 2, got rank {audio.ndim}")
    if audio.shape[1] > CLIP_SAMPLES:
        raise ValueError(f"Input audio must have 32000 samples, got {audio.shape[1]}")
    dev = audio.device
    x = audio.detach().to(device="cuda", dtype=torch.float32).contiguous()
    out = hear_plan()(x).unsqueeze(1)
    return out if dev.type == "cuda" else out.to(dev)
