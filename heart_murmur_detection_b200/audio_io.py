"""WAV decode + GPU resample: the ``librosa.load(path, sr=16000)`` step of the reference
(src/util.py:153-155,222-224,323-325,391-393,805-807; extract_feature.py:214-216).

Decode follows soundfile's float32 conventions (PCM16 / 32768, PCM24 / 2^23, PCM32 / 2^31,
unsigned 8-bit (x-128)/128, IEEE float as is); multi-channel audio is averaged to mono like
``librosa.to_mono``.  Rate conversion runs on the GPU with the polyphase windowed-sinc kernel
(``frontend.ResamplePlan``); output length is ``ceil(n * sr / sr_native)`` as in librosa.
librosa's own resampler (libsoxr HQ) is not reproducible here - see DESIGN.md.
"""
from __future__ import annotations

import struct

import numpy as np
import torch

from . import frontend as fe


def read_wav(path: str):
    """Returns (float32 array [n] or [n, channels], sample_rate)."""
    with open(path, "rb") as f:
        data = f.read()
    if len(data) < 12 or data[:4] not in (b"RIFF", b"RF64") or data[8:12] != b"WAVE":
        raise ValueError(f"{path}: not a RIFF/WAVE file")
    pos, fmt, raw = 12, None, None
    while pos + 8 <= len(data):
        cid, size = data[pos : pos + 4], struct.unpack("<I", data[pos + 4 : pos + 8])[0]
        body = data[pos + 8 : pos + 8 + size]
        if cid == b"fmt ":
            tag, ch, sr, _, _, bits = struct.unpack("<HHIIHH", body[:16])
            if tag == 0xFFFE and len(body) >= 26:  # WAVE_FORMAT_EXTENSIBLE: sub-format GUID starts with the tag
                tag = struct.unpack("<H", body[24:26])[0]
            fmt = (tag, ch, sr, bits)
        elif cid == b"data":
            raw = body
        pos += 8 + size + (size & 1)
    if fmt is None or raw is None:
        raise ValueError(f"{path}: missing fmt or data chunk")
    tag, ch, sr, bits = fmt
    if tag == 1:
        if bits == 16:
            x = np.frombuffer(raw, dtype="<i2").astype(np.float32) / np.float32(32768.0)
        elif bits == 8:
            x = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - 128.0) / np.float32(128.0)
        elif bits == 32:
            x = (np.frombuffer(raw, dtype="<i4").astype(np.float64) / 2147483648.0).astype(np.float32)
        elif bits == 24:
            b = np.frombuffer(raw[: len(raw) // 3 * 3], dtype=np.uint8).reshape(-1, 3).astype(np.int32)
            v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
            v = np.where(v & 0x800000, v - (1 << 24), v)
            x = (v.astype(np.float64) / 8388608.0).astype(np.float32)
        else:
            raise ValueError(f"{path}: unsupported PCM width {bits}")
    elif tag == 3:
        x = np.frombuffer(raw, dtype="<f4" if bits == 32 else "<f8").astype(np.float32)
    else:
        raise ValueError(f"{path}: unsupported WAVE format tag {tag}")
    if ch > 1:
        x = x[: len(x) // ch * ch].reshape(-1, ch)
    return x, int(sr)


def load(path: str, sr: int | None = 16000, device=None):
    """``librosa.load`` shaped: returns (float32 numpy array, sample_rate)."""
    x, native = read_wav(path)
    if x.ndim == 2:
        x = x.mean(axis=1, dtype=np.float32)
    if sr is None or native == sr:
        return np.ascontiguousarray(x, dtype=np.float32), native
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    plan = fe.resample_plan(native, sr)
    y, _ = plan(torch.from_numpy(np.ascontiguousarray(x)).to(dev), np.array([0, len(x)], dtype=np.int64))
    return y.cpu().numpy(), sr


def write_wav_pcm16(path: str, x: np.ndarray, sr: int):
    """Minimal PCM16 writer (tests and examples)."""
    x = np.asarray(x)
    pcm = np.clip(np.round(x * 32768.0), -32768, 32767).astype("<i2")
    ch = 1 if pcm.ndim == 1 else pcm.shape[1]
    raw = pcm.tobytes()
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 36 + len(raw)) + b"WAVE")
        f.write(b"fmt " + struct.pack("<IHHIIHH", 16, 1, ch, sr, sr * ch * 2, ch * 2, 16))
        f.write(b"data" + struct.pack("<I", len(raw)) + raw)


def write_wav_f32(path: str, x: np.ndarray, sr: int):
    """IEEE float32 WAV writer (lossless for float32 arrays; tests and examples)."""
    x = np.ascontiguousarray(x, dtype="<f4")
    ch = 1 if x.ndim == 1 else x.shape[1]
    raw = x.tobytes()
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 36 + len(raw)) + b"WAVE")
        f.write(b"fmt " + struct.pack("<IHHIIHH", 16, 3, ch, sr, sr * ch * 4, ch * 4, 32))
        f.write(b"data" + struct.pack("<I", len(raw)) + raw)
