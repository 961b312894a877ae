"""numpy restatement of the librosa 0.10.1 functions on the reference's hot path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  parity unpinned against
librosa itself (not installable here); semantics follow SURVEY.md Appendix A and
are cross-checked against torchaudio in ``tests/test_oracle.py``.

Reference call sites (``/root/reference``):
  librosa.effects.trim            src/util.py:172,242,340,822; extract_feature.py:221
  librosa.get_duration            src/util.py:175,250,370
  librosa.util.frame              src/util.py:600
  librosa.feature.melspectrogram  src/util.py:484
  librosa.power_to_db             src/util.py:494

The module exposes a ``librosa``-shaped namespace (``effects.trim``,
``feature.melspectrogram`` ...) so that the golden generator can execute the
unmodified reference ``src/util.py`` on top of it.
"""
from __future__ import annotations

import types

import numpy as np

# --------------------------------------------------------------------------- mel scale


def hz_to_mel(frequencies, htk=False):
    frequencies = np.asanyarray(frequencies, dtype=np.float64)
    if htk:
        return 2595.0 * np.log10(1.0 + frequencies / 700.0)
    f_min, f_sp = 0.0, 200.0 / 3
    mels = (frequencies - f_min) / f_sp
    min_log_hz = 1000.0
    min_log_mel = (min_log_hz - f_min) / f_sp
    logstep = np.log(6.4) / 27.0
    if frequencies.ndim:
        log_t = frequencies >= min_log_hz
        mels[log_t] = min_log_mel + np.log(frequencies[log_t] / min_log_hz) / logstep
    elif frequencies >= min_log_hz:
        mels = min_log_mel + np.log(frequencies / min_log_hz) / logstep
    return mels


def mel_to_hz(mels, htk=False):
    mels = np.asanyarray(mels, dtype=np.float64)
    if htk:
        return 700.0 * (10.0 ** (mels / 2595.0) - 1.0)
    f_min, f_sp = 0.0, 200.0 / 3
    freqs = f_min + f_sp * mels
    min_log_hz = 1000.0
    min_log_mel = (min_log_hz - f_min) / f_sp
    logstep = np.log(6.4) / 27.0
    if mels.ndim:
        log_t = mels >= min_log_mel
        freqs[log_t] = min_log_hz * np.exp(logstep * (mels[log_t] - min_log_mel))
    elif mels >= min_log_mel:
        freqs = min_log_hz * np.exp(logstep * (mels - min_log_mel))
    return freqs


def mel_frequencies(n_mels=128, fmin=0.0, fmax=11025.0, htk=False):
    min_mel = hz_to_mel(fmin, htk=htk)
    max_mel = hz_to_mel(fmax, htk=htk)
    mels = np.linspace(min_mel, max_mel, n_mels)
    return mel_to_hz(mels, htk=htk)


def mel_filterbank(sr, n_fft, n_mels=128, fmin=0.0, fmax=None, htk=False, norm="slaney", dtype=np.float32):
    """``librosa.filters.mel``: [n_mels, 1 + n_fft//2] triangular filters, Slaney-normalised."""
    if fmax is None:
        fmax = float(sr) / 2
    n_mels = int(n_mels)
    weights = np.zeros((n_mels, int(1 + n_fft // 2)), dtype=dtype)
    fftfreqs = np.fft.rfftfreq(n=n_fft, d=1.0 / sr)
    mel_f = mel_frequencies(n_mels + 2, fmin=fmin, fmax=fmax, htk=htk)
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    if norm == "slaney":
        enorm = 2.0 / (mel_f[2 : n_mels + 2] - mel_f[:n_mels])
        weights *= enorm[:, np.newaxis]
    return weights


# --------------------------------------------------------------------------- framing / stft


def frame(x, *, frame_length, hop_length, axis=-1):
    """``librosa.util.frame`` for 1-D input.

    axis=-1 -> shape (frame_length, n_frames); axis=0 -> shape (n_frames, frame_length).
    """
    x = np.asarray(x)
    if x.ndim != 1:
        raise ValueError("restatement handles 1-D input only")
    n = x.shape[0]
    if n < frame_length:
        raise ValueError(f"Input is too short (n={n}) for frame_length={frame_length}")
    if hop_length < 1:
        raise ValueError(f"Invalid hop_length: {hop_length}")
    n_frames = 1 + (n - frame_length) // hop_length
    s = x.strides[0]
    out = np.lib.stride_tricks.as_strided(
        x, shape=(n_frames, frame_length), strides=(hop_length * s, s), writeable=False
    )
    if axis in (-1, 1):
        return out.T
    return out


def hann_periodic(n):
    """scipy.signal.get_window('hann', n, fftbins=True), float64."""
    k = np.arange(n, dtype=np.float64)
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * k / n)


def stft(y, n_fft=2048, hop_length=None, center=True, pad_mode="constant"):
    """``librosa.stft`` (window='hann', win_length=n_fft): [1+n_fft//2, 1+N//hop].

    The window (float64) multiplies the float32 frames in float64, the FFT runs in
    float64 and the result is stored as complex64 for float32 input, complex128 for
    float64 input (librosa ``util.dtype_r2c``).
    """
    y = np.asarray(y)
    if hop_length is None:
        hop_length = n_fft // 4
    if pad_mode not in ("constant", "reflect"):
        raise NotImplementedError(pad_mode)
    if center:
        y = np.pad(y, (n_fft // 2, n_fft // 2), mode=pad_mode)
    fft_window = hann_periodic(n_fft).reshape(-1, 1)
    y_frames = frame(y, frame_length=n_fft, hop_length=hop_length)  # (n_fft, T)
    out_dtype = np.complex64 if y.dtype == np.float32 else np.complex128
    T = y_frames.shape[1]
    D = np.empty((1 + n_fft // 2, T), dtype=out_dtype, order="F")
    blk = 512
    for s in range(0, T, blk):
        e = min(T, s + blk)
        D[:, s:e] = np.fft.rfft(fft_window * y_frames[:, s:e].astype(np.float64), axis=0)
    return D


def melspectrogram(*, y, sr=22050, n_fft=2048, hop_length=512, power=2.0, n_mels=128, fmin=0.0, fmax=None):
    """``librosa.feature.melspectrogram`` -> [n_mels, T] (dtype follows ``y``)."""
    y = np.asarray(y)
    D = stft(y, n_fft=n_fft, hop_length=hop_length, center=True, pad_mode="constant")
    S = np.abs(D) ** power
    mel_basis = mel_filterbank(sr, n_fft, n_mels=n_mels, fmin=fmin, fmax=fmax)
    return np.einsum("ft,mf->mt", S, mel_basis, optimize=True)


def power_to_db(S, ref=1.0, amin=1e-10, top_db=80.0):
    """``librosa.power_to_db``; ``ref`` may be a callable such as ``np.max``."""
    S = np.asarray(S)
    magnitude = S
    ref_value = ref(magnitude) if callable(ref) else np.abs(ref)
    log_spec = 10.0 * np.log10(np.maximum(amin, magnitude))
    # numpy 1.26 (the reference pin) evaluates the scalar reference term in float64
    # and subtracts it from the float32 array after rounding once.
    ref_db = 10.0 * np.log10(np.maximum(np.float64(amin), np.float64(ref_value)))
    log_spec = (log_spec - log_spec.dtype.type(ref_db)).astype(log_spec.dtype)
    if top_db is not None:
        if top_db < 0:
            raise ValueError("top_db must be non-negative")
        log_spec = np.maximum(log_spec, log_spec.max() - log_spec.dtype.type(top_db))
    return log_spec


def amplitude_to_db(S, ref=1.0, amin=1e-5, top_db=80.0):
    magnitude = np.abs(np.asarray(S))
    ref_value = ref(magnitude) if callable(ref) else np.abs(ref)
    power = np.square(magnitude)
    return power_to_db(power, ref=ref_value**2, amin=amin**2, top_db=top_db)


def rms(*, y, frame_length=2048, hop_length=512, center=True):
    """``librosa.feature.rms`` (pad_mode='constant'): shape (1, n_frames)."""
    y = np.asarray(y)
    if center:
        y = np.pad(y, (frame_length // 2, frame_length // 2), mode="constant")
    x = frame(y, frame_length=frame_length, hop_length=hop_length)  # (L, T)
    # librosa 0.10.1: np.mean(util.abs2(x, dtype=np.float32), axis=-2) - always float32
    power = np.mean(np.square(x, dtype=np.float32), axis=-2, keepdims=True)
    return np.sqrt(power)


def trim(y, *, top_db=60, ref=np.max, frame_length=2048, hop_length=512):
    """``librosa.effects.trim`` -> (y[start:end], np.array([start, end]))."""
    y = np.asarray(y)
    mse = rms(y=y, frame_length=frame_length, hop_length=hop_length)
    db = amplitude_to_db(mse[0, :], ref=ref, top_db=None)
    non_silent = db > -top_db
    nonzero = np.flatnonzero(non_silent)
    if nonzero.size > 0:
        start = int(nonzero[0] * hop_length)
        end = min(y.shape[-1], int((nonzero[-1] + 1) * hop_length))
    else:
        start, end = 0, 0
    return y[start:end], np.asarray([start, end])


def get_duration(*, y, sr=22050):
    return float(np.asarray(y).shape[-1]) / sr


# --------------------------------------------------------------------------- librosa-shaped namespace


def as_librosa_module(load_fn=None):
    """Return a module object shaped like ``librosa`` for the golden generator.

    ``load_fn(path, sr)`` supplies the decoded + resampled audio (the real
    ``librosa.load`` needs soundfile + soxr, both absent here).
    """
    m = types.ModuleType("librosa")
    m.__version__ = "0.10.1-restated"
    m.effects = types.SimpleNamespace(trim=trim)
    m.feature = types.SimpleNamespace(melspectrogram=melspectrogram, rms=rms)
    m.util = types.SimpleNamespace(frame=frame)
    m.filters = types.SimpleNamespace(mel=mel_filterbank)
    m.power_to_db = power_to_db
    m.amplitude_to_db = amplitude_to_db
    m.get_duration = get_duration
    m.stft = stft

    def _load(path, sr=22050, **kw):
        if load_fn is None:
            raise RuntimeError("no loader installed")
        return load_fn(path, sr)

    m.load = _load
    return m
