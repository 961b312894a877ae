"""CPU restatement of the reference's own front-end Python, on in-memory arrays.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Every function cites the
reference lines it follows (``/root/reference``).  File decode + resample
(``librosa.load``) is outside these functions: they start from the float32 array
``librosa.load`` would have returned.

Pinned by ``tests/golden/*.npz``: fixtures produced by executing the real
``src/util.py`` (``tests/golden/make_golden.py``).
"""
from __future__ import annotations

import math
import random

import numpy as np

from . import librosa_restated as lr

# ----------------------------------------------------------------------------- band-pass


def butter_bandpass(lowcut, highcut, fs, order=5):
    """src/util.py:113-119 - scipy Butterworth band-pass, transfer-function form."""
    from scipy.signal import butter

    nyq = 0.5 * fs
    return butter(order, [lowcut / nyq, highcut / nyq], btype="band")


def butter_bandpass_filter(data, lowcut, highcut, fs, order=5):
    """src/util.py:122-126 - causal single-pass ``lfilter``; float64 result."""
    from scipy.signal import lfilter

    b, a = butter_bandpass(lowcut, highcut, fs, order=order)
    return lfilter(b, a, data)


# ----------------------------------------------------------------------------- silence trim


def trim_silence(data, sample_rate=16000):
    """src/util.py:170-172,237-244,338-340,820-822 - 100 ms frames, 50 ms hop, top_db 60."""
    frame_len = int(sample_rate / 10)
    hop = int(frame_len / 2)
    return lr.trim(data, frame_length=frame_len, hop_length=hop)


# ----------------------------------------------------------------------------- pad / split


def zero_padding(source, output_length):
    """src/util.py:504-519 - right zero-pad, or tile forward while a whole copy still fits
    strictly inside, when the source is shorter than half the target."""
    out = np.zeros(output_length, dtype=np.float32)
    n = len(source)
    if n / output_length < 0.5:
        pos = 0
        while pos + n < output_length:
            out[pos : pos + n] = source
            pos += n
    else:
        out[:n] = source
    return out


def equally_slice_pad(clip, desired_length, sample_rate):
    """src/util.py:522-547 - ceil(duration/desired) equal slices, each through zero_padding."""
    L = int(desired_length * sample_rate)
    clip = np.array(clip, copy=True)
    n = len(clip)
    n_slices = int(math.ceil((n / sample_rate) / desired_length))
    per = n // n_slices
    out = []
    lo = 0
    for _ in range(n_slices):
        hi = min(lo + per, n)
        out.append(zero_padding(clip[lo:hi], L))
        lo = hi
    return out


def duplicate_padding(clip, source, output_length):
    """src/util.py:550-575 - 'repeat' padding.

    The reference reseeds Python's global RNG with 7456 on every call and draws one
    number (0.0617...) so the branch taken is always "source at the end, tail of the
    doubled clip in front".  The reseed is a visible side effect and is reproduced.
    """
    out = np.zeros(output_length, dtype=np.float32)
    left = output_length - len(source)
    aug = clip
    while len(aug) < left:
        aug = np.concatenate([aug, aug])
    random.seed(7456)
    if random.random() < 0.5:
        out[left:] = source
        out[:left] = aug[len(aug) - left :]
    else:
        out[: len(source)] = source
        out[len(source) :] = aug[:left]
    return out


def split_pad_sample(clip, desired_length, sample_rate, types="repeat"):
    """src/util.py:578-620 - returns the list of padded chunks (arrays only)."""
    if types == "zero":
        return equally_slice_pad(clip, desired_length, sample_rate)
    L = int(desired_length * sample_rate)
    clip = np.array(clip, copy=True)
    n = len(clip)
    out = []
    if n > L:
        hop = L // 2
        n_full = 1 + (n - L) // hop
        for j in range(n_full):
            out.append(clip[j * hop : j * hop + L])
        tail = clip[n_full * hop :]
        out.append(duplicate_padding(clip, tail, L))
    else:
        out.append(duplicate_padding(clip, clip, L))
    return out


def split_sample(clip, desired_length, sample_rate):
    """extract_feature.py:250-259 - non-overlapping chunks, last one short."""
    L = int(desired_length * sample_rate)
    clip = np.array(clip, copy=True)
    n_chunks = int(np.ceil(len(clip) / L))
    return [clip[L * i : L * (i + 1)] for i in range(n_chunks)]


def decide_droplast(yt, sr, input_sec):
    """src/util.py:369-371."""
    duration = lr.get_duration(y=yt, sr=sr)
    return duration > input_sec and (duration % input_sec) * 2 < input_sec


# ----------------------------------------------------------------------------- log-mel


def log_mel(audio, sample_rate=16000, n_mels=64, f_min=50, f_max=2000, nfft=1024, hop=512, return_parts=False):
    """src/util.py:481-501 - mel power -> dB re clip max (floor -80) -> clip min-max -> [T, n_mels]."""
    S = lr.melspectrogram(y=audio, sr=sample_rate, n_mels=n_mels, fmin=f_min, fmax=f_max, n_fft=nfft, hop_length=hop)
    db = lr.power_to_db(S, ref=np.max)
    lo, hi = db.min(), db.max()
    if hi != lo:
        norm = (db - lo) / (hi - lo)
    else:
        norm = db
    if return_parts:
        return norm.T, db.T, S.T
    return norm.T


# ----------------------------------------------------------------------------- kaldi fbank


def kaldi_fbank_chunk(chunk, sample_rate=16000):
    """src/util.py:841-856 / extract_feature.py:228-243 - mean-removed chunk -> kaldi fbank
    (live ``torchaudio.compliance.kaldi.fbank``, the library the reference calls).
    Returns ``None`` for chunks of <= 400 samples (skipped by the reference)."""
    import torch
    import torchaudio

    w = chunk - chunk.mean()
    w = torch.tensor(w).reshape([1, -1])
    if w.shape[1] <= 400:
        return None
    return torchaudio.compliance.kaldi.fbank(
        w,
        channel=0,
        frame_length=25,
        htk_compat=True,
        sample_frequency=sample_rate,
        use_energy=False,
        window_type="hanning",
        num_mel_bins=128,
        dither=0.0,
        frame_shift=10,
    )


def pad_to_model(x, n_frames=1024, n_mels=128):
    """audioMAE/models_mae.py:1178-1181, mae_training.py:88-109 - zero rows/cols appended, crop if longer."""
    x = np.asarray(x)
    x = x[:n_frames, :n_mels]
    return np.pad(x, ((0, n_frames - x.shape[0]), (0, n_mels - x.shape[1])))


# ----------------------------------------------------------------------------- composite entry points


def _maybe_filter(data, butterworth_filter, lowcut, highcut, sample_rate):
    if butterworth_filter:
        return butter_bandpass_filter(data, lowcut, highcut, sample_rate, order=butterworth_filter)
    return data


def entire_signal(
    data,
    input_sec=8,
    sample_rate=16000,
    butterworth_filter=None,
    spectrogram=False,
    pad=False,
    types="repeat",
    lowcut=200,
    highcut=1800,
    max_sec=None,
):
    """src/util.py:205-267 (get_entire_signal_librosa) after ``librosa.load``."""
    data = _maybe_filter(data, butterworth_filter, lowcut, highcut, sample_rate)
    yt, _ = trim_silence(data, sample_rate)
    duration = lr.get_duration(y=yt, sr=sample_rate)
    if duration < input_sec:
        if not pad:
            return None
        yt = split_pad_sample(yt, input_sec, sample_rate, types)[0]
    if max_sec and duration > max_sec:
        yt = yt[: int(max_sec * sample_rate)]
    if spectrogram:
        return log_mel(yt.squeeze(), f_max=8000)
    return yt


def split_signal(
    data,
    input_sec=8,
    sample_rate=16000,
    butterworth_filter=None,
    spectrogram=False,
    trim_tail=False,
    lowcut=200,
    highcut=1800,
):
    """src/util.py:309-364 (get_split_signal_librosa) after ``librosa.load``."""
    data = _maybe_filter(data, butterworth_filter, lowcut, highcut, sample_rate)
    yt, _ = trim_silence(data, sample_rate)
    drop_last = decide_droplast(yt, sample_rate, input_sec) if trim_tail else False
    chunks = split_pad_sample(yt, input_sec, sample_rate)
    if drop_last:
        chunks.pop()
    if not spectrogram:
        return chunks
    return [log_mel(c.squeeze(), f_max=8000) for c in chunks]


def split_signal_fbank_pad(
    data, input_sec=8, sample_rate=16000, butterworth_filter=None, spectrogram=False, trim_tail=False
):
    """src/util.py:794-860 (get_split_signal_fbank_pad) after ``librosa.load``."""
    data = _maybe_filter(data, butterworth_filter, 200, 1800, sample_rate)
    yt, _ = trim_silence(data, sample_rate)
    drop_last = decide_droplast(yt, sample_rate, input_sec) if trim_tail else False
    chunks = split_pad_sample(yt, input_sec, sample_rate)
    if drop_last:
        chunks.pop()
    if not spectrogram:
        return chunks
    out = []
    for c in chunks:
        fb = kaldi_fbank_chunk(c, sample_rate)
        if fb is not None:
            out.append(fb)
    return out


def split_signal_fbank(data, input_sec=10, sample_rate=16000):
    """extract_feature.py:213-247 (get_split_signal_fbank) after ``librosa.load``."""
    yt, _ = trim_silence(data, sample_rate)
    out = []
    for c in split_sample(yt, input_sec, sample_rate):
        fb = kaldi_fbank_chunk(c, sample_rate)
        if fb is not None:
            out.append(fb)
    return out


def individual_segments(
    data, input_sec=8, sample_rate=16000, hop_sec=2, butterworth_filter=None, spectrogram=False
):
    """src/util.py:141-202 (get_individual_segments_librosa) after ``librosa.load``."""
    data = _maybe_filter(data, butterworth_filter, 200, 1800, sample_rate)
    yt, _ = trim_silence(data, sample_rate)
    duration = lr.get_duration(y=yt, sr=sample_rate)
    if duration < 2:
        return []

    def cut(t0, t1):  # src/util.py:129-138
        a = min(int(t0 * sample_rate), len(yt))
        b = min(int(t1 * sample_rate), len(yt))
        return yt[a:b]

    segs = []
    start, end = 0, input_sec
    while end <= duration:
        segs.append(cut(start, end))
        start += hop_sec
        end += hop_sec
    if start + 2 < duration:
        segs.append(split_pad_sample(cut(start, end), 8, sample_rate)[0])
    if spectrogram:
        return [log_mel(s.squeeze()) for s in segs]
    return segs


# ----------------------------------------------------------------------------- spectrogram-domain ops


def crop_first(data, crop_size=128):
    """src/util.py:26-27."""
    return data[0:crop_size, :]


def random_crop(data, crop_size=128):
    """src/util.py:30-32 - one draw from Python's global RNG."""
    start = int(random.random() * (data.shape[0] - crop_size))
    return data[start : start + crop_size, :]


def random_mask(data, rate_start=0.1, rate_seq=0.2):
    """src/util.py:35-46 - Markov frame masking with the global mean; the second draw
    happens only when the first fails and the previous frame was masked."""
    out = data.copy()
    mean = out.mean()
    prev = False
    for i in range(out.shape[0]):
        if random.random() < rate_start or (prev and random.random() < rate_seq):
            prev = True
            out[i, :] = mean
        else:
            prev = False
    return out


def random_multiply(data):
    """src/util.py:49-51."""
    return data.copy() * (0.9 + random.random() / 5.0)


def resample_torchaudio(wave, orig_sr, new_sr):
    """Pinned resampler oracle: ``torchaudio.functional.resample`` (Hann-windowed sinc,
    the resampler used at src/model/models_eval.py:964-968).  ``librosa.load``'s soxr_hq
    resampler (src/util.py:153,222,...) is not installable here: parity unpinned."""
    import torch
    import torchaudio

    w = torch.as_tensor(np.asarray(wave, dtype=np.float32))
    return torchaudio.functional.resample(w, int(orig_sr), int(new_sr)).numpy()


def htsat_input(x, bn_weight, bn_bias, bn_mean, bn_var, eps=1e-5, spec_size=256):
    """src/model/htsat/htsat.py:889-891 (bn0, inference) + :829-858 (reshape_wav2img) on one
    spectrogram ``x [T, F]`` -> ``[spec_size, spec_size]`` (torch CPU ops, the library the reference calls)."""
    import torch

    x = torch.as_tensor(np.asarray(x, dtype=np.float32))[None, None]  # B C T F
    F_ = x.shape[3]
    bn = torch.nn.BatchNorm2d(F_, eps=eps)
    with torch.no_grad():
        bn.weight.copy_(torch.as_tensor(np.asarray(bn_weight, dtype=np.float32)))
        bn.bias.copy_(torch.as_tensor(np.asarray(bn_bias, dtype=np.float32)))
        bn.running_mean.copy_(torch.as_tensor(np.asarray(bn_mean, dtype=np.float32)))
        bn.running_var.copy_(torch.as_tensor(np.asarray(bn_var, dtype=np.float32)))
    bn.eval()
    with torch.no_grad():
        x = bn(x.transpose(1, 3)).transpose(1, 3)
        ratio = spec_size // F_
        target_T = spec_size * ratio
        if x.shape[2] < target_T:
            x = torch.nn.functional.interpolate(x, (target_T, F_), mode="bicubic", align_corners=True)
        x = x.permute(0, 1, 3, 2).contiguous()
        x = x.reshape(1, 1, F_, ratio, target_T // ratio).permute(0, 1, 3, 2, 4).contiguous()
        x = x.reshape(1, 1, ratio * F_, target_T // ratio)
    return x[0, 0].numpy()


# ------------------------------------------------------------------------------------------------
# CLAP baseline front-end (sibling front-end).  torchlibrosa, which the reference
# imports (msclap/models/audio.py:5), is not installed here: parity unpinned against torchlibrosa itself;
# tests/test_oracle.py cross-checks this restatement against torchaudio's MelSpectrogram with the same
# constants (reflect padding, Slaney mel, power 2).
# ------------------------------------------------------------------------------------------------
def clap_fixed_duration(x, audio_duration=5, sample_rate=44100):
    """load_audio_into_tensor after read_audio (msclap/CLAPWrapper.py:279-299); draws from Python's global RNG."""
    x = np.asarray(x).reshape(-1)
    L = audio_duration * sample_rate
    if L >= x.shape[0]:
        repeat_factor = int(np.ceil(L / x.shape[0]))
        return np.tile(x, repeat_factor)[0:L].astype(np.float32)
    start_index = random.randrange(x.shape[0] - L)
    return x[start_index : start_index + L].astype(np.float32)


def clap_logmel(x, sample_rate=44100, window_size=1024, hop_size=320, mel_bins=64, fmin=50, fmax=14000, return_power=False):
    """Cnn14 input stage (msclap/models/audio.py:146-175,190-192; configs/config_2022.yml:10-17):
    torchlibrosa Spectrogram (Hann, centre, reflect, power 2) -> librosa mel -> 10 log10(max(1e-10, .)), [T, mel_bins]."""
    x = np.asarray(x, dtype=np.float32)
    D = lr.stft(x, n_fft=window_size, hop_length=hop_size, center=True, pad_mode="reflect")
    S = (D.real.astype(np.float64) ** 2 + D.imag.astype(np.float64) ** 2).T
    mel = S @ lr.mel_filterbank(sample_rate, window_size, n_mels=mel_bins, fmin=fmin, fmax=fmax).T.astype(np.float64)
    mel = mel.astype(np.float32)
    db = (10.0 * np.log10(np.maximum(np.float32(1e-10), mel))).astype(np.float32)
    return (db, mel) if return_power else db


# ------------------------------------------------------------------------------------------------
# HeAR baseline front-end (sibling front-end): restatement of
# src/benchmark/baseline/hear/python/data_processing/audio_utils.py with the torch CPU ops the reference
# calls.  Pinned: tests/golden/ref_hear.npz holds outputs of the reference's own preprocess_audio /
# _linear_to_mel_weight_matrix (tests/golden/make_golden_hear.py executes them).
# ------------------------------------------------------------------------------------------------
def hear_mel_matrix(num_mel_bins=128, num_spectrogram_bins=201, sample_rate=16000.0, lower_edge_hertz=0.0,
                    upper_edge_hertz=8000.0):
    """_linear_to_mel_weight_matrix (audio_utils.py:264-358): [num_spectrogram_bins, num_mel_bins] float32, HTK mel,
    triangles in the mel domain, DC row zero, all arithmetic in float32 torch ops as in the reference."""
    import torch

    dt = torch.float32
    to_mel = lambda f: 2595.0 * torch.log10(1.0 + f / 700.0)  # noqa: E731  (:249-261)
    nyquist = torch.tensor(sample_rate, dtype=dt) / 2.0
    lin = torch.linspace(torch.tensor(0.0, dtype=dt), nyquist, num_spectrogram_bins, dtype=dt)[1:]
    bins_mel = to_mel(lin).unsqueeze(1)
    edges = torch.linspace(to_mel(torch.tensor(lower_edge_hertz, dtype=dt)), to_mel(torch.tensor(upper_edge_hertz, dtype=dt)),
                           num_mel_bins + 2, dtype=dt).unfold(0, 3, 1)
    lower, center, upper = (edges[:, i].unsqueeze(0) for i in range(3))
    lower_slopes = (bins_mel - lower) / (center - lower)
    upper_slopes = (upper - bins_mel) / (upper - center)
    w = torch.maximum(torch.tensor(0.0, dtype=dt), torch.minimum(lower_slopes, upper_slopes))
    return torch.nn.functional.pad(w, (0, 0, 1, 0)).numpy()


def hear_mel_power(audio):
    """Scaling + STFT + mel of _mel_pcen (audio_utils.py:357-382): [B, n] -> [B, ceil(n/160), 128] float32."""
    import torch

    x = torch.as_tensor(np.asarray(audio, dtype=np.float32)).clone()
    x -= torch.min(x)
    x = x / (torch.max(x) + 1e-8)
    x = (x * 2) - 1
    n = x.shape[-1]
    n_frames = math.ceil(n / 160) if n > 0 else 0
    padded = max(0, (n_frames - 1) * 160 + 400) if n_frames > 0 else 400
    if padded > n:
        x = torch.nn.functional.pad(x, (0, padded - n))
    frames = x.unfold(-1, 400, 160) * torch.hann_window(400)
    spec = torch.square(torch.abs(torch.fft.rfft(frames, n=400, dim=-1)))
    return torch.matmul(spec, torch.as_tensor(hear_mel_matrix())).numpy()


def hear_pcen(mel, alpha=0.8, smooth_coef=0.04, delta=2.0, root=2.0, floor=1e-8):
    """_pcen_function with its _ema (audio_utils.py:121-246): [B, T, C] -> [B, T, C].  The reference's two matmuls
    with diagonal kernels are two separately rounded float32 products."""
    import torch

    x = torch.as_tensor(np.asarray(mel, dtype=np.float32))
    c_in = torch.tensor(smooth_coef, dtype=torch.float32)
    c_state = torch.tensor(1.0 - smooth_coef, dtype=torch.float32)
    state = x[:, 0]
    ema = [state]
    for t in range(1, x.shape[1]):
        state = x[:, t] * c_in + state * c_state
        ema.append(state)
    ema = torch.stack(ema, dim=1)
    a = torch.ones(x.shape[-1]) * min(alpha, 1.0)
    r = 1.0 / (torch.ones(x.shape[-1]) * max(root, 1.0))
    d = torch.ones(x.shape[-1]) * delta
    return ((x / (floor + ema) ** a + d) ** r - d**r).numpy()


def hear_preprocess_audio(audio):
    """preprocess_audio (audio_utils.py:448-476): [B, n <= 32000] -> [B, 1, 192, 128] float32."""
    import torch

    audio = np.asarray(audio, dtype=np.float32)
    if audio.ndim != 2:
        raise ValueError(f"Input audio must have rank 2, got rank {audio.ndim}")
    if audio.shape[1] > 32000:
        raise ValueError(f"Input audio must have 32000 samples, got {audio.shape[1]}")
    if audio.shape[1] < 32000:
        audio = np.pad(audio, ((0, 0), (0, 32000 - audio.shape[1])))
    img = torch.as_tensor(hear_pcen(hear_mel_power(audio))).unsqueeze(1)
    return torch.nn.functional.interpolate(img, size=(192, 128), mode="bilinear", align_corners=False, antialias=False).numpy()


class SpecAugmentation:
    """torchlibrosa.augmentation.SpecAugmentation (torchlibrosa 0.1.0, un-vendored dependency of
    src/benchmark/other_eval/finetuning.py:14,64-69,104-116; restated from its published source - parity unpinned
    for the library itself, torchlibrosa is not installable here).  Training-mode DropStripes on a 4-D tensor
    (batch, channels, time, freq), in place: per batch element and stripe ``distance = torch.randint(0, drop_width)``,
    ``bgn = torch.randint(0, total - distance)``, time stripes (dim 2) first, then frequency stripes (dim 3); the
    draws come from torch's global generator."""

    def __init__(self, time_drop_width, time_stripes_num, freq_drop_width, freq_stripes_num):
        self.cfg = ((2, time_drop_width, time_stripes_num), (3, freq_drop_width, freq_stripes_num))
        self.training = True

    def __call__(self, x):
        import torch

        assert x.ndimension() == 4
        if not self.training:
            return x
        for dim, width, num in self.cfg:
            total = x.shape[dim]
            for n in range(x.shape[0]):
                e = x[n]
                for _ in range(num):
                    distance = torch.randint(low=0, high=width, size=(1,))[0]
                    bgn = torch.randint(low=0, high=total - distance, size=(1,))[0]
                    if dim == 2:
                        e[:, bgn : bgn + distance, :] = 0
                    else:
                        e[:, :, bgn : bgn + distance] = 0
        return x


def dataset_cola_item(x, max_len=251, augment=True, windowing=False):
    """AudioDataset.__getitem__, method 'cola' (src/pretrain/cola_training.py:56-80; src/pretrain/mae_training.py:64-79
    adds the ``windowing`` pre-crop to 3 * max_len)."""
    if windowing and x.shape[0] > max_len * 3:
        x = random_crop(x, crop_size=max_len * 3)
    if augment:
        x = random_mask(x)
    x1 = random_crop(x, crop_size=max_len)
    x2 = random_crop(x, crop_size=max_len)
    if augment:
        x1, x2 = random_multiply(x1), random_multiply(x2)
    return np.asarray(x1, dtype=np.float32), np.asarray(x2, dtype=np.float32)


def dataset_mae_item(x, max_len):
    """AudioDataset.__getitem__, methods 'mae' / 'audiomae' (src/pretrain/mae_training.py:82-109)."""
    p = max_len - x.shape[0]
    if p < 0:
        x = random_crop(x, crop_size=max_len)
    elif p > 0:
        x = np.pad(x, ((0, p), (0, 0)), mode="constant")
    return x.astype(np.float32)


def dataset_finetune_item(x, max_len=256, augment=True, crop_mode="first", spec_augment=False, time_drop_width=100,
                          time_stripes_num=2, freq_drop_width=20, freq_stripes_num=2):
    """AudioDataset.__getitem__ of src/benchmark/other_eval/finetuning.py:74-123 (spectrogram branch)."""
    import torch

    if max_len:
        x = random_crop(x, crop_size=max_len) if crop_mode == "random" else crop_first(x, crop_size=max_len)
    if augment:
        x = random_mask(x)
        x = random_multiply(x)
    x = torch.tensor(x, dtype=torch.float)
    if spec_augment:
        aug = SpecAugmentation(time_drop_width, time_stripes_num, freq_drop_width, freq_stripes_num)
        x = aug(x.unsqueeze(0).unsqueeze(0)).squeeze(0).squeeze(0)
    return x.numpy()
