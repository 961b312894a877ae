"""Drop-in mirror of the reference's ``src/util.py`` front-end call surface.

Same names, positional order, defaults, return types and "too short" behaviour (a printed
warning and ``None`` / ``[]``, never an exception) as ``/root/reference/src/util.py``; the
arithmetic runs on the GPU through libhmfe.so.  A maintainer switches the reference over with

    from heart_murmur_detection_b200.util import (get_entire_signal_librosa, get_split_signal_librosa,
        get_split_signal_fbank_pad, pre_process_audio_mel_t, split_pad_sample, crop_first, random_crop,
        random_mask, random_multiply)

Batch entry points for rewritten extractor loops live in ``pipeline`` / ``frontend``.
There is no CPU fallback: without a CUDA device these functions raise.
"""
from __future__ import annotations

import os
import random

import numpy as np
import torch

from . import audio_io
from . import frontend as fe
from . import pipeline as pl


def _dev():
    if not torch.cuda.is_available():
        raise RuntimeError("heart_murmur_detection_b200 needs a CUDA device (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def _to_dev(x) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(_dev())


def _one(x):
    return _to_dev(x), np.array([0, len(x)], dtype=np.int64)


# ---------------------------------------------------------------------------------------------
# spectrogram-domain ops (src/util.py:26-51): cheap host views, kept verbatim in behaviour.
# GPU batch versions: frontend.spec_crop / datasets.ColaBatcher.
# ---------------------------------------------------------------------------------------------


def crop_first(data, crop_size=128):
    return data[0:crop_size, :]


def random_crop(data, crop_size=128):
    start = int(random.random() * (data.shape[0] - crop_size))
    return data[start : (start + crop_size), :]


def draw_mask_rows(n_rows, rate_start=0.1, rate_seq=0.2):
    """The Markov frame mask of random_mask: consumes Python's global RNG exactly as
    src/util.py:35-46 does (second draw only when the first fails after a masked frame)."""
    rows = np.zeros(n_rows, dtype=np.uint8)
    prev = False
    for i in range(n_rows):
        if random.random() < rate_start or (prev and random.random() < rate_seq):
            prev = True
            rows[i] = 1
        else:
            prev = False
    return rows


def random_mask(data, rate_start=0.1, rate_seq=0.2):
    new_data = data.copy()
    mean = new_data.mean()
    rows = draw_mask_rows(new_data.shape[0], rate_start, rate_seq)
    new_data[rows.astype(bool), :] = mean
    return new_data


def random_multiply(data):
    new_data = data.copy()
    return new_data * (0.9 + random.random() / 5.0)


# ---------------------------------------------------------------------------------------------
# band-pass (src/util.py:113-126)
# ---------------------------------------------------------------------------------------------


def _butter_bandpass(lowcut, highcut, fs, order=5):
    return fe.butter_bandpass_ba(lowcut, highcut, fs, order=order)


def _butter_bandpass_filter(data, lowcut, highcut, fs, order=5):
    """float64 result like ``scipy.signal.lfilter`` (the cascade runs in float64 on the GPU)."""
    data = np.asarray(data)
    wav, off = _one(data)
    sos = fe.butter_bandpass_sos(lowcut, highcut, fs, order=order)
    y = fe.iir_sos(wav, off, sos, out_dtype=torch.float64)
    return y.cpu().numpy()


# ---------------------------------------------------------------------------------------------
# log-mel (src/util.py:481-501)
# ---------------------------------------------------------------------------------------------


def pre_process_audio_mel_t(audio, sample_rate=16000, n_mels=64, f_min=50, f_max=2000, nfft=1024, hop=512):
    audio = np.asarray(audio)
    out_dtype = np.float64 if audio.dtype == np.float64 else np.float32  # librosa keeps the input precision
    plan = fe.logmel_plan(sample_rate, n_mels, f_min, f_max, nfft, hop)
    wav, off = _one(audio)
    out, _ = plan(wav, off)
    res = out.cpu().numpy()
    if res.size and res.max() == res.min():
        print("warning in producing spectrogram!")
    return res.astype(out_dtype, copy=False)


# ---------------------------------------------------------------------------------------------
# pad / split (src/util.py:504-620)
# ---------------------------------------------------------------------------------------------


def _chunks_to_numpy(cb: pl.ChunkBatch):
    return [cb.samples(k).cpu().numpy() for k in range(len(cb.starts))]


def split_pad_sample(sample, desired_length, sample_rate, types="repeat"):
    clip = np.asarray(sample[0])
    chunks = fe.plan_split_pad(len(clip), desired_length, sample_rate, types)
    work, off = _one(clip)
    work, starts, lengths, _, _ = fe.materialise_chunks(work, len(clip), [0], [chunks])
    if types != "zero":
        fe.reseed_like_reference()
    return [(work[int(s) : int(s) + int(n)].cpu().numpy(), sample[1], sample[2]) for s, n in zip(starts, lengths)]


def decide_droplast(yt, sr, input_sec):
    duration = len(yt) / sr
    return duration > input_sec and (duration % input_sec) * 2 < input_sec


# ---------------------------------------------------------------------------------------------
# composite entry points
# ---------------------------------------------------------------------------------------------


def _load(data_folder, filename, sample_rate):
    return audio_io.load(os.path.join(data_folder, filename + ".wav"), sr=sample_rate)


def get_entire_signal_librosa(
    data_folder,
    filename,
    input_sec=8,
    sample_rate=16000,
    butterworth_filter=None,
    spectrogram=False,
    pad=False,
    from_cycle=False,
    yt=None,
    types="repeat",
    lowcut=200,
    highcut=1800,
    max_sec=None,
):
    if from_cycle:
        # the reference skips load / filter / trim and uses the caller's yt (src/util.py:220,250)
        n = len(yt)
        chunker = pl.entire_signal_chunker(input_sec, sample_rate, pad, types, max_sec)
        chunker.dup_called = False
        chunks = chunker(n)
        if chunks is None:
            print("Warning: audio too short, skipped")
            return None
        work, _ = _one(np.asarray(yt))
        work, starts, lengths, _, is_view = fe.materialise_chunks(work, n, [0], [chunks])
        cb = pl.ChunkBatch(work, starts, lengths, np.zeros(1, np.int64), 1, np.ones(1, bool), np.array([[0, n]]),
                           chunker.dup_called, 0, is_view, np.asarray(yt).dtype == np.float64)
        res = pl.log_mel_features(cb, f_max=8000) if spectrogram else cb  # mel basis at the 16 kHz default (src/util.py:261-263)
    else:
        data, _ = _load(data_folder, filename, sample_rate)
        wav, off = _one(data)
        res = pl.entire_signal_batch(wav, off, input_sec, sample_rate, butterworth_filter, spectrogram, pad, types,
                                     lowcut, highcut, max_sec)
    cb = res.chunks if spectrogram else res
    if cb.used_duplicate_padding:
        fe.reseed_like_reference()
    if not cb.valid[0]:
        print("Warning: audio too short, skipped")
        return None
    duration = (cb.trim[0, 1] - cb.trim[0, 0]) / sample_rate
    if max_sec and duration > max_sec:
        print(f"Trimmed audio to {max_sec} seconds")
    # padded chunks are float32 in the reference (np.zeros(..., float32)); untouched views of
    # band-passed audio keep lfilter's float64
    if spectrogram:
        return res.chunk(0).cpu().numpy().astype(cb.ref_dtype(0), copy=False)
    return _chunks_to_numpy(cb)[0].astype(cb.ref_dtype(0), copy=False)


def _split_common(data_folder, filename, input_sec, sample_rate, butterworth_filter, trim_tail, lowcut, highcut):
    data, rate = _load(data_folder, filename, sample_rate)
    wav, off = _one(data)
    return wav, off, rate


def get_split_signal_librosa(
    data_folder,
    filename,
    input_sec=8,
    sample_rate=16000,
    butterworth_filter=None,
    spectrogram=False,
    trim_tail=False,
    lowcut=200,
    highcut=1800,
):
    wav, off, rate = _split_common(data_folder, filename, input_sec, sample_rate, butterworth_filter, trim_tail, lowcut,
                                   highcut)
    res = pl.split_signal_batch(wav, off, input_sec, rate, butterworth_filter, spectrogram, trim_tail, lowcut, highcut)
    cb = res.chunks if spectrogram else res
    if cb.used_duplicate_padding:
        fe.reseed_like_reference()
    if not spectrogram:
        return [a.astype(cb.ref_dtype(k), copy=False) for k, a in enumerate(_chunks_to_numpy(cb))]
    return [res.chunk(k).cpu().numpy().astype(cb.ref_dtype(k), copy=False) for k in range(len(cb.starts))]


def get_split_signal_fbank_pad(
    data_folder,
    filename,
    input_sec=8,
    sample_rate=16000,
    butterworth_filter=None,
    spectrogram=False,
    trim_tail=False,
):
    data, rate = _load(data_folder, filename, sample_rate)
    wav, off = _one(data)
    res = pl.split_signal_fbank_pad_batch(wav, off, input_sec, rate, butterworth_filter, spectrogram, trim_tail)
    cb = res.chunks if spectrogram else res
    if cb.used_duplicate_padding:
        fe.reseed_like_reference()
    if not spectrogram:
        return _chunks_to_numpy(cb)
    return [res.chunk(k).cpu() for k in range(len(cb.starts))]


def get_individual_segments_librosa(
    data_folder,
    filename,
    input_sec=8,
    sample_rate=16000,
    hop_sec=2,
    butterworth_filter=None,
    spectrogram=False,
):
    data, rate = _load(data_folder, filename, sample_rate)
    wav, off = _one(data)
    res = pl.individual_segments_batch(wav, off, input_sec, rate, hop_sec, butterworth_filter, spectrogram)
    cb = res.chunks if spectrogram else res
    if cb.used_duplicate_padding:
        fe.reseed_like_reference()
    if not cb.valid[0]:
        print("Warning: audio too short, skipped")
        return []
    if not spectrogram:
        return [a.astype(cb.ref_dtype(k), copy=False) for k, a in enumerate(_chunks_to_numpy(cb))]
    return [res.chunk(k).cpu().numpy().astype(cb.ref_dtype(k), copy=False) for k in range(len(cb.starts))]


# ---------------------------------------------------------------------------------------------
# ICBHI respiratory-cycle slicing (src/util.py:129-138, 374-422, 447-478).  Index work on the host;
# the optional band-pass runs on the GPU and, as in the reference, makes the slices float64.
# ---------------------------------------------------------------------------------------------


def _slice_data_librosa(start, end, data, sample_rate):
    """data[int(start*sr) : int(end*sr)], both ends clamped to the recording (src/util.py:129-138)."""
    n = len(data)
    return data[min(int(start * sample_rate), n) : min(int(end * sample_rate), n)]


def _get_lungsound_label(crackle, wheeze, n_cls):
    """src/util.py:447-462: 4 classes = crackle + 2 * wheeze, 2 classes = any adventitious sound."""
    if n_cls == 4:
        table = {(0, 0): 0, (1, 0): 1, (0, 1): 2, (1, 1): 3}
        return table.get((crackle, wheeze))
    if n_cls == 2:
        return 0 if (crackle == 0 and wheeze == 0) else 1
    return None


def _get_diagnosis_label(disease, n_cls):
    """src/util.py:465-478."""
    if n_cls == 3:
        if disease in ("COPD", "Bronchiectasis", "Asthma"):
            return 1
        if disease in ("URTI", "LRTI", "Pneumonia", "Bronchiolitis"):
            return 2
        return 0
    if n_cls == 2:
        return 0 if disease == "Healthy" else 1
    return None


def get_individual_cycles_librosa(
    class_split,
    recording_annotations,
    data_folder,
    filename,
    sample_rate,
    n_cls,
    butterworth_filter=None,
):
    """[(audio_chunk, label), ...], one entry per annotated respiratory cycle (src/util.py:374-422).
    ``recording_annotations`` is the pandas frame the reference builds (columns Start, End and
    Crackles / Wheezes or Disease)."""
    data, rate = _load(data_folder, filename, sample_rate)
    if butterworth_filter:
        data = _butter_bandpass_filter(data=data, lowcut=200, highcut=1800, fs=sample_rate, order=butterworth_filter)
    sample_data = []
    for idx in recording_annotations.index:
        row = recording_annotations.loc[idx]
        chunk = _slice_data_librosa(row["Start"], row["End"], data, rate)
        if class_split == "cycle":
            sample_data.append((chunk, _get_lungsound_label(row["Crackles"], row["Wheezes"], n_cls)))
        elif class_split == "diagnosis":
            sample_data.append((chunk, _get_diagnosis_label(row["Disease"], n_cls)))
    return sample_data
