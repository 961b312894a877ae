"""Batched writers of the reference's on-disk feature caches: the loops that
call the front-end one file at a time and ``np.save`` the result.

    entire_spec_npy/<id>.npy + entire_spec_filenames.npy   heart_pressl.py:58-99   [T, 64] per recording
    spectrogram_pad8.npy                                   finetuning.py:1120-1138 [N, 256, 64]
    fbank_audiomae.npy                                     finetuning.py:967-980   [N, 998, 128]

Files are decoded on the host (``audio_io``), packed into ragged batches of about ``batch_bytes`` of
samples, and run through the batched pipelines; arrays are written with ``numpy.save`` so the files
are byte-compatible with what the reference's loaders ``np.load``.
"""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

from . import audio_io
from . import pipeline as pl


def _batches(paths, sample_rate, batch_bytes, workers):
    """Yield (indices, clips at ``sample_rate``) for consecutive groups of files.

    The worker threads only DECODE (host work, native rate).  Rate conversion runs on the GPU from the calling thread,
    one ``ResamplePlan`` call per native rate and batch: an ``hmfe`` plan / context belongs to one host thread
    (include/hmfe.h) and worker threads would all default to CUDA device 0."""
    with ThreadPoolExecutor(max_workers=workers) as ex:
        pending, idx, size = [], [], 0
        for i, (clip, native) in enumerate(ex.map(lambda p: audio_io.load(p, sr=None), paths)):
            pending.append((clip, native))
            idx.append(i)
            size += int(clip.size * 4 * (sample_rate / native if native != sample_rate else 1))
            if size >= batch_bytes:
                yield idx, _resample_group(pending, sample_rate)
                pending, idx, size = [], [], 0
        if pending:
            yield idx, _resample_group(pending, sample_rate)


def _resample_group(decoded, sample_rate):
    """[(clip, native rate)] -> [clip at sample_rate]; one ragged GPU call per distinct native rate."""
    out = [None] * len(decoded)
    by_rate = {}
    for k, (clip, native) in enumerate(decoded):
        if native == sample_rate:
            out[k] = clip
        else:
            by_rate.setdefault(native, []).append(k)
    for native, ks in by_rate.items():
        wav, off = _to_device([decoded[k][0] for k in ks])
        y, no = audio_io.fe.resample_plan(native, sample_rate)(wav, off)
        y = y.cpu().numpy()
        for j, k in enumerate(ks):
            out[k] = y[int(no[j]) : int(no[j + 1])]
    return out


def _to_device(clips):
    off = np.zeros(len(clips) + 1, dtype=np.int64)
    np.cumsum([len(c) for c in clips], out=off[1:])
    host = torch.empty(int(off[-1]), dtype=torch.float32, pin_memory=True)
    for c, a, b in zip(clips, off[:-1], off[1:]):
        host[a:b] = torch.from_numpy(np.ascontiguousarray(c, dtype=np.float32))
    return host.cuda(non_blocking=True), off


def _wav_path(audio_file):
    """The reference passes ``audio_file[:-4]`` and re-appends ".wav" (heart_pressl.py:35,79)."""
    return audio_file if audio_file.endswith(".wav") else audio_file + ".wav"


def write_entire_spec_cache(sound_files, feature_dir, input_sec=8, spec_dir="entire_spec_npy",
                            filenames_file="entire_spec", sample_rate=16000, batch_bytes=256 << 20, workers=8):
    """preprocess_spectrogram_SSL (heart_pressl.py:58-99): one ``[T, 64]`` log-mel .npy per recording
    that is at least ``input_sec`` long after the silence trim, plus the list of written paths.
    Returns (written path stems, number of skipped recordings)."""
    out_dir = os.path.join(feature_dir, spec_dir)
    written, invalid = [], 0
    for idx, clips in _batches([_wav_path(f) for f in sound_files], sample_rate, batch_bytes, workers):
        wav, off = _to_device(clips)
        res = pl.entire_signal_batch(wav, off, input_sec=input_sec, sample_rate=sample_rate, spectrogram=True)
        feats = res.features[: int(res.row_offsets[-1])].cpu().numpy()
        k = 0
        for j, i in enumerate(idx):
            if not res.chunks.valid[j]:
                print("Warning: audio too short, skipped")
                invalid += 1
                continue
            file_id = os.path.basename(sound_files[i])
            file_id = file_id[:-4] if file_id.endswith(".wav") else file_id
            os.makedirs(out_dir, exist_ok=True)
            stem = os.path.join(out_dir, file_id)
            np.save(stem + ".npy", feats[int(res.row_offsets[k]) : int(res.row_offsets[k + 1])])
            written.append(stem)
            k += 1
    np.save(os.path.join(feature_dir, filenames_file + "_filenames.npy"), written)
    return written, invalid


def build_spectrogram_pad_cache(sound_files, input_sec=8.18, sample_rate=16000, batch_bytes=256 << 20, workers=8,
                                save_to=None):
    """finetuning.py:1120-1138: ``np.array([get_split_signal_librosa(f, spectrogram=True, input_sec)[0]])``
    -> ``[N, 1 + int(input_sec*sr)//512, 64]`` float32 (256 rows for 8.18 s)."""
    rows = []
    for idx, clips in _batches([_wav_path(f) for f in sound_files], sample_rate, batch_bytes, workers):
        wav, off = _to_device(clips)
        res = pl.split_signal_batch(wav, off, input_sec=input_sec, sample_rate=sample_rate, spectrogram=True,
                                    first_only=True)
        T = 1 + int(input_sec * sample_rate) // 512
        feats = res.features[: int(res.row_offsets[-1])].cpu().numpy().reshape(-1, T, 64)
        if feats.shape[0] != len(idx):
            raise ValueError("an empty recording has no first chunk (the reference raises IndexError here)")
        rows.append(feats)
    x = np.concatenate(rows) if rows else np.zeros((0, 1 + int(input_sec * sample_rate) // 512, 64), np.float32)
    if save_to:
        np.save(save_to, x)
    return x


def build_fbank_cache(sound_files, input_sec=10, sample_rate=16000, batch_bytes=256 << 20, workers=8, save_to=None):
    """finetuning.py:967-980: ``np.array([get_split_signal_fbank_pad(f, spectrogram=True, input_sec=10)[0]])``
    -> ``[N, 998, 128]`` float32."""
    rows = []
    m = 1 + (int(input_sec * sample_rate) - 400) // 160
    for idx, clips in _batches([_wav_path(f) for f in sound_files], sample_rate, batch_bytes, workers):
        wav, off = _to_device(clips)
        res = pl.split_signal_fbank_pad_batch(wav, off, input_sec=input_sec, sample_rate=sample_rate, spectrogram=True,
                                              first_only=True)
        feats = res.features[: int(res.row_offsets[-1])].cpu().numpy().reshape(-1, m, 128)
        if feats.shape[0] != len(idx):
            raise ValueError("an empty recording has no first chunk (the reference raises IndexError here)")
        rows.append(feats)
    x = np.concatenate(rows) if rows else np.zeros((0, m, 128), np.float32)
    if save_to:
        np.save(save_to, x)
    return x
