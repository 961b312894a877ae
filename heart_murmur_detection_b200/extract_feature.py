"""Drop-in mirror of the front-end functions of the reference's
``src/benchmark/baseline/extract_feature.py`` (get_split_signal_fbank :213-247,
split_sample :250-259).  The extractor loops themselves (model forward) stay in the reference."""
from __future__ import annotations

import os

import numpy as np

from . import audio_io
from . import frontend as fe
from . import pipeline as pl
from .util import _one


def split_sample(sample, desired_length, sample_rate, hop_len=0):
    clip = np.asarray(sample).copy()
    return [clip[s : s + n] for _, s, n in fe.plan_split_sample(len(clip), desired_length, sample_rate)]


def get_split_signal_fbank(data_folder, filename, input_sec=10, sample_rate=16000):
    data, rate = audio_io.load(os.path.join(data_folder, filename + ".wav"), sr=sample_rate)
    wav, off = _one(data)
    res = pl.split_signal_fbank_batch(wav, off, input_sec, rate)
    return [res.chunk(k).cpu() for k in range(len(res.chunks.starts))]
