"""GPU batchers for the spectrogram-domain Dataset ops of the reference's SSL training
(COLA: src/pretrain/cola_training.py:56-80; MAE: src/pretrain/mae_training.py:57-109;
fine-tuning: src/benchmark/other_eval/finetuning.py:74-123).

The reference does, per item and in this order (Python ``random`` stream):
    x = random_mask(x)                     T draws (+1 per frame following a masked frame)
    x1 = random_crop(x, max_len); x2 = random_crop(x, max_len)      1 draw each
    x1 = random_multiply(x1);   x2 = random_multiply(x2)            1 draw each
Here the draws are made on the host in exactly that order (bit-exact crop starts, mask rows
and gains) and applied to whole batches by one kernel (frontend.spec_crop).
"""
from __future__ import annotations

import random

import numpy as np
import torch

from . import frontend as fe
from .util import draw_mask_rows


class SpecStore:
    """Ragged set of spectrograms resident on the GPU: [sum T_i, n_cols] + row offsets."""

    def __init__(self, specs, device=None):
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        T = [int(s.shape[0]) for s in specs]
        self.row_offsets = np.zeros(len(T) + 1, dtype=np.int64)
        np.cumsum(T, out=self.row_offsets[1:])
        self.n_cols = int(specs[0].shape[1])
        self.data = torch.cat([torch.as_tensor(np.asarray(s), dtype=torch.float32) for s in specs]).to(dev).contiguous()
        self.means = fe.spec_means(self.data, self.row_offsets)

    @classmethod
    def from_features(cls, features: torch.Tensor, row_offsets):
        """Wrap spectrograms that are already on the GPU (a pipeline.FeatureBatch): the on-the-fly
        form of producer + Dataset (heart_pressl.py:58-99 followed by cola_training.py:56-80) with no
        .npy round trip.  Whole-recording normalisation is preserved because the rows come from
        the whole-recording log-mel."""
        self = cls.__new__(cls)
        self.row_offsets = np.ascontiguousarray(row_offsets, dtype=np.int64)
        self.n_cols = int(features.shape[1])
        self.data = features[: int(self.row_offsets[-1])].contiguous()
        self.means = fe.spec_means(self.data, self.row_offsets)
        return self

    def rows(self, i):
        return int(self.row_offsets[i + 1] - self.row_offsets[i])


def cola_batch(store: SpecStore, indices, max_len=251, augment=True):
    """AudioDataset.__getitem__ (method='cola') for a batch of item indices.
    Returns (x1, x2): float32 CUDA tensors [B, max_len, n_cols].  An index may appear several
    times in one batch: every occurrence gets its own mask, crops and gains, as in the reference."""
    masks, mask_base = [], 0
    d1 = np.zeros(len(indices), dtype=fe.CROP_DTYPE)
    d2 = np.zeros(len(indices), dtype=fe.CROP_DTYPE)
    for k, idx in enumerate(indices):
        r0, T = int(store.row_offsets[idx]), store.rows(idx)
        if augment:
            masks.append(draw_mask_rows(T))  # one fresh mask per item, shared by its two crops
        s1 = int(random.random() * (T - max_len))
        s2 = int(random.random() * (T - max_len))
        g1 = 0.9 + random.random() / 5.0 if augment else 1.0
        g2 = 0.9 + random.random() / 5.0 if augment else 1.0
        n = min(max_len, T)
        s1, s2 = max(s1, 0), max(s2, 0)
        d1[k] = (r0 + s1, min(n, T - s1), idx, np.float32(g1), mask_base + s1)
        d2[k] = (r0 + s2, min(n, T - s2), idx, np.float32(g2), mask_base + s2)
        mask_base += T if augment else 0
    dmask = torch.from_numpy(np.concatenate(masks)).to(store.data.device) if augment and masks else None
    means = store.means if augment else None
    x1 = fe.spec_crop(store.data, d1, max_len, dmask, means)
    x2 = fe.spec_crop(store.data, d2, max_len, dmask, means)
    return x1, x2


def pad_or_crop_batch(store: SpecStore, indices, max_len=1024, starts=None):
    """Pad-to-model-size / crop_first (audioMAE/models_mae.py:1178-1181, mae_training.py:88-109):
    rows [start, start+max_len) of each item, zero rows appended when the item is shorter."""
    d = np.zeros(len(indices), dtype=fe.CROP_DTYPE)
    for k, idx in enumerate(indices):
        r0, T = int(store.row_offsets[idx]), store.rows(idx)
        s = 0 if starts is None else int(starts[k])
        d[k] = (r0 + s, max(0, min(max_len, T - s)), idx, np.float32(1.0), 0)
    return fe.spec_crop(store.data, d, max_len)
