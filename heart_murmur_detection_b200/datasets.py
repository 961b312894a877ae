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


class _PyRandomStream:
    """Python's global Mersenne Twister read in bulk: the generator state is transplanted into numpy's legacy MT19937
    (``RandomState.random_sample`` and ``random.random`` are the same genrand_res53 formula), uniforms are drawn
    vectorised, and ``commit(k)`` leaves Python's generator exactly where k calls of ``random.random()`` would."""

    def __init__(self):
        self._py = random.getstate()
        ver, internal, self._gauss = self._py
        if ver != 3 or len(internal) != 625:
            raise RuntimeError("unexpected random.getstate() layout")
        self._np = ("MT19937", np.array(internal[:-1], dtype=np.uint32), int(internal[-1]), 0, 0.0)

    def uniforms(self, n: int) -> np.ndarray:
        rs = np.random.RandomState()
        rs.set_state(self._np)
        return rs.random_sample(int(n))

    def commit(self, k: int):
        rs = np.random.RandomState()
        rs.set_state(self._np)
        if k:
            rs.random_sample(int(k))
        _, keys, pos, _, _ = rs.get_state()
        random.setstate((3, tuple(int(x) for x in keys) + (int(pos),), self._gauss))


def cola_draws(rows, max_len=251, augment=True, windowing=False, rate_start=0.1, rate_seq=0.2):
    """All random draws of a batch of COLA items (cola_training.py:56-80; mae_training.py:64-79 with ``windowing``),
    consumed from Python's global ``random`` stream in the reference's order by the native planner
    ``hmfe_cola_draws``.  Returns dict(mask uint8, mask_off, win_start, start1, start2, gain1, gain2, rows_eff)."""
    import ctypes as C

    from . import _lib

    rows = np.ascontiguousarray(rows, dtype=np.int64)
    n = rows.size
    rows_eff = np.where(windowing & (rows > 3 * max_len), 3 * max_len, rows) if windowing else rows
    out = dict(mask=np.zeros(int(rows_eff.sum()) if augment else 0, dtype=np.uint8), mask_off=np.zeros(n, np.int64),
               win_start=np.zeros(n, np.int64), start1=np.zeros(n, np.int64), start2=np.zeros(n, np.int64),
               gain1=np.ones(n, np.float32), gain2=np.ones(n, np.float32), rows_eff=rows_eff)
    stream = _PyRandomStream()
    bound = int((2 * rows_eff.sum() if augment else 0) + 5 * n)
    guess = min(bound, int((1.25 * rows_eff.sum() if augment else 0) + 5 * n + 256))
    for n_u in (guess, bound):
        u = stream.uniforms(n_u)
        ptr = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
        k = int(_lib.hmfe_cola_draws(ptr(u), n_u, ptr(rows), n, int(max_len), int(bool(windowing)), int(bool(augment)),
                                     float(rate_start), float(rate_seq), ptr(out["mask"]), ptr(out["mask_off"]),
                                     ptr(out["win_start"]), ptr(out["start1"]), ptr(out["start2"]), ptr(out["gain1"]),
                                     ptr(out["gain2"])))
        if k >= 0:
            stream.commit(k)
            out["consumed"] = k
            return out
        if k != -100:
            _lib.check(int(k), "hmfe_cola_draws")
    raise RuntimeError("hmfe_cola_draws: uniform stream too short (internal error)")


def cola_batch(store: SpecStore, indices, max_len=251, augment=True, windowing=False):
    """AudioDataset.__getitem__ (method='cola') for a batch of item indices.
    Returns (x1, x2): float32 CUDA tensors [B, max_len, n_cols].  An index may appear several
    times in one batch: every occurrence gets its own mask, crops and gains, as in the reference.
    ``windowing`` is the pre-crop to 3 * max_len of mae_training.py:64-67."""
    idx = np.ascontiguousarray(indices, dtype=np.int64)
    r0 = store.row_offsets[idx]
    T = store.row_offsets[idx + 1] - r0
    dr = cola_draws(T, max_len, augment, windowing)
    Te, w0 = dr["rows_eff"], dr["win_start"]

    windowed = bool(windowing) and bool((T > Te).any())

    def descs(start, gain):
        s = np.maximum(start, 0)
        d = np.zeros(idx.size, dtype=fe.CROP_DTYPE)
        d["src_row"] = r0 + w0 + s
        d["n_rows"] = np.minimum(np.minimum(max_len, Te), Te - s)
        d["spec_id"] = np.arange(idx.size) if windowed else idx
        d["gain"] = gain
        d["mask_off"] = dr["mask_off"] + s if augment else 0
        return d

    dmask = torch.from_numpy(dr["mask"]).to(store.data.device, non_blocking=True) if augment and dr["mask"].size else None
    means = store.means if augment else None
    if augment and windowed:  # random_mask runs on the WINDOW: its fill value is the window's mean, per item
        means = fe.spec_mean_ranges(store.data, r0 + w0, r0 + w0 + Te)
    x1 = fe.spec_crop(store.data, descs(dr["start1"], dr["gain1"]), max_len, dmask, means)
    x2 = fe.spec_crop(store.data, descs(dr["start2"], dr["gain2"]), max_len, dmask, means)
    return x1, x2


def cola_batch_reference_order(store: SpecStore, indices, max_len=251, augment=True):
    """Same as ``cola_batch`` with the draws made one ``random.random()`` at a time in Python (the first
    implementation; kept as the in-package cross-check of the native planner, tests/test_planners_cpu.py)."""
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


def mae_batch(store: SpecStore, indices, max_len=256):
    """AudioDataset.__getitem__ for method 'mae' / 'audiomae' (mae_training.py:82-109) over a batch: items longer
    than ``max_len`` frames get ``random_crop`` (one ``random.random()`` draw each, in batch order - items that are
    padded consume no draw), shorter ones are zero padded at the end.  Returns float32 CUDA [B, max_len, n_cols]."""
    starts = []
    for idx in indices:
        T = store.rows(idx)
        starts.append(int(random.random() * (T - max_len)) if T > max_len else 0)
    return pad_or_crop_batch(store, indices, max_len, starts)


def finetune_batch(store: SpecStore, indices, max_len=256, crop_mode="first", augment=True, spec_augment=False,
                   time_drop_width=100, time_stripes_num=2, freq_drop_width=20, freq_stripes_num=2):
    """AudioDataset.__getitem__ of the fine-tuning script (finetuning.py:74-123) over a batch, per item and in the
    reference's order: crop (``random_crop`` draws from Python's ``random``; ``crop_first`` draws nothing) ->
    ``random_mask`` on the CROPPED item (its mean is the crop's mean) -> ``random_multiply`` -> SpecAugmentation
    (torchlibrosa DropStripes on a batch of one, training mode: per stripe ``distance = torch.randint(0, width)``,
    ``bgn = torch.randint(0, total - distance)``, first the time stripes then the frequency stripes, drawn from
    torch's global generator).  Items must share one length after the crop.  Returns float32 CUDA [B, rows, n_cols]."""
    import ctypes as C

    from . import _lib

    idx = np.ascontiguousarray(indices, dtype=np.int64)
    B = idx.size
    r0 = store.row_offsets[idx]
    T = store.row_offsets[idx + 1] - r0
    rows = int(max_len) if max_len else int(T.max(initial=0))
    starts, masks, gains, rects = np.zeros(B, np.int64), [], np.ones(B, np.float32), []
    n_eff = np.minimum(T, rows) if max_len else T
    if not max_len and B and not (T == rows).all():
        raise ValueError("without max_len every item must have the same number of frames (torch.stack would fail)")
    for k in range(B):
        if max_len and crop_mode == "random":
            starts[k] = max(int(random.random() * (int(T[k]) - rows)), 0)
        if augment:
            masks.append(draw_mask_rows(int(n_eff[k])))
            gains[k] = np.float32(0.9 + random.random() / 5.0)
        if spec_augment:
            for dim, width, num in ((0, time_drop_width, time_stripes_num), (1, freq_drop_width, freq_stripes_num)):
                total = int(n_eff[k]) if dim == 0 else store.n_cols
                for _ in range(num):
                    distance = int(torch.randint(low=0, high=width, size=(1,))[0])
                    bgn = int(torch.randint(low=0, high=total - distance, size=(1,))[0])
                    if distance:
                        rects.append((k, bgn, distance, 0, store.n_cols) if dim == 0 else (k, 0, int(n_eff[k]), bgn, distance))
    d = np.zeros(B, dtype=fe.CROP_DTYPE)
    d["src_row"], d["n_rows"], d["spec_id"], d["gain"] = r0 + starts, n_eff, np.arange(B), gains
    dmask = means = None
    if augment:
        d["mask_off"] = np.concatenate([[0], np.cumsum(n_eff)[:-1]]) if B else 0
        dmask = torch.from_numpy(np.concatenate(masks) if masks else np.zeros(0, np.uint8)).to(store.data.device)
        # random_mask runs AFTER the crop: the fill value is the mean of the cropped item, not of the recording
        means = fe.spec_mean_ranges(store.data, r0 + starts, r0 + starts + n_eff)
    out = fe.spec_crop(store.data, d, rows, dmask, means)
    if rects:
        ctx = fe.default_ctx()
        ra = np.array(rects, dtype=np.dtype([("item", "<i8"), ("row0", "<i4"), ("n_rows", "<i4"), ("col0", "<i4"), ("n_cols", "<i4")]))
        with torch.cuda.device(out.device):
            _lib.check(_lib.hmfe_spec_zero_rects(ctx._h, C.c_void_p(out.data_ptr()), rows, store.n_cols, B,
                                                 ra.ctypes.data_as(C.c_void_p), ra.size, fe._stream_ptr()),
                       "hmfe_spec_zero_rects")
    return out


def pad_or_crop_batch(store: SpecStore, indices, max_len=1024, starts=None):
    """Pad-to-model-size / crop_first (audioMAE/models_mae.py:1178-1181, mae_training.py:88-109):
    rows [start, start+max_len) of each item, zero rows appended when the item is shorter."""
    d = np.zeros(len(indices), dtype=fe.CROP_DTYPE)
    for k, idx in enumerate(indices):
        r0, T = int(store.row_offsets[idx]), store.rows(idx)
        s = 0 if starts is None else int(starts[k])
        d[k] = (r0 + s, max(0, min(max_len, T - s)), idx, np.float32(1.0), 0)
    return fe.spec_crop(store.data, d, max_len)
