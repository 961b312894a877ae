"""Multi-GPU plumbing: shard recordings across ranks, run the front-end locally, all-gather
the feature tensors (the path's single collective) and restore the original clip order.

Recordings are independent (only intra-clip coupling: IIR state, clip-wide max/min), so the path
shards with no data-path collective; the reference itself is single-GPU
(src/pretrain/cola_training.py:275-278).  One process per GPU, torch.distributed for the
plumbing (NCCL on GPUs; the same code runs on gloo/CPU tensors, which is how the tests cover it).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_contiguous(n_clips: int, world: int):
    """Equal contiguous index ranges (fixed-length configs)."""
    bounds = np.linspace(0, n_clips, world + 1).round().astype(np.int64)
    return [np.arange(bounds[r], bounds[r + 1], dtype=np.int64) for r in range(world)]


def shard_by_length(lengths, world: int):
    """Greedy longest-first assignment to the least-loaded rank: every rank gets (nearly) the
    same number of SAMPLES.  Deterministic; indices inside a shard are kept in ascending order."""
    lengths = np.asarray(lengths, dtype=np.int64)
    order = np.argsort(-lengths, kind="stable")
    load = np.zeros(world, dtype=np.int64)
    shards = [[] for _ in range(world)]
    for i in order:
        r = int(np.argmin(load))
        shards[r].append(int(i))
        load[r] += int(lengths[i])
    return [np.array(sorted(s), dtype=np.int64) for s in shards]


def local_batch(wav_host: np.ndarray, offsets, shard):
    """Concatenate this rank's clips: returns (float32 array, offsets)."""
    offsets = np.asarray(offsets, dtype=np.int64)
    lens = offsets[shard + 1] - offsets[shard]
    off = np.zeros(len(shard) + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    out = np.empty(int(off[-1]), dtype=np.float32)
    for k, i in enumerate(shard):
        out[off[k] : off[k + 1]] = wav_host[offsets[i] : offsets[i + 1]]
    return out, off


def all_gather_features(local: torch.Tensor, rows_per_clip, shard, n_clips_total: int, group=None):
    """All-gather ragged per-clip feature blocks and return them in GLOBAL clip order.

    local          [sum rows, C] features of this rank's clips, in shard order
    rows_per_clip  rows of each local clip (0 for clips the front-end dropped)
    shard          global indices of this rank's clips
    Returns (features [total rows, C] on every rank, row_offsets[n_clips_total + 1]).
    One data collective (all_gather of the row-padded blocks) + one tiny metadata all_gather.
    """
    world = dist.get_world_size(group)
    dev = local.device
    rows_per_clip = np.asarray(rows_per_clip, dtype=np.int64)
    shard = np.asarray(shard, dtype=np.int64)
    # metadata: (clip id, rows) pairs, padded to the largest shard
    n_local = torch.tensor([len(shard), int(rows_per_clip.sum())], dtype=torch.int64, device=dev)
    sizes = [torch.zeros(2, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, n_local, group=group)
    sizes = torch.stack(sizes).cpu().numpy()
    max_clips, max_rows = int(sizes[:, 0].max()), int(sizes[:, 1].max())
    meta = torch.full((max_clips, 2), -1, dtype=torch.int64, device=dev)
    if len(shard):
        meta[: len(shard), 0] = torch.from_numpy(shard).to(dev)
        meta[: len(shard), 1] = torch.from_numpy(rows_per_clip).to(dev)
    metas = [torch.empty_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta, group=group)
    # data: row-padded blocks
    C = local.shape[1]
    send = torch.zeros((max_rows, C), dtype=local.dtype, device=dev)
    send[: local.shape[0]] = local
    gathered = torch.empty((world * max_rows, C), dtype=local.dtype, device=dev)
    if dev.type == "cuda":
        dist.all_gather_into_tensor(gathered, send, group=group)
    else:  # gloo
        parts = [torch.empty_like(send) for _ in range(world)]
        dist.all_gather(parts, send, group=group)
        gathered = torch.cat(parts)
    # restore global order with one gather (index_select) over rows
    rows_global = np.zeros(n_clips_total, dtype=np.int64)
    src_start = np.zeros(n_clips_total, dtype=np.int64)
    for r in range(world):
        m = metas[r].cpu().numpy()
        m = m[m[:, 0] >= 0]
        starts = r * max_rows + np.concatenate([[0], np.cumsum(m[:, 1])[:-1]]) if len(m) else np.zeros(0, np.int64)
        rows_global[m[:, 0]] = m[:, 1]
        src_start[m[:, 0]] = starts
    row_offsets = np.zeros(n_clips_total + 1, dtype=np.int64)
    np.cumsum(rows_global, out=row_offsets[1:])
    idx = np.concatenate([np.arange(s, s + n, dtype=np.int64) for s, n in zip(src_start, rows_global)]) if n_clips_total else np.zeros(0, np.int64)
    out = gathered.index_select(0, torch.from_numpy(idx).to(dev))
    return out, row_offsets
