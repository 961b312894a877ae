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


def chunk_rounds(n_clips: int, chunk: int, world: int):
    """Deal chunks of ``chunk`` consecutive clips round-robin over the ranks: in round j rank r owns the clips
    ``[(j * world + r) * chunk, (j * world + r + 1) * chunk)``.  The blocks a round's all-gather collects are then
    contiguous AND in global clip order, so every rank can produce its features in place inside the global-order
    tensor and the collective needs neither a staging buffer nor an order-restoring pass afterwards (contrast
    ``all_gather_features``, which restores the order of length-balanced ragged shards with an index_select).
    Returns the number of rounds; ``n_clips`` must be a multiple of ``chunk * world``."""
    if chunk <= 0 or world <= 0 or n_clips % (chunk * world):
        raise ValueError(f"n_clips={n_clips} is not a multiple of chunk * world = {chunk} * {world}")
    return n_clips // (chunk * world)


def all_gather_round_inplace(final: torch.Tensor, round_index: int, rows_per_chunk: int, group=None):
    """In-place all-gather of one round of ``chunk_rounds``: ``final`` is the global-order feature tensor
    ``[n_clips * rows_per_clip, C]``; this rank has already written its chunk of the round into it.  One
    ``all_gather_into_tensor`` whose input is this rank's slice of its output (NCCL's in-place form)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo = round_index * world * rows_per_chunk
    block = final[lo : lo + world * rows_per_chunk]
    mine = block[rank * rows_per_chunk : (rank + 1) * rows_per_chunk]
    if final.device.type == "cuda":
        dist.all_gather_into_tensor(block, mine, group=group)
    else:  # gloo (CPU tests)
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine.clone(), group=group)
        for r, p in enumerate(parts):
            block[r * rows_per_chunk : (r + 1) * rows_per_chunk].copy_(p)
    return block


def bind_to_gpu_numa_node(device_index: int | None = None) -> dict:
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, BEFORE pinned host buffers are allocated
    (first touch decides where they live): with one process per GPU on a two-socket box, host<->device copies of
    ranks whose staging memory sits on the far socket cross the inter-socket link and contend with each other.
    Returns {"numa_node", "cpus"}; a no-op (node -1) when the topology cannot be read."""
    import os

    idx = torch.cuda.current_device() if device_index is None else int(device_index)
    try:
        bus = torch.cuda.get_device_properties(idx).pci_bus_id
        dom = torch.cuda.get_device_properties(idx).pci_domain_id
        dev = torch.cuda.get_device_properties(idx).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return {"numa_node": -1, "cpus": 0}
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
        return {"numa_node": node, "cpus": len(allowed)}
    except Exception:  # no sysfs / no permission: leave the affinity alone
        return {"numa_node": -1, "cpus": 0}


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


class PeerAllGather:
    """All-gather of equally sized row blocks through peer memory, without SM time.

    Every rank owns a ``[world * rows, cols]`` buffer in symmetric memory (torch.distributed
    ._symmetric_memory: the same allocation mapped into every process of the node over NVLink).  A rank
    produces its block IN PLACE in its own slot and pushes it into the same slot of the other ranks'
    buffers with device-to-device copies (copy engines over NVLink / NVSwitch), one copy stream per
    peer, then joins a device-side barrier: when the barrier releases, every rank holds all blocks.
    Compared with an SM-driven collective the kernels of the next step keep all SMs; the reference has
    no counterpart (single GPU, src/pretrain/cola_training.py:275-278).

    Buffers are double (``n_buffers``) so that the gather of step i overlaps the compute of step i+1.
    """

    def __init__(self, rows: int, cols: int, group=None, n_buffers: int = 2, dtype=torch.float32, mode: str = "auto",
                 push_ctas: int = 16):
        """mode: "multicast" (one multimem.st push through the NVLink switch), "copy" (one copy-engine
        transfer per peer) or "auto" (multicast when the fabric offers it)."""
        import torch.distributed._symmetric_memory as symm_mem

        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.rows, self.cols = int(rows), int(cols)
        dev = torch.device("cuda", torch.cuda.current_device())
        shape = (self.world * self.rows, self.cols)
        self.buffers = [symm_mem.empty(shape, dtype=dtype, device=dev) for _ in range(n_buffers)]
        self.handles = [symm_mem.rendezvous(b, self.group) for b in self.buffers]
        self.peer = [[h.get_buffer(r, shape, dtype) for r in range(self.world)] for h in self.handles]
        self.comm = torch.cuda.Stream(device=dev)
        self.copy_streams = [torch.cuda.Stream(device=dev) for _ in range(max(1, self.world - 1))]
        self.ready = [torch.cuda.Event() for _ in range(n_buffers)]
        self.done = [torch.cuda.Event() for _ in range(n_buffers)]
        self.used = [False] * n_buffers
        self.push_ctas = int(push_ctas)
        has_mc = all(int(getattr(h, "multicast_ptr", 0) or 0) != 0 for h in self.handles)
        if mode == "multicast" and not has_mc:
            raise RuntimeError("NVLink multicast is not available for this group")
        self.mode = "multicast" if (mode in ("auto", "multicast") and has_mc and dtype == torch.float32) else "copy"

    def slot(self, i: int) -> torch.Tensor:
        """This rank's block of buffer i: produce the features here."""
        return self.buffers[i][self.rank * self.rows : (self.rank + 1) * self.rows]

    def wait_reusable(self, i: int, stream=None):
        """Make ``stream`` wait until the previous gather of buffer i has completed on every rank."""
        if self.used[i]:
            (stream or torch.cuda.current_stream()).wait_event(self.done[i])

    def gather(self, i: int, stream=None):
        """Push slot(i) (produced on ``stream``) to every peer and barrier; asynchronous."""
        main = stream or torch.cuda.current_stream()
        self.ready[i].record(main)
        lo, hi = self.rank * self.rows, (self.rank + 1) * self.rows
        src = self.buffers[i][lo:hi]
        if self.mode == "multicast":
            import ctypes as C

            from . import _lib

            n = src.numel()
            mc = int(self.handles[i].multicast_ptr) + lo * self.cols * 4
            with torch.cuda.stream(self.comm):
                self.comm.wait_event(self.ready[i])
                self.handles[i].barrier(channel=0)  # every rank has produced its block, nobody still reads buffer i
                _lib.check(_lib.hmfe_multicast_push(C.c_void_p(src.data_ptr()), C.c_void_p(mc), (n // 4) * 4, self.push_ctas,
                                                    C.c_void_p(self.comm.cuda_stream)), "hmfe_multicast_push")
                self.handles[i].barrier(channel=1)  # every rank's push has landed
                self.done[i].record(self.comm)
            self.used[i] = True
            return self.buffers[i]
        # Rendezvous BEFORE the push (as the multicast path does): a rank that runs ahead must not overwrite buffer i
        # of a peer that has not finished with the previous contents - a peer joins this barrier only after its own
        # ready[i], i.e. after everything it enqueued on its producing stream (including reads of the old buffer i).
        with torch.cuda.stream(self.comm):
            self.comm.wait_event(self.ready[i])
            self.handles[i].barrier(channel=0)
            go = torch.cuda.Event()
            go.record(self.comm)
        evs = []
        for k in range(self.world - 1):
            r = (self.rank + 1 + k) % self.world  # stagger the targets: rank r sends to r+1, r+2, ...
            cs = self.copy_streams[k]
            cs.wait_event(go)
            with torch.cuda.stream(cs):
                self.peer[i][r][lo:hi].copy_(src, non_blocking=True)
                e = torch.cuda.Event()
                e.record(cs)
                evs.append(e)
        with torch.cuda.stream(self.comm):
            for e in evs:
                self.comm.wait_event(e)
            self.handles[i].barrier(channel=1)  # every rank's copies have landed
            self.done[i].record(self.comm)
        self.used[i] = True
        return self.buffers[i]

    def finish(self, stream=None):
        (stream or torch.cuda.current_stream()).wait_stream(self.comm)
