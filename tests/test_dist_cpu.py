"""N>1 path on CPU: world_size-2 gloo processes shard a ragged batch, compute a stand-in
"feature" per clip, all-gather and restore the global order; the result must be identical to the
single-process result (gathered tensor bit-identical to P=1, order restored)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _fake_features(x: np.ndarray) -> np.ndarray:
    """Deterministic ragged stand-in for the GPU front-end: rows = 1 + n // 512, 4 columns."""
    T = 1 + len(x) // 512
    t = np.arange(T, dtype=np.float32)[:, None]
    return (t * np.float32(0.5) + np.array([x.sum(), x.min(), x.max(), len(x)], dtype=np.float32)[None, :]).astype(np.float32)


def _make():
    rng = np.random.default_rng(5)
    lens = rng.integers(1, 6000, size=23)
    lens[3] = 0  # dropped clip -> zero rows
    off = np.zeros(len(lens) + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    wav = rng.standard_normal(int(off[-1])).astype(np.float32)
    return wav, off


def _single():
    wav, off = _make()
    blocks = [_fake_features(wav[off[i] : off[i + 1]]) if off[i + 1] > off[i] else np.zeros((0, 4), np.float32) for i in range(len(off) - 1)]
    rows = np.array([len(b) for b in blocks])
    return np.concatenate(blocks), np.concatenate([[0], np.cumsum(rows)])


def _worker(rank, world, port, mode, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from heart_murmur_detection_b200 import dist as D

    wav, off = _make()
    lens = np.diff(off)
    shards = D.shard_by_length(lens, world) if mode == "length" else D.shard_contiguous(len(lens), world)
    shard = shards[rank]
    lw, lo = D.local_batch(wav, off, shard)
    blocks = [_fake_features(lw[lo[k] : lo[k + 1]]) if lo[k + 1] > lo[k] else np.zeros((0, 4), np.float32) for k in range(len(shard))]
    rows = [len(b) for b in blocks]
    local = torch.from_numpy(np.concatenate(blocks) if blocks else np.zeros((0, 4), np.float32))
    out, ro = D.all_gather_features(local, rows, shard, len(lens))
    q.put((rank, out.numpy(), ro))
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("mode", ["length", "contiguous"])
def test_two_rank_gather_matches_single_process(mode):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, mode, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref, ref_ro = _single()
    for rank, out, ro in results:
        np.testing.assert_array_equal(ro, ref_ro)
        np.testing.assert_array_equal(out, ref)  # bit identical, original order


def test_shard_by_length_balances_samples():
    from heart_murmur_detection_b200 import dist as D
    from heart_murmur_detection_b200 import synth

    lens = synth.clip_lengths("c2", 5272)
    for world in (2, 4, 8):
        shards = D.shard_by_length(lens, world)
        assert sorted(np.concatenate(shards).tolist()) == list(range(len(lens)))
        loads = np.array([lens[s].sum() for s in shards])
        assert loads.max() / loads.mean() < 1.001
    c = D.shard_contiguous(1000, 8)
    assert [len(x) for x in c] == [125] * 8 and np.concatenate(c).tolist() == list(range(1000))


def _rr_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from heart_murmur_detection_b200 import dist as D

    n_clips, chunk, rows, cols = 24, 3, 5, 4
    rounds = D.chunk_rounds(n_clips, chunk, world)
    final = torch.full((n_clips * rows, cols), -1.0)
    for j in range(rounds):
        c0 = (j * world + rank) * chunk  # first global clip of this rank's chunk in round j
        for i in range(chunk):
            final[(c0 + i) * rows : (c0 + i + 1) * rows] = float(c0 + i) + torch.arange(rows * cols).view(rows, cols) / 100.0
        D.all_gather_round_inplace(final, j, chunk * rows)
    q.put((rank, final.numpy()))
    dist.destroy_process_group()


def test_round_robin_chunks_gather_in_place_in_global_order():
    """c5's sharding: chunks dealt round-robin, features produced in place, one in-place all-gather per round; every
    rank ends with the single-process tensor, no order-restoring pass."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rr_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = np.concatenate([c + np.arange(20, dtype=np.float32).reshape(5, 4) / 100.0 for c in range(24)]).astype(np.float32)
    for _, out in results:
        np.testing.assert_array_equal(out, ref)
    from heart_murmur_detection_b200 import dist as D

    with pytest.raises(ValueError):
        D.chunk_rounds(25, 3, 2)
