"""torchrun --nproc-per-node N tools/check_peer_gather.py : PeerAllGather == NCCL all_gather, several rounds."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from heart_murmur_detection_b200.dist import PeerAllGather

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", torch.cuda.current_device())
dist.init_process_group("nccl", device_id=dev)
rows, cols = 100003, 64
ok = True
modes = ["copy"] + (["multicast"] if "--no-multicast" not in sys.argv else [])
for mode in modes:
  try:
    ag = PeerAllGather(rows, cols, mode=mode)
  except RuntimeError as e:
    if rank == 0: print("mode", mode, "unavailable:", e)
    continue
  if rank == 0: print("mode", ag.mode)
  for step in range(6):
    i = step % 2
    ag.wait_reusable(i)
    g = torch.Generator(device=dev); g.manual_seed(1000 * step + rank)
    ag.slot(i).copy_(torch.rand((rows, cols), device=dev, generator=g))
    ref_in = ag.slot(i).clone()
    out = ag.gather(i)
    ag.finish()
    torch.cuda.synchronize()
    ref = torch.empty((world * rows, cols), device=dev)
    dist.all_gather_into_tensor(ref, ref_in)
    ok = ok and torch.equal(out, ref)
  # skewed ranks, delayed consumer, no host synchronisation between steps: rank 0 reads each gathered buffer late
  # (a long spin kernel in front of the read) while the other ranks run two steps ahead.  Without the rendezvous
  # before the push a fast rank would overwrite buffer i of the slow rank before it was read.
  steps = 8
  sums = torch.zeros((steps, world), device=dev, dtype=torch.float64)
  for step in range(steps):
    i = step % 2
    ag.wait_reusable(i)
    ag.slot(i).fill_(float(100 * step + rank + 1))
    out = ag.gather(i)
    ag.finish()
    if rank == 0:
        torch.cuda._sleep(200_000_000)  # ~0.1 s
    sums[step] = out.view(world, rows * cols)[:, ::4097].double().mean(dim=1)
  torch.cuda.synchronize()
  want = torch.tensor([[100 * s_ + r + 1 for r in range(world)] for s_ in range(steps)], device=dev, dtype=torch.float64)
  skew_ok = bool(torch.equal(sums, want))
  if not skew_ok and rank == 0: print("mode", mode, "skewed-consumer check FAILED:", sums.tolist())
  ok = ok and skew_ok
t = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("peer gather == nccl all_gather on every rank:", bool(t.item()))
dist.destroy_process_group()
sys.exit(0 if t.item() else 1)
