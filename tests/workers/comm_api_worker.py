"""communication-group API beyond the graph ops: coalesced all-reduce, reduce, gather, scatter
(ref: hetu/impl/communication/nccl_comm_group.cu AllReduceCoalesce / Reduce / Gather / Scatter)"""
import json
import os

import torch
import torch.distributed as dist

import hetu_b200 as ht

ht.init_comm_group()
rank, world = dist.get_rank(), dist.get_world_size()
use_cuda = torch.cuda.is_available() and os.environ.get("HETU_B200_FORCE_CPU", "0") != "1"
dev = torch.device("cuda", torch.cuda.current_device()) if use_cuda else torch.device("cpu")
C = ht._C
ranks = list(range(world))
xs = [torch.full((3, 2), float(rank + 1), device=dev), torch.arange(5, dtype=torch.float32, device=dev) * (rank + 1), torch.ones((), device=dev)]
before = C.comm_stats()[1].get("all_reduce_coalesce", 0)
ys = C.comm_all_reduce_coalesce(xs, ranks, "sum")
tot = sum(range(1, world + 1))
ok = torch.allclose(ys[0].cpu(), torch.full((3, 2), float(tot))) and torch.allclose(ys[1].cpu(), torch.arange(5.0) * tot) and float(ys[2]) == world
ok = ok and C.comm_stats()[1]["all_reduce_coalesce"] == before + 1          # ONE collective for the whole bucket
r = C.comm_reduce(torch.full((4,), float(rank + 1), device=dev), ranks, world - 1, "sum")
if rank == world - 1:
    ok = ok and torch.allclose(r.cpu(), torch.full((4,), float(tot)))
g = C.comm_gather(torch.full((2,), float(rank), device=dev), ranks, 0)
if rank == 0:
    ok = ok and g.shape == (world, 2) and torch.allclose(g[:, 0].cpu(), torch.arange(world, dtype=torch.float32))
else:
    ok = ok and g.numel() == 0
src = torch.arange(world * 3, dtype=torch.float32, device=dev).reshape(world, 3) if rank == 1 % world else torch.zeros(world, 3, device=dev)
s = C.comm_scatter(src, ranks, 1 % world)
ok = ok and torch.allclose(s.cpu(), torch.arange(3, dtype=torch.float32) + 3 * rank)
t = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("COMMAPI " + json.dumps({"ok": bool(int(t.item())), "world": world}))
dist.barrier()
dist.destroy_process_group()
