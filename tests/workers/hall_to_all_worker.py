"""hierarchical all-to-all (intra-node stage, layout transform, inter-node stage) must equal the flat all-to-all, forward and
backward (ref: hetu/v1 halltoall_op, mpi_nccl_communication.cu:152-243)."""
import json
import os
import sys

import torch
import torch.distributed as dist

import hetu_b200 as ht

local = int(sys.argv[1]) if len(sys.argv) > 1 else 2
ht.init_comm_group()
rank, world = dist.get_rank(), dist.get_world_size()
use_cuda = torch.cuda.is_available() and os.environ.get("HETU_B200_FORCE_CPU", "0") != "1"
dev = torch.device("cuda", torch.cuda.current_device()) if use_cuda else torch.device("cpu")
ranks = list(range(world))
g = torch.Generator().manual_seed(11 + rank)
x = torch.randn(world * 3, 8, generator=g)
if use_cuda:
    x = x.to(torch.bfloat16)
X = ht.from_numpy(x.to(dev), requires_grad=True)
flat = ht.all_to_all(X, ranks)
hier = ht.hall_to_all(X, ranks, local)
a, b = torch.as_tensor(flat.numpy()).float().cpu(), torch.as_tensor(hier.numpy()).float().cpu()
w = torch.arange(b.numel(), dtype=torch.float32).reshape(b.shape) * 0.01
ht.sum(hier * ht.from_numpy(w.to(dev).to(x.dtype))).backward()
gh = torch.as_tensor(X.grad.numpy()).float().cpu()
# reference gradient: the exchange is its own transpose
ref = torch.empty_like(w)
dist.all_to_all_single(ref, w.to(dev).contiguous() if dist.get_backend() == "nccl" else w.contiguous())
ref = ref.float().cpu()
ok = bool(torch.equal(a, b)) and bool(torch.allclose(gh, ref.to(gh.dtype), atol=0.1 if use_cuda else 1e-6))
res = torch.tensor([1 if ok else 0], device=dev if dist.get_backend() == "nccl" else "cpu")
dist.all_reduce(res, op=dist.ReduceOp.MIN)
if rank == 0:
    print("HA2A " + json.dumps({"ok": bool(int(res.item())), "world": world, "gpus_per_node": local}))
dist.barrier()
dist.destroy_process_group()
