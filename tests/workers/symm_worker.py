"""2..8-GPU worker: numerics + timing of the symmetric-memory collectives and the fused GEMM->reduce-scatter against
NCCL (launched by tests/test_symm_gpu.py or scripts/gpu_symm.sh through torchrun-style env)."""
import json
import os
import sys

import torch
import torch.distributed as dist

import hetu_b200 as ht
from hetu_b200.parallel import SymmetricBuffer

ht.init_comm_group()
rank, world = dist.get_rank(), dist.get_world_size()
dev = torch.device("cuda", torch.cuda.current_device())
res = {}


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


torch.manual_seed(100 + rank)
n = 8 * 1024 * 1024          # elements per rank (bf16: 16 MiB)
buf = SymmetricBuffer("t", 2 * n * world * 2 + 1024)
x = buf.tensor([n * world], "bfloat16")
src = (torch.randn(n * world, device=dev) * 0.5).to(torch.bfloat16)

# --- all-gather (each rank contributes its first n elements)
x.copy_(src)
out = torch.empty(n * world, dtype=torch.bfloat16, device=dev)
buf.all_gather(n * 2, out)
ref = [torch.empty(n, dtype=torch.bfloat16, device=dev) for _ in range(world)]
dist.all_gather(ref, src[:n].contiguous())
res["all_gather_ok"] = bool(torch.equal(out, torch.cat(ref)))
# --- reduce-scatter
x.copy_(src)
rs = torch.empty(n, dtype=torch.bfloat16, device=dev)
buf.reduce_scatter(n, rs)
ref_rs = torch.empty(n, dtype=torch.float32, device=dev)
dist.reduce_scatter_tensor(ref_rs, src.float())
res["reduce_scatter_err"] = float((rs.float() - ref_rs).abs().max())
# --- all-reduce (in place)
x.copy_(src)
buf.all_reduce_(n * world, True)
ref_ar = src.float().clone()
dist.all_reduce(ref_ar)
res["all_reduce_err"] = float((x.float() - ref_ar).abs().max())
# --- all-to-all
x.copy_(src)
a2a = torch.empty(n * world, dtype=torch.bfloat16, device=dev)
buf.all_to_all(n * 2, a2a)
ref_a2a = torch.empty_like(src)
dist.all_to_all_single(ref_a2a, src)
res["all_to_all_ok"] = bool(torch.equal(a2a, ref_a2a))
# --- timing vs NCCL
x.copy_(src)
res["ms_symm_all_gather"] = timed(lambda: buf.all_gather(n * 2, out))
res["ms_nccl_all_gather"] = timed(lambda: dist.all_gather_into_tensor(out, src[:n]))
res["ms_symm_reduce_scatter"] = timed(lambda: buf.reduce_scatter(n, rs))
res["ms_nccl_reduce_scatter"] = timed(lambda: dist.reduce_scatter_tensor(rs, src))
res["ms_symm_all_reduce"] = timed(lambda: buf.all_reduce_(n * world, True))
tmp = src.clone()
res["ms_nccl_all_reduce"] = timed(lambda: dist.all_reduce(tmp))
res["payload_MiB_per_rank"] = n * 2 / 2**20

# --- NVLS: VMM allocation + multicast mapping; the NVSwitch reduces (multimem.ld_reduce) and replicates (multimem.st)
from hetu_b200.parallel.symm import multicast_available
res["multicast_supported"] = bool(multicast_available())
if res["multicast_supported"] and os.environ.get("HETU_NVLS", "1") != "0":
    try:
        mbuf = SymmetricBuffer("mc", n * world * 2 + 1024, multicast=True)
        res["multicast_mapped"] = bool(mbuf.has_multicast)
    except Exception as e:                      # noqa: BLE001 -- reported, the IPC path above is still valid
        res["multicast_mapped"] = False
        res["multicast_error"] = f"{type(e).__name__}: {str(e)[:300]}"
    if res.get("multicast_mapped"):
        mx = mbuf.tensor([n * world], "bfloat16")
        # all-reduce in place
        mx.copy_(src)
        mbuf.mc_all_reduce_(n * world, True)
        res["mc_all_reduce_err"] = float((mx.float() - ref_ar).abs().max())
        # reduce-scatter
        mx.copy_(src)
        mrs = torch.empty(n, dtype=torch.bfloat16, device=dev)
        mbuf.mc_reduce_scatter(n, mrs)
        res["mc_reduce_scatter_err"] = float((mrs.float() - ref_rs).abs().max())
        # all-gather: every rank multicasts its first n elements into slot `rank` of every buffer
        mx.zero_()
        torch.cuda.synchronize(); dist.barrier()
        mbuf.mc_all_gather(src[:n].contiguous())
        res["mc_all_gather_ok"] = bool(torch.equal(mx, torch.cat(ref)))
        mx.copy_(src)
        res["ms_mc_all_reduce"] = timed(lambda: mbuf.mc_all_reduce_(n * world, True))
        res["ms_mc_reduce_scatter"] = timed(lambda: mbuf.mc_reduce_scatter(n, mrs))
        shard = src[:n].contiguous()
        res["ms_mc_all_gather"] = timed(lambda: mbuf.mc_all_gather(shard))

# --- fused GEMM -> reduce-scatter (row-parallel linear of a TP group = all ranks)
T, K, N = 8192, 8192 // world, 2048
stage = SymmetricBuffer("stage", T * N * 2)
xa = (torch.randn(T, K, device=dev) * 0.5).to(torch.bfloat16)
w = (torch.randn(N, K, device=dev) * 0.02).to(torch.bfloat16)
bias = torch.randn(N, device=dev).to(torch.bfloat16)
resid = torch.randn(T // world, N, device=dev).to(torch.bfloat16)
y = stage.gemm_reduce_scatter(xa, w, bias, resid)
part = (xa.float() @ w.float().t())
ref_y = torch.empty(T // world, N, dtype=torch.float32, device=dev)
dist.reduce_scatter_tensor(ref_y, part)
ref_y = ref_y + bias.float() + resid.float()
res["gemm_rs_err"] = float((y.float() - ref_y).abs().max())
res["gemm_rs_ref_max"] = float(ref_y.abs().max())


def unfused():
    p = _C_gemm(xa, w)
    o = torch.empty(T // world, N, dtype=torch.bfloat16, device=dev)
    dist.reduce_scatter_tensor(o, p)
    return o + bias + resid


_C_gemm = lambda a, b: ht._C.gemm(a, b)
res["ms_fused_gemm_rs"] = timed(lambda: stage.gemm_reduce_scatter(xa, w, bias, resid))
res["ms_unfused_gemm_nccl_rs"] = timed(unfused)
res["ms_gemm_only"] = timed(lambda: _C_gemm(xa, w))
flops = 2.0 * T * K * N
res["gemm_rs_roofline_ms"] = max(flops / 1.409e15 * 1e3, (world - 1) / world * T * N * 2 / 770e9 * 1e3)
if rank == 0:
    print("SYMM " + json.dumps(res))
dist.barrier()
dist.destroy_process_group()
