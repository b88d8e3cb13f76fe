"""Single-GPU driver for ncu captures of the fused kernels: the GEMM -> reduce-scatter epilogue (peer pointers mapped onto
this GPU, world = 1: the kernel and its code path are the ones used across ranks), the fp8 GEMM, and flash attention."""
import math
import sys

import torch

import hetu_b200 as ht
from hetu_b200 import _C

which = sys.argv[1] if len(sys.argv) > 1 else "all"
torch.manual_seed(0)
T, K, N = 8192, 8192, 2048
x = (torch.randn(T, K, device="cuda")).to(torch.bfloat16)
w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).to(torch.bfloat16)
if which in ("all", "gemm_rs"):
    h = _C.symm_alloc("ncu_stage", T * N * 2, 0, 1)
    _C.symm_open("ncu_stage", [h])
    for _ in range(3):
        y = _C.gemm_reduce_scatter(x, w, "ncu_stage", None, None)
    torch.cuda.synchronize()
    print("gemm_rs ok", float((y.float() - x.float() @ w.float().t()).abs().max()))
if which in ("all", "fp8"):
    qx, sx = _C.quantize_rowwise_e4m3(x)
    qw, sw = _C.quantize_rowwise_e4m3(w)
    for _ in range(3):
        y = _C.gemm_fp8(qx, sx, qw, sw, False, 2)
    torch.cuda.synchronize()
    print("fp8 ok")
if which in ("all", "attn"):
    B, S, H, D = 16, 1024, 16, 128
    q, k, v = (torch.randn(B, S, H, D, device="cuda").to(torch.bfloat16) for _ in range(3))
    Q, Kt, V = (ht.from_numpy(t, requires_grad=True) for t in (q, k, v))
    o = ht.attn(Q, Kt, V, is_causal=True)
    ht.sum(o).backward()
    torch.cuda.synchronize()
    print("attn ok")
