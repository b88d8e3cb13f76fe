"""Single-GPU driver for ncu captures of the round-2 kernels: fused dropout + residual + LayerNorm forward, norm backward with
the folded residual-gradient add, single-launch variable-length attention (forward), blockwise nf4 quantisation, and the
generic reduce / softmax / strided-copy kernels."""
import sys

import torch

import hetu_b200 as ht
from hetu_b200 import _C

which = sys.argv[1] if len(sys.argv) > 1 else "all"
torch.manual_seed(0)
bf = lambda *s: torch.randn(*s, device="cuda").to(torch.bfloat16)
leaf = lambda t, g=True: ht.from_numpy(t, requires_grad=g)
rows, cols = 16384, 2048
if which in ("all", "norm"):
    x, r, gamma, beta = bf(rows, cols), bf(rows, cols), bf(cols), bf(cols)
    X, R = leaf(x), leaf(r)
    y, z = ht.dropout_add_norm(X, leaf(gamma), leaf(beta), residual=R, p=0.1, eps=1e-5)
    ht.sum(y * leaf(bf(rows, cols), False) + z).backward()       # norm backward with dx_add (gradient through z)
    torch.cuda.synchronize()
    print("norm ok")
if which in ("all", "varlen"):
    T, H, D = 16384, 16, 128
    bounds = [0, 1000, 1800, 5000, 5064, 9000, 12000, 16384]
    q, k, v = (bf(T, H, D) for _ in range(3))
    cu = ht.from_numpy(torch.tensor(bounds, dtype=torch.int32, device="cuda"))
    Q, K, V = leaf(q), leaf(k), leaf(v)
    o = ht.attn_varlen(Q, K, V, cu, cu, max(b - a for a, b in zip(bounds, bounds[1:])), is_causal=True)
    ht.sum(o).backward()
    torch.cuda.synchronize()
    print("varlen ok")
if which in ("all", "quant"):
    w = bf(8192, 8192)
    q, absmax = ht.quantization(ht.from_numpy(w), "nf4", 64)
    _ = ht.dequantization(q, absmax, "bfloat16", 64, shape=[8192, 8192], quant_type="nf4")
    torch.cuda.synchronize()
    print("quant ok")
if which in ("all", "generic"):
    x = bf(rows, cols)
    _C.g_reduce(_C.GENERIC_REDUCE["sum"], x, [1], False)
    _C.g_reduce(_C.GENERIC_REDUCE["sum"], x, [0], False)
    _C.g_softmax(False, x, 1)
    _C.g_contiguous(x.t())
    _C.g_unary(_C.GENERIC_UNARY["tanh"], x)
    torch.cuda.synchronize()
    print("generic ok")
