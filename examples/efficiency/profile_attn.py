"""Micro-benchmark of the attention paths: dense causal and packed variable-length (cu_seqlens), forward + backward,
device-timed.  On a machine without a GPU it runs a tiny CPU configuration (correctness only).

    python examples/efficiency/profile_attn.py --seq 8192 --batch 2
(ring attention over a context-parallel group is exercised by tests/workers/cp_worker.py)

(ref: examples/efficiency/profile_attn.py)"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import hetu_b200 as ht
from hetu_b200.ops_extra import attn_packed

ap = argparse.ArgumentParser()
ap.add_argument("--seq", type=int, default=0)
ap.add_argument("--batch", type=int, default=0)
ap.add_argument("--heads", type=int, default=16)
ap.add_argument("--head-dim", type=int, default=128)
ap.add_argument("--docs", type=int, default=8, help="documents per packed row in the varlen case")
ap.add_argument("--iters", type=int, default=10)
args = ap.parse_args()
cuda = torch.cuda.is_available() and os.environ.get("HETU_B200_FORCE_CPU") != "1"
S = args.seq or (4096 if cuda else 64)
B = args.batch or (4 if cuda else 2)
H, D = (args.heads, args.head_dim) if cuda else (2, 16)
world = int(os.environ.get("WORLD_SIZE", "1"))
ht.init_comm_group(world)
rank = int(os.environ.get("RANK", "0"))
dtype = "bfloat16" if cuda else "float32"


def bench(name, cu=None):
    """eager forward + backward of packed-QKV attention over T = B*S tokens; `cu` = document boundaries over the T tokens"""
    T = B * S
    x = torch.randn(T, 3 * H * D) * 0.5
    gy = torch.randn(T, H * D) * 0.1
    if cuda:
        x, gy = x.cuda().bfloat16(), gy.cuda().bfloat16()
    G = ht.from_numpy(gy)
    cu_t = None
    if cu is not None:
        cu_t = torch.as_tensor(cu, dtype=torch.int32)
        cu_t = ht.from_numpy(cu_t.cuda() if cuda else cu_t)

    def step():
        X = ht.from_numpy(x, requires_grad=True)
        o = attn_packed(X, S if cu is None else T, H, H, D, is_causal=True, layout="hqkv", cu_seqlens=cu_t)
        ht.sum(o * G).backward()

    for _ in range(2):
        step()
    if cuda:
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    t0 = time.perf_counter()
    for _ in range(args.iters):
        step()
    if cuda:
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.iters
    else:
        ms = (time.perf_counter() - t0) * 1e3 / args.iters
    # causal flops: 2 forward + 5 backward GEMMs of 2 * len^2 * D per head, halved by the mask, summed over documents
    lens = np.diff(cu) if cu is not None else np.full(B, S)
    flops = 7 * 2 * float((lens.astype(np.float64) ** 2).sum()) * D * H / 2
    if rank == 0:
        print(f"{name:30s} {ms:9.3f} ms   {flops / ms / 1e9:8.2f} TFLOP/s (fwd+bwd, incl. the elementwise loss)")


bench(f"dense causal  B{B} S{S}")
# packed rows: every row of S tokens holds `docs` documents of random (16-aligned) lengths
rs = np.random.RandomState(0)
cu = [0]
for r in range(B):
    n_edges = min(args.docs - 1, S // 16 - 1)
    edges = np.sort(rs.choice(np.arange(1, S // 16), n_edges, replace=False)) * 16
    cu += [r * S + int(e) for e in edges] + [(r + 1) * S]
bench(f"packed varlen {args.docs} docs/row", cu=np.array(cu, dtype=np.int64))
