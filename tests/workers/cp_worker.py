"""Context-parallel (ring) attention worker: every rank holds its SYM / NORMAL chunks of the sequence; outputs and input
gradients must equal full causal attention on the gathered sequence.  argv: cp pattern"""
import os
import sys

import numpy as np
import torch

import hetu_b200 as ht

cp = int(sys.argv[1])
pattern = sys.argv[2] if len(sys.argv) > 2 else "SYM"
varlen = len(sys.argv) > 3 and sys.argv[3] == "varlen"
bounds = [0, 5, 6, 19, 32] if varlen else None      # documents of a packed row: they straddle chunk and rank borders
ht.init_comm_group(cp)
rank = int(os.environ.get("RANK", "0"))
dev = torch.device("cuda", torch.cuda.current_device()) if (torch.cuda.is_available() and os.environ.get("HETU_B200_FORCE_CPU", "0") != "1") else torch.device("cpu")
B, S, H, D = 2, 32, 2, 8
rng = np.random.RandomState(0)
q, k, v, g = (rng.randn(B, S, H, D).astype(np.float32) for _ in range(4))


def my_rows():
    parts = 2 * cp if pattern == "SYM" else cp
    chunks = np.split(np.arange(S), parts)
    return np.concatenate([chunks[rank], chunks[2 * cp - 1 - rank]]) if pattern == "SYM" else chunks[rank]


rows = my_rows()
os.environ["HETU_PARALLEL_ATTN_SPLIT_PATTERN"] = pattern
Q, K, V = (ht.from_numpy(torch.as_tensor(t[:, rows]).contiguous().to(dev), requires_grad=True) for t in (q, k, v))
cu = ht.from_numpy(torch.tensor(bounds + [32, 32], dtype=torch.int32).to(dev)) if varlen else None     # trailing repeats = padding entries
o = ht.parallel_attn(Q, K, V, list(range(cp)), is_causal=True, split_pattern=pattern, cu_seqlens=cu)
ht.sum(o * ht.from_numpy(torch.as_tensor(g[:, rows]).contiguous().to(dev))).backward()
qr, kr, vr = (torch.as_tensor(t).requires_grad_() for t in (q, k, v))
if varlen:
    ref = torch.cat([torch.nn.functional.scaled_dot_product_attention(qr[:, a:b].transpose(1, 2), kr[:, a:b].transpose(1, 2), vr[:, a:b].transpose(1, 2),
                                                                      is_causal=True).transpose(1, 2) for a, b in zip(bounds[:-1], bounds[1:])], dim=1)
else:
    ref = torch.nn.functional.scaled_dot_product_attention(qr.transpose(1, 2), kr.transpose(1, 2), vr.transpose(1, 2), is_causal=True).transpose(1, 2)
(ref * torch.as_tensor(g)).sum().backward()
errs = [float((torch.as_tensor(o.numpy()).cpu() - ref.detach()[:, rows]).abs().max()),
        float((torch.as_tensor(Q.grad.numpy()).cpu() - qr.grad[:, rows]).abs().max()),
        float((torch.as_tensor(K.grad.numpy()).cpu() - kr.grad[:, rows]).abs().max()),
        float((torch.as_tensor(V.grad.numpy()).cpu() - vr.grad[:, rows]).abs().max())]
print("CPERR", rank, errs, flush=True)
assert max(errs) < (1e-4 if dev.type == "cpu" else 2e-3), errs       # GPU fp32 matmuls may run in TF32
