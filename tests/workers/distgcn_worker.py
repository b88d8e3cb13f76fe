"""p ranks train a 2-layer GCN on one random graph with the 1.5-D partitioning (replication c); rank 0 also trains the same model on
the whole graph alone: losses and final weights must agree."""
import json
import sys

import numpy as np
import torch
import torch.distributed as dist

import hetu_b200 as ht
from hetu_b200.models.gnn import DistGCN15D, normalise_adjacency

c = int(sys.argv[1])
comm = sys.argv[2] if len(sys.argv) > 2 else "groups"
ht.init_comm_group()
rank, p = dist.get_rank(), dist.get_world_size()
rng = np.random.RandomState(0)
n, f, hdim, k = 96, 12, 16, 4
edges = rng.randint(0, n, (2, 400))
idx, val = normalise_adjacency(edges, n)
x = rng.randn(n, f).astype(np.float32)
y = rng.randint(0, k, n)
mask = rng.rand(n) < 0.7
for q in range(p // c):                      # pre-create the groups collectively (same order on every rank)
    ht._C.comm_create_group([q * c + j for j in range(c)])
for j in range(c):
    ht._C.comm_create_group([q * c + j for q in range(p // c)])
m = DistGCN15D(idx, val, x, y, [f, hdim, k], rank, p, c, lr=0.5, seed=3, train_mask=mask, comm=comm)
losses = [m.step() for _ in range(8)]
if rank == 0:
    # single-process reference: plain dense training of the same model
    A = torch.sparse_coo_tensor(torch.as_tensor(idx), torch.as_tensor(val), (n, n)).to_dense()
    g = torch.Generator().manual_seed(3)
    ws = [(torch.randn(a, b, generator=g) * (1.0 / np.sqrt(a))).requires_grad_() for a, b in ((f, hdim), (hdim, k))]
    ref = []
    X, Y, M = torch.as_tensor(x), torch.as_tensor(y), torch.as_tensor(mask)
    for _ in range(8):
        h = torch.relu(A @ (X @ ws[0]))
        z = A @ (h @ ws[1])
        l = torch.nn.functional.cross_entropy(z[M], Y[M])
        ref.append(float(l))
        gs = torch.autograd.grad(l, ws)
        with torch.no_grad():
            for w, gg in zip(ws, gs):
                w -= 0.5 * gg
    wdiff = max(float((a - b.detach()).abs().max()) for a, b in zip(m.weights, ws))
    print("DISTGCN " + json.dumps({"p": p, "c": c, "losses": losses, "ref": ref, "wdiff": wdiff, "bytes": m.bytes_moved}), flush=True)
dist.barrier()
