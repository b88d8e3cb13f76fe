"""GCN on a random community graph + the 1.5-D partition used by DistGCN."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import hetu_b200 as ht
from hetu_b200.models import GCN, normalise_adjacency, partition_15d

rng = np.random.RandomState(0)
n, f, c = 600, 32, 3
labels = rng.randint(0, c, n)
src = rng.randint(0, n, 6000)
dst = np.where(rng.rand(6000) < 0.8, rng.permutation(n)[np.searchsorted(np.sort(labels), labels[src]) % n], rng.randint(0, n, 6000))
idx, val = normalise_adjacency(np.stack([src, dst]), n)
x = (np.eye(c)[labels] @ rng.randn(c, f) + rng.randn(n, f)).astype(np.float32)
with ht.graph("define_and_run", create_new=True) as g:
    model = GCN(f, 64, c)
    I, V = ht.from_numpy(torch.as_tensor(idx)), ht.from_numpy(torch.as_tensor(val))
    X, Y = ht.placeholder("float32", [n, f], name="x"), ht.placeholder("int64", [n], name="y")
    loss, logits = model(I, V, X, n, Y)
    train = ht.AdamOptimizer(lr=0.02).minimize(loss)
for step in range(60):
    out = g.run(loss, [loss, logits, train], {X: torch.as_tensor(x), Y: torch.as_tensor(labels)})
    if step % 20 == 0:
        print(f"step {step} loss {float(out[0]):.4f} acc {(out[1].argmax(-1).numpy() == labels).mean():.3f}", flush=True)
print("1.5-D layout for 8 devices, replication 2:", [(d['device'], d['rows'], d['col_chunk']) for d in partition_15d(n, 8, 2)][:4], "...")
