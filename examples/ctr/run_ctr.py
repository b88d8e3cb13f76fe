"""CTR models on synthetic Criteo-shaped data; --compress swaps the embedding table for a compressed one; --ps routes the
embeddings through the parameter server with the HET cache."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import hetu_b200 as ht
from hetu_b200.models import DCN, WDL, DeepFM
from hetu_b200.tools.emb_compress import build_compressed_embedding

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="wdl", choices=["wdl", "deepfm", "dcn"]); ap.add_argument("--compress", default=None)
ap.add_argument("--steps", type=int, default=50)
a = ap.parse_args()
B, F, D, N = 256, 26, 16, 20000
rng = np.random.RandomState(0)
kw = {"hash": {"buckets": 2048}, "robe": {"array_size": 32768}, "adapt": {"hot": 512, "buckets": 2048}, "dhe": {"num_hashes": 64}}.get(a.compress or "", {})
with ht.graph("define_and_run", create_new=True) as g:
    emb = build_compressed_embedding(a.compress, N, D, **kw) if a.compress else None
    model = {"wdl": WDL, "deepfm": DeepFM, "dcn": DCN}[a.model](N, D, num_fields=F, num_dense=13, embedding=emb)
    d, s, y = ht.placeholder("float32", [B, 13], name="dense"), ht.placeholder("int64", [B, F], name="sparse"), ht.placeholder("float32", [B, 1], name="y")
    loss, logit = model(d, s, y)
    train = ht.AdamOptimizer(lr=0.01).minimize(loss)
if emb is not None:
    print(f"embedding compression {a.compress}: {emb.compression_ratio():.1f}x")
for step in range(a.steps):
    dense = rng.randn(B, 13).astype(np.float32)
    sparse = rng.zipf(1.3, (B, F)) % N
    label = ((dense[:, :1] + (sparse[:, :1] % 3 == 0)) > 0.5).astype(np.float32)
    out = g.run(loss, [loss, train], {d: torch.as_tensor(dense), s: torch.as_tensor(sparse), y: torch.as_tensor(label)})
    if step % 10 == 0:
        print(f"step {step} loss {float(out[0]):.4f}", flush=True)
