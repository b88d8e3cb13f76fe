"""Neural collaborative filtering on synthetic implicit feedback (users like the items of their own latent group).

    python examples/rec/train_ncf.py --steps 200

(ref: hetu/v1/examples/rec -- hetu_ncf.py, run_hetu.py)"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import hetu_b200 as ht
from hetu_b200.models import NCF
from hetu_b200.v1.metrics import auc

ap = argparse.ArgumentParser()
ap.add_argument("--users", type=int, default=200); ap.add_argument("--items", type=int, default=300); ap.add_argument("--groups", type=int, default=6)
ap.add_argument("--batch", type=int, default=256); ap.add_argument("--steps", type=int, default=200); ap.add_argument("--lr", type=float, default=1e-2)
a = ap.parse_args()
rng = np.random.RandomState(0)
ug, ig = rng.randint(0, a.groups, a.users), rng.randint(0, a.groups, a.items)
like = (ug[:, None] == ig[None, :]) ^ (rng.rand(a.users, a.items) < 0.05)        # group structure + 5 % noise
with ht.graph("define_and_run", create_new=True) as g:
    model = NCF(a.users, a.items, factors=8, mlp_layers=(32, 16, 8))
    U, I = ht.placeholder("int64", [a.batch], name="users"), ht.placeholder("int64", [a.batch], name="items")
    Y = ht.placeholder("float32", [a.batch, 1], name="label")
    loss, logit = model(U, I, Y)
    train = ht.AdamOptimizer(lr=a.lr).minimize(loss)
for step in range(a.steps):
    u, i = rng.randint(0, a.users, a.batch), rng.randint(0, a.items, a.batch)
    y = like[u, i].astype(np.float32).reshape(-1, 1)
    out = g.run(loss, [loss, logit, train], {U: torch.as_tensor(u), I: torch.as_tensor(i), Y: torch.as_tensor(y)})
    if step % 50 == 0 or step == a.steps - 1:
        print(f"step {step} loss {float(out[0]):.4f} auc {auc(y.reshape(-1), 1 / (1 + np.exp(-out[1].float().cpu().numpy().reshape(-1)))):.3f}", flush=True)
