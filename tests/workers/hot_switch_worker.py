"""2-rank worker: alternate between a data-parallel and a tensor-parallel strategy every step (hot switching re-shards
parameters and Adam states in place); the loss curve must equal training under the data-parallel strategy alone.
argv: mode (switch | single)"""
import os
import sys

import numpy as np
import torch

import hetu_b200 as ht
from hetu_b200.models import GPTConfig, GPTLMHeadModel, generate_ds_parallel_config

mode = sys.argv[1]
world = 2
ht.init_comm_group(world)
rank = int(os.environ.get("RANK", "0"))
ht.set_seed(7)
S, Bg = 16, 8
strategies = [generate_ds_parallel_config(2, world, 2, 1, 1, zero=False), generate_ds_parallel_config(2, world, 1, 2, 1, zero=False)]
cfg = GPTConfig(vocab_size=128, n_positions=S, n_embd=32, n_layer=2, n_head=4)
with ht.graph("define_and_run", create_new=True, num_strategy=2) as g:
    model = GPTLMHeadModel(cfg, strategies)
    hs = [ht.nn.parallel.config2ds(s["input"]) for s in strategies]
    mk = lambda n: ht.parallel_placeholder("int64", [Bg * S], [h[0] for h in hs], device_group_hierarchy=[h[1] for h in hs], name=n)   # noqa: E731
    ids, pos, lab = mk("ids"), mk("pos"), mk("lab")
    loss = model(ids, pos, lab, seq_len=S)
    train_op = ht.AdamOptimizer(lr=1e-2).minimize(loss)
rng = np.random.RandomState(0)
losses = []
for step in range(6):
    X = rng.randint(0, 128, (Bg, S))
    L = np.roll(X, -1, axis=1)
    P = np.tile(np.arange(S), (Bg, 1))
    sid = step % 2 if mode == "switch" else 0
    dp = 2 if sid == 0 else 1
    sl = slice(rank * (Bg // 2), (rank + 1) * (Bg // 2)) if dp == 2 else slice(0, Bg)
    out = g.run(loss, [loss, train_op], {ids: torch.as_tensor(X[sl].reshape(-1)), pos: torch.as_tensor(P[sl].reshape(-1)),
                                         lab: torch.as_tensor(L[sl].reshape(-1))}, cur_strategy_id=sid, grad_scale=1.0 / dp)
    lv = out[0].float().mean()
    if dp == 2:
        lv = (ht._C.comm_all_reduce(lv.reshape(1), [0, 1], "sum") / 2)[0]
    losses.append(float(lv))
if rank == 0:
    print("LOSSES", losses)
