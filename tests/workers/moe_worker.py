"""EP/DP worker: GPT-MoE with experts sharded over the data-parallel ranks must reproduce the single-device loss curve
(capacity large enough that no token is dropped).  argv: world [gate]"""
import os
import sys

import numpy as np
import torch

import hetu_b200 as ht
from hetu_b200.models import GPTMoELMHeadModel, MoEConfig, generate_ds_parallel_config

world = int(sys.argv[1])
gate = sys.argv[2] if len(sys.argv) > 2 else "topk"
ht.init_comm_group(world)
rank = int(os.environ.get("RANK", "0"))
ht.set_seed(7)
S, Bg = 16, 8
cfg = MoEConfig(vocab_size=128, n_positions=S, n_embd=32, n_layer=2, n_head=4, num_experts=4, top_k=2, capacity_factor=float(world) * 4.0,
                gate_type=gate, moe_every=1, ep_ranks=tuple(range(world)) if world > 1 else (), aux_loss_weight=0.0)
with ht.graph("define_and_run", create_new=True) as g:
    dsc = [generate_ds_parallel_config(cfg.n_layer, world, world, 1, 1, zero=False)]
    model = GPTMoELMHeadModel(cfg, dsc)
    in_ds, in_dg = ht.nn.parallel.config2ds(dsc[0]["input"])
    T = Bg * S
    ids = ht.parallel_placeholder("int64", [T], [in_ds], device_group_hierarchy=[in_dg], name="ids")
    pos = ht.parallel_placeholder("int64", [T], [in_ds], device_group_hierarchy=[in_dg], name="pos")
    lab = ht.parallel_placeholder("int64", [T], [in_ds], device_group_hierarchy=[in_dg], name="lab")
    loss = model(ids, pos, lab, seq_len=S)
    train_op = ht.AdamOptimizer(lr=1e-2).minimize(loss)
rng = np.random.RandomState(0)
X = rng.randint(0, 128, (Bg, S))
L = np.roll(X, -1, axis=1)
P = np.tile(np.arange(S), (Bg, 1))
per = Bg // world
sl = slice(rank * per, (rank + 1) * per)
losses = []
for step in range(4):
    out = g.run(loss, [loss, train_op], {ids: torch.as_tensor(X[sl].reshape(-1)), pos: torch.as_tensor(P[sl].reshape(-1)),
                                         lab: torch.as_tensor(L[sl].reshape(-1))}, grad_scale=1.0 / world)
    lv = out[0].float().mean()
    if world > 1:
        lv = (ht._C.comm_all_reduce(lv.reshape(1), list(range(world)), "sum") / world)[0]
    losses.append(float(lv))
if rank == 0:
    print("LOSSES", losses)
