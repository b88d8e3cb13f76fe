"""Whole-model context parallelism: a tiny Llama trains with every sequence split over a CP ring (SYM zig-zag chunks), ring
attention inside every layer, rotary embeddings at the original positions, parameter gradients reduced over the ring.
The loss curve must equal the single-device run of tests/workers/gpt_parallel_worker.py (model kind llama).
argv: cp degree, tp degree"""
import json
import os
import sys

import numpy as np
import torch

import hetu_b200 as ht
from hetu_b200.models import LlamaConfig, LlamaLMHeadModel, generate_ds_parallel_config

cp, tp = int(sys.argv[1]), int(sys.argv[2]) if len(sys.argv) > 2 else 1
world = cp * tp
ht.init_comm_group(world)
rank = int(os.environ.get("RANK", "0"))
ht.set_seed(7)
S, Bg = 16, 8
cp_idx, tp_idx = rank // tp, rank % tp                      # device layout [dcp][tp]
ring = tuple(c * tp + tp_idx for c in range(cp))
cfg = LlamaConfig(vocab_size=128, hidden_size=32, intermediate_size=64, num_hidden_layers=4, num_attention_heads=4, num_key_value_heads=2,
                  cp_ranks=ring)
chunks = np.split(np.arange(S), 2 * cp)
rows = np.concatenate([chunks[cp_idx], chunks[2 * cp - 1 - cp_idx]])     # SYM: chunk i and its mirror
Sl = len(rows)
with ht.graph("define_and_run", create_new=True) as g:
    dsc = [generate_ds_parallel_config(cfg.num_hidden_layers, world, 1, tp, 1, cp=cp, zero=False)]
    model = LlamaLMHeadModel(cfg, dsc)
    in_ds, in_dg = ht.nn.parallel.config2ds(dsc[0]["input"])
    lb_ds, lb_dg = ht.nn.parallel.config2ds(dsc[0]["label"])
    T = Bg * S                                                          # global tokens; every ring member feeds T / cp
    ids = ht.parallel_placeholder("int64", [T], [in_ds], device_group_hierarchy=[in_dg], name="ids")
    pos = ht.parallel_placeholder("int64", [T], [in_ds], device_group_hierarchy=[in_dg], name="pos")
    lab = ht.parallel_placeholder("int64", [T], [lb_ds], device_group_hierarchy=[lb_dg], name="lab")
    loss = model(ids, pos, lab, seq_len=Sl)
    train_op = (ht.SGDOptimizer(lr=0.5) if os.environ.get("WORKER_OPT") == "sgd" else ht.AdamOptimizer(lr=1e-2)).minimize(loss)
rng = np.random.RandomState(0)
X = rng.randint(0, 128, (Bg, S))
L = np.roll(X, -1, axis=1)
P = np.tile(np.arange(S), (Bg, 1))
feed = {ids: [torch.as_tensor(X[:, rows].reshape(-1))], pos: [torch.as_tensor(P[:, rows].reshape(-1))], lab: [torch.as_tensor(L[:, rows].reshape(-1))]}
losses = []
for step in range(4):
    out = g.run(loss, [loss, train_op], feed, num_micro_batches=1, grad_scale=1.0 / cp)
    lv = ht._C.comm_all_reduce(out[0].float().mean().reshape(1), list(ring), "sum") / cp
    losses.append(float(lv[0]))
if rank == 0:
    print("LOSSES " + json.dumps(losses))
