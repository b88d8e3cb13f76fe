"""Galvatron layer-wise hybrid parallelism: layers of ONE model run under different (tp, dp) degrees on the same devices
(e.g. layers 0-1 tp2 x dp1, layers 2-3 tp1 x dp2); activations are relocated where the layout changes.  Prints the loss
curve, which must equal the single-device one.  argv: comma-separated tp sizes per layer, world size."""
import json
import os
import sys

import numpy as np
import torch

import hetu_b200 as ht
from hetu_b200.models import GPTConfig, GPTLMHeadModel
from hetu_b200.planner.runtime import build_hybrid_parallel_model, load_plan

tps = [int(v) for v in sys.argv[1].split(",")]
world = int(sys.argv[2])
ht.init_comm_group(world)
rank = int(os.environ.get("RANK", "0"))
ht.set_seed(7)
S, Bg = 16, 8
cfg = GPTConfig(vocab_size=128, n_positions=S, n_embd=32, n_layer=len(tps), n_head=4)
plan = load_plan({"pp_deg": 1, "tp_sizes_enc": ",".join(map(str, tps)), "tp_consecutive_flags": ",".join("1" for _ in tps),
                  "dp_types_enc": ",".join("0" for _ in tps), "checkpoint": ",".join("0" for _ in tps), "world_size": world, "global_bsz": Bg,
                  "chunks": 1})
with ht.graph("define_and_run", create_new=True) as g:
    model, dsc, run_kw = build_hybrid_parallel_model(plan, world, GPTLMHeadModel, cfg)
    in_ds, in_dg = ht.nn.parallel.config2ds(dsc["input"])
    lb_ds, lb_dg = ht.nn.parallel.config2ds(dsc["label"])
    dp_in, dp_lab = in_ds.get(0).get_dim(0), lb_ds.get(0).get_dim(0)
    T = Bg * S
    ids = ht.parallel_placeholder("int64", [T], [in_ds], device_group_hierarchy=[in_dg], name="ids")
    pos = ht.parallel_placeholder("int64", [T], [in_ds], device_group_hierarchy=[in_dg], name="pos")
    lab = ht.parallel_placeholder("int64", [T], [lb_ds], device_group_hierarchy=[lb_dg], name="lab")
    loss = model(ids, pos, lab, seq_len=S)
    train_op = ht.AdamOptimizer(lr=1e-2).minimize(loss)
rng = np.random.RandomState(0)
X = rng.randint(0, 128, (Bg, S))
L = np.roll(X, -1, axis=1)
P = np.tile(np.arange(S), (Bg, 1))


def shard(a, ds):
    dp = ds.get(0).get_dim(0)
    idx = ds.get(0).map_device_to_state_index(rank).get(0, 0) if dp > 1 else 0
    per = Bg // dp
    return [torch.as_tensor(a[idx * per:(idx + 1) * per].reshape(-1))]


losses = []
for step in range(4):
    out = g.run(loss, [loss, train_op], {ids: shard(X, in_ds), pos: shard(P, in_ds), lab: shard(L, lb_ds)}, num_micro_batches=1,
                grad_scale=1.0 / dp_lab)
    lv = out[0].float().mean().reshape(1)
    if dp_lab > 1:
        lv = ht._C.comm_all_reduce(lv, list(range(world)), "sum") / world
    losses.append(float(lv[0]))
if rank == 0:
    print("LOSSES " + json.dumps(losses))
