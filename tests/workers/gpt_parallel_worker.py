"""Worker for the strategy-equivalence test: trains a tiny GPT for a few steps under (dp, tp, pp, zero, sp, mbs) and
prints the per-step losses.  Every strategy must reproduce the single-device loss curve (the reference's CI check:
tests/ci_test/train_hetu_gpt_ds_parallel.py compares the loss across 19 strategy files)."""
import json
import os
import sys

import numpy as np
import torch

import hetu_b200 as ht
from hetu_b200.models import GPTConfig, GPTLMHeadModel, generate_ds_parallel_config

dp, tp, pp = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
zero = int(sys.argv[4]) if len(sys.argv) > 4 else 0
sp = int(sys.argv[5]) if len(sys.argv) > 5 else 0
num_mb = int(sys.argv[6]) if len(sys.argv) > 6 else 1
model_kind = sys.argv[7] if len(sys.argv) > 7 else "gpt"
layer_split = [int(v) for v in sys.argv[8].split(",")] if len(sys.argv) > 8 else None
world = dp * tp * pp
ht.init_comm_group(world)
rank = int(os.environ.get("RANK", "0"))
ht.set_seed(7)
S, Bg = 16, 8
if model_kind == "llama":
    from hetu_b200.models import LlamaConfig, LlamaLMHeadModel
    cfg = LlamaConfig(vocab_size=128, hidden_size=32, intermediate_size=64, num_hidden_layers=4, num_attention_heads=4,
                      num_key_value_heads=2, sequence_parallel=bool(sp))
    n_layer = cfg.num_hidden_layers
else:
    cfg = GPTConfig(vocab_size=128, n_positions=S, n_embd=32, n_layer=4, n_head=4, sequence_parallel=bool(sp))
    n_layer = cfg.n_layer
with ht.graph("define_and_run", create_new=True) as g:
    dsc = [generate_ds_parallel_config(n_layer, world, dp, tp, pp, zero=bool(zero), layer_split=layer_split)]
    model = (LlamaLMHeadModel if model_kind == "llama" else GPTLMHeadModel)(cfg, dsc)
    in_ds, in_dg = ht.nn.parallel.config2ds(dsc[0]["input"])
    lb_ds, lb_dg = ht.nn.parallel.config2ds(dsc[0]["label"])
    T = Bg * S // num_mb
    ids = ht.parallel_placeholder("int64", [T], [in_ds], device_group_hierarchy=[in_dg], name="ids")
    pos = ht.parallel_placeholder("int64", [T], [in_ds], device_group_hierarchy=[in_dg], name="pos")
    lab = ht.parallel_placeholder("int64", [T], [lb_ds], device_group_hierarchy=[lb_dg], name="lab")
    loss = model(ids, pos, lab, seq_len=S)
    opt = ht.SGDOptimizer(lr=0.5) if os.environ.get("WORKER_OPT") == "sgd" else ht.AdamOptimizer(lr=1e-2)
    train_op = opt.minimize(loss)

rng = np.random.RandomState(0)
X = rng.randint(0, 128, (Bg, S))
L = np.roll(X, -1, axis=1)
P = np.tile(np.arange(S), (Bg, 1))
# which data-parallel slice does this rank own?  (device layout [pp][dp][tp])
stage = rank // (dp * tp)
d_idx = (rank % (dp * tp)) // tp
per_mb = Bg // num_mb
per_dp = per_mb // dp


def shard(a):
    out = []
    for m in range(num_mb):
        blk = a[m * per_mb:(m + 1) * per_mb]
        out.append(torch.as_tensor(blk[d_idx * per_dp:(d_idx + 1) * per_dp].reshape(-1)))
    return out


losses = []
for step in range(4):
    out = g.run(loss, [loss, train_op], {ids: shard(X), pos: shard(P), lab: shard(L)}, num_micro_batches=num_mb,
                grad_scale=1.0 / dp)
    if out[0] is not None:
        lv = out[0].float().mean()
        if dp > 1:
            # per-replica mean losses -> global mean
            import torch.distributed as dist
            t = lv.clone()
            ranks = [stage * dp * tp + d * tp + (rank % tp) for d in range(dp)]
            grp = ht._C  # noqa
            t = ht._C.comm_all_reduce(t.reshape(1), ranks, "sum") / dp
            lv = t[0]
        losses.append(float(lv))
    else:
        losses.append(None)
if any(l is not None for l in losses) and rank == world - 1:
    print("LOSSES " + json.dumps(losses))
