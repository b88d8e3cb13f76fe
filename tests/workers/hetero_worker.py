"""Worker for heterogeneous (Malleus / Ampelos style) strategies: pipelines with DIFFERENT tensor-parallel degrees train
the same tiny GPT on unequal shares of the global batch; every rank builds the member-local graph of its own pipeline and
parameter gradients are synchronised slice-by-slice across pipelines (grouped_all_reduce).  The loss curve must match
the single-device run.  argv: layout name."""
import json
import os
import sys

import numpy as np
import torch

import hetu_b200 as ht
from hetu_b200.models import GPTConfig, GPTLMHeadModel
from hetu_b200.engine.hetero import HeteroSession
from hetu_b200.models.parallel_config import generate_hetero_ds_parallel_config

layout = sys.argv[1] if len(sys.argv) > 1 else "tp2_tp1"
S, Bg, NL = 16, 8, 4
LAYOUTS = {
    # pipelines (stages: devices + layer range), sequences of the global batch per pipeline
    "tp2_tp1": ([{"stages": [{"devices": [0, 1], "layers": [0, 3]}]}, {"stages": [{"devices": [2], "layers": [0, 3]}]}], [6, 2]),
    "tp2pp2_tp1": ([{"stages": [{"devices": [0, 1], "layers": [0, 1]}, {"devices": [2, 3], "layers": [2, 3]}]},
                    {"stages": [{"devices": [4], "layers": [0, 3]}]}], [5, 3]),
    # one pipeline whose stages have DIFFERENT tensor-parallel degrees (a straggler left stage 1 with a single device)
    "tp2to1_tp1": ([{"stages": [{"devices": [0, 1], "layers": [0, 1]}, {"devices": [2], "layers": [2, 3]}]},
                    {"stages": [{"devices": [3], "layers": [0, 3]}]}], [5, 3]),
    "tp4_tp2_tp1": ([{"stages": [{"devices": [0, 1, 2, 3], "layers": [0, 3]}]}, {"stages": [{"devices": [4, 5], "layers": [0, 3]}]},
                     {"stages": [{"devices": [6], "layers": [0, 3]}]}], [4, 3, 1]),
}
pipelines, shares = LAYOUTS[layout]
num_mb = int(sys.argv[2]) if len(sys.argv) > 2 else 1
if num_mb > 1:
    shares = [6, 2] if len(shares) == 2 else shares      # divisible into micro-batches
world = sum(len(st["devices"]) for p in pipelines for st in p["stages"])
ht.init_comm_group(world)
rank = int(os.environ.get("RANK", "0"))
ht.set_seed(7)
kind = sys.argv[3] if len(sys.argv) > 3 else "gpt"
if kind == "llama":
    from hetu_b200.models import LlamaConfig, LlamaLMHeadModel
    cfg = LlamaConfig(vocab_size=128, hidden_size=32, intermediate_size=64, num_hidden_layers=NL, num_attention_heads=4, num_key_value_heads=2,
                      sequence_parallel=False)
else:
    cfg = GPTConfig(vocab_size=128, n_positions=S, n_embd=32, n_layer=NL, n_head=4, sequence_parallel=os.environ.get("HETERO_SP") == "1")
hetero = generate_hetero_ds_parallel_config(NL, pipelines, zero=False)
sess = HeteroSession(hetero, rank, shares=shares)
local, me = sess.local_cfg, sess.pipeline
assert sess.split_batch(Bg) == shares
bs = sess.batch_slice(Bg)
n_seq = bs.stop - bs.start
with ht.graph("define_and_run", create_new=True) as g:
    model = (LlamaLMHeadModel if kind == "llama" else GPTLMHeadModel)(cfg, [local])
    in_ds, in_dg = ht.nn.parallel.config2ds(local["input"])
    lb_ds, lb_dg = ht.nn.parallel.config2ds(local["label"])
    T = n_seq * S // num_mb
    ids = ht.parallel_placeholder("int64", [T], [in_ds], device_group_hierarchy=[in_dg], name="ids")
    pos = ht.parallel_placeholder("int64", [T], [in_ds], device_group_hierarchy=[in_dg], name="pos")
    lab = ht.parallel_placeholder("int64", [T], [lb_ds], device_group_hierarchy=[lb_dg], name="lab")
    loss = model(ids, pos, lab, seq_len=S)
    opt = ht.AdamOptimizer(lr=1e-2)
    train_op = opt.minimize(loss)
sess.precreate_groups()

rng = np.random.RandomState(0)
X = rng.randint(0, 128, (Bg, S))
L = np.roll(X, -1, axis=1)
P = np.tile(np.arange(S), (Bg, 1))
sl = bs
def mbs(a):
    rows = a[sl]
    per = len(rows) // num_mb
    return [torch.as_tensor(rows[m * per:(m + 1) * per].reshape(-1)) for m in range(num_mb)]


feed = {ids: mbs(X), pos: mbs(P), lab: mbs(L)}
losses = []
for step in range(4):
    out = g.run(loss, [loss, train_op], feed, num_micro_batches=num_mb, grad_scale=sess.grad_scale(n_seq, Bg))
    if rank in sess.last_stage_ranks:
        losses.append(sess.reduce_loss(out[0].float().mean(), n_seq, Bg))
if rank == sess.last_stage_ranks[0]:
    print("LOSSES " + json.dumps(losses))
