"""Heterogeneous training: pipelines with different tensor-parallel degrees / stage counts train one GPT on unequal shares
of the global batch (what Malleus plans around stragglers and LobRA deploys for mixed sequence lengths).

    python -m torch.distributed.run --nproc-per-node 5 --master-addr 127.0.0.1 examples/hetero/train_hetero.py
    (CPU / gloo: add HETU_B200_FORCE_CPU=1)

Every rank builds the homogeneous graph of its own pipeline; parameter gradients are synchronised across pipelines by
`grouped_all_reduce` (see hetu_b200/engine/hetero.py).
(ref: examples/hetero/train_hetu.py with hetero ds_parallel_config, examples/malleus)"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import hetu_b200 as ht
from hetu_b200.engine import HeteroSession
from hetu_b200.models import GPTConfig, GPTLMHeadModel
from hetu_b200.models.parallel_config import generate_hetero_ds_parallel_config

world = int(os.environ.get("WORLD_SIZE", "1"))
assert world == 5, "this example lays 5 ranks out as (tp2 x pp2) + tp1; edit `pipelines` for other sizes"
S, GLOBAL_BATCH, LAYERS = 64, 16, 4
pipelines = [{"stages": [{"devices": [0, 1], "layers": [0, 1]}, {"devices": [2, 3], "layers": [2, 3]}]},
             {"stages": [{"devices": [4], "layers": [0, 3]}]}]
shares = [3, 1]                       # the 4-GPU pipeline takes 3/4 of every global batch
ht.init_comm_group(world)
ht.set_seed(1)
hetero_cfg = generate_hetero_ds_parallel_config(LAYERS, pipelines, zero=False)
sess = HeteroSession(hetero_cfg, shares=shares)
cfg = GPTConfig(vocab_size=512, n_positions=S, n_embd=128, n_layer=LAYERS, n_head=4, dtype="bfloat16" if torch.cuda.is_available() else "float32")
rows = sess.batch_slice(GLOBAL_BATCH)
n_seq = rows.stop - rows.start
with ht.graph("define_and_run", create_new=True) as g:
    model = GPTLMHeadModel(cfg, [sess.local_cfg])
    in_ds, in_dg = ht.nn.parallel.config2ds(sess.local_cfg["input"])
    lb_ds, lb_dg = ht.nn.parallel.config2ds(sess.local_cfg["label"])
    ids = ht.parallel_placeholder("int64", [n_seq * S], [in_ds], device_group_hierarchy=[in_dg], name="ids")
    pos = ht.parallel_placeholder("int64", [n_seq * S], [in_ds], device_group_hierarchy=[in_dg], name="pos")
    lab = ht.parallel_placeholder("int64", [n_seq * S], [lb_ds], device_group_hierarchy=[lb_dg], name="lab")
    loss = model(ids, pos, lab, seq_len=S)
    train_op = ht.AdamOptimizer(lr=3e-3).minimize(loss)
sess.precreate_groups()
rng = np.random.RandomState(0)
P = np.tile(np.arange(S), (GLOBAL_BATCH, 1))
for step in range(10):
    X = rng.randint(0, 512, (GLOBAL_BATCH, S))
    X[:, 1::2] = (X[:, 0::2] + 1) % 512          # learnable structure
    L = np.roll(X, -1, axis=1)
    feed = {ids: [torch.as_tensor(X[rows].reshape(-1))], pos: [torch.as_tensor(P[rows].reshape(-1))], lab: [torch.as_tensor(L[rows].reshape(-1))]}
    out = g.run(loss, [loss, train_op], feed, num_micro_batches=1, grad_scale=sess.grad_scale(n_seq, GLOBAL_BATCH))
    if sess.rank in sess.last_stage_ranks:
        lv = sess.reduce_loss(out[0].float().mean(), n_seq, GLOBAL_BATCH)
        if sess.rank == sess.last_stage_ranks[0]:
            print(f"step {step}: loss {lv:.4f}  (pipeline shares {sess.split_batch(GLOBAL_BATCH)})", flush=True)
