"""HotSPa: one model, several parallel strategies, switched on the fly by the step's sequence-length bucket.  Long
sequences run under a tensor-parallel strategy, short ones under pure data parallelism; parameters and optimizer states are
re-sharded between the layouts by the executor's hot switch (no checkpoint round trip)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import hetu_b200 as ht
from hetu_b200.models import GPTConfig, GPTLMHeadModel, generate_ds_parallel_config

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
ht.init_comm_group(world)
tp = 2 if world % 2 == 0 else 1
strategies = [generate_ds_parallel_config(2, world, world, 1, 1, zero=False),            # 0: short sequences, data parallel
              generate_ds_parallel_config(2, world, world // tp, tp, 1, zero=False)]      # 1: long sequences, tensor parallel
cfg = GPTConfig(vocab_size=512, n_positions=256, n_embd=64, n_layer=2, n_head=4)
seq = ht.IntSymbol(64)
with ht.graph("define_and_run", create_new=True, num_strategy=2) as g:
    model = GPTLMHeadModel(cfg, strategies)
    in_h = [ht.nn.parallel.config2ds(s["input"]) for s in strategies]
    ids = ht.parallel_placeholder("int64", [world * 64], [h[0] for h in in_h], device_group_hierarchy=[h[1] for h in in_h], name="ids")
    pos = ht.parallel_placeholder("int64", [world * 64], [h[0] for h in in_h], device_group_hierarchy=[h[1] for h in in_h], name="pos")
    lab = ht.parallel_placeholder("int64", [world * 64], [h[0] for h in in_h], device_group_hierarchy=[h[1] for h in in_h], name="lab")
    loss = model(ids, pos, lab, seq_len=seq)
    train_op = ht.AdamOptimizer(lr=1e-3).minimize(loss)
rng = np.random.RandomState(rank)
for step in range(8):
    long_step = step % 4 == 3
    sid, S = (1, 256) if long_step else (0, 64)
    dp = world // tp if sid == 1 else world
    B = 2
    x = torch.as_tensor(rng.randint(0, 512, (B * S,)))
    seq.set_data(S)
    out = g.run(loss, [loss, train_op], {ids: x, pos: torch.arange(S).repeat(B), lab: torch.roll(x, -1)}, cur_strategy_id=sid, grad_scale=1.0 / dp)
    if rank == 0 and out[0] is not None:
        print(f"step {step} strategy {sid} seq {S} loss {float(out[0].float().mean()):.4f}", flush=True)
