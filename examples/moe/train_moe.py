"""GPT-MoE: experts sharded over the data-parallel ranks; dispatch / combine all-to-all fused into the layout transform."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import hetu_b200 as ht
from hetu_b200.models import GPTMoELMHeadModel, MoEConfig, generate_ds_parallel_config

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
ht.init_comm_group(world)
S, B = 128, 4
cfg = MoEConfig(vocab_size=1024, n_positions=S, n_embd=256, n_layer=4, n_head=8, num_experts=max(8, world), top_k=2, capacity_factor=1.5,
                moe_every=2, ep_ranks=tuple(range(world)) if world > 1 else ())
use_cuda = torch.cuda.is_available() and not os.environ.get("HETU_B200_FORCE_CPU")
with ht.graph("define_and_run", create_new=True) as g, (ht.autocast("bfloat16") if use_cuda else ht.autocast(None)):
    dsc = [generate_ds_parallel_config(cfg.n_layer, world, world, 1, 1, zero=False)]
    model = GPTMoELMHeadModel(cfg, dsc)
    ic = ht.nn.parallel.config2ds(dsc[0]["input"])
    ids, pos, lab = (ht.parallel_placeholder("int64", [B * S * world], [ic[0]], device_group_hierarchy=[ic[1]], name=n) for n in ("ids", "pos", "lab"))
    loss = model(ids, pos, lab, seq_len=S)
    train_op = ht.AdamOptimizer(lr=1e-3).minimize(loss)
gen = torch.Generator().manual_seed(rank)
for step in range(10):
    x = torch.randint(0, cfg.vocab_size, (B * S,), generator=gen)
    out = g.run(loss, [loss, train_op], {ids: x, pos: torch.arange(S).repeat(B), lab: torch.roll(x, -1)}, grad_scale=1.0 / world)
    if rank == 0:
        print(f"step {step} loss {float(out[0].float().mean()):.4f}", flush=True)
