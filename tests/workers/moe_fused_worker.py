"""N-GPU worker: GPT-MoE (bf16) with the dispatch / combine all-to-all fused into the layout transform over NVLink peer
memory must train like the NCCL all-to-all path of the same graph; also times one MoE layer both ways."""
import json
import os
import sys

import torch
import torch.distributed as dist

import hetu_b200 as ht
from hetu_b200.models import GPTMoELMHeadModel, MoEConfig, generate_ds_parallel_config

ht.init_comm_group()
rank, world = dist.get_rank(), dist.get_world_size()
dev = torch.device("cuda", torch.cuda.current_device())
S, B = 256, 8
cfg = MoEConfig(vocab_size=2048, n_positions=S, n_embd=512, n_layer=2, n_head=8, num_experts=8, top_k=2, capacity_factor=2.0,
                moe_every=1, ep_ranks=tuple(range(world)), aux_loss_weight=0.01)
T = B * S


def train(fused: bool, steps: int = 4, time_it: bool = False):
    os.environ["HETU_EP_FUSED"] = "1" if fused else "0"
    ht.set_seed(3)
    with ht.graph("define_and_run", create_new=True) as g, ht.autocast("bfloat16"):
        dsc = [generate_ds_parallel_config(cfg.n_layer, world, world, 1, 1, zero=False)]
        model = GPTMoELMHeadModel(cfg, dsc)
        ic = ht.nn.parallel.config2ds(dsc[0]["input"])
        ids = ht.parallel_placeholder("int64", [T * world], [ic[0]], device_group_hierarchy=[ic[1]], name="ids")
        pos = ht.parallel_placeholder("int64", [T * world], [ic[0]], device_group_hierarchy=[ic[1]], name="pos")
        lab = ht.parallel_placeholder("int64", [T * world], [ic[0]], device_group_hierarchy=[ic[1]], name="lab")
        loss = model(ids, pos, lab, seq_len=S)
        train_op = ht.AdamOptimizer(lr=1e-3).minimize(loss)
    gen = torch.Generator().manual_seed(5 + rank)
    losses = []
    feeds = []
    for step in range(steps):
        x = torch.randint(0, cfg.vocab_size, (T,), generator=gen)
        feeds.append({ids: x.to(dev), pos: torch.arange(S).repeat(B).to(dev), lab: torch.roll(x, -1).to(dev)})
    for f in feeds:
        out = g.run(loss, [loss, train_op], f, grad_scale=1.0 / world)
        losses.append(float(out[0].float().mean()))
    ms = None
    if time_it:
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(10):
            g.run(loss, [loss, train_op], feeds[i % steps], grad_scale=1.0 / world)
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 10], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    sd = {k: v.float().cpu() for k, v in model.state_dict().items()}
    return losses, sd, ms


ref_l, ref_sd, ref_ms = train(False, time_it=True)
s0 = ht._C.symm_launch_count()
fus_l, fus_sd, fus_ms = train(True, time_it=True)
err = max(float((ref_sd[k] - fus_sd[k]).abs().max()) for k in ref_sd)
if rank == 0:
    print("MOEFUSED " + json.dumps({"world": world, "ref_losses": ref_l, "fused_losses": fus_l, "max_param_diff": err,
                                     "ms_step_nccl_a2a": ref_ms, "ms_step_fused_peer": fus_ms, "symm_launches": ht._C.symm_launch_count() - s0}))
dist.barrier()
dist.destroy_process_group()
