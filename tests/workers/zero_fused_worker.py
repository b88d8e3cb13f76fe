"""N-GPU worker: the symmetric-memory fused ZeRO path (wgrad GEMM -> peer slots -> fused reduce+AdamW+all-gather kernel)
must train exactly like the NCCL reduce-scatter / all-gather path of the same graph."""
import json
import os
import sys

import torch
import torch.distributed as dist

import hetu_b200 as ht
from hetu_b200.models import GPTConfig, GPTLMHeadModel, generate_ds_parallel_config

ht.init_comm_group()
rank, world = dist.get_rank(), dist.get_world_size()
dev = torch.device("cuda", torch.cuda.current_device())
os.environ["HETU_B200_STRICT"] = "1"
S, B = 128, 4
cfg = GPTConfig(vocab_size=1024, n_positions=S, n_embd=256 * world // 2 if world > 2 else 256, n_layer=2, n_head=max(4, (256 * world // 2 if world > 2 else 256) // 64))
T = B * S


def train(fused: bool, micro_batches: int = 1, steps: int = 4):
    os.environ["HETU_ZERO_FUSED"] = "1" if fused else "0"
    ht.set_seed(11)
    with ht.graph("define_and_run", create_new=True) as g, ht.autocast("bfloat16"):
        dsc = [generate_ds_parallel_config(cfg.n_layer, world, world, 1, 1, zero=True)]
        model = GPTLMHeadModel(cfg, dsc)
        ic = ht.nn.parallel.config2ds(dsc[0]["input"])
        ids = ht.parallel_placeholder("int64", [T * world], [ic[0]], device_group_hierarchy=[ic[1]], name="ids")
        pos = ht.parallel_placeholder("int64", [T * world], [ic[0]], device_group_hierarchy=[ic[1]], name="pos")
        lab = ht.parallel_placeholder("int64", [T * world], [ic[0]], device_group_hierarchy=[ic[1]], name="lab")
        loss = model(ids, pos, lab, seq_len=S)
        train_op = ht.AdamOptimizer(lr=1e-3, weight_decay=0.01).minimize(loss)
    gen = torch.Generator().manual_seed(5 + rank)
    losses = []
    n0 = ht._C.kernel_launch_count()
    for step in range(steps):
        x = torch.randint(0, cfg.vocab_size, (T,), generator=gen)
        feed = {ids: x.to(dev), pos: torch.arange(S).repeat(B).to(dev), lab: torch.roll(x, -1).to(dev)}
        if micro_batches > 1:
            feed = {k: list(v.chunk(micro_batches)) for k, v in feed.items()}
        out = g.run(loss, [loss, train_op], feed, num_micro_batches=micro_batches, grad_scale=1.0 / world)
        losses.append(float(out[0].float().mean()))
    sd = {k: v.float().cpu() for k, v in model.state_dict().items()}
    torch.cuda.synchronize()
    return losses, sd, ht._C.kernel_launch_count() - n0


ref_l, ref_sd, ref_launch = train(False)
fus_l, fus_sd, fus_launch = train(True)
err = max(float((ref_sd[k] - fus_sd[k]).abs().max()) for k in ref_sd)
# every rank must hold identical parameters after the fused all-gather-by-peer-stores
chk = torch.tensor([sum(float(v.double().sum()) for v in fus_sd.values())], device=dev, dtype=torch.float64)
lo, hi = chk.clone(), chk.clone()
dist.all_reduce(lo, op=dist.ReduceOp.MIN)
dist.all_reduce(hi, op=dist.ReduceOp.MAX)
acc_l, acc_sd, _ = train(True, micro_batches=2)      # accumulation path (flat in-kernel all-reduce + legacy for fused entries)
acc_ref_l, acc_ref_sd, _ = train(False, micro_batches=2)
err_acc = max(float((acc_ref_sd[k] - acc_sd[k]).abs().max()) for k in acc_sd)
if rank == 0:
    print("ZEROFUSED " + json.dumps({"ref_losses": ref_l, "fused_losses": fus_l, "max_param_diff": err, "ranks_identical": bool(lo.item() == hi.item()),
                                      "ref_launches": ref_launch, "fused_launches": fus_launch, "acc_losses": acc_l, "acc_ref_losses": acc_ref_l,
                                      "max_param_diff_accum": err_acc, "symm_launches": ht._C.symm_launch_count() if hasattr(ht._C, "symm_launch_count") else -1}))
dist.barrier()
dist.destroy_process_group()
