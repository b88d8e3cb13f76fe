"""N-GPU worker: tensor-parallel GPT/Llama training with the fused GEMM->reduce-scatter path (symmetric memory) must match
the NCCL reduce-scatter path of the same graph.  argv: tp [model]"""
import json
import os
import sys

import torch
import torch.distributed as dist

import hetu_b200 as ht
from hetu_b200.models import GPTConfig, GPTLMHeadModel, LlamaConfig, LlamaLMHeadModel, generate_ds_parallel_config

ht.init_comm_group()
rank, world = dist.get_rank(), dist.get_world_size()
tp = int(sys.argv[1]) if len(sys.argv) > 1 else world
kind = sys.argv[2] if len(sys.argv) > 2 else "gpt"
dp = world // tp
dev = torch.device("cuda", torch.cuda.current_device())
os.environ["HETU_B200_STRICT"] = "1"
S, B = 256, 4
if kind == "llama":
    cfg = LlamaConfig(vocab_size=2048, hidden_size=512, intermediate_size=1024, num_hidden_layers=2, num_attention_heads=8,
                      num_key_value_heads=4, sequence_parallel=True)
    n_layer = cfg.num_hidden_layers
else:
    cfg = GPTConfig(vocab_size=2048, n_positions=S, n_embd=512, n_layer=2, n_head=8, sequence_parallel=True)
    n_layer = cfg.n_layer
T = B * S


def train(fused: bool, steps: int = 4):
    os.environ["HETU_TP_FUSED"] = "1" if fused else "0"
    os.environ["HETU_ZERO_FUSED"] = "1" if fused else "0"
    ht.set_seed(3)
    with ht.graph("define_and_run", create_new=True) as g, ht.autocast("bfloat16"):
        dsc = [generate_ds_parallel_config(n_layer, world, dp, tp, 1, zero=True)]
        model = (LlamaLMHeadModel if kind == "llama" else GPTLMHeadModel)(cfg, dsc)
        ic = ht.nn.parallel.config2ds(dsc[0]["input"])
        ids = ht.parallel_placeholder("int64", [T * dp], [ic[0]], device_group_hierarchy=[ic[1]], name="ids")
        pos = ht.parallel_placeholder("int64", [T * dp], [ic[0]], device_group_hierarchy=[ic[1]], name="pos")
        lab = ht.parallel_placeholder("int64", [T * dp], [ic[0]], device_group_hierarchy=[ic[1]], name="lab")
        loss = model(ids, pos, lab, seq_len=S)
        train_op = ht.AdamOptimizer(lr=1e-3).minimize(loss)
    gen = torch.Generator().manual_seed(5 + rank // tp)
    losses = []
    s0 = ht._C.symm_launch_count()
    for step in range(steps):
        x = torch.randint(0, cfg.vocab_size, (T,), generator=gen)
        feed = {ids: x.to(dev), pos: torch.arange(S).repeat(B).to(dev), lab: torch.roll(x, -1).to(dev)}
        out = g.run(loss, [loss, train_op], feed, grad_scale=1.0 / dp)
        losses.append(float(out[0].float().mean()) if out[0] is not None else None)
    sd = {k: v.float().cpu() for k, v in model.state_dict().items()}
    torch.cuda.synchronize()
    return losses, sd, ht._C.symm_launch_count() - s0


ref_l, ref_sd, _ = train(False)
fus_l, fus_sd, symm = train(True)
err = max(float((ref_sd[k] - fus_sd[k]).abs().max()) for k in ref_sd)
if rank == 0:
    print("TPFUSED " + json.dumps({"model": kind, "tp": tp, "dp": dp, "ref_losses": ref_l, "fused_losses": fus_l, "max_param_diff": err,
                                    "symm_kernel_launches_fused_run": symm}))
dist.barrier()
dist.destroy_process_group()
