"""Throughput of the other BASELINE.json configurations (device-timed, max over ranks), one JSON line each:
    llama2-7b   Llama-2 7B, bf16, dp x tp2 x pp2 (+ sequence parallel, ZeRO, 1F1B micro-batches)
    llama3-8b   Llama-3 8B, block-scaled fp8 projections, TP = 8 + sequence parallel
    gpt-moe     GPT-MoE 350M x 8 experts, expert parallel over the data-parallel ranks
usage (torchrun): bench_configs.py <config> [--steps K] [--warmup W] [--layers L]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import hetu_b200 as ht
from hetu_b200.models import (GPTMoELMHeadModel, LlamaConfig, LlamaLMHeadModel, MoEConfig, generate_ds_parallel_config)

ap = argparse.ArgumentParser()
ap.add_argument("config", choices=["llama2-7b", "llama3-8b", "gpt-moe"])
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--layers", type=int, default=0, help="override the layer count (0 = full model)")
ap.add_argument("--seq", type=int, default=2048)
ap.add_argument("--batch", type=int, default=4, help="sequences per data-parallel replica and step")
ap.add_argument("--micro-batches", type=int, default=0)
ap.add_argument("--no-fp8", action="store_true")
ap.add_argument("--tp", type=int, default=0, help="override the tensor-parallel degree of the config")
ap.add_argument("--pp", type=int, default=0, help="override the pipeline-parallel degree of the config")
a = ap.parse_args()
ht.init_comm_group()
world, rank = dist.get_world_size(), dist.get_rank()
dev = torch.device("cuda", torch.cuda.current_device())
os.environ.setdefault("HETU_B200_STRICT", "1")
S, B = a.seq, a.batch
if a.config == "llama2-7b":
    tp, pp = a.tp or 2, a.pp or 2
    dp = world // (tp * pp)
    cfg = LlamaConfig.llama2_7b(sequence_parallel=True)
    mbs = a.micro_batches or 4
    cls, name = LlamaLMHeadModel, "Llama-2 7B"
elif a.config == "llama3-8b":
    tp, pp, dp = world, 1, 1
    cfg = LlamaConfig.llama3_8b(sequence_parallel=True, fp8=not a.no_fp8)
    mbs = a.micro_batches or 1
    cls, name = LlamaLMHeadModel, "Llama-3 8B" + ("" if a.no_fp8 else " fp8")
else:
    tp, pp, dp = 1, 1, world
    S = min(S, 1024)
    cfg = MoEConfig.gpt_moe_350m_8e(ep_ranks=tuple(range(world)), n_positions=S, capacity_factor=1.25)
    mbs = a.micro_batches or 1
    cls, name = GPTMoELMHeadModel, "GPT-MoE 350M x 8E"
if a.layers:
    if hasattr(cfg, "num_hidden_layers"):
        cfg.num_hidden_layers = a.layers
    else:
        cfg.n_layer = a.layers
L = getattr(cfg, "num_hidden_layers", getattr(cfg, "n_layer", 0))
cfg.max_position_embeddings = S if hasattr(cfg, "max_position_embeddings") else None
T = B * S
with ht.graph("define_and_run", create_new=True) as g, ht.autocast("bfloat16"):
    dsc = [generate_ds_parallel_config(L, world, dp, tp, pp, zero=(a.config != "gpt-moe"))]
    model = cls(cfg, dsc)
    ic = ht.nn.parallel.config2ds(dsc[0]["input"])
    lc = ht.nn.parallel.config2ds(dsc[0]["label"])
    per_mb = T // mbs
    ids = ht.parallel_placeholder("int64", [per_mb * dp], [ic[0]], device_group_hierarchy=[ic[1]], name="ids")
    pos = ht.parallel_placeholder("int64", [per_mb * dp], [ic[0]], device_group_hierarchy=[ic[1]], name="pos")
    lab = ht.parallel_placeholder("int64", [per_mb * dp], [lc[0]], device_group_hierarchy=[lc[1]], name="lab")
    loss = model(ids, pos, lab, seq_len=S)
    train_op = ht.AdamOptimizer(lr=1e-4, weight_decay=0.1).minimize(loss)
vocab = cfg.vocab_size
gen = torch.Generator().manual_seed(7 + rank // max(tp, 1))
x = torch.randint(0, vocab, (T,), generator=gen).to(dev)
p = torch.arange(S).repeat(B).to(dev)
y = torch.roll(x, -1)
feed = {ids: list(x.chunk(mbs)), pos: list(p.chunk(mbs)), lab: list(y.chunk(mbs))} if mbs > 1 else {ids: x, pos: p, lab: y}


def step():
    return g.run(loss, [loss, train_op], feed, num_micro_batches=mbs, grad_scale=1.0 / dp)[0]


for _ in range(max(a.warmup, 3)):
    last = step()
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    last = step()
e1.record(); torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
lv = torch.tensor([float(last.float().mean()) if last is not None else 0.0, 1.0 if last is not None else 0.0], device=dev, dtype=torch.float64)
dist.all_reduce(lv)
ms = float(t) / a.steps
tokens = T * dp
if rank == 0:
    print("CONFIG " + json.dumps({"config": name, "n_gpus": world, "parallelism": f"dp{dp} tp{tp} pp{pp}" + (" ep%d" % world if a.config == "gpt-moe" else ""),
                                  "layers": L, "seq_len": S, "global_batch": B * dp, "micro_batches": mbs, "ms_per_step": ms,
                                  "tokens_per_s": tokens / ms * 1e3, "loss": float(lv[0] / max(float(lv[1]), 1.0)), "dtype": "bf16" if a.no_fp8 or a.config != "llama3-8b" else "fp8+bf16",
                                  "gpu_launches_per_step": None, "mem_gb": torch.cuda.max_memory_allocated() / 2**30}), flush=True)
dist.barrier()
dist.destroy_process_group()
