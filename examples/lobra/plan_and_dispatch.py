"""LobRA: plan a heterogeneous replica mix for several LoRA fine-tuning tasks, then dispatch one step's global batch.

    python examples/lobra/plan_and_dispatch.py --ngpus 16 --planner prune

(ref: examples/lobra/scripts/deploy_strategy_plan.py, llama_lora_multi_task.py)"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from hetu_b200.engine import lobra as L

ap = argparse.ArgumentParser()
ap.add_argument("--ngpus", type=int, default=16)
ap.add_argument("--layers", type=int, default=32)
ap.add_argument("--hidden", type=int, default=4096)
ap.add_argument("--ffn", type=int, default=11008)
ap.add_argument("--planner", choices=["group", "balance", "prune"], default="prune")
args = ap.parse_args()

cm = L.LoraCostModel.analytic(args.hidden, args.ffn)
# (tp, pp) schemes with the micro-batch capacity a memory profile would give them (activation memory ~ tokens / tp / pp)
cands = [{"tp": tp, "pp": pp, "max_tokens": 2048 * tp * pp, "throughput_per_gpu": 1.0 / (1.0 + 0.06 * (tp - 1) + 0.03 * (pp - 1))}
         for tp in (1, 2, 4, 8) for pp in (1, 2) if tp * pp <= args.ngpus]
# three tenants: chat (short), code (medium), long-document summarisation; bucket -> sequences per global batch
tasks = [{256: 400, 512: 160, 1024: 30}, {512: 80, 2048: 40, 4096: 10}, {2048: 20, 8192: 8, 16384: 2}]
planner = {"group": L.GroupStaticPlanner, "balance": L.BalanceStaticPlanner, "prune": L.PruneStaticPlanner}[args.planner](
    cm, args.layers, len(tasks), [sum(t.values()) for t in tasks], args.ngpus, cands)
plan = planner.schedule(tasks)
print(f"scheme pool: {[(s.tp, s.pp, s.max_tokens) for s in planner.schemes]}")
print(f"deployed (dp, tp, pp): {plan.strategy()}  on {plan.gpus} GPUs, {planner.evaluated} candidates evaluated")
for sc, d, disp, t in zip(plan.schemes, plan.dp, plan.dispatch, plan.scheme_times):
    if d:
        print(f"  {d} x (tp{sc.tp}, pp{sc.pp}, {sc.max_tokens} tok): {dict(sorted(disp.items()))}  est {t / 1e3:.2f} s")
print(f"estimated step time {plan.time / 1e3:.2f} s")

# one step: sample a global batch, dispatch it on the deployed mix, build the per-replica micro-batches
rng = np.random.RandomState(0)
buckets = sorted({b for t in tasks for b in t})
batches = []
for t in tasks:
    seqs = []
    for b, n in t.items():
        for _ in range(max(n // 8, 1)):
            seqs.append(list(rng.randint(1, 1000, rng.randint(b // 2 + 1, b + 1))))
    batches.append(seqs)
disp = L.BalanceDynamicDispatcher(cm, args.layers, plan.strategy(), [s.max_tokens for s, d in zip(plan.schemes, plan.dp) if d], len(tasks))
step = disp.schedule(L.seq_distribution(batches, buckets))
per = L.global_batch_scheduler(batches, step, buckets)
for j, sch in enumerate(per):
    for r, rows in enumerate(sch):
        pad = L.greedy_local_batch_scheduler(rows, step.schemes[j].max_tokens, len(tasks))
        pack = L.local_batch_pack_scheduler(rows, step.schemes[j].max_tokens, len(tasks))
        print(f"  scheme {j} replica {r}: {len(rows)} sequences -> {len(pad)} padded / {len(pack)} packed micro-batches")
print(f"step estimate {step.time / 1e3:.3f} s; heterogeneous pipelines for engine.hetero.HeteroSession: {len(plan.pipelines(args.layers))}")
