"""Malleus end to end: train a GPT under dp x tp x pp, let a device become a straggler, detect it, re-plan and move the RUNNING
job onto the new plan.

    # 4 processes (CPU / gloo or 4 GPUs); device 3 is made 3x slower after step 4
    python -m torch.distributed.run --nproc-per-node 4 --master-addr 127.0.0.1 examples/malleus/train_malleus.py --slow-rank 3 --slowdown 3.0
    # measure real slow-down ratios with the busy-loop workload instead of injecting them
    ... examples/malleus/train_malleus.py --measure

What happens (ref: examples/malleus/pretrain_gpt.py + helper.py, python/hetu/engine/{strategy,straggler}.py):
  1. `MalleusTrainer` trains under the homogeneous strategy (--dp/--tp/--pp);
  2. every --replan-interval steps it obtains per-device slow-down ratios -- `engine.Straggler.run_profile()` (a fixed GEMM
     workload timed on every rank) or the injected report of this script;
  3. `engine.StrategyModel` solves the Malleus plan: tensor-parallel groups are regrouped around slow devices, layers and
     micro-batches move away from slow pipelines (`plans_log` records it);
  4. with --auto-apply the trainer rebuilds onto the plan through a split checkpoint: a homogeneous plan becomes an ordinary
     strategy, a heterogeneous one (unequal batch shares, different tp per pipeline, idle ranks) runs through
     `engine.HeteroSession`;
  5. the loss curve continues (compare with --no-straggler: same data order, so curves agree up to reduction order).
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import hetu_b200 as ht
from hetu_b200 import distributed
from hetu_b200.data import ByteTokenizer, SyntheticDataset
from hetu_b200.engine import MalleusTrainer, ModelWrapper, OptimizerWrapper, TrainerCtxs, TrainerStrategyArgs, TrainingConfig
from hetu_b200.models import GPTConfig, GPTLMHeadModel, generate_ds_parallel_config

ap = argparse.ArgumentParser()
ap.add_argument("--dp", type=int, default=0, help="0: world / (tp * pp)")
ap.add_argument("--tp", type=int, default=2)
ap.add_argument("--pp", type=int, default=1)
ap.add_argument("--layers", type=int, default=4)
ap.add_argument("--hidden", type=int, default=64)
ap.add_argument("--heads", type=int, default=4)
ap.add_argument("--seq", type=int, default=32)
ap.add_argument("--global-batch", type=int, default=8)
ap.add_argument("--micro-batch", type=int, default=2)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--replan-interval", type=int, default=4)
ap.add_argument("--slow-rank", type=int, default=-1, help="rank reported as a straggler (injected report)")
ap.add_argument("--slowdown", type=float, default=3.0)
ap.add_argument("--measure", action="store_true", help="measure ratios with the Straggler workload instead of injecting them")
ap.add_argument("--no-auto-apply", action="store_true", help="only log the plans")
ap.add_argument("--out", default=os.environ.get("TRAINER_OUT", "/tmp/hetu_b200_malleus"))
a = ap.parse_args()

world = int(os.environ.get("WORLD_SIZE", "1"))
ht.init_comm_group(world)
ht.set_seed(3)
tp = min(a.tp, world)
pp = min(a.pp, max(world // tp, 1))
dp = a.dp or max(world // (tp * pp), 1)
assert dp * tp * pp == world, f"dp{dp} x tp{tp} x pp{pp} != world {world}"
mcfg = GPTConfig(vocab_size=260, n_positions=a.seq, n_embd=a.hidden, n_layer=a.layers, n_head=a.heads)
data = SyntheticDataset(max(64, a.global_batch * a.steps), 259, a.seq, seed=1, length_distribution="fixed")
cfg = TrainingConfig(packing=False, micro_batch_size=a.micro_batch, global_load_size=a.global_batch, max_seq_length=a.seq, steps=a.steps,
                     learning_rate=1e-2, log_interval=1, pack_alignment=16, output_dir=a.out)
ratios = {r: 1.0 for r in range(world)}
if 0 <= a.slow_rank < world:
    ratios[a.slow_rank] = a.slowdown
trainer = MalleusTrainer(cfg, ModelWrapper(GPTLMHeadModel, mcfg), ByteTokenizer(), OptimizerWrapper({"type": "adam", "lr": 1e-2}), data,
                         ds_parallel_configs=[generate_ds_parallel_config(a.layers, world, dp, tp, pp, zero=False)],
                         ctxs=TrainerCtxs(normal_layers=a.layers // pp, normal_mbn=max(a.global_batch // a.micro_batch // dp, 1)),
                         strategy_args=TrainerStrategyArgs(dp=dp, tp=tp, pp=pp, rank_to_device_mapping={i: i for i in range(world)}),
                         replan_interval=a.replan_interval, ratio_source=None if a.measure else (lambda: dict(ratios)),
                         auto_apply=not a.no_auto_apply and world > 1)
losses = trainer.train(steps=a.steps)
rank = distributed.rank()
for rec in trainer.plans_log:
    if rank == 0:
        print("PLAN " + json.dumps({"step": rec["step"], "ratios": {str(k): round(v, 2) for k, v in rec["ratios"].items()}, "layers per stage": rec["hetero_layers"],
                                    "micro-batches per pipeline": rec["micro_batches"], "homogeneous": rec["executable"], "applied": rec.get("applied"),
                                    "estimated step": round(rec["estimated_time"], 2)}), flush=True)
if losses and rank in (trainer._loss_ranks() if hasattr(trainer, "_loss_ranks") else [0]):
    print(f"rank {rank}: loss {losses[0]:.4f} -> {losses[-1]:.4f} over {len(losses)} steps; hetero path: {trainer.hetero is not None}; idle: {getattr(trainer, 'idle', False)}",
          flush=True)
