"""MalleusTrainer(auto_apply=True): 4 ranks start as dp2 x tp2; after 2 steps the (injected) straggler report says device 3 is
3x slower, the planner moves batch share away from its pipeline (6 : 2) and the job continues under the heterogeneous plan."""
import json
import os
import sys

import torch

import hetu_b200 as ht
from hetu_b200 import distributed
from hetu_b200.data import ByteTokenizer, SyntheticDataset
from hetu_b200.engine import MalleusTrainer, ModelWrapper, OptimizerWrapper, Trainer, TrainerCtxs, TrainerStrategyArgs, TrainingConfig
from hetu_b200.models import GPTConfig, GPTLMHeadModel, generate_ds_parallel_config

mode = sys.argv[1]                 # single | malleus (device 3 slow: batch shares 6 : 2) | shrink (device 1 dead slow: its tp group is dissolved)
world = 1 if mode == "single" else 4
ht.init_comm_group(world)
ht.set_seed(3)
mcfg = GPTConfig(vocab_size=260, n_positions=32, n_embd=32, n_layer=4, n_head=4)
ds = SyntheticDataset(64, 259, 32, seed=1, length_distribution="fixed")
cfg = TrainingConfig(packing=False, micro_batch_size=2, global_load_size=8, max_seq_length=32, steps=4, learning_rate=1e-2, log_interval=0,
                     pack_alignment=16, output_dir=os.environ.get("TRAINER_OUT", "/tmp/hb_trainer_malleus"))
common = (cfg, ModelWrapper(GPTLMHeadModel, mcfg), ByteTokenizer(), OptimizerWrapper({"type": "adam", "lr": 1e-2}), ds)
if mode == "single":
    losses = Trainer(*common, ds_parallel_configs=[generate_ds_parallel_config(4, 1, 1, 1, 1, zero=False)]).train(steps=4)
    t = torch.tensor([float(v) for v in losses], dtype=torch.float64)
else:
    ratios = {0: 1.0, 1: 1.0, 2: 1.0, 3: 3.0} if mode == "malleus" else {0: 1.0, 1: 50.0, 2: 1.0, 3: 1.0}
    tr = MalleusTrainer(*common, ds_parallel_configs=[generate_ds_parallel_config(4, 4, 2, 2, 1, zero=False)],
                        ctxs=TrainerCtxs(normal_layers=4, normal_mbn=4),
                        strategy_args=TrainerStrategyArgs(dp=2, tp=2, pp=1, rank_to_device_mapping={i: i for i in range(4)}),
                        replan_interval=2, ratio_source=lambda: ratios, auto_apply=True)
    losses = tr.train(steps=4)
    print("PLANS", distributed.rank(), [(r["step"], r.get("applied"), r["micro_batches"]) for r in tr.plans_log], tr.hetero is not None, tr.idle, flush=True)
    rec = tr.plans_log[0]
    if mode == "malleus":
        assert rec["applied"] == "hetero" and rec["micro_batches"] == [6, 2] and tr.hetero is not None and tr.hetero.split_batch(8) == [6, 2], rec
    else:
        # one tensor-parallel pipeline over the healthy devices {0, 2}; ranks 1 and 3 idle through the rest of the job
        assert rec["applied"] == "hetero" and tr.hetero.pipelines == [[[0, 2]]] and tr.idle == (distributed.rank() in (1, 3)), (rec, tr.hetero.pipelines)
    losses = list(losses) + [float("nan")] * (4 - len(losses))       # idle ranks report nothing
    t = torch.tensor([float(v) for v in losses], dtype=torch.float64)
    # dp2 phase: per-replica means on the loss ranks {0, 2}; hetero phase: global mean on the first last-stage leader
    head = ht._C.comm_all_reduce(t[:2].clone(), [0, 2], "sum") / 2 if distributed.rank() in (0, 2) else t[:2]
    t = torch.cat([head, t[2:]])
if distributed.rank() == 0:
    print("LOSSES " + json.dumps([float(v) for v in t]))
