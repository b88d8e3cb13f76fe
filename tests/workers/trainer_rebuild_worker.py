"""Trainer.rebuild: 3 ranks train 2 steps data-parallel (dp3), re-plan to a heterogeneous strategy (tp2 pipeline + 1-device
pipeline, shares 2:1) through a split checkpoint, train 2 more steps.  'single' = the 4 steps on one device."""
import json
import os
import sys

import hetu_b200 as ht
from hetu_b200 import distributed
from hetu_b200.data import ByteTokenizer, SyntheticDataset
from hetu_b200.engine import ModelWrapper, OptimizerWrapper, Trainer, TrainingConfig
from hetu_b200.models import GPTConfig, GPTLMHeadModel, generate_ds_parallel_config
from hetu_b200.models.parallel_config import generate_hetero_ds_parallel_config

mode = sys.argv[1]
world = 3 if mode == "rebuild" else 1
ht.init_comm_group(world)
ht.set_seed(3)
mcfg = GPTConfig(vocab_size=260, n_positions=32, n_embd=32, n_layer=2, n_head=4)
ds = SyntheticDataset(64, 259, 32, seed=1, length_distribution="fixed")
cfg = TrainingConfig(packing=False, micro_batch_size=2, global_load_size=6, max_seq_length=32, steps=4, learning_rate=1e-2, log_interval=0,
                     pack_alignment=16, output_dir=os.environ.get("TRAINER_OUT", "/tmp/hb_trainer_rebuild"))
first = [generate_ds_parallel_config(2, world, world, 1, 1, zero=False)]
tr = Trainer(cfg, ModelWrapper(GPTLMHeadModel, mcfg), ByteTokenizer(), OptimizerWrapper({"type": "adam", "lr": 1e-2}), ds, ds_parallel_configs=first)
if mode == "single":
    losses = tr.train(steps=5)
else:
    tr.train(steps=2)
    pipelines = [{"stages": [{"devices": [0, 1], "layers": [0, 1]}]}, {"stages": [{"devices": [2], "layers": [0, 1]}]}]
    tr.rebuild([generate_hetero_ds_parallel_config(2, pipelines, zero=False)], hetero_shares=[2, 1])
    assert tr.hetero is not None and tr.global_step == 2
    tr.train(steps=2)
    # ... and back: the checkpoint written by the heterogeneous pipelines (each saves its own shards) re-shards onto plain dp3
    tr.rebuild(first)
    assert tr.hetero is None and tr.global_step == 4
    losses = tr.train(steps=1)
import torch
t = torch.tensor([float(v) for v in losses], dtype=torch.float64)
if mode == "rebuild":
    # dp3 phases report per-replica means: average them; the hetero phase already reports the global mean (on the loss ranks)
    t = torch.where(torch.isnan(t), torch.zeros_like(t), t) if len(t) == 5 else torch.cat([t, torch.zeros(5 - len(t), dtype=t.dtype)])
    head = ht._C.comm_all_reduce(t[:2].clone(), [0, 1, 2], "sum") / 3
    tail = ht._C.comm_all_reduce(t[4:].clone(), [0, 1, 2], "sum") / 3
    t = torch.cat([head, t[2:4], tail])
if distributed.rank() == 0:
    print("LOSSES " + json.dumps([float(v) for v in t]))
