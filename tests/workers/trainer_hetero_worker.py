"""Trainer on a heterogeneous strategy: argv[1] = 'single' (one device) or 'hetero' (5 ranks: pipeline 0 = tp2 x pp2 with
3/4 of every batch, pipeline 1 = one device with 1/4)"""
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
world = 5 if mode == "hetero" else 1
ht.init_comm_group(world)
ht.set_seed(3)
mcfg = GPTConfig(vocab_size=260, n_positions=32, n_embd=32, n_layer=4, n_head=4)
ds = SyntheticDataset(64, 259, 32, seed=1, length_distribution="fixed")
PACK = os.environ.get("TRAINER_PACK") == "1"
cfg = TrainingConfig(packing=PACK, micro_batch_size=None if PACK else 2, global_load_size=8, max_seq_length=32, steps=4, learning_rate=1e-2, log_interval=0,
                     pack_alignment=16, output_dir=os.environ.get("TRAINER_OUT", "/tmp/hb_trainer_hetero"), save_interval=int(os.environ.get("TRAINER_SAVE", "0")))
if mode == "hetero":
    pipelines = [{"stages": [{"devices": [0, 1], "layers": [0, 1]}, {"devices": [2, 3], "layers": [2, 3]}]}, {"stages": [{"devices": [4], "layers": [0, 3]}]}]
    dsc, kw = [generate_hetero_ds_parallel_config(4, pipelines, zero=False)], {"hetero_shares": [3, 1]}
else:
    dsc, kw = [generate_ds_parallel_config(4, 1, 1, 1, 1, zero=False)], {}
tr = Trainer(cfg, ModelWrapper(GPTLMHeadModel, mcfg), ByteTokenizer(), OptimizerWrapper({"type": "adam", "lr": 1e-2}), ds, ds_parallel_configs=dsc, **kw)
losses = tr.train()
if distributed.rank() in tr._loss_ranks():
    print("LOSSES " + json.dumps([float(v) for v in losses]))
print("INFO", distributed.rank(), tr.hetero is not None, (tr.hetero.pipeline, tr.hetero.split_batch(8)) if tr.hetero else None, len(tr.loss_history), flush=True)
