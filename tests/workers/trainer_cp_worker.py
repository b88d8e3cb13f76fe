"""Trainer with a context-parallel strategy: `Trainer` slices every padded row into this ring member's SYM chunks, sets the
model's ring and scales gradients by 1 / (dp * cp).  Prints the loss history; argv: dp cp tp"""
import json
import os
import sys

import hetu_b200 as ht
from hetu_b200 import distributed
from hetu_b200.data import ByteTokenizer, SyntheticDataset
from hetu_b200.engine import ModelWrapper, OptimizerWrapper, Trainer, TrainingConfig
from hetu_b200.models import LlamaConfig, LlamaLMHeadModel, generate_ds_parallel_config

dp, cp, tp = (int(v) for v in sys.argv[1:4])
packing = len(sys.argv) > 4 and sys.argv[4] == "pack"
world = dp * cp * tp
ht.init_comm_group(world)
ht.set_seed(3)
mcfg = LlamaConfig(vocab_size=260, hidden_size=32, intermediate_size=64, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2)
ds = SyntheticDataset(64, 259, 32, seed=1, length_distribution="fixed") if not packing else \
    SyntheticDataset(64, 259, 8, seed=1, length_distribution="fixed")      # 8-token documents cross the 12-token chunk borders; rows pack without padding, so every
                                                                            # ring member sees the same number of valid labels
cfg = TrainingConfig(packing=packing, micro_batch_size=None if packing else 2, global_load_size=12 if packing else 8, max_seq_length=48 if packing else 32, steps=4, learning_rate=1e-2, log_interval=0,
                     pack_alignment=4 if packing else 16, output_dir=os.environ.get("TRAINER_OUT", "/tmp/hb_trainer_cp"))
tr = Trainer(cfg, ModelWrapper(LlamaLMHeadModel, mcfg), ByteTokenizer(), OptimizerWrapper({"type": "adam", "lr": 1e-2}), ds,
             ds_parallel_configs=[generate_ds_parallel_config(2, world, dp, tp, 1, cp=cp, zero=False)])
losses = tr.train()
# every rank of the last stage holds the mean over its own tokens: average over the dp x cp replicas
import torch
t = torch.tensor(losses, dtype=torch.float64)
if world > 1:
    t = ht._C.comm_all_reduce(t, list(range(world)), "sum") / world
if distributed.rank() == 0:
    print("LOSSES " + json.dumps([float(v) for v in t]))
