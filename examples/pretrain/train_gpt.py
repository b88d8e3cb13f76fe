"""Pre-train a GPT with the Trainer under a (dp, tp, pp, zero, sp) strategy on synthetic tokens."""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import hetu_b200 as ht
from hetu_b200.data import ByteTokenizer, SyntheticDataset
from hetu_b200.engine import ModelWrapper, OptimizerWrapper, Trainer, TrainingConfig
from hetu_b200.models import GPTConfig, GPTLMHeadModel, generate_ds_parallel_config

ap = argparse.ArgumentParser()
ap.add_argument("--dp", type=int, default=1); ap.add_argument("--tp", type=int, default=1); ap.add_argument("--pp", type=int, default=1)
ap.add_argument("--sp", action="store_true"); ap.add_argument("--no-zero", action="store_true")
ap.add_argument("--layers", type=int, default=4); ap.add_argument("--hidden", type=int, default=256); ap.add_argument("--heads", type=int, default=8)
ap.add_argument("--seq", type=int, default=256); ap.add_argument("--global-batch", type=int, default=16); ap.add_argument("--micro-batch", type=int, default=4)
ap.add_argument("--steps", type=int, default=20); ap.add_argument("--bf16", action="store_true"); ap.add_argument("--packing", action="store_true")
a = ap.parse_args()
world = a.dp * a.tp * a.pp
ht.init_comm_group(world)
cfg = GPTConfig(vocab_size=259, n_positions=a.seq, n_embd=a.hidden, n_layer=a.layers, n_head=a.heads, sequence_parallel=a.sp)
dsc = [generate_ds_parallel_config(a.layers, world, a.dp, a.tp, a.pp, zero=not a.no_zero)]
tc = TrainingConfig(packing=a.packing, micro_batch_size=None if a.packing else a.micro_batch, global_load_size=a.global_batch, max_seq_length=a.seq,
                    steps=a.steps, learning_rate=3e-4, bf16=a.bf16, log_interval=1)
data = SyntheticDataset(4096, 259, a.seq, min_seq_len=a.seq // 4, length_distribution="longtail")
trainer = Trainer(tc, ModelWrapper(GPTLMHeadModel, cfg), ByteTokenizer(), OptimizerWrapper({"type": "adam", "lr": 3e-4, "weight_decay": 0.01}),
                  data, ds_parallel_configs=dsc)
losses = trainer.train()
print("final loss", losses[-1] if losses else None)
