"""YAML-driven pre-training entry point (the reference's examples/pretrain/train_hetu.py with hydra configs):

    python examples/pretrain/train_hetu.py --config-path examples/pretrain/config --config-name llama_pad_cp trainer.steps=5
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 examples/pretrain/train_hetu.py \\
        --config-path examples/pretrain/config --config-name llama_pad_cp

configs: gpt_small_dp2_tp2 (padding, dp x tp + SP), llama_pack_tp (packing, tp + ZeRO), llama_pad_cp (context parallel),
gpt_hetero (heterogeneous pipelines).  Data is synthetic (no network)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import hetu_b200 as ht
from hetu_b200.data import SyntheticDataset
from hetu_b200.engine import build_trainer
from hetu_b200.utils.parallel import distributed_init

ap = argparse.ArgumentParser()
ap.add_argument("--config-path", default=os.path.join(os.path.dirname(__file__), "config"))
ap.add_argument("--config-name", default="gpt_small_dp2_tp2")
ap.add_argument("overrides", nargs="*")
a = ap.parse_args()
path = os.path.join(a.config_path, a.config_name + ("" if a.config_name.endswith((".yaml", ".yml")) else ".yaml"))
distributed_init()
trainer = build_trainer(path, a.overrides)
cfg = trainer.pretrain_config
vocab = int(getattr(trainer.model_wrapper.model_config, "vocab_size", 259))
trainer.train_dataset = SyntheticDataset(4096, min(vocab, 259), int(cfg.max_seq_length or 1024), min_seq_len=max(int(cfg.max_seq_length or 1024) // 4, 8),
                                         length_distribution="fixed" if not cfg.packing else "longtail")
losses = trainer.train()
if ht.distributed.rank() in trainer._loss_ranks():
    print(f"steps {len(losses)}  loss {losses[0]:.4f} -> {losses[-1]:.4f}  ({sum(trainer.step_times[1:]) / max(len(trainer.step_times) - 1, 1) * 1e3:.1f} ms/step)")
