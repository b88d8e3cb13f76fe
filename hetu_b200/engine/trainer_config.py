"""Training configuration dataclasses (ref: python/hetu/engine/trainer_config.py, sft_config.py)."""
from __future__ import annotations

import enum
from dataclasses import dataclass
from typing import Optional

from ..utils.parallel import StrategyConfig


class DataLoadLevel(str, enum.Enum):
    SAMPLE = "SAMPLE"
    TOKEN = "TOKEN"


@dataclass
class TrainingConfig:
    output_dir: str = "./output"
    overwrite_output_dir: bool = False
    plot_loss: bool = False
    plot_update_freq: int = 10
    bf16: bool = False
    packing: bool = True                       # pack variable-length samples into rows (varlen attention)
    micro_batch_size: Optional[int] = None     # padding mode only
    dynamic_micro_batch_padding: bool = False  # padding mode: sort by length and pad every micro-batch to its OWN longest row
    ds_parallel: Optional[StrategyConfig] = None
    global_load_size: int = 64                 # samples (SAMPLE level) or tokens (TOKEN level) per step
    data_load_level: DataLoadLevel = DataLoadLevel.SAMPLE
    torch_profile: bool = False
    start_profile_step: int = 1
    end_profile_step: int = 5
    profile_save_path: str = "./trace"
    train_dataset_path: Optional[str] = None
    dataset_text_field: Optional[str] = None
    max_seq_length: Optional[int] = None
    steps: int = 1000
    learning_rate: float = 1e-3
    weight_decay: float = 0.0
    warmup_steps: int = 0
    lr_decay_style: str = "constant"
    min_lr: float = 0.0
    clip_grad: float = 0.0
    save_interval: int = 0
    eval_interval: int = 0                # evaluate on Trainer.eval_dataset every N steps (0 = never)
    eval_iters: int = 0                   # batches per evaluation (0 = one pass over the evaluation set)
    log_interval: int = 1
    seed: int = 1234
    pack_alignment: int = 128


@dataclass
class SFTConfig(TrainingConfig):
    """supervised fine-tuning: chat/prompt template + loss only on the response tokens"""
    prompt_template: Optional[str] = None
    train_on_prompt: bool = False
    dataset_format: str = "messages"           # "messages" | "alpaca" | "text"
    lora_rank: int = 0
    lora_alpha: float = 16.0
    lora_dropout: float = 0.0
    lora_target_modules: tuple = ("qkv_dense", "dense", "dense_h_to_4h", "dense_4h_to_h")
