"""(ref: python/hetu/peft/lora/config.py)"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List


@dataclass
class LoraConfig:
    rank: int = 8
    lora_alpha: float = 16.0
    lora_dropout: float = 0.0
    target_modules: List[str] = field(default_factory=lambda: ["qkv_dense", "dense", "dense_h_to_4h", "dense_4h_to_h"])
    num_tasks: int = 1                 # > 1: multi-task LoRA, one adapter pair per task, tokens routed by task id
    init_std: float = 0.01

    @property
    def scaling(self) -> float:
        return self.lora_alpha / max(self.rank, 1)
