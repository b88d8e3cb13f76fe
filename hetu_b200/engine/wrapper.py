"""Deferred construction of model / optimizer / dataset so that the trainer can (re)build them inside a graph and under
a new parallel strategy (ref: python/hetu/engine/wrapper.py)."""
from __future__ import annotations

from typing import Any, Dict


class ModelWrapper:
    def __init__(self, model_class, model_config):
        self.model_class, self.model_config = model_class, model_config

    def create_model(self, ds_parallel_configs):
        return self.model_class(self.model_config, ds_parallel_configs)


class ModelWrapperFromConfig:
    """{"type": "gpt" | "llama" | "moe", **config fields}"""

    def __init__(self, config: Dict[str, Any]):
        from .. import models
        self.config = dict(config)
        cfg = dict(self.config)
        self.kind = cfg.pop("type", "gpt").lower()
        table = {"gpt": (models.GPTConfig, models.GPTLMHeadModel), "llama": (models.LlamaConfig, models.LlamaLMHeadModel),
                 "moe": (models.MoEConfig, models.GPTMoELMHeadModel), "gpt_moe": (models.MoEConfig, models.GPTMoELMHeadModel)}
        if self.kind not in table:
            raise ValueError(f"unknown model type {self.kind}")
        cfg_cls, self.model_class = table[self.kind]
        # the config object exists before the model: the Trainer adjusts it (e.g. the context-parallel ring) prior to building
        self.model_config = cfg_cls(**cfg)

    def create_model(self, ds_parallel_configs):
        return self.model_class(self.model_config, ds_parallel_configs)


class OptimizerWrapper:
    def __init__(self, optimizer_config: Dict[str, Any]):
        self.config = dict(optimizer_config)

    def create_optimizer(self, **kwargs):
        from .. import optim
        cfg = {**self.config, **kwargs}
        kind = cfg.pop("type", "adam").lower()
        if kind in ("adam", "adamw"):
            return optim.AdamOptimizer(**cfg)
        if kind == "sgd":
            return optim.SGDOptimizer(**cfg)
        raise ValueError(f"unknown optimizer {kind}")


class DatasetWrapper:
    def __init__(self, dataset_class):
        self.dataset_class = dataset_class

    def create_dataset(self, **kwargs):
        return self.dataset_class(**kwargs)
