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
        self.config = dict(config)

    def create_model(self, ds_parallel_configs):
        from .. import models
        cfg = dict(self.config)
        kind = cfg.pop("type", "gpt").lower()
        if kind == "gpt":
            self.model_config = models.GPTConfig(**cfg)
            return models.GPTLMHeadModel(self.model_config, ds_parallel_configs)
        if kind == "llama":
            self.model_config = models.LlamaConfig(**cfg)
            return models.LlamaLMHeadModel(self.model_config, ds_parallel_configs)
        if kind in ("moe", "gpt_moe"):
            self.model_config = models.MoEConfig(**cfg)
            return models.MoELMHeadModel(self.model_config, ds_parallel_configs)
        raise ValueError(f"unknown model type {kind}")


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
