"""Apply LoRA to a model: replace the targeted parallel linears, freeze everything else, save / merge the adapters.
(ref: python/hetu/peft/lora/model.py -- LoraModel / MultiLoraModel)"""
from __future__ import annotations

from typing import Dict

import torch

from ...nn.module import Module
from ...nn.parallel import HtMultiColumnParallelLinear, HtMultiRowParallelLinear
from .config import LoraConfig
from .layer import (LoraColumnParallelLinear, LoraRowParallelLinear, MultiLoraColumnParallelLinear, MultiLoraRowParallelLinear)


def _replace(model: Module, config: LoraConfig):
    multi = config.num_tasks > 1
    n = 0
    for _, mod in list(model.named_modules()):
        for cname, child in list(mod._modules.items()):
            if not any(cname == t or cname.endswith(t) for t in config.target_modules):
                continue
            if isinstance(child, HtMultiColumnParallelLinear):
                new = (MultiLoraColumnParallelLinear if multi else LoraColumnParallelLinear)(child, config)
            elif isinstance(child, HtMultiRowParallelLinear):
                new = (MultiLoraRowParallelLinear if multi else LoraRowParallelLinear)(child, config)
            else:
                continue
            mod._modules[cname] = new
            object.__setattr__(mod, cname, new)
            n += 1
    return n


class LoraModel(Module):
    def __init__(self, model: Module, config: LoraConfig):
        super().__init__()
        for _, p in model.named_parameters():
            p.requires_grad = False
        self.num_adapted = _replace(model, config)
        assert self.num_adapted > 0, f"no module matched target_modules={config.target_modules}"
        for name, p in model.named_parameters():
            if "lora_" in name:
                p.requires_grad = True
        self.model, self.peft_config = model, config
        self.config = getattr(model, "config", None)

    def forward(self, *a, **kw):
        return self.model(*a, **kw)

    def __getattr__(self, k):
        try:
            return super().__getattr__(k)
        except AttributeError:
            return getattr(self.__dict__["_modules"]["model"], k)


class MultiLoraModel(LoraModel):
    def set_task_mask(self, mask):
        """mask [tokens, num_tasks]: one-hot task membership of every token of the step"""
        for _, m in self.model.named_modules():
            if hasattr(m, "set_task_mask"):
                m.set_task_mask(mask)


def get_peft_model(model: Module, config: LoraConfig) -> LoraModel:
    return (MultiLoraModel if config.num_tasks > 1 else LoraModel)(model, config)


class _LoraFactory:
    def __init__(self, inner, config):
        self.inner, self.peft_config = inner, config
        self.model_config = getattr(inner, "model_config", None)

    def create_model(self, ds_parallel_configs):
        base = self.inner.create_model(ds_parallel_configs) if hasattr(self.inner, "create_model") else self.inner
        self.model_config = getattr(self.inner, "model_config", getattr(base, "config", None))
        return get_peft_model(base, self.peft_config)


def wrap_model_factory(model_or_wrapper, config: LoraConfig):
    return _LoraFactory(model_or_wrapper, config)


def lora_state_dict(model: Module) -> Dict[str, torch.Tensor]:
    return {k: v for k, v in model.state_dict().items() if "lora_" in k}


def merge_lora_weights(model: LoraModel) -> Dict[str, torch.Tensor]:
    """-> state dict of the base model with W <- W + (alpha / r) B A folded in (single-task adapters)"""
    sd = model.model.state_dict()
    scale = model.peft_config.scaling
    out = {}
    for k, v in sd.items():
        if "lora_" in k:
            continue
        if k.endswith("base.weight"):
            stem = k[: -len("base.weight")]
            a, b = sd.get(stem + "lora_A"), sd.get(stem + "lora_B")
            if a is not None and b is not None:
                v = v.float() + scale * (b.float() @ a.float())
            out[stem + "weight"] = v
        elif ".base." in k:
            out[k.replace(".base.", ".")] = v
        else:
            out[k] = v
    return out
