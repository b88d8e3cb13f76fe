"""YAML experiment configs (the reference merges hydra / OmegaConf YAML onto TrainingConfig and StrategyConfig:
examples/pretrain/config/*.yaml with sections rpc / ds_parallel / trainer / model{tokenizer, optimizer}).  The same
layout is read here with plain PyYAML; `key=value` overrides mimic the hydra command line."""
from __future__ import annotations

import copy
import dataclasses
from typing import Any, Dict, Sequence

import yaml

from ..utils.parallel import RecomputeConfig, StrategyConfig
from .trainer_config import DataLoadLevel, SFTConfig, TrainingConfig


def _set(d: dict, dotted: str, value: Any):
    keys = dotted.split(".")
    for k in keys[:-1]:
        d = d.setdefault(k, {})
    d[keys[-1]] = value


def _from_dict(cls, d: dict):
    fields = {f.name for f in dataclasses.fields(cls)}
    return cls(**{k: v for k, v in d.items() if k in fields})


def load_experiment(path: str, overrides: Sequence[str] = ()) -> Dict[str, Any]:
    """-> {"trainer": TrainingConfig | SFTConfig, "strategy": StrategyConfig, "model": dict, "optimizer": dict, "tokenizer": dict,
    "rpc": dict, "raw": dict}"""
    # hydra-style loading: `defaults` composition with config groups next to the file, dotted overrides (+add, ~delete,
    # group=option), ${...} interpolation (utils/hydra_lite.py)
    import os
    from ..utils import hydra_lite
    cfg_dir, cfg_name = os.path.split(os.path.abspath(path))
    choices, values = {}, []
    for ov in overrides:
        k, eq, v = ov.partition("=")
        if eq and not ov.startswith(("+", "~")) and "." not in k and os.path.isdir(os.path.join(cfg_dir, k)):
            choices[k] = v
        else:
            values.append(ov)
    raw = hydra_lite.compose(cfg_dir, cfg_name, choices)
    raw = hydra_lite.resolve(hydra_lite.apply_overrides(copy.deepcopy(raw), values))
    ds = dict(raw.get("ds_parallel", {}))
    rc = ds.pop("recompute", None) or {}
    strat = _from_dict(StrategyConfig, {**ds, "recompute": _from_dict(RecomputeConfig, {
        "recompute_granularity": rc.get("granularity"), "recompute_method": rc.get("method"), "recompute_num_layers": rc.get("num_layers"),
        "recompute_layer_idxs_list": rc.get("layer_idxs") or []})})
    tr = dict(raw.get("trainer", {}))
    if "data_load_level" in tr:
        tr["data_load_level"] = DataLoadLevel(str(tr["data_load_level"]).upper())
    cls = SFTConfig if raw.get("sft") or tr.get("lora_rank") else TrainingConfig
    trainer = _from_dict(cls, {**tr, **(raw.get("sft") or {})})
    trainer.ds_parallel = strat
    model = dict(raw.get("model", {}))
    return {"trainer": trainer, "strategy": strat, "model": {k: v for k, v in model.items() if k not in ("tokenizer", "optimizer")},
            "optimizer": dict(model.get("optimizer", raw.get("optimizer", {"type": "adam"}))),
            "tokenizer": dict(model.get("tokenizer", raw.get("tokenizer", {"type": "byte"}))), "rpc": dict(raw.get("rpc", {})), "raw": raw}


def build_trainer(path: str, overrides: Sequence[str] = (), train_dataset=None):
    """one call from a YAML file to a ready `Trainer` (model / optimizer / tokenizer wrappers included)"""
    from .trainer import Trainer
    from .sft_trainer import SFTTrainer
    from .wrapper import ModelWrapperFromConfig, OptimizerWrapper
    exp = load_experiment(path, overrides)
    cls = SFTTrainer if isinstance(exp["trainer"], SFTConfig) else Trainer
    return cls(exp["trainer"], ModelWrapperFromConfig(exp["model"]), exp["tokenizer"], OptimizerWrapper(exp["optimizer"]), train_dataset)
