"""HuggingFace-style persistence for configs and models: `config.save_pretrained(dir)` / `Config.from_pretrained(dir)` (config.json),
`model.save_pretrained(dir, max_shard_size=...)` (safetensors shards + `model.safetensors.index.json`) and
`Model.from_pretrained(dir, ds_parallel_configs=...)` which rebuilds the model in the current graph under ANY parallel strategy and
loads the weights (every rank slices its shard by the parameter's DistributedStates).
(ref: python/hetu/models/utils/{config_utils.py PreTrainedConfig, model_utils.py PreTrainedModel / load_state_dict /
get_state_dict_dtype / get_parameter_dtype, common_utils.py split_hetu_state_dict_into_shards, hub.py is_remote_url})"""
from __future__ import annotations

import dataclasses
import json
import os
import re
from typing import Dict, List, Optional, Tuple

import torch

CONFIG_NAME = "config.json"
WEIGHTS_INDEX_NAME = "model.safetensors.index.json"


def is_remote_url(path: str) -> bool:
    return bool(re.match(r"^(https?|s3|gs|hdfs)://", str(path)))


def get_state_dict_dtype(state: Dict[str, torch.Tensor]) -> torch.dtype:
    for v in state.values():
        if v.is_floating_point():
            return v.dtype
    return next(iter(state.values())).dtype


def get_parameter_dtype(module) -> Optional[str]:
    for _, p in module.named_parameters():
        return str(p.dtype)
    return None


def _nbytes(t: torch.Tensor) -> int:
    return t.numel() * t.element_size()


def parse_size(size) -> int:
    if isinstance(size, int):
        return size
    m = re.match(r"^\s*([0-9.]+)\s*([KMGT]?)(i?)B\s*$", str(size), re.I)
    if not m:
        raise ValueError(f"size {size!r}: expected forms like '5GB', '200MiB'")
    base = 1024 if m.group(3) else 1000
    return int(float(m.group(1)) * base ** " KMGT".index(m.group(2).upper() or " "))


def split_state_dict_into_shards(state: Dict[str, torch.Tensor], max_shard_size="5GB", weights_name: str = "model.safetensors"):
    """-> ({file name: {tensor name: tensor}}, index or None); tensors keep their order, a tensor larger than the limit gets a file of its own"""
    limit = parse_size(max_shard_size)
    shards: List[Dict[str, torch.Tensor]] = [{}]
    used = 0
    for k, v in state.items():
        n = _nbytes(v)
        if shards[-1] and used + n > limit:
            shards.append({})
            used = 0
        shards[-1][k] = v
        used += n
    if len(shards) == 1:
        return {weights_name: shards[0]}, None
    stem, ext = os.path.splitext(weights_name)
    files = {f"{stem}-{i + 1:05d}-of-{len(shards):05d}{ext}": s for i, s in enumerate(shards)}
    index = {"metadata": {"total_size": sum(_nbytes(v) for v in state.values())}, "weight_map": {k: f for f, s in files.items() for k in s}}
    return files, index


split_hetu_state_dict_into_shards = split_state_dict_into_shards


class PreTrainedConfig:
    """mixin for the dataclass configs of the model families"""
    model_type: str = ""

    def to_dict(self) -> Dict:
        d = dataclasses.asdict(self) if dataclasses.is_dataclass(self) else dict(self.__dict__)
        d = {k: (list(v) if isinstance(v, tuple) else v) for k, v in d.items()}
        d["model_type"] = getattr(self, "model_type", "") or type(self).__name__.replace("Config", "").lower()
        return d

    @classmethod
    def from_dict(cls, d: Dict):
        names = {f.name for f in dataclasses.fields(cls)} if dataclasses.is_dataclass(cls) else set(d)
        kw = {k: (tuple(v) if isinstance(v, list) and isinstance(getattr(cls, k, None), tuple) else v) for k, v in d.items() if k in names}
        return cls(**kw)

    def save_pretrained(self, save_directory: str):
        os.makedirs(save_directory, exist_ok=True)
        with open(os.path.join(save_directory, CONFIG_NAME), "w") as f:
            json.dump(self.to_dict(), f, indent=2, sort_keys=True)

    @classmethod
    def from_pretrained(cls, path: str, **overrides):
        if is_remote_url(path):
            raise ValueError(f"{path}: remote model hubs are not reachable; pass a local directory")
        file = os.path.join(path, CONFIG_NAME) if os.path.isdir(path) else path
        with open(file) as f:
            d = json.load(f)
        d.update(overrides)
        return cls.from_dict(d)


class PreTrainedModel:
    """mixin for the *LMHeadModel classes (`config_class` names the config dataclass)"""
    config_class = None

    def save_pretrained(self, save_directory: str, max_shard_size="5GB", state_dict: Optional[Dict[str, torch.Tensor]] = None):
        from ...utils.checkpoint import save_file
        os.makedirs(save_directory, exist_ok=True)
        cfg = self.config
        (cfg.save_pretrained(save_directory) if hasattr(cfg, "save_pretrained") else
         json.dump(dataclasses.asdict(cfg), open(os.path.join(save_directory, CONFIG_NAME), "w"), indent=2))
        state = state_dict if state_dict is not None else {k: v.detach().cpu().contiguous() for k, v in self.state_dict().items()}
        files, index = split_state_dict_into_shards(state, max_shard_size)
        for name, shard in files.items():
            save_file(shard, os.path.join(save_directory, name))
        if index is not None:
            with open(os.path.join(save_directory, WEIGHTS_INDEX_NAME), "w") as f:
                json.dump(index, f, indent=2)
        return sorted(files)

    @staticmethod
    def load_weights(path: str) -> Dict[str, torch.Tensor]:
        from ...utils.checkpoint import load_file
        index = os.path.join(path, WEIGHTS_INDEX_NAME)
        names = sorted(set(json.load(open(index))["weight_map"].values())) if os.path.exists(index) else \
            sorted(f for f in os.listdir(path) if f.endswith(".safetensors"))
        if not names:
            pt = [f for f in os.listdir(path) if f.endswith((".bin", ".pt"))]
            if not pt:
                raise FileNotFoundError(f"no weight files under {path}")
            return torch.load(os.path.join(path, pt[0]), map_location="cpu", weights_only=False)
        state: Dict[str, torch.Tensor] = {}
        for n in names:
            state.update(load_file(os.path.join(path, n)))
        return state

    @classmethod
    def from_pretrained(cls, path: str, ds_parallel_configs=None, config=None, strict: bool = True, **config_overrides) -> Tuple:
        """build the model inside the CURRENT graph context and load the weights -> model"""
        if is_remote_url(path):
            raise ValueError(f"{path}: remote model hubs are not reachable; pass a local directory")
        import inspect
        cfg = config if config is not None else cls.config_class.from_pretrained(path, **config_overrides)
        if "ds_parallel_configs" in inspect.signature(cls.__init__).parameters:
            if ds_parallel_configs is None:
                from ..parallel_config import generate_ds_parallel_config
                layers = getattr(cfg, "n_layer", None) or getattr(cfg, "num_hidden_layers")
                ds_parallel_configs = [generate_ds_parallel_config(int(layers), 1, 1, 1, 1, zero=False)]
            model = cls(cfg, ds_parallel_configs)
        else:
            model = cls(cfg)                          # model families without a parallel layout argument (BERT, seq2seq, ...)
        model.load_state_dict(cls.load_weights(path), strict=strict)
        return model
