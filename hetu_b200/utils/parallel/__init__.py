"""Strategy / ds_parallel_config helpers (ref: python/hetu/utils/parallel/{read_ds,generate_ds,distributed,ds_config}.py)."""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

from ...models.parallel_config import (generate_ds_parallel_config, generate_hetero_ds_parallel_config, read_ds_parallel_config,
                                       save_ds_parallel_config)
from ...nn.parallel import config2ds, get_multi_ds_parallel_config
from ...data.dataloader import parallel_data_provider


@dataclass
class RecomputeConfig:
    recompute_granularity: Optional[str] = None      # None | "full" | "selective"
    recompute_layer_idxs_list: List[List[int]] = field(default_factory=list)
    recompute_method: Optional[str] = None
    recompute_num_layers: Optional[int] = None
    cpu_offload: bool = False


RecomputeStrategy = RecomputeConfig


@dataclass
class StrategyConfig:
    """one parallel strategy: sizes + where the ds_parallel_config JSON lives (ref: utils/parallel/ds_config.py)"""
    dp: int = 1
    tp: int = 1
    pp: int = 1
    cp: int = 1
    zero: bool = True
    sequence_parallel: bool = False
    num_gpus: Optional[int] = None
    ds_parallel_config_path: Optional[str] = None
    ds_parallel_config_name: Optional[str] = None
    recompute: RecomputeConfig = field(default_factory=RecomputeConfig)
    hetero: bool = False
    hetero_layers: Optional[List[List[int]]] = None            # layers per stage, one list per pipeline
    hetero_tp: Optional[List[int]] = None                      # tensor-parallel degree per pipeline (default: tp)
    micro_batch_num_list: Optional[List[int]] = None           # micro-batches (= batch share) per pipeline
    seq_len_list: Optional[List[int]] = None
    cp_list: Optional[List[int]] = None
    rank_to_device_mapping: Optional[Dict[int, int]] = None
    unused_rank: List[int] = field(default_factory=list)

    def world(self):
        return self.num_gpus or self.dp * self.tp * self.pp * self.cp


def convert_strategy(strategy: StrategyConfig, num_layers: int) -> dict:
    """StrategyConfig -> ds_parallel_config dict"""
    if strategy.hetero_layers:
        mapping = {int(k): int(v) for k, v in (strategy.rank_to_device_mapping or {}).items()}
        pipelines, rank = [], 0
        for p, stages in enumerate(strategy.hetero_layers):
            tp = strategy.hetero_tp[p] if strategy.hetero_tp else strategy.tp
            pl, lo = [], 0
            for nl in stages:
                devs = [mapping.get(r, r) for r in range(rank, rank + tp) if r not in strategy.unused_rank]
                pl.append({"layers": [lo, lo + int(nl) - 1], "devices": devs})
                lo += int(nl)
                rank += tp
            assert lo == num_layers, f"pipeline {p} covers {lo} layers, the model has {num_layers}"
            pipelines.append({"stages": pl})
        return generate_hetero_ds_parallel_config(num_layers, pipelines, zero=strategy.zero)
    return generate_ds_parallel_config(num_layers, strategy.world(), strategy.dp, strategy.tp, strategy.pp, strategy.cp, zero=strategy.zero)


def generate_recompute_config(dp: int, num_layers: int, hetero_layers: Sequence[Sequence[int]], recompute_granularity=None,
                              recompute_method=None, recompute_num_layers=None, recompute_layer_idxs_list=None) -> RecomputeConfig:
    """pick the layers to recompute per pipeline: 'uniform' every k-th layer, 'block' the first k of each stage"""
    idxs: List[List[int]] = []
    for p, stages in enumerate(hetero_layers):
        if recompute_layer_idxs_list:
            idxs.append(list(recompute_layer_idxs_list[min(p, len(recompute_layer_idxs_list) - 1)]))
            continue
        chosen, base = [], 0
        k = recompute_num_layers or 0
        for nl in stages:
            if recompute_granularity is None or k <= 0:
                pass
            elif recompute_method == "block":
                chosen += list(range(base, base + min(k, nl)))
            else:   # uniform
                chosen += list(range(base, base + nl))
            base += nl
        idxs.append(chosen)
    return RecomputeConfig(recompute_granularity, idxs, recompute_method, recompute_num_layers)


def get_multi_recompute_from(recompute_configs: Sequence[RecomputeConfig], layer_idx: int) -> List[bool]:
    """per strategy: is `layer_idx` recomputed (in any pipeline)?"""
    return [any(layer_idx in p for p in rc.recompute_layer_idxs_list) for rc in recompute_configs]


def parse_multi_ds_parallel_config(ds_parallel_configs: List[dict], module_name: str, _range: int = -1):
    """-> (ds_hierarchy, dg_hierarchy) for one leaf across strategies"""
    leafs = get_multi_ds_parallel_config(ds_parallel_configs, module_name, _range)
    dsh, dgh = [], []
    for leaf in leafs:
        ds_union, dg_union = config2ds(leaf)
        dsh.append(ds_union)
        dgh.append(dg_union)
    return dsh, dgh


def config_spread_zero(ds_parallel_config: dict) -> dict:
    """propagate the top-level `zero` flag onto every variable leaf (ref: read_ds.config_spread_zero)"""
    zero = ds_parallel_config.get("zero", False)

    def walk(node):
        if isinstance(node, dict):
            if node.get("type") == "variable" and "zero" not in node:
                node["zero"] = zero
            for v in node.values():
                walk(v)
        elif isinstance(node, list):
            for v in node:
                walk(v)
    walk(ds_parallel_config)
    return ds_parallel_config


def distributed_init(ngpus: Optional[int] = None, server_addr: str = "127.0.0.1", server_port: str = "23457", need_kv_store: bool = False):
    """set up this rank: torch.distributed (NCCL on GPU boxes, gloo on CPU) + the framework's comm runtime"""
    from ... import distributed as dist
    os.environ.setdefault("HETU_LOCAL_HOSTNAME", "127.0.0.1")
    dist.init_comm_group(ngpus, server_address=f"{server_addr}:{server_port}")
    return dist.local_device(), dist.global_device_group()


def get_device_index(device_group) -> int:
    from ... import distributed as dist
    return device_group.get_index(dist.local_device()) if device_group.contains(dist.local_device()) else -1


def get_local_index(device_group_or_union) -> int:
    from ... import distributed as dist
    dev = dist.local_device()
    groups = list(device_group_or_union) if not hasattr(device_group_or_union, "contains") else [device_group_or_union]
    for i, g in enumerate(groups):
        if g.contains(dev):
            return i
    return -1


def get_dg_from_union(device, dg_union):
    for i, g in enumerate(dg_union):
        if g.contains(device):
            return i, g
    return None, None


def parallel_multi_data_provider(global_data, multi_ds, device_groups):
    from ... import distributed as dist
    dev = dist.local_device()
    for ds, dg in zip(multi_ds, device_groups):
        if dg.contains(dev):
            return parallel_data_provider(global_data, ds, dg.get_index(dev))
    raise RuntimeError("local device not in any device group of the union")
