"""Strategy JSON (`ds_parallel_config`) generators for GPT / Llama style models: homogeneous dp x tp x pp (x cp)
plus the heterogeneous form (per-pipeline tp / layer ranges).  Same schema as the reference
(ref: python/hetu/models/{gpt,llama}/generate_*_4d_config.py, engine/parallel_config.py:5).
"""
from __future__ import annotations

import json
from typing import Dict, List, Optional, Sequence


def _leaf(split: Dict[str, List[int]], dup: List[int], groups: List[List[int]], kind: str, zero: bool = False,
          recompute: Optional[List[bool]] = None, cpu_offload: Optional[List[bool]] = None) -> dict:
    n = len(groups)
    leaf = {"split": split, "dup": dup, "device_group_union": groups, "type": kind}
    if kind == "variable":
        leaf["zero"] = zero
    leaf["recompute"] = recompute or [False] * n
    leaf["cpu_offload"] = cpu_offload or [False] * n
    return leaf


def generate_ds_parallel_config(num_layers: int, num_gpus: int, dp: int, tp: int, pp: int, cp: int = 1, zero: bool = True,
                                recompute_layers: Sequence[int] = (), model: str = "gpt", devices: Optional[List[int]] = None,
                                layer_split: Optional[Sequence[int]] = None) -> dict:
    """devices are laid out [pp][dp*cp][tp] (tp innermost = NVLink neighbours).  layer_split: layers per pipeline stage
    (uneven splits move work away from slow stages -- the homogeneous-tp form of Malleus' hetero_layers)"""
    assert dp * cp * tp * pp == num_gpus, f"dp{dp} x cp{cp} x tp{tp} x pp{pp} != {num_gpus}"
    devices = list(devices) if devices is not None else list(range(num_gpus))
    dcp = dp * cp
    per_stage = dcp * tp
    stage_devs = [devices[s * per_stage:(s + 1) * per_stage] for s in range(pp)]
    base, rem = divmod(num_layers, pp)
    if layer_split is not None:
        assert len(layer_split) == pp and sum(layer_split) == num_layers and all(n > 0 for n in layer_split), "bad layer_split"
    ranges, lo = [], 0
    for s in range(pp):
        n = layer_split[s] if layer_split is not None else base + (1 if s < rem else 0)
        ranges.append([lo, lo + n - 1])
        lo += n

    def var(g, split_dim=None):
        return _leaf({str(split_dim): [tp]} if (split_dim is not None and tp > 1) else {}, [dcp if split_dim is not None else dcp * tp],
                     [g], "variable", zero)

    cfg = {
        "zero": zero, "devices": devices, "dp": dp, "tp": tp, "pp": pp, "cp": cp,
        "input": _leaf({"0": [dcp]}, [tp], [stage_devs[0]], "placeholder"),
        "wte": var(stage_devs[0], 0),
        "wpe": var(stage_devs[0]),
        "blocks": {},
        "layernorm_final": var(stage_devs[-1]),
        "lm_head": var(stage_devs[-1], 0),
        "label": _leaf({"0": [dcp]}, [tp], [stage_devs[-1]], "placeholder"),
    }
    for s in range(pp):
        g = stage_devs[s]
        lo, hi = ranges[s]
        rc = [any(lo <= l <= hi for l in recompute_layers)]
        blk = {
            "range": [lo, hi], "recompute": rc, "cpu_offload": [False],
            "layernorm1": var(g), "layernorm2": var(g),
            "attn": {"qkv": var(g, 0), "dense": var(g, 1)},
            "mlp": {"dense_h_to_4h": var(g, 0), "dense_4h_to_h": var(g, 1)},
        }
        cfg["blocks"][f"blocks{lo}-{hi}"] = blk
    return cfg


def generate_hetero_ds_parallel_config(num_layers: int, pipelines: List[dict], zero: bool = True) -> dict:
    """Heterogeneous strategy: every pipeline = {"stages": [{"devices": [...], "layers": [lo, hi]}, ...]} with its own
    TP degree per stage (Malleus / Ampelos form).  Members of a union are the pipelines."""
    n_pipe = len(pipelines)
    first = [p["stages"][0]["devices"] for p in pipelines]
    last = [p["stages"][-1]["devices"] for p in pipelines]

    def var_u(groups, split_dim=None):
        tps = [len(g) for g in groups]
        split = {str(split_dim): tps} if split_dim is not None else {}
        dup = [n_pipe] * n_pipe if split_dim is not None else [n_pipe * t for t in tps]
        return _leaf(split, dup, groups, "variable", zero)

    cfg = {"zero": zero, "hetero": True,
           "input": _leaf({"0": [n_pipe] * n_pipe}, [len(g) for g in first], first, "placeholder"),
           "wte": var_u(first, 0), "wpe": var_u(first), "blocks": {},
           "layernorm_final": var_u(last), "lm_head": var_u(last, 0),
           "label": _leaf({"0": [n_pipe] * n_pipe}, [len(g) for g in last], last, "placeholder")}
    for layer in range(num_layers):
        groups = []
        for p in pipelines:
            for st in p["stages"]:
                if st["layers"][0] <= layer <= st["layers"][1]:
                    groups.append(st["devices"])
        cfg["blocks"][f"blocks{layer}"] = {
            "range": [layer, layer], "recompute": [False] * n_pipe, "cpu_offload": [False] * n_pipe,
            "layernorm1": var_u(groups), "layernorm2": var_u(groups),
            "attn": {"qkv": var_u(groups, 0), "dense": var_u(groups, 1)},
            "mlp": {"dense_h_to_4h": var_u(groups, 0), "dense_4h_to_h": var_u(groups, 1)}}
    return cfg


def localize_hetero_config(cfg: dict, rank: int) -> dict:
    """Member-local view of a heterogeneous strategy: the pipeline (union member) that contains `rank` as an ordinary
    homogeneous (dp = 1, tp_m, pp_m) ds_parallel_config on its own devices.  Every localized leaf keeps a reference to the
    original leaf (`_hetero_orig`) and its member index, from which the cross-pipeline gradient synchronisation groups are
    derived (hetu_b200.nn.parallel.hetero_grad_sync_spec)."""
    import copy

    pipe = [None]

    def find_pipe(node):
        if isinstance(node, dict):
            if "device_group_union" in node and "type" in node:
                for m, devs in enumerate(node["device_group_union"]):
                    if rank in devs and pipe[0] is None:
                        pipe[0] = m
            else:
                for v in node.values():
                    find_pipe(v)

    find_pipe(cfg)
    assert pipe[0] is not None, f"rank {rank} does not appear in the heterogeneous strategy"

    def member_of(leaf):
        return pipe[0]

    def walk(node):
        if isinstance(node, dict) and "device_group_union" in node and "type" in node:
            m = member_of(node)
            devs = list(node["device_group_union"][m])
            split = {k: [v[m] if isinstance(v, (list, tuple)) else v] for k, v in node.get("split", {}).items()}
            tp_m = len(devs)
            if node["type"] == "placeholder":
                loc = {"split": {"0": [1]}, "dup": [tp_m], "device_group_union": [devs], "type": "placeholder"}
            else:
                has_split = any(v[0] > 1 for v in split.values())
                loc = {"split": {k: v for k, v in split.items() if v[0] > 1}, "dup": [1 if has_split else tp_m], "device_group_union": [devs],
                       "type": "variable", "zero": False}
            loc["_hetero_orig"], loc["_hetero_member"] = node, m
            return loc
        if isinstance(node, dict):
            return {k: (walk(v) if isinstance(v, (dict, list)) else copy.copy(v)) for k, v in node.items()}
        if isinstance(node, list):
            return [walk(v) if isinstance(v, (dict, list)) else v for v in node]
        return node

    # blocks: a member only instantiates each layer once, on the stage that holds it in THIS member
    out = walk(cfg)
    out["hetero"], out["zero"] = False, False
    for name, blk in list(out.get("blocks", {}).items()):
        blk["recompute"], blk["cpu_offload"] = [False], [False]
    return out


def read_ds_parallel_config(path_or_list) -> List[dict]:
    """one file per strategy (comma separated) -> list of configs"""
    if isinstance(path_or_list, (list, tuple)) and path_or_list and isinstance(path_or_list[0], dict):
        return list(path_or_list)
    paths = path_or_list.split(",") if isinstance(path_or_list, str) else list(path_or_list)
    out = []
    for p in paths:
        with open(p) as f:
            out.append(json.load(f))
    return out


def save_ds_parallel_config(cfg: dict, path: str):
    with open(path, "w") as f:
        json.dump(cfg, f, indent=2)


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser(description="emit a ds_parallel_config JSON")
    ap.add_argument("--num_layers", type=int, default=32)
    ap.add_argument("--num_gpus", type=int, default=8)
    ap.add_argument("--dp", type=int, default=2)
    ap.add_argument("--tp", type=int, default=2)
    ap.add_argument("--pp", type=int, default=2)
    ap.add_argument("--cp", type=int, default=1)
    ap.add_argument("--zero", action="store_true")
    ap.add_argument("--out", type=str, default="")
    a = ap.parse_args()
    c = generate_ds_parallel_config(a.num_layers, a.num_gpus, a.dp, a.tp, a.pp, a.cp, a.zero)
    s = json.dumps(c, indent=2)
    if a.out:
        open(a.out, "w").write(s)
    else:
        print(s)
