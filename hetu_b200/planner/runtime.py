"""Galvatron runtime side: turn a searched plan (JSON: per-layer tp sizes, consecutive flags, dp types, checkpoint flags,
pp division) into process-group layouts and a runnable model.  In the reference this is a PyTorch FSDP / Megatron-TP
runtime of its own (tools/Galvatron/galvatron/core/{comm_groups,parallel,pipeline/*,grad_reduce}.py); here the plan is
lowered onto the framework's executor, which already implements the pieces: per-layer TP/DP groups = DistributedStates of
each block, 1F1B / GPipe schedules, ZeRO-2/3 sharding, activation recompute, relocation of activations between layers of
different TP degree (the parallel modules insert the re-sharding comm op where the layout changes)."""
from __future__ import annotations

import json
from typing import Dict, List, Sequence

from .cost_model import Strategy
from .search_engine import galvatron_plan_to_ds_parallel_config


def load_plan(path_or_dict) -> dict:
    """galvatron_config_*.json (comma-separated per-layer lists) -> plan dict with `strategies`"""
    js = json.load(open(path_or_dict)) if isinstance(path_or_dict, str) else dict(path_or_dict)
    ints = lambda k: [int(v) for v in str(js[k]).split(",")]   # noqa: E731
    tps, consec, dpt, ck = ints("tp_sizes_enc"), ints("tp_consecutive_flags"), ints("dp_types_enc"), ints("checkpoint")
    pp = int(js["pp_deg"])
    split = ints("pp_division") if "pp_division" in js else [len(tps) // pp] * pp
    world = int(js.get("world_size", 0)) or None
    default_sdp = 2 if js.get("default_dp_type", "ddp") == "zero2" else 0
    strategies = []
    for tp, c, d, k in zip(tps, consec, dpt, ck):
        per_stage = (world // pp) if world else None
        dp = (per_stage // tp) if per_stage else 1
        strategies.append(Strategy(pp, tp, dp, 3 if d else default_sdp, bool(k), bool(c), sp=tp > 1))
    return {"pp": pp, "global_bsz": int(js.get("global_bsz", 8)), "chunks": int(js.get("chunks", 1)), "layer_split": split,
            "strategies": strategies}


def gen_comm_groups(world_size: int, pp: int, tp_sizes: Sequence[int], tp_consecutive: Sequence[int]) -> List[Dict[str, List[List[int]]]]:
    """per layer: {"tp": [[ranks]...], "dp": [[ranks]...], "pp": [[ranks]...]} -- consecutive TP puts a TP group on
    neighbouring ranks (same NVLink island first), strided TP interleaves it with the data-parallel dimension"""
    per_stage = world_size // pp
    out = []
    for li, (tp, consec) in enumerate(zip(tp_sizes, tp_consecutive)):
        dp = per_stage // tp
        tp_groups, dp_groups = [], []
        for s in range(pp):
            base = s * per_stage
            if consec:
                tp_groups += [[base + d * tp + t for t in range(tp)] for d in range(dp)]
                dp_groups += [[base + d * tp + t for d in range(dp)] for t in range(tp)]
            else:
                tp_groups += [[base + t * dp + d for t in range(tp)] for d in range(dp)]
                dp_groups += [[base + t * dp + d for d in range(dp)] for t in range(tp)]
        pp_groups = [[s * per_stage + i for s in range(pp)] for i in range(per_stage)]
        out.append({"tp": tp_groups, "dp": dp_groups, "pp": pp_groups})
    return out


def build_hybrid_parallel_model(plan: dict, num_gpus: int, model_class, model_config):
    """-> (model, ds_parallel_config, run_kwargs): construct the model under the plan inside the current graph"""
    for s in plan["strategies"]:
        if s.dp * s.tp * plan["pp"] != num_gpus:       # plans loaded without world_size: fill dp from the device count
            object.__setattr__(s, "dp", num_gpus // plan["pp"] // s.tp)
    cfg = galvatron_plan_to_ds_parallel_config(plan, num_gpus)
    model = model_class(model_config, [cfg])
    dp = plan["strategies"][0].dp
    return model, cfg, {"num_micro_batches": int(plan["chunks"]), "grad_scale": 1.0 / dp}
