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


class GalvatronRuntime:
    """A searched (or hand-written) plan as an object: validated against the device count and a memory budget, printable per layer,
    buildable into a model + run arguments, and checkable against the cost model once step times have been measured.

        rt = GalvatronRuntime.from_json("galvatron_config_llama2-7b_8gpus.json", num_gpus=8, layer=LayerProfile.transformer(...))
        print(rt.describe())
        with ht.graph("define_and_run", create_new=True) as g, ht.autocast("bfloat16"):
            model, cfg, run_kw = rt.build(LlamaLMHeadModel, LlamaConfig.llama2_7b())
        ...
        rt.record_step(ms); print(rt.cost_model_report())

    (ref: tools/Galvatron/galvatron/core/hybrid_parallel_config.py get_hybrid_parallel_configs_api / check_hp_config /
    print_hp_config, core/hybrid_parallel_model.py construct_hybrid_parallel_model_api, profile hooks of core/profiler.py)"""

    def __init__(self, plan: dict, num_gpus: int, layer=None, hardware=None, vocab: int = 50304, hidden: int = 2048, seq: int = 1024,
                 memory_mb: float = None):
        self.plan, self.num_gpus = plan, int(num_gpus)
        self.layer, self.hw = layer, hardware
        self.vocab, self.hidden, self.seq, self.memory_mb = vocab, hidden, seq, memory_mb
        self.measured_ms: List[float] = []
        for s in plan["strategies"]:
            if s.dp * s.tp * plan["pp"] != self.num_gpus:
                object.__setattr__(s, "dp", self.num_gpus // plan["pp"] // s.tp)
        self.check()

    @classmethod
    def from_json(cls, path_or_dict, num_gpus: int, **kw):
        js = json.load(open(path_or_dict)) if isinstance(path_or_dict, str) else dict(path_or_dict)
        js.setdefault("world_size", num_gpus)
        return cls(load_plan(js), num_gpus, **kw)

    # -- validation ---------------------------------------------------------------------------------------------------
    def check(self):
        """structural rules of a hybrid-parallel plan (ref: check_hp_config)"""
        p, ss = self.plan, self.plan["strategies"]
        pp = int(p["pp"])
        if self.num_gpus % pp:
            raise ValueError(f"pp_deg {pp} does not divide {self.num_gpus} GPUs")
        per_stage = self.num_gpus // pp
        split = list(p["layer_split"])
        if len(split) != pp or sum(split) != len(ss) or any(n <= 0 for n in split):
            raise ValueError(f"pp_division {split} does not partition {len(ss)} layers into {pp} non-empty stages")
        for i, s in enumerate(ss):
            if s.tp < 1 or s.tp & (s.tp - 1) or per_stage % s.tp:
                raise ValueError(f"layer {i}: tp {s.tp} must be a power of two dividing the {per_stage} GPUs of a stage")
            if s.tp * s.dp != per_stage:
                raise ValueError(f"layer {i}: tp {s.tp} x dp {s.dp} != {per_stage} GPUs per stage")
            if s.sdp not in (0, 2, 3):
                raise ValueError(f"layer {i}: dp type {s.sdp} is not ddp (0) / zero-2 (2) / zero-3 (3)")
        gbs, chunks = int(p.get("global_bsz", 0)), max(int(p.get("chunks", 1)), 1)
        min_dp = min(s.dp for s in ss)
        if gbs and (gbs % chunks or (gbs // chunks) % min_dp):
            raise ValueError(f"global batch {gbs} / {chunks} micro-batches is not divisible by the data-parallel degree {min_dp}")
        if self.memory_mb is not None and self.layer is not None:
            worst = max(self.memory_per_stage_mb())
            if worst > self.memory_mb:
                raise ValueError(f"plan needs {worst:.0f} MB on its fullest stage, budget is {self.memory_mb:.0f} MB")
        return True

    # -- cost model views ---------------------------------------------------------------------------------------------
    def _models(self):
        from .cost_model import HardwareProfile, MemoryCostModel, TimeCostModel
        hw = self.hw or HardwareProfile()
        return MemoryCostModel(self.layer), TimeCostModel(self.layer, hw), hw

    def memory_per_stage_mb(self) -> List[float]:
        mem, _, _ = self._models()
        p, ss = self.plan, self.plan["strategies"]
        chunks = max(int(p.get("chunks", 1)), 1)
        out, lo = [], 0
        for st, nl in enumerate(p["layer_split"]):
            in_flight = min(chunks, p["pp"] - st)                    # 1F1B keeps (pp - stage) micro-batches alive
            total = 0.0
            for s in ss[lo:lo + nl]:
                micro = max(int(p.get("global_bsz", s.dp)), 1) / chunks / s.dp
                total += mem.layer_mb(s, micro, in_flight)
            out.append(total)
            lo += nl
        return out

    def predicted_step_ms(self) -> float:
        _, tm, _ = self._models()
        p, ss = self.plan, self.plan["strategies"]
        chunks = max(int(p.get("chunks", 1)), 1)
        gbs = max(int(p.get("global_bsz", ss[0].dp)), 1)
        stage_ms, lo = [], 0
        for nl in p["layer_split"]:
            t = 0.0
            for i in range(lo, lo + nl):
                t += tm.layer_ms(ss[i], gbs, chunks)
                if i + 1 < lo + nl:
                    t += tm.transition_ms(ss[i], ss[i + 1], gbs / chunks / ss[i].dp) * chunks
            stage_ms.append(t)
            lo += nl
        boundary = self.layer.boundary_mb * gbs / chunks / max(ss[0].dp, 1)
        return tm.pipeline_ms(stage_ms, chunks, boundary, int(p["pp"]))

    def record_step(self, ms: float):
        self.measured_ms.append(float(ms))

    def cost_model_report(self) -> Dict:
        """measured vs predicted step time (the reference's profile-vs-model check); ratio > 1: the model is optimistic"""
        pred = self.predicted_step_ms() if self.layer is not None else None
        med = sorted(self.measured_ms)[len(self.measured_ms) // 2] if self.measured_ms else None
        return {"predicted_ms": pred, "measured_ms": med, "ratio": (med / pred) if (pred and med) else None,
                "memory_per_stage_mb": self.memory_per_stage_mb() if self.layer is not None else None}

    # -- presentation -------------------------------------------------------------------------------------------------
    def describe(self) -> str:
        """one line per run of equal layers (ref: print_hp_config)"""
        p, ss = self.plan, self.plan["strategies"]
        lines = [f"hybrid parallel plan: {self.num_gpus} GPUs, pp {p['pp']} (layers per stage {list(p['layer_split'])}), "
                 f"global batch {p.get('global_bsz')}, {p.get('chunks', 1)} micro-batches"]
        lo = 0
        stage_of = [st for st, nl in enumerate(p["layer_split"]) for _ in range(nl)]
        while lo < len(ss):
            hi = lo
            while hi + 1 < len(ss) and ss[hi + 1].key() == ss[lo].key() and stage_of[hi + 1] == stage_of[lo]:
                hi += 1
            s = ss[lo]
            dpt = {0: "ddp", 2: "zero-2", 3: "zero-3"}[s.sdp]
            lines.append(f"  layers {lo:>3}-{hi:<3} stage {stage_of[lo]}  tp {s.tp} ({'consecutive' if s.tp_consec else 'strided'})"
                         f"{' +sp' if s.sp else ''}  dp {s.dp} ({dpt})  recompute {'on' if s.ckpt else 'off'}")
            lo = hi + 1
        return "\n".join(lines)

    def comm_groups(self):
        ss = self.plan["strategies"]
        return gen_comm_groups(self.num_gpus, self.plan["pp"], [s.tp for s in ss], [int(s.tp_consec) for s in ss])

    def to_json(self) -> dict:
        from .search_engine import GalvatronSearchEngine
        js = GalvatronSearchEngine.to_json(self.plan)
        js["world_size"] = self.num_gpus
        return js

    def build(self, model_class, model_config):
        return build_hybrid_parallel_model(self.plan, self.num_gpus, model_class, model_config)
