"""Galvatron search: enumerate pp degree, global batch / chunks; per pipeline stage pick a per-layer strategy with the C++
dynamic-programming core (knapsack over memory with inter-layer transition cost); emit the plan JSON and the equivalent
ds_parallel_config.  (ref: tools/Galvatron/galvatron/core/search_engine.py, csrc/dp_core.cpp:22-90)"""
from __future__ import annotations

import json
import math
from typing import List, Optional, Sequence

from .. import _C
from .cost_model import HardwareProfile, LayerProfile, MemoryCostModel, Strategy, TimeCostModel


def _pow2_upto(n):
    return [1 << i for i in range(int(math.log2(n)) + 1) if (1 << i) <= n]


class GalvatronSearchEngine:
    def __init__(self, num_layers: int, num_gpus: int, layer: LayerProfile, hw: Optional[HardwareProfile] = None, vocab: int = 50304,
                 hidden: int = 2048, seq: int = 1024, memory_mb: Optional[float] = None, mixed_precision: bool = True,
                 max_tp: int = 8, allow_ckpt: bool = True, sdp_types: Sequence[int] = (0, 2, 3), mem_unit_mb: int = 64):
        self.L, self.N, self.layer, self.hw = num_layers, num_gpus, layer, hw or HardwareProfile()
        self.vocab, self.hidden, self.seq = vocab, hidden, seq
        self.mem_mb = memory_mb or self.hw.mem_mb
        self.mem_model, self.time_model = MemoryCostModel(layer, mixed_precision), TimeCostModel(layer, self.hw)
        self.max_tp, self.allow_ckpt, self.sdp_types, self.unit = max_tp, allow_ckpt, tuple(sdp_types), mem_unit_mb

    # ------------------------------------------------------------------ strategy space
    def strategies(self, pp: int) -> List[Strategy]:
        per_stage = self.N // pp
        out = []
        for tp in _pow2_upto(min(per_stage, self.max_tp)):
            dp = per_stage // tp
            for sdp in (self.sdp_types if dp > 1 else (0,)):
                for ck in ((False, True) if self.allow_ckpt else (False,)):
                    for consec in ((True, False) if 1 < tp < per_stage else (True,)):
                        out.append(Strategy(pp, tp, dp, sdp, ck, consec, sp=tp > 1))
        return out

    # ------------------------------------------------------------------ one (pp, bsz, chunks) point
    def _solve_stage(self, n_layers: int, strategies: List[Strategy], gbs: int, chunks: int, pp: int, stage: int,
                     budget_mb: float):
        S = len(strategies)
        micro = gbs / chunks
        in_flight = min(chunks, pp - stage) if pp > 1 else 1        # 1F1B keeps (pp - stage) micro-batches alive
        mem, intra = [], []
        for _ in range(n_layers):
            for s in strategies:
                mem.append(max(1, int(math.ceil(self.mem_model.layer_mb(s, micro / s.dp, in_flight) / self.unit))))
                intra.append(self.time_model.layer_ms(s, gbs, chunks))
        inter = []
        for _ in range(n_layers):
            for a in strategies:
                for b in strategies:
                    inter.append(self.time_model.transition_ms(a, b, micro / max(a.dp, 1)) * chunks)
        cost, picks, rem = _C.galvatron_dp(n_layers, int(budget_mb // self.unit), S, mem, intra, inter)
        if not picks:
            return None
        return cost, [strategies[i] for i in picks], rem * self.unit

    def evaluate(self, pp: int, gbs: int, chunks: int, layer_split: Optional[List[int]] = None):
        if self.N % pp or self.L < pp:
            return None
        strategies = self.strategies(pp)
        split = layer_split or [self.L // pp + (1 if i < self.L % pp else 0) for i in range(pp)]
        stage_ms, plan = [], []
        for st, nl in enumerate(split):
            budget = self.mem_mb
            if st == 0 or st == pp - 1:
                budget -= min(self.mem_model.other_mb(self.vocab, self.hidden, self.seq, s, gbs / chunks / s.dp) for s in strategies)
            r = self._solve_stage(nl, strategies, gbs, chunks, pp, st, budget)
            if r is None:
                return None
            stage_ms.append(r[0])
            plan += r[1]
        p2p_mb = self.layer.boundary_mb * gbs / chunks
        total = self.time_model.pipeline_ms(stage_ms, chunks, p2p_mb, pp)
        return {"pp": pp, "global_bsz": gbs, "chunks": chunks, "layer_split": split, "time_ms": total,
                "throughput_samples_per_s": gbs / total * 1e3, "strategies": plan}

    # ------------------------------------------------------------------ full search
    def search(self, batch_sizes: Sequence[int] = (8, 16, 32, 64, 128), max_chunks: int = 32) -> Optional[dict]:
        best = None
        for pp in _pow2_upto(self.N):
            for gbs in batch_sizes:
                for chunks in [c for c in _pow2_upto(max_chunks) if c <= gbs and (pp == 1 and c <= 8 or pp > 1 and c >= pp)] or [1]:
                    r = self.evaluate(pp, gbs, chunks)
                    if r and (best is None or r["throughput_samples_per_s"] > best["throughput_samples_per_s"]):
                        best = r
        return best

    @staticmethod
    def to_json(plan: dict) -> dict:
        """the reference's galvatron_config_*.json schema: comma-separated per-layer lists"""
        ss: List[Strategy] = plan["strategies"]
        return {"pp_deg": plan["pp"], "tp_sizes_enc": ",".join(str(s.tp) for s in ss),
                "tp_consecutive_flags": ",".join(str(int(s.tp_consec)) for s in ss),
                "dp_types_enc": ",".join(str(int(s.sdp == 3)) for s in ss), "use_sp": ",".join(str(int(s.sp)) for s in ss),
                "checkpoint": ",".join(str(int(s.ckpt)) for s in ss), "global_bsz": plan["global_bsz"], "chunks": plan["chunks"],
                "pp_division": ",".join(str(n) for n in plan["layer_split"]), "pipeline_type": "pipedream_flush",
                "default_dp_type": "zero2" if any(s.sdp for s in ss) else "ddp"}

    @staticmethod
    def save(plan: dict, path: str):
        with open(path, "w") as f:
            json.dump(GalvatronSearchEngine.to_json(plan), f, indent=2)


def galvatron_plan_to_ds_parallel_config(plan: dict, num_gpus: int) -> dict:
    """per-layer (tp, dp, zero, recompute) plan -> the executor's ds_parallel_config (layers may differ in tp degree: the
    parallel modules insert the relocation comm where the layout changes)"""
    from ..models.parallel_config import generate_ds_parallel_config
    ss: List[Strategy] = plan["strategies"]
    pp = plan["pp"]
    L = len(ss)
    # start from the most common (tp, dp) and patch the blocks that differ
    common = max({(s.tp, s.dp) for s in ss}, key=lambda k: sum(1 for s in ss if (s.tp, s.dp) == k))
    cfg = generate_ds_parallel_config(L, num_gpus, common[1], common[0], pp, zero=any(s.sdp for s in ss))
    per_stage = num_gpus // pp
    lo = 0
    for st, nl in enumerate(plan["layer_split"]):
        devs = list(range(st * per_stage, (st + 1) * per_stage))
        for li in range(lo, lo + nl):
            s = ss[li]
            tmpl = generate_ds_parallel_config(1, per_stage, s.dp, s.tp, 1, zero=bool(s.sdp))["blocks"]["blocks0-0"] \
                if "blocks0-0" in generate_ds_parallel_config(1, per_stage, s.dp, s.tp, 1, zero=bool(s.sdp))["blocks"] \
                else list(generate_ds_parallel_config(1, per_stage, s.dp, s.tp, 1, zero=bool(s.sdp))["blocks"].values())[0]
            blk = json.loads(json.dumps(tmpl))

            def remap(node):
                if isinstance(node, dict):
                    if "device_group_union" in node:
                        node["device_group_union"] = [[devs[i] for i in g] for g in node["device_group_union"]]
                    for v in node.values():
                        remap(v)
                elif isinstance(node, list):
                    for v in node:
                        remap(v)
            remap(blk)
            blk["range"] = [li, li]
            blk["recompute"] = [bool(s.ckpt)]
            cfg["blocks"][f"blocks{li}"] = blk
        lo += nl
    cfg["blocks"] = {k: v for k, v in cfg["blocks"].items() if "-" not in k} or cfg["blocks"]
    return cfg
