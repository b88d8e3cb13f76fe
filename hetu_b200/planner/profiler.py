"""Profilers feeding the cost model (ref: tools/Galvatron/galvatron/core/profiler.py:243-533, profile_hardware/*):
* ModelProfiler: per-layer forward ms by layer-count differencing and activation / parameter memory by allocator deltas,
  measured through this framework's own executor;
* HardwareProfiler: all-reduce / p2p bandwidth per group size (consecutive vs strided ranks) over the live process groups;
* profile_overlap_coefficient: compute slowdown while a collective is in flight."""
from __future__ import annotations

import json
import time
from typing import Dict, Optional, Sequence

import torch

from .cost_model import HardwareProfile, LayerProfile


def _sync():
    if torch.cuda.is_available():
        torch.cuda.synchronize()


class ModelProfiler:
    def __init__(self, build_model_fn, seq_len: int, bsz: int = 1, warmup: int = 2, iters: int = 5):
        """build_model_fn(num_layers) -> (graph, loss tensor, feed dict builder(bsz))"""
        self.build, self.seq, self.bsz, self.warmup, self.iters = build_model_fn, seq_len, bsz, warmup, iters

    def _time(self, num_layers: int) -> Dict[str, float]:
        g, loss, feed_fn = self.build(num_layers)
        feed = feed_fn(self.bsz)
        if torch.cuda.is_available():
            torch.cuda.reset_peak_memory_stats()
            base = torch.cuda.memory_allocated()
        for _ in range(self.warmup):
            g.run(loss, [loss], feed)
        _sync()
        t0 = time.perf_counter()
        for _ in range(self.iters):
            g.run(loss, [loss], feed)
        _sync()
        ms = (time.perf_counter() - t0) / self.iters * 1e3
        mem = (torch.cuda.max_memory_allocated() - base) / 2**20 if torch.cuda.is_available() else 0.0
        return {"fwd_ms": ms, "peak_mb": mem}

    def profile(self, layernums: Sequence[int] = (2, 4)) -> Dict[str, float]:
        """layer-count differencing: (t(n2) - t(n1)) / (n2 - n1) removes embedding / head / launch constants"""
        a, b = self._time(layernums[0]), self._time(layernums[1])
        d = layernums[1] - layernums[0]
        return {"fwd_ms_per_layer": (b["fwd_ms"] - a["fwd_ms"]) / d / self.bsz, "act_mb_per_layer": (b["peak_mb"] - a["peak_mb"]) / d / self.bsz,
                "other_fwd_ms": a["fwd_ms"] - layernums[0] * (b["fwd_ms"] - a["fwd_ms"]) / d}

    def to_layer_profile(self, hidden: int, ffn: int, heads: int, measured: Optional[Dict[str, float]] = None, bytes_per_el: int = 2):
        lp = LayerProfile.transformer(hidden, ffn, self.seq, heads, bytes_per_el)
        if measured:
            lp.fwd_ms = measured["fwd_ms_per_layer"]
            if measured.get("act_mb_per_layer", 0) > 0:
                lp.act_mb = measured["act_mb_per_layer"]
        return lp


class HardwareProfiler:
    def __init__(self, size_mb: int = 64, iters: int = 10):
        self.size_mb, self.iters = size_mb, iters

    def allreduce_bandwidth(self, ranks: Sequence[int]) -> float:
        """bus bandwidth GB/s of an all-reduce over `ranks` (every member must call this)"""
        from .. import _C
        dev = "cuda" if torch.cuda.is_available() else "cpu"
        x = torch.ones(self.size_mb * 2**20 // 4, device=dev)
        for _ in range(2):
            _C.comm_all_reduce(x, list(ranks), "sum")
        _sync()
        t0 = time.perf_counter()
        for _ in range(self.iters):
            _C.comm_all_reduce(x, list(ranks), "sum")
        _sync()
        dt = (time.perf_counter() - t0) / self.iters
        n = len(ranks)
        return self.size_mb / 1024 * 2 * (n - 1) / n / dt

    def profile(self, world: int, gpus_per_node: int = 8) -> HardwareProfile:
        hw = HardwareProfile(gpus_per_node=gpus_per_node)
        from .. import _C, distributed
        r = distributed.rank()
        n = 2
        while n <= world:
            consecutive = [list(range(i, i + n)) for i in range(0, world, n)]
            if world > 1 and _C.comm_initialized():
                for g in consecutive:          # process groups are created collectively: every rank walks the same list
                    _C.comm_create_group(g)
            mine = next(g for g in consecutive if r in g)
            hw.allreduce_bw[n] = self.allreduce_bandwidth(mine)
            n *= 2
        return hw

    @staticmethod
    def save(hw: HardwareProfile, path: str):
        with open(path, "w") as f:
            json.dump({"allreduce_bw": hw.allreduce_bw, "p2p_bw": hw.p2p_bw, "overlap_coe": hw.overlap_coe,
                       "gpus_per_node": hw.gpus_per_node, "mem_mb": hw.mem_mb}, f, indent=2)

    @staticmethod
    def load(path: str) -> HardwareProfile:
        d = json.load(open(path))
        return HardwareProfile({int(k): v for k, v in d["allreduce_bw"].items()}, d["p2p_bw"], d.get("inter_node_bw", 45.0),
                               d["overlap_coe"], d["gpus_per_node"], d["mem_mb"])


def profile_overlap_coefficient(ranks: Sequence[int], n: int = 4096, iters: int = 10) -> float:
    """t(compute while all-reduce runs) / t(compute alone) on this rank"""
    from .. import _C
    if not torch.cuda.is_available():
        return 1.0
    a = torch.randn(n, n, device="cuda", dtype=torch.bfloat16)
    buf = torch.ones(64 * 2**20 // 4, device="cuda")
    side = torch.cuda.Stream()

    def compute():
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            a @ a
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)
    compute()
    alone = compute()
    with torch.cuda.stream(side):
        for _ in range(iters):
            _C.comm_all_reduce(buf, list(ranks), "sum")
    both = compute()
    return max(1.0, both / alone)
