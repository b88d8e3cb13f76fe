"""Straggler detection for Malleus: times a fixed GEMM workload (or the real step) on every device and reports the
slow-down ratio against the fastest (ref: python/hetu/engine/straggler.py:9-93)."""
from __future__ import annotations

import time
from dataclasses import dataclass
from typing import Dict, List, Optional

import torch


@dataclass
class WorkloadInfo:
    mbs: int = 4
    seq_len: int = 1024
    hidden_size: int = 4096
    iters: int = 8


class Straggler:
    def __init__(self, num_devices: int, log_file: Optional[str] = None, workload: Optional[WorkloadInfo] = None):
        self.num_devices, self.log_file, self.workload = num_devices, log_file, workload or WorkloadInfo()
        self._t0 = None
        self.records: List[float] = []

    @staticmethod
    def read_profile(log_file: str, length: int, ignore_first: bool = True) -> List[float]:
        vals = [float(l.split()[-1]) for l in open(log_file) if l.strip()]
        vals = vals[1:] if ignore_first and len(vals) > 1 else vals
        return vals[-length:]

    def begin_profile(self):
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        self._t0 = time.perf_counter()

    def end_profile(self) -> float:
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        dt = time.perf_counter() - self._t0
        self.records.append(dt)
        if self.log_file:
            with open(self.log_file, "a") as f:
                f.write(f"step_time {dt}\n")
        return dt

    def run_workload(self) -> float:
        """a fixed bf16 GEMM chain through the framework's own tcgen05 GEMM (ATen matmul on CPU)"""
        from .. import ops, core
        w = self.workload
        dev = "cuda" if torch.cuda.is_available() else "cpu"
        dt = torch.bfloat16 if dev == "cuda" else torch.float32
        x = torch.randn(w.mbs * w.seq_len, w.hidden_size, device=dev, dtype=dt)
        wt = torch.randn(w.hidden_size, w.hidden_size, device=dev, dtype=dt) * 0.01
        X, W = core.from_numpy(x), core.from_numpy(wt)
        self.begin_profile()
        for _ in range(w.iters):
            X = ops.linear(X, W, None)
        return self.end_profile()

    def run_profile(self) -> Dict[int, float]:
        """all-gather every device's workload time -> {device: slow-down ratio vs the fastest}"""
        from .. import distributed, _C
        t = self.run_workload()
        world = max(distributed.world_size(), 1)
        if world == 1:
            return {0: 1.0}
        ts = _C.comm_all_gather(torch.tensor([t], dtype=torch.float32), list(range(world)))
        ts = [float(v) for v in ts.reshape(-1)]
        best = min(ts)
        return {i: v / best for i, v in enumerate(ts)}
