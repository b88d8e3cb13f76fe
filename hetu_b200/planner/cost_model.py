"""Analytic cost model of one transformer layer under a (pp, tp, dp, sdp/zero, ckpt) strategy
(ref: tools/Galvatron/galvatron/core/cost_model.py MemoryCostModel / TimeCostModel)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Optional, Sequence


@dataclass(frozen=True)
class Strategy:
    pp: int
    tp: int
    dp: int
    sdp: int = 0          # 0 ddp, 2 zero-2 (grads+optimizer sharded), 3 zero-3 (parameters too)
    ckpt: bool = False
    tp_consec: bool = True
    sp: bool = False

    def key(self):
        return (self.pp, self.tp, self.dp, self.sdp, int(self.ckpt), int(self.tp_consec), int(self.sp))

    def to_json(self):
        return {"pp": self.pp, "tp": self.tp, "dp": self.dp, "sdp": self.sdp, "ckpt": int(self.ckpt), "tp_consec": int(self.tp_consec),
                "sp": int(self.sp)}


@dataclass
class LayerProfile:
    """measured (or analytic) per-layer numbers at batch size 1"""
    fwd_ms: float                   # forward time of one layer, one sample, tp = 1
    param_mb: float                 # parameter size of one layer (in the compute dtype)
    act_mb: float                   # activation memory kept for backward, one sample, tp = 1
    act_ckpt_mb: float              # activation memory with recompute (layer input only)
    boundary_mb: float = 0.0        # layer input/output activation (p2p between stages, relocation between tp degrees)
    tp_comm_mb: float = 0.0         # bytes all-reduced (or RS+AG'd) per sample per layer forward by tensor parallelism

    @staticmethod
    def transformer(hidden: int, ffn: int, seq: int, heads: int, bytes_per_el: int = 2, tflops: float = 1400.0, swiglu: bool = False,
                    kv_heads: Optional[int] = None):
        kv = kv_heads or heads
        hd = hidden // heads
        qkv = hidden * (hidden + 2 * kv * hd)
        mlp = (3 if swiglu else 2) * hidden * ffn
        params = qkv + hidden * hidden + mlp
        flops = 2 * seq * params + 4 * seq * seq * hidden / 2
        act = seq * (hidden * 10 + ffn * (3 if swiglu else 2)) * bytes_per_el     # saved tensors of one block
        return LayerProfile(fwd_ms=flops / (tflops * 1e9), param_mb=params * bytes_per_el / 2**20, act_mb=act / 2**20,
                            act_ckpt_mb=seq * hidden * bytes_per_el / 2**20, boundary_mb=seq * hidden * bytes_per_el / 2**20,
                            tp_comm_mb=2 * seq * hidden * bytes_per_el / 2**20)


@dataclass
class HardwareProfile:
    """bandwidths in GB/s; NVSwitch gives every group size the full per-GPU bandwidth (no per-link scaling)"""
    allreduce_bw: Dict[int, float] = field(default_factory=lambda: {2: 600.0, 4: 650.0, 8: 680.0})
    p2p_bw: float = 700.0
    inter_node_bw: float = 45.0
    overlap_coe: float = 1.1          # slowdown of compute when a collective runs concurrently
    gpus_per_node: int = 8
    mem_mb: float = 180 * 1024 * 0.92

    def ar_bw(self, n: int, consecutive: bool = True) -> float:
        if n <= 1:
            return float("inf")
        if n > self.gpus_per_node:
            return self.inter_node_bw
        keys = sorted(self.allreduce_bw)
        k = min(keys, key=lambda x: abs(x - n))
        return self.allreduce_bw[k] * (1.0 if consecutive else 0.97)


class MemoryCostModel:
    """MB per GPU of one layer: parameters + gradients + optimizer states (fp32 master, m, v) + activations"""

    def __init__(self, layer: LayerProfile, mixed_precision: bool = True):
        self.l, self.mp = layer, mixed_precision

    def layer_mb(self, s: Strategy, micro_bsz: float, chunks_in_flight: int) -> float:
        l = self.l
        p = l.param_mb / s.tp
        states = p * (1 + 1 + (6 if self.mp else 2))       # bf16 param + bf16 grad + fp32 master/m/v  (fp32: param+grad+m+v)
        if s.sdp == 2:
            states = p * 1 + p * (1 + (6 if self.mp else 2)) / s.dp
        elif s.sdp == 3:
            states = states / s.dp
        act = (l.act_ckpt_mb if s.ckpt else l.act_mb) * micro_bsz / (s.tp if (s.sp or True) else 1)
        return states + act * chunks_in_flight

    def other_mb(self, vocab: int, hidden: int, seq: int, s: Strategy, micro_bsz: float, bytes_per_el: int = 2) -> float:
        """embedding + head + logits on the first / last stage"""
        emb = vocab * hidden * bytes_per_el / 2**20 / s.tp
        states = emb * (8 if self.mp else 4)
        if s.sdp:
            states = emb + emb * 7 / s.dp
        logits = micro_bsz * seq * vocab * (bytes_per_el + 4) / 2**20 / s.tp
        return states + logits


class TimeCostModel:
    def __init__(self, layer: LayerProfile, hw: HardwareProfile):
        self.l, self.hw = layer, hw

    def layer_ms(self, s: Strategy, global_bsz: int, chunks: int) -> float:
        """one optimizer step's time spent in this layer (all micro-batches, fwd + bwd), including the exposed part of the
        tensor-parallel and data-parallel collectives"""
        l, hw = self.l, self.hw
        bsz_per_dp = global_bsz / s.dp
        fwd = l.fwd_ms * bsz_per_dp / s.tp
        comp = fwd * (3.0 + (1.0 if s.ckpt else 0.0))                     # bwd = 2x fwd, recompute = +1 fwd
        tp_bytes = l.tp_comm_mb * bsz_per_dp * (3 + (1 if s.ckpt else 0)) * 2 * (s.tp - 1) / max(s.tp, 1)
        tp_ms = tp_bytes / 1024 / hw.ar_bw(s.tp, s.tp_consec) * 1e3 if s.tp > 1 else 0.0
        grad_mb = l.param_mb / s.tp
        dp_bytes = grad_mb * 2 * (s.dp - 1) / max(s.dp, 1) * (1.5 if s.sdp == 3 else 1.0)
        dp_ms = dp_bytes / 1024 / hw.ar_bw(s.dp * (1 if s.tp_consec else 1), not s.tp_consec) * 1e3 if s.dp > 1 else 0.0
        # gradient sync overlaps with the backward of the other layers; compute slows by overlap_coe while it does
        bwd = comp - fwd
        overlapped = min(dp_ms, bwd)
        return fwd + bwd + overlapped * (hw.overlap_coe - 1.0) + (dp_ms - overlapped) + tp_ms

    def transition_ms(self, a: Strategy, b: Strategy, micro_bsz: float) -> float:
        """activation relocation between consecutive layers with different tensor-parallel layouts"""
        if a.tp == b.tp and a.tp_consec == b.tp_consec:
            return 0.0
        return self.l.boundary_mb * micro_bsz / 1024 / self.hw.ar_bw(max(a.tp, b.tp)) * 1e3 * 2

    def pipeline_ms(self, per_stage_ms: Sequence[float], chunks: int, p2p_mb: float, pp: int) -> float:
        """1F1B: (chunks + pp - 1) * bottleneck micro-batch time + p2p"""
        if pp == 1:
            return per_stage_ms[0]
        mb = [t / chunks for t in per_stage_ms]
        p2p = p2p_mb / 1024 / self.hw.p2p_bw * 1e3
        return (chunks + pp - 1) * (max(mb) + 2 * p2p)
