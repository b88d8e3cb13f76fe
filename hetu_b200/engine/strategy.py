"""Malleus straggler-resilient strategy planner: given per-device slow-down ratios, regroup the devices into
heterogeneous tensor-parallel groups, arrange the groups into `dp` pipelines, split the layers over each pipeline's
stages in proportion to stage speed and the micro-batches over pipelines in proportion to pipeline throughput.
The result is a heterogeneous ds_parallel_config the executor can hot-switch to.
(ref: python/hetu/engine/strategy.py:14-984 -- LayersProp, TPGroup, StrategyModel.make_plans/solve_tp_arrangments/
solve_pp_arrangement; engine/utils.py TrainerCtxs/TrainerStrategyArgs)
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

from ..models.parallel_config import generate_ds_parallel_config, generate_hetero_ds_parallel_config

DEVICES_PER_NODE = 8


@dataclass
class TrainerCtxs:
    bf16: bool = True
    hetero_tp_alpha: Sequence[float] = (1.0, 1.9, 3.6, 7.0)   # time of a layer on tp = full/1, /2, /4, /8 relative to full tp
    hetero_tp_weight: Sequence[float] = (1.0, 1.0, 1.0, 1.0)
    normal_layers: int = 8            # layers per stage in the homogeneous plan
    normal_mbn: int = 8               # micro-batches per pipeline in the homogeneous plan
    normal_compute_time: float = 1.0  # time of one layer on a healthy full-tp group
    memory_k: Sequence[float] = (1.0,)
    memory_embedding: float = 0.0
    memory_extra: float = 0.0
    memory_d: Sequence[float] = (1.0,)
    memory_bound: float = math.inf    # max layers (weighted) a stage may hold
    memory_safe_gap: float = 0.0
    straggler_threshold: float = 1.2
    straggler_safe_gap: float = 0.3
    top_k: int = 3


class Args:
    """attribute bundle printed field by field (ref: python/hetu/engine/utils.py Args)"""

    def print_args(self):
        for k, v in vars(self).items():
            print(f"{type(self).__name__}.{k} = {v}")


@dataclass
class TrainerDatasetArgs(Args):
    dataset: object = None
    consumed_samples: int = 0
    steps: int = 0
    epochs: int = 1
    step: int = 0
    epoch: int = 0


@dataclass
class TrainerCommArgs(Args):
    """one communicator of the straggler report: its members and how long they spent in it"""
    input_ds_union: object = None
    output_ds_union: object = None
    device_group_union: object = None
    local_device: object = None
    is_hetero: bool = False


@dataclass
class TrainerCommAllArgs(Args):
    comm_args_list: List[TrainerCommArgs] = field(default_factory=list)


@dataclass
class TrainerEnvs(Args):
    """environment knobs a (re)launched trainer process is started with"""
    run_straggler_experiment: bool = False
    run_memory_experiment: bool = False
    straggler_file: str = ""
    memory_file: str = ""
    elastic: bool = False
    event_timing: bool = True

    def as_env(self) -> Dict[str, str]:
        e = {}
        if self.run_straggler_experiment:
            e["HETU_STRAGGLER"] = "EXP"
            if self.straggler_file:
                e["HETU_STRAGGLER_LOG_FILE"] = self.straggler_file
        if self.run_memory_experiment:
            e["HETU_MEMORY_PROFILE"] = "MICRO_BATCH"
            if self.memory_file:
                e["HETU_MEMORY_LOG_FILE"] = self.memory_file
        if not self.event_timing:
            e["HETU_EVENT_TIMING"] = "OFF"
        return e


@dataclass
class TrainerStrategyArgs:
    dp: int = 1
    tp: int = 1
    pp: int = 1
    zero: bool = True
    rank_to_device_mapping: Dict[int, int] = field(default_factory=dict)
    suspended_rank_list: List[int] = field(default_factory=list)
    unused_rank_list: List[int] = field(default_factory=list)
    hetero_data: bool = False
    hetero_layers: List[List[int]] = field(default_factory=list)
    hetero_stages: List[int] = field(default_factory=list)
    hetero_micro_batch_num_list: List[int] = field(default_factory=list)


@dataclass
class LayersProp:
    """which layers (and which tensor slice) a device currently holds -- used to keep the re-plan close to the old
    placement so that hot switching moves little memory"""
    start_layer: int
    end_layer: int
    slice_idx: int
    slice_num: int

    @staticmethod
    def calculate_iou(r1, r2) -> float:
        inter = max(0, min(r1[1], r2[1]) - max(r1[0], r2[0]))
        union = max(r1[1], r2[1]) - min(r1[0], r2[0])
        return inter / union if union > 0 else 0.0


@dataclass
class TPGroup:
    devices: List[int]
    sr: float                      # slow-down ratio of the slowest member
    full_tp: int
    alpha: float = 1.0             # cost multiplier of running a layer on this (possibly reduced) tp degree

    @property
    def tp(self):
        return len(self.devices)

    @property
    def layer_time(self) -> float:
        return self.sr * self.alpha

    def copy(self):
        return TPGroup(list(self.devices), self.sr, self.full_tp, self.alpha)


class StrategyModel:
    def __init__(self, ctxs: TrainerCtxs, old_strategy_args: TrainerStrategyArgs, used_devices_sr: Dict[int, float],
                 suspended_devices_sr: Optional[Dict[int, float]] = None, unused_devices: Optional[List[int]] = None):
        self.ctxs, self.old = ctxs, old_strategy_args
        self.dp, self.tp, self.pp, self.zero = old_strategy_args.dp, old_strategy_args.tp, old_strategy_args.pp, old_strategy_args.zero
        self.used_devices_sr = dict(used_devices_sr)
        self.suspended_devices_sr = dict(suspended_devices_sr or {})
        self.unused_devices = list(unused_devices or [])
        self.all_devices_num = self.dp * self.tp * self.pp
        total = len(self.used_devices_sr) + len(self.suspended_devices_sr) + len(self.unused_devices)
        assert total == self.all_devices_num, f"{total} devices described, strategy has {self.all_devices_num}"
        self.total_layers = ctxs.normal_layers * self.pp
        self.device_to_layers_prop: Dict[int, LayersProp] = {}
        rank = 0
        for pipe in old_strategy_args.hetero_layers or [[ctxs.normal_layers] * self.pp] * self.dp:
            lo = 0
            for nl in pipe:
                devs = [old_strategy_args.rank_to_device_mapping.get(r, r) for r in range(rank, rank + self.tp)
                        if r not in old_strategy_args.unused_rank_list]
                for i, d in enumerate(devs):
                    self.device_to_layers_prop[d] = LayersProp(lo, lo + nl, i, len(devs))
                lo += nl
                rank += self.tp
        self.strategies: Optional[TrainerStrategyArgs] = None
        self.ds_parallel_configs: Optional[dict] = None

    def __eq__(self, other):
        if not isinstance(other, StrategyModel):
            return False
        if (self.dp, self.tp, self.pp, self.zero, sorted(self.unused_devices)) != \
                (other.dp, other.tp, other.pp, other.zero, sorted(other.unused_devices)):
            return False
        gap, thr = self.ctxs.straggler_safe_gap, self.ctxs.straggler_threshold
        for k, v in self.used_devices_sr.items():
            if k not in other.used_devices_sr or abs(v - other.used_devices_sr[k]) >= gap:
                return False
        for k, v in self.suspended_devices_sr.items():
            o = other.suspended_devices_sr.get(k)
            if o is None:
                return False
            if not (v >= thr and o >= thr) and abs(v - o) >= gap:
                return False
        return True

    # ------------------------------------------------------------------ tensor-parallel groups
    def _alpha(self, tp_size: int) -> float:
        k = int(round(math.log2(self.tp / tp_size))) if tp_size > 0 else 0
        a = self.ctxs.hetero_tp_alpha
        return a[min(k, len(a) - 1)]

    def solve_tp_arrangments(self) -> Tuple[List[TPGroup], List[int], List[int]]:
        """Per node: healthy devices form full-tp groups (sorted so similar speeds share a group -- a group runs at the
        speed of its slowest member); stragglers are isolated into smaller power-of-two groups, or suspended when even a
        tp=1 group of them would be slower than dropping them."""
        thr = self.ctxs.straggler_threshold
        sr = {**self.used_devices_sr, **self.suspended_devices_sr}
        groups: List[TPGroup] = []
        suspended: List[int] = []
        nodes: Dict[int, List[int]] = {}
        for d in sr:
            nodes.setdefault(d // DEVICES_PER_NODE, []).append(d)
        for _, devs in sorted(nodes.items()):
            devs.sort(key=lambda d: sr[d])
            healthy = [d for d in devs if sr[d] < thr]
            slow = [d for d in devs if sr[d] >= thr]
            while len(healthy) >= self.tp:
                g, healthy = healthy[:self.tp], healthy[self.tp:]
                groups.append(TPGroup(g, max(sr[d] for d in g), self.tp, 1.0))
            rest = healthy + slow          # leftovers: split into the largest power-of-two groups of similar speed
            rest.sort(key=lambda d: sr[d])
            size = self.tp
            while rest and size >= 1:
                if len(rest) >= size and size < self.tp or (size == self.tp and len(rest) >= size):
                    g, rest = rest[:size], rest[size:]
                    cand = TPGroup(g, max(sr[d] for d in g), self.tp, self._alpha(size))
                    # a group slower than 4x a healthy full group only lengthens the pipeline: suspend it
                    if cand.layer_time > 4.0 * self.ctxs.hetero_tp_alpha[0] * max(1.0, thr):
                        suspended += g
                    else:
                        groups.append(cand)
                else:
                    size //= 2
            suspended += rest
        return groups, suspended, list(self.unused_devices)

    # ------------------------------------------------------------------ pipelines
    def solve_pp_arrangement(self, groups: List[TPGroup]):
        """LPT assignment of tp groups to `dp` pipelines (balance total speed), stages ordered to maximise the overlap with
        the layers each device already holds, then layers split in proportion to stage speed under the memory bound."""
        groups = sorted(groups, key=lambda g: g.layer_time)
        pipes: List[List[TPGroup]] = [[] for _ in range(self.dp)]
        speed = [0.0] * self.dp
        for g in groups:
            i = min(range(self.dp), key=lambda p: (speed[p], len(pipes[p])))
            pipes[i].append(g)
            speed[i] += 1.0 / g.layer_time
        pipes = [p for p in pipes if p]
        plans = []
        for p in pipes:
            # order the stages by where their devices' old layers sat (minimises hot-switch traffic)
            def old_pos(g):
                props = [self.device_to_layers_prop[d] for d in g.devices if d in self.device_to_layers_prop]
                return sum(q.start_layer for q in props) / len(props) if props else 0.0
            p.sort(key=old_pos)
            inv = [1.0 / g.layer_time for g in p]
            tot = sum(inv)
            raw = [self.total_layers * v / tot for v in inv]
            layers = [max(1, int(math.floor(r))) for r in raw]
            # memory bound (first stage also holds the embedding, the last the head)
            bound = self.ctxs.memory_bound
            while sum(layers) < self.total_layers:
                cand = [i for i in range(len(p)) if layers[i] + 1 <= bound]
                assert cand, "memory bound too tight for the remaining devices"
                i = max(cand, key=lambda i: raw[i] - layers[i])
                layers[i] += 1
            while sum(layers) > self.total_layers:
                i = max(range(len(p)), key=lambda i: layers[i] - raw[i] if layers[i] > 1 else -1e9)
                layers[i] -= 1
            stage_time = max(l * g.layer_time for l, g in zip(layers, p))
            plans.append({"groups": p, "layers": layers, "stage_time": stage_time})
        # micro-batches ~ pipeline throughput (1 / bottleneck stage time)
        total_mb = self.ctxs.normal_mbn * self.dp
        thr = [1.0 / pl["stage_time"] for pl in plans]
        raw = [total_mb * t / sum(thr) for t in thr]
        mbs = [max(1, int(r)) for r in raw]
        while sum(mbs) < total_mb:
            i = max(range(len(mbs)), key=lambda i: raw[i] - mbs[i])
            mbs[i] += 1
        while sum(mbs) > total_mb:
            i = max(range(len(mbs)), key=lambda i: mbs[i] - raw[i] if mbs[i] > 1 else -1e9)
            mbs[i] -= 1
        for pl, m in zip(plans, mbs):
            pl["micro_batches"] = m
            # 1F1B pipeline time: (m + stages - 1) * bottleneck stage
            pl["time"] = (m + len(pl["groups"]) - 1) * pl["stage_time"] * self.ctxs.normal_compute_time
        return plans

    def estimate_time(self, plans) -> float:
        return max(pl["time"] for pl in plans)

    def make_plans(self):
        if self.strategies is not None:
            return self.strategies, self.ds_parallel_configs
        groups, suspended, unused = self.solve_tp_arrangments()
        assert groups, "no usable device left"
        plans = self.solve_pp_arrangement(groups)
        # emit
        mapping: Dict[int, int] = {}
        pipelines, hetero_layers, hetero_stages, rank = [], [], [], 0
        unused_ranks: List[int] = []
        for pl in plans:
            lo = 0
            stages = []
            for g, nl in zip(pl["groups"], pl["layers"]):
                stages.append({"devices": list(g.devices), "layers": [lo, lo + nl - 1]})
                for i in range(self.tp):
                    if i < len(g.devices):
                        mapping[rank + i] = g.devices[i]
                    else:
                        unused_ranks.append(rank + i)
                rank += self.tp
                lo += nl
            pipelines.append({"stages": stages})
            hetero_layers.append(list(pl["layers"]))
            hetero_stages.append(len(pl["groups"]))
        spare = iter(suspended + unused)
        for r in unused_ranks:
            d = next(spare, None)
            if d is not None:
                mapping[r] = d
        self.strategies = TrainerStrategyArgs(
            dp=len(plans), tp=self.tp, pp=max(hetero_stages), zero=self.zero, rank_to_device_mapping=mapping,
            suspended_rank_list=[r for r, d in mapping.items() if d in suspended], unused_rank_list=unused_ranks,
            hetero_data=len({pl["micro_batches"] for pl in plans}) > 1, hetero_layers=hetero_layers, hetero_stages=hetero_stages,
            hetero_micro_batch_num_list=[pl["micro_batches"] for pl in plans])
        self.ds_parallel_configs = generate_hetero_ds_parallel_config(self.total_layers, pipelines, zero=self.zero)
        self.plans = plans
        # When all groups kept the full tp degree and the pipelines ended up identical in depth and layer split, the plan is
        # also expressible as a homogeneous (dp, tp, pp) strategy with uneven stages on re-ordered devices, which hot
        # switching can adopt in place (`executable_config`).  Every other plan runs through the member-local path:
        # engine.hetero.HeteroSession(self.ds_parallel_configs, shares=micro-batch counts) gives each rank the homogeneous
        # graph of its own pipeline plus the cross-pipeline gradient synchronisation.
        self.executable_config = None
        full = all(g.tp == self.tp for pl in plans for g in pl["groups"])
        same = len({(tuple(pl["layers"]), len(pl["groups"])) for pl in plans}) == 1
        if full and same:
            pp = len(plans[0]["groups"])
            devices = [d for s in range(pp) for pl in plans for d in pl["groups"][s].devices]
            self.executable_config = generate_ds_parallel_config(self.total_layers, len(devices), len(plans), self.tp, pp, zero=self.zero,
                                                                 devices=devices, layer_split=plans[0]["layers"])
        return self.strategies, self.ds_parallel_configs


# ---------------------------------------------------------------------------------------------------------------------
# Ampelos-style sequence dispatch over heterogeneous pipelines (ref: python/hetu/engine/strategy_ampelos.py)
def dispatch_sequences(seq_lens: Sequence[int], pipeline_speeds: Sequence[float], quadratic_coeff: float = 1.0 / 4096) -> List[List[int]]:
    """Assign variable-length sequences to pipelines so that the estimated time (linear + attention-quadratic cost divided
    by pipeline speed) is balanced: longest-processing-time-first greedy.  Returns the sample indices per pipeline."""
    cost = [n * (1.0 + quadratic_coeff * n) for n in seq_lens]
    order = sorted(range(len(seq_lens)), key=lambda i: -cost[i])
    loads = [0.0] * len(pipeline_speeds)
    out: List[List[int]] = [[] for _ in pipeline_speeds]
    for i in order:
        p = min(range(len(loads)), key=lambda p: (loads[p] + cost[i]) / pipeline_speeds[p])
        out[p].append(i)
        loads[p] += cost[i]
    return out
