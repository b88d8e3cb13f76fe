"""Ampelos elastic strategy planner: re-plan a running (dp, tp, pp) job after devices *disappeared* (dead GPUs / lost
nodes), slowed down (stragglers) or came back -- the elastic generalisation of the Malleus planner in `strategy.py`.

Compared with `StrategyModel` it
  * accepts dead devices (the surviving device count is no longer dp * tp * pp),
  * searches over the NUMBER of pipelines and lets pipelines differ in depth (heterogeneous stage counts),
  * explores tensor-parallel regroupings by repeatedly splitting the slowest group in two (more, narrower stages),
  * assigns the groups to pipelines with a balanced k-way partition (LPT seed + move / swap refinement) instead of one
    greedy pass, and
  * ranks all candidates by estimated 1F1B step time, breaking ties by how much state a hot switch would have to move
    (IoU of the layers a device holds before and after), returning the `top_k` best.

(ref: python/hetu/engine/strategy_ampelos.py -- LayersProp/TPGroup.split/HMP :14-143, StrategyModel.make_plans :241,
solve_tp_arrangments_new :589, enumerate_pp_pattern :906 (hetero stages), enumerate_balanced_pp_pattern :1270
(partition_into_k_groups), solve_pp_arrangement :1566; python/hetu/rpc/heturpc_elastic_server.py ElasticStrategy)
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

from ..models.parallel_config import generate_hetero_ds_parallel_config
from .strategy import DEVICES_PER_NODE, LayersProp, StrategyModel, TPGroup, TrainerCtxs, TrainerStrategyArgs


@dataclass
class HMP:
    """one heterogeneous model-parallel candidate: pipelines of tensor-parallel groups with their layer / micro-batch split"""
    pipelines: List[List[TPGroup]]
    layers: List[List[int]] = field(default_factory=list)
    micro_batches: List[int] = field(default_factory=list)
    time: float = math.inf            # estimated 1F1B step time (units of normal_compute_time)
    moved: float = 0.0                # fraction of (device, layer) state that changes owner in a hot switch
    note: str = ""

    def __lt__(self, other):
        return (self.time, self.moved) < (other.time, other.moved)

    def describe(self) -> str:
        parts = []
        for p, ls, m in zip(self.pipelines, self.layers, self.micro_batches):
            parts.append(" | ".join(f"tp{g.tp}x{l}L" for g, l in zip(p, ls)) + f"  (mb {m})")
        return f"time {self.time:.3f}, moved {self.moved:.2f}: " + " ;; ".join(parts)


def partition_into_k_groups(weights: Sequence[float], k: int, max_rounds: int = 200) -> List[List[int]]:
    """balanced k-way partition of item indices (maximise the minimum group weight sum == balance pipeline throughput):
    longest-processing-time seed, then single-item moves and pairwise swaps while the spread shrinks"""
    n = len(weights)
    assert 1 <= k <= n
    order = sorted(range(n), key=lambda i: -weights[i])
    groups: List[List[int]] = [[] for _ in range(k)]
    sums = [0.0] * k
    for i in order:
        j = min(range(k), key=lambda g: (sums[g], len(groups[g])))
        groups[j].append(i)
        sums[j] += weights[i]
    for _ in range(max_rounds):
        hi, lo = max(range(k), key=lambda g: sums[g]), min(range(k), key=lambda g: sums[g])
        spread = sums[hi] - sums[lo]
        if spread <= 1e-12:
            break
        best = None
        for i in groups[hi]:                       # move one item hi -> lo
            if len(groups[hi]) == 1:
                break
            new = max(sums[hi] - weights[i], sums[lo] + weights[i]) - min(sums[hi] - weights[i], sums[lo] + weights[i])
            if new < spread - 1e-12 and (best is None or new < best[0]):
                best = (new, "move", i, None)
        for i in groups[hi]:                       # swap a pair
            for j in groups[lo]:
                d = weights[i] - weights[j]
                if d <= 0:
                    continue
                new = abs((sums[hi] - d) - (sums[lo] + d))
                if new < spread - 1e-12 and (best is None or new < best[0]):
                    best = (new, "swap", i, j)
        if best is None:
            break
        _, kind, i, j = best
        groups[hi].remove(i); groups[lo].append(i)
        sums[hi] -= weights[i]; sums[lo] += weights[i]
        if kind == "swap":
            groups[lo].remove(j); groups[hi].append(j)
            sums[lo] -= weights[j]; sums[hi] += weights[j]
    return groups


class AmpelosStrategyModel(StrategyModel):
    def __init__(self, ctxs: TrainerCtxs, old_strategy_args: TrainerStrategyArgs, used_devices_sr: Dict[int, float],
                 suspended_devices_sr: Optional[Dict[int, float]] = None, unused_devices: Optional[List[int]] = None,
                 dead_devices: Optional[Sequence[int]] = None, split_iters: int = 2, devices_per_node: int = DEVICES_PER_NODE):
        self.dead_devices = sorted(set(dead_devices or []))
        self.split_iters = split_iters
        self.devices_per_node = devices_per_node
        dead = set(self.dead_devices)
        used = {d: s for d, s in used_devices_sr.items() if d not in dead}
        susp = {d: s for d, s in (suspended_devices_sr or {}).items() if d not in dead}
        unused = [d for d in (unused_devices or []) if d not in dead]
        # the parent asserts that all dp*tp*pp devices are described; the dead ones count as unused there
        super().__init__(ctxs, old_strategy_args, used, susp, unused + self.dead_devices)
        self.unused_devices = unused
        self.candidates: List[HMP] = []

    # ------------------------------------------------------------------ tensor-parallel groups
    def _sr(self, d: int) -> float:
        return self.used_devices_sr.get(d, self.suspended_devices_sr.get(d, 1.0))

    def _group(self, devs: Sequence[int]) -> TPGroup:
        devs = sorted(devs, key=self._sr)
        return TPGroup(list(devs), max(self._sr(d) for d in devs), self.tp, self._alpha(len(devs)))

    def solve_tp_arrangments(self):
        """per node (tensor parallelism never crosses a node): sort the surviving devices by speed, cut them into the
        largest power-of-two groups <= tp; a straggler that would throttle a healthy group gets a narrower group of its own"""
        thr = self.ctxs.straggler_threshold
        nodes: Dict[int, List[int]] = {}
        for d in list(self.used_devices_sr) + list(self.suspended_devices_sr):
            nodes.setdefault(d // self.devices_per_node, []).append(d)
        groups: List[TPGroup] = []
        suspended: List[int] = []
        for _, devs in sorted(nodes.items()):
            healthy = sorted((d for d in devs if self._sr(d) < thr), key=self._sr)
            slow = sorted((d for d in devs if self._sr(d) >= thr), key=self._sr)
            for pool in (healthy, slow):
                size = self.tp
                while pool and size >= 1:
                    if len(pool) >= size:
                        g, pool = pool[:size], pool[size:]
                        cand = self._group(g)
                        # a group 4x slower than a healthy full-width one only lengthens its pipeline: park its devices
                        if cand.layer_time > 4.0 * self.ctxs.hetero_tp_alpha[0] * max(1.0, thr):
                            suspended += g
                        else:
                            groups.append(cand)
                    else:
                        size //= 2
        return groups, suspended, list(self.unused_devices)

    def tp_variants(self, groups: List[TPGroup]) -> List[List[TPGroup]]:
        """the base arrangement plus `split_iters` refinements, each splitting the currently slowest splittable group into two
        narrower ones (more stages of less work each -- pays when one wide group is the pipeline bottleneck)"""
        out = [list(groups)]
        cur = list(groups)
        for _ in range(self.split_iters):
            cand = [g for g in cur if g.tp >= 2 and g.tp % 2 == 0]
            if not cand:
                break
            worst = max(cand, key=lambda g: (g.sr, g.tp))
            devs = sorted(worst.devices, key=self._sr)
            half = worst.tp // 2
            cur = [g for g in cur if g is not worst] + [self._group(devs[:half]), self._group(devs[half:])]
            out.append(list(cur))
        return out

    # ------------------------------------------------------------------ pipelines
    def _split_layers(self, pipe: List[TPGroup]) -> Optional[List[int]]:
        """layers ~ stage speed under the per-stage memory bound; None when the pipeline cannot hold the model"""
        n = len(pipe)
        if n > self.total_layers:
            return None
        inv = [1.0 / g.layer_time for g in pipe]
        raw = [self.total_layers * v / sum(inv) for v in inv]
        # a narrower group holds 1 / tp of each layer per device: its memory bound scales with its width
        bound = [max(1.0, self.ctxs.memory_bound * g.tp / self.tp) if math.isfinite(self.ctxs.memory_bound) else math.inf for g in pipe]
        if sum(min(b, self.total_layers) for b in bound) < self.total_layers:
            return None
        layers = [max(1, min(int(math.floor(r)), int(min(b, self.total_layers)))) for r, b in zip(raw, bound)]
        while sum(layers) < self.total_layers:
            cand = [i for i in range(n) if layers[i] + 1 <= bound[i]]
            if not cand:
                return None
            i = min(cand, key=lambda i: (layers[i] + 1) * pipe[i].layer_time)     # grow the stage that stays fastest
            layers[i] += 1
        while sum(layers) > self.total_layers:
            i = max((i for i in range(n) if layers[i] > 1), key=lambda i: layers[i] * pipe[i].layer_time, default=None)
            if i is None:
                return None
            layers[i] -= 1
        return layers

    def _order_stages(self, pipe: List[TPGroup]) -> List[TPGroup]:
        """stages in the order of the layers their devices hold today (less state to move in the hot switch)"""
        def old_pos(g):
            props = [self.device_to_layers_prop[d] for d in g.devices if d in self.device_to_layers_prop]
            return sum(q.start_layer for q in props) / len(props) if props else float(self.total_layers)
        return sorted(pipe, key=old_pos)

    def _moved_fraction(self, hmp: HMP) -> float:
        tot, keep = 0.0, 0.0
        for pipe, layers in zip(hmp.pipelines, hmp.layers):
            lo = 0
            for g, nl in zip(pipe, layers):
                for d in g.devices:
                    tot += 1.0
                    q = self.device_to_layers_prop.get(d)
                    if q is not None:
                        keep += LayersProp.calculate_iou((q.start_layer, q.end_layer), (lo, lo + nl)) * min(1.0, q.slice_num / max(g.tp, 1))
                lo += nl
        return 1.0 - keep / tot if tot else 0.0

    def evaluate(self, pipelines: List[List[TPGroup]], note: str = "") -> Optional[HMP]:
        pipes, layers = [], []
        for p in pipelines:
            p = self._order_stages(p)
            ls = self._split_layers(p)
            if ls is None:
                return None
            pipes.append(p); layers.append(ls)
        stage_t = [max(l * g.layer_time for l, g in zip(ls, p)) for p, ls in zip(pipes, layers)]
        total_mb = self.ctxs.normal_mbn * self.dp            # the global batch does not shrink when devices die
        if total_mb < len(pipes):
            return None
        # micro-batches: minimise max_p (m_p + depth_p - 1) * t_p  -- greedy water-filling from one micro-batch each
        mbs = [1] * len(pipes)
        for _ in range(total_mb - len(pipes)):
            i = min(range(len(pipes)), key=lambda i: (mbs[i] + 1 + len(pipes[i]) - 1) * stage_t[i])
            mbs[i] += 1
        time = max((m + len(p) - 1) * t for m, p, t in zip(mbs, pipes, stage_t))
        hmp = HMP(pipes, layers, mbs, time * self.ctxs.normal_compute_time, 0.0, note)
        hmp.moved = self._moved_fraction(hmp)
        return hmp

    def enumerate_balanced_pp_pattern(self, groups: List[TPGroup], note: str = "") -> List[HMP]:
        """for every feasible pipeline count k: balanced partition of the groups by speed (1 / layer_time), so the
        pipelines get similar throughput; pipelines may end up with different depths (heterogeneous stages)"""
        out = []
        speeds = [1.0 / g.layer_time for g in groups]
        k_max = min(len(groups), self.ctxs.normal_mbn * self.dp)
        for k in range(1, k_max + 1):
            parts = partition_into_k_groups(speeds, k)
            hmp = self.evaluate([[groups[i] for i in part] for part in parts], f"{note} k={k}")
            if hmp is not None:
                out.append(hmp)
        return out

    def make_plans(self):
        if self.strategies is not None:
            return self.strategies, self.ds_parallel_configs
        groups, suspended, unused = self.solve_tp_arrangments()
        assert groups, "no usable device left"
        cands: List[HMP] = []
        for vi, variant in enumerate(self.tp_variants(groups)):
            cands += self.enumerate_balanced_pp_pattern(variant, f"split{vi}")
        assert cands, "no feasible plan: the memory bound cannot hold the model on the surviving devices"
        cands.sort()
        self.candidates = cands[:max(self.ctxs.top_k, 1)]
        best = self.candidates[0]
        self.plans = [{"groups": p, "layers": ls, "micro_batches": m,
                       "stage_time": max(l * g.layer_time for l, g in zip(ls, p)), "time": best.time}
                      for p, ls, m in zip(best.pipelines, best.layers, best.micro_batches)]
        # emit: ranks are numbered pipeline by pipeline, stage by stage, tp slot by tp slot (slots a narrow group leaves empty
        # are "unused ranks"; parked / spare devices fill them so every rank still maps to a physical device)
        mapping: Dict[int, int] = {}
        pipelines, hetero_layers, hetero_stages, unused_ranks, rank = [], [], [], [], 0
        for p, ls in zip(best.pipelines, best.layers):
            lo, stages = 0, []
            for g, nl in zip(p, ls):
                stages.append({"devices": list(g.devices), "layers": [lo, lo + nl - 1]})
                for i in range(self.tp):
                    if i < g.tp:
                        mapping[rank + i] = g.devices[i]
                    else:
                        unused_ranks.append(rank + i)
                rank += self.tp
                lo += nl
            pipelines.append({"stages": stages})
            hetero_layers.append(list(ls))
            hetero_stages.append(len(p))
        spare = iter(suspended + unused)
        for r in unused_ranks:
            d = next(spare, None)
            if d is not None:
                mapping[r] = d
        self.strategies = TrainerStrategyArgs(
            dp=len(best.pipelines), tp=self.tp, pp=max(hetero_stages), zero=self.zero, rank_to_device_mapping=mapping,
            suspended_rank_list=[r for r, d in mapping.items() if d in suspended], unused_rank_list=unused_ranks,
            hetero_data=len(set(best.micro_batches)) > 1, hetero_layers=hetero_layers, hetero_stages=hetero_stages,
            hetero_micro_batch_num_list=list(best.micro_batches))
        self.ds_parallel_configs = generate_hetero_ds_parallel_config(self.total_layers, pipelines, zero=self.zero)
        self.executable_config = None
        return self.strategies, self.ds_parallel_configs

    def estimate_time(self, plans=None) -> float:
        return self.candidates[0].time if self.candidates else math.inf


def replan_after_failure(ctxs: TrainerCtxs, old: TrainerStrategyArgs, alive_devices: Sequence[int],
                         straggler_ratios: Optional[Dict[int, float]] = None, **kw) -> AmpelosStrategyModel:
    """convenience entry used by the elastic server: everything that is not alive is dead; unknown speeds count as healthy"""
    total = old.dp * old.tp * old.pp
    alive = set(alive_devices)
    sr = {d: float((straggler_ratios or {}).get(d, 1.0)) for d in range(total) if d in alive}
    return AmpelosStrategyModel(ctxs, old, sr, {}, [], dead_devices=[d for d in range(total) if d not in alive], **kw)
