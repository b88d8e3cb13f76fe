"""v1 distributed strategies and auto-parallel search (ref: hetu/v1/python/hetu/distributed_strategies/{base,simple,
flexflow,optcnn,gpipe,pipedream,pipeopt}.py).  A strategy maps every layer of a layer-graph description to a
(device list, split) placement; the searching strategies optimise that mapping against the analytic cost model of
hetu_b200.planner and emit the same ds_parallel_config the executor consumes.

Layer description: [{"name", "type": "conv"|"linear"|"attention"|"embedding"|..., "flops", "params" (bytes),
"act" (bytes of activation output), "splittable": ["batch", "out", "in"]}].
"""
from __future__ import annotations

import math
import random
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple


@dataclass
class LayerSpec:
    name: str
    type: str = "linear"
    flops: float = 1.0
    params: float = 0.0
    act: float = 0.0
    splittable: Tuple[str, ...] = ("batch", "out", "in")


@dataclass
class Placement:
    devices: List[int]
    split: Dict[str, int] = field(default_factory=lambda: {"batch": 1})     # axis -> parts (product == len(devices))

    def key(self):
        return (tuple(self.devices), tuple(sorted(self.split.items())))


@dataclass
class HardwareSpec:
    tflops: float = 1400.0
    bw_gbs: float = 600.0        # all-reduce bus bandwidth
    p2p_gbs: float = 700.0
    mem_gb: float = 180.0


class Strategy:
    """base: assigns a Placement to every layer"""

    def __init__(self, num_devices: int, hw: Optional[HardwareSpec] = None):
        self.n, self.hw = num_devices, hw or HardwareSpec()

    def assign(self, layers: Sequence[LayerSpec]) -> List[Placement]:
        raise NotImplementedError

    # ---- cost model shared by all strategies
    def layer_time(self, l: LayerSpec, p: Placement, batch_scale: float = 1.0) -> float:
        parts = max(1, math.prod(p.split.values()))
        comp = 3.0 * l.flops * batch_scale / parts / (self.hw.tflops * 1e12)
        comm = 0.0
        dp = p.split.get("batch", 1)
        if dp > 1:                                     # gradient all-reduce of the (possibly sharded) parameters
            shard = l.params / max(1, parts // dp)
            comm += 2.0 * shard * (dp - 1) / dp / (self.hw.bw_gbs * 1e9)
        if p.split.get("in", 1) > 1:                   # partial sums of a row split are all-reduced (fwd + bwd)
            k = p.split["in"]
            comm += 2.0 * 2.0 * l.act * batch_scale / max(dp, 1) * (k - 1) / k / (self.hw.bw_gbs * 1e9)
        return comp + comm

    def transition_time(self, a: Placement, b: Placement, act_bytes: float) -> float:
        if a.key() == b.key():
            return 0.0
        if set(a.devices) != set(b.devices):
            return act_bytes / (self.hw.p2p_gbs * 1e9)
        return 2.0 * act_bytes / max(len(b.devices), 1) / (self.hw.bw_gbs * 1e9)

    def memory(self, layers, placements) -> Dict[int, float]:
        mem: Dict[int, float] = {}
        for l, p in zip(layers, placements):
            parts = max(1, math.prod(p.split.values()))
            shard = max(1, parts // p.split.get("batch", 1))
            per = l.params * 8 / shard + l.act / parts          # params + grads + adam states (fp32 master) ~ 8x bf16 bytes
            for d in p.devices:
                mem[d] = mem.get(d, 0.0) + per
        return mem

    def total_time(self, layers, placements) -> float:
        t = 0.0
        for i, (l, p) in enumerate(zip(layers, placements)):
            t += self.layer_time(l, p)
            if i:
                t += self.transition_time(placements[i - 1], p, layers[i - 1].act)
        return t

    def feasible(self, layers, placements) -> bool:
        return all(v <= self.hw.mem_gb * 1e9 for v in self.memory(layers, placements).values())


class DataParallel(Strategy):
    def assign(self, layers):
        return [Placement(list(range(self.n)), {"batch": self.n}) for _ in layers]


class ModelParallel4CNN(Strategy):
    """convolutions data parallel, fully connected layers split over their output features"""

    def assign(self, layers):
        return [Placement(list(range(self.n)), {"out": self.n} if l.type in ("linear", "fc") else {"batch": self.n}) for l in layers]


class OneWeirdTrick4CNN(ModelParallel4CNN):
    """Krizhevsky's 'one weird trick': data-parallel convolutions, model-parallel classifier (same mapping, kept as a
    distinct strategy for API parity)"""


class ModelParallel4LM(Strategy):
    def assign(self, layers):
        out = []
        for l in layers:
            if l.type == "embedding":
                out.append(Placement(list(range(self.n)), {"out": self.n}))
            elif l.type in ("linear", "attention"):
                out.append(Placement(list(range(self.n)), {"out": self.n}))
            else:
                out.append(Placement(list(range(self.n)), {"batch": 1, "dup": self.n}))
        return out


class MegatronLM(Strategy):
    """alternate column (out) / row (in) splits inside each transformer block, data parallel across `dp` replicas"""

    def __init__(self, num_devices, tp: int, hw=None):
        super().__init__(num_devices, hw)
        self.tp, self.dp = tp, num_devices // tp

    def assign(self, layers):
        out, col = [], True
        for l in layers:
            if l.type == "attention":      # heads follow the column split of the qkv projection
                out.append(Placement(list(range(self.n)), {"batch": self.dp, "out": self.tp}))
            elif l.type == "linear":
                out.append(Placement(list(range(self.n)), {"batch": self.dp, "out" if col else "in": self.tp}))
                col = not col
            else:
                out.append(Placement(list(range(self.n)), {"batch": self.dp, "dup": self.tp}))
        return out


def _candidates(l: LayerSpec, n: int) -> List[Placement]:
    cands = []
    devs = list(range(n))
    d = 1
    while d <= n:
        rest = n // d
        if "batch" in l.splittable or d == 1:
            if rest == 1:
                cands.append(Placement(devs, {"batch": d}))
            else:
                for ax in ("out", "in"):
                    if ax in l.splittable:
                        cands.append(Placement(devs, {"batch": d, ax: rest}))
        d *= 2
    return cands or [Placement(devs, {"batch": n})]


class FlexFlowSearching(Strategy):
    """MCMC over per-layer placements (Metropolis acceptance on the simulated iteration time)"""

    def __init__(self, num_devices, hw=None, budget: int = 2000, beta: float = 50.0, seed: int = 0):
        super().__init__(num_devices, hw)
        self.budget, self.beta, self.rng = budget, beta, random.Random(seed)

    def assign(self, layers):
        cands = [_candidates(l, self.n) for l in layers]
        cur = DataParallel(self.n, self.hw).assign(layers)
        cur_t = self.total_time(layers, cur)
        best, best_t = list(cur), cur_t
        for _ in range(self.budget):
            i = self.rng.randrange(len(layers))
            prop = list(cur)
            prop[i] = self.rng.choice(cands[i])
            if not self.feasible(layers, prop):
                continue
            t = self.total_time(layers, prop)
            if t < cur_t or self.rng.random() < math.exp(-self.beta * (t - cur_t) / max(cur_t, 1e-12)):
                cur, cur_t = prop, t
                if t < best_t:
                    best, best_t = list(prop), t
        self.best_time = best_t
        return best


class OptCNNSearching(Strategy):
    """exact dynamic programming over a chain of layers: best[i][c] = min over previous candidate of cost + transition"""

    def assign(self, layers):
        cands = [_candidates(l, self.n) for l in layers]
        best = [[self.layer_time(layers[0], c) for c in cands[0]]]
        back: List[List[int]] = [[-1] * len(cands[0])]
        for i in range(1, len(layers)):
            row, brow = [], []
            for c in cands[i]:
                opts = [best[i - 1][j] + self.transition_time(p, c, layers[i - 1].act) for j, p in enumerate(cands[i - 1])]
                j = min(range(len(opts)), key=opts.__getitem__)
                row.append(opts[j] + self.layer_time(layers[i], c))
                brow.append(j)
            best.append(row)
            back.append(brow)
        j = min(range(len(best[-1])), key=best[-1].__getitem__)
        self.best_time = best[-1][j]
        out = [None] * len(layers)
        for i in range(len(layers) - 1, -1, -1):
            out[i] = cands[i][j]
            j = back[i][j]
        return out


class _PipeBase(Strategy):
    def __init__(self, num_devices, num_stages: Optional[int] = None, micro_batches: int = 8, hw=None):
        super().__init__(num_devices, hw)
        self.stages, self.mb = num_stages or num_devices, micro_batches

    def partition(self, layers, stages) -> List[int]:
        """contiguous partition minimising the bottleneck stage time (DP over prefix sums) -> stage index per layer"""
        t = [3.0 * l.flops / (self.hw.tflops * 1e12) for l in layers]
        n = len(layers)
        pre = [0.0]
        for v in t:
            pre.append(pre[-1] + v)
        INF = float("inf")
        f = [[INF] * (stages + 1) for _ in range(n + 1)]
        cut = [[0] * (stages + 1) for _ in range(n + 1)]
        f[0][0] = 0.0
        for i in range(1, n + 1):
            for s in range(1, min(stages, i) + 1):
                for k in range(s - 1, i):
                    c = max(f[k][s - 1], pre[i] - pre[k] + (layers[k - 1].act / (self.hw.p2p_gbs * 1e9) if k else 0.0))
                    if c < f[i][s]:
                        f[i][s], cut[i][s] = c, k
        self.bottleneck = f[n][stages]
        assign, i, s = [0] * n, n, stages
        while s > 0:
            k = cut[i][s]
            for j in range(k, i):
                assign[j] = s - 1
            i, s = k, s - 1
        return assign

    def assign(self, layers):
        stages = min(self.stages, len(layers))
        per = self.n // stages
        st = self.partition(layers, stages)
        return [Placement(list(range(s * per, (s + 1) * per)), {"batch": per}) for s in st]


class GPipeSearching(_PipeBase):
    """balanced contiguous stages, all forwards then all backwards: time = (mb + stages - 1) * bottleneck"""

    def estimate(self, layers):
        self.assign(layers)
        return (self.mb + min(self.stages, len(layers)) - 1) * self.bottleneck / self.mb


class PipeDreamSearching(_PipeBase):
    """PipeDream partitioner: also tries replicating stages (data parallel inside a stage) for every stage count"""

    def assign(self, layers):
        best = None
        s = 1
        while s <= min(self.n, len(layers)):
            self.stages = s
            pl = super().assign(layers)
            t = self.bottleneck / max(self.n // s, 1)        # a stage replicated r times processes r micro-batches at once
            if best is None or t < best[0]:
                best = (t, pl, s)
            s *= 2
        self.best_time, self.stages = best[0], best[2]
        return best[1]


class PipeOptSearching(PipeDreamSearching):
    """pipeline partition + per-stage intra-layer search: OptCNN's DP runs inside every stage's device group"""

    def assign(self, layers):
        base = super().assign(layers)
        out = list(base)
        groups: Dict[Tuple[int, ...], List[int]] = {}
        for i, p in enumerate(base):
            groups.setdefault(tuple(p.devices), []).append(i)
        for devs, idxs in groups.items():
            sub = OptCNNSearching(len(devs), self.hw).assign([layers[i] for i in idxs])
            for i, p in zip(idxs, sub):
                out[i] = Placement([devs[d] for d in p.devices], p.split)
        return out


def transformer_layers(num_layers: int, hidden: int, ffn: int, seq: int, batch: int, vocab: int = 50304, bytes_per_el: int = 2) -> List[LayerSpec]:
    """layer-graph description of a GPT for the strategies above"""
    tok = batch * seq
    L = [LayerSpec("embedding", "embedding", 0.0, vocab * hidden * bytes_per_el, tok * hidden * bytes_per_el, ("batch", "out"))]
    for i in range(num_layers):
        L.append(LayerSpec(f"qkv{i}", "linear", 2.0 * tok * hidden * 3 * hidden, 3 * hidden * hidden * bytes_per_el, tok * 3 * hidden * bytes_per_el))
        L.append(LayerSpec(f"attn{i}", "attention", 4.0 * batch * seq * seq * hidden, 0.0, tok * hidden * bytes_per_el, ("batch", "out")))
        L.append(LayerSpec(f"proj{i}", "linear", 2.0 * tok * hidden * hidden, hidden * hidden * bytes_per_el, tok * hidden * bytes_per_el))
        L.append(LayerSpec(f"fc1_{i}", "linear", 2.0 * tok * hidden * ffn, hidden * ffn * bytes_per_el, tok * ffn * bytes_per_el))
        L.append(LayerSpec(f"fc2_{i}", "linear", 2.0 * tok * hidden * ffn, hidden * ffn * bytes_per_el, tok * hidden * bytes_per_el))
    L.append(LayerSpec("head", "linear", 2.0 * tok * hidden * vocab, 0.0, tok * vocab * bytes_per_el, ("batch", "out")))
    return L


# ---------------------------------------------------------------------------------------------------- strategy -> executable plan
def summarize_placements(layers: Sequence[LayerSpec], placements: Sequence[Placement]) -> Dict:
    """collapse a per-layer placement list of a transformer layer graph into the (dp, tp, pp, layers per stage) form the executor's
    ds_parallel_config expresses: pipeline stages = maximal runs of layers on one device set; inside a stage the tensor-parallel
    degree is the FLOP-weighted majority of the linear layers' non-batch split, data parallel takes the rest of the group"""
    stages: List[Dict] = []
    for l, p in zip(layers, placements):
        devs = tuple(sorted(p.devices))
        if not stages or stages[-1]["devices"] != devs:
            stages.append({"devices": devs, "layers": [], "votes": {}})
        st = stages[-1]
        st["layers"].append(l.name)
        if l.type == "linear":
            tp = 1
            for axis, parts in p.split.items():
                if axis != "batch":
                    tp *= parts
            st["votes"][tp] = st["votes"].get(tp, 0.0) + max(l.flops, 1.0)
    sizes = {len(s["devices"]) for s in stages}
    assert len(sizes) == 1, f"pipeline stages of different widths {sorted(sizes)} need a heterogeneous plan (engine.HeteroSession)"
    width = sizes.pop()
    votes: Dict[int, float] = {}
    for s in stages:
        for tp, w in s["votes"].items():
            votes[tp] = votes.get(tp, 0.0) + w
    tp = max(votes, key=votes.get) if votes else 1
    tp = max(1, min(tp, width))
    while width % tp:
        tp //= 2
    blocks_per_stage = []
    for s in stages:
        ids = {int("".join(ch for ch in n if ch.isdigit())) for n in s["layers"] if n.startswith("qkv")}
        blocks_per_stage.append(len(ids))
    return {"pp": len(stages), "tp": tp, "dp": width // tp, "layer_split": blocks_per_stage,
            "devices": [d for s in stages for d in s["devices"]]}


def strategy_to_ds_parallel_config(strategy: Strategy, num_layers: int, hidden: int, ffn: int, seq: int, batch: int, model: str = "gpt",
                                   zero: bool = True) -> Dict:
    """run a v1 strategy (fixed or searching) on the layer graph of a GPT / Llama and emit the ds_parallel_config the graph
    executor consumes -- the search result becomes a runnable plan (ref: distributed_strategies/base.py Strategy.set_raw_ctxs_n_states:
    the reference writes contexts into the v1 graph; here the plan targets the DistributedStates executor)"""
    from ..models.parallel_config import generate_ds_parallel_config
    layers = transformer_layers(num_layers, hidden, ffn, seq, batch)
    placements = strategy.assign(layers)
    s = summarize_placements(layers, placements)
    pp, tp = s["pp"], s["tp"]
    split = s["layer_split"]
    if sum(split) != num_layers or any(n <= 0 for n in split):      # a stage holding only the embedding / head: spread the blocks evenly
        if num_layers >= pp:
            base, rem = divmod(num_layers, pp)
            split = [base + (1 if i < rem else 0) for i in range(pp)]
        else:
            pp, split = 1, None
    dp = strategy.n // (tp * pp)
    cfg = generate_ds_parallel_config(num_layers, strategy.n, dp, tp, pp, zero=zero, model=model, layer_split=split if pp > 1 else None)
    cfg["searched_by"] = type(strategy).__name__
    cfg["estimated_step_s"] = strategy.total_time(layers, placements)
    return cfg


class BaseSearchingStrategy(Strategy):
    """what the searching strategies share in v1: a found plan can be saved and a saved plan loaded instead of searching again
    (`save_path` / `load_path`).  Wraps any searcher: `BaseSearchingStrategy(FlexFlowSearching(8), save_path=...)`.
    (ref: hetu/v1/python/hetu/distributed_strategies/base.py BaseSearchingStrategy)"""

    def __init__(self, searcher: Strategy, save_path: Optional[str] = None, load_path: Optional[str] = None):
        super().__init__(searcher.n, searcher.hw)
        self.searcher, self.save_path, self.load_path = searcher, save_path, load_path
        self.loaded = False

    def assign(self, layers):
        import json
        import os
        if self.load_path and os.path.exists(self.load_path):
            raw = json.load(open(self.load_path))
            assert len(raw["placements"]) == len(layers), "the saved plan belongs to a model with a different number of layers"
            self.loaded = True
            return [Placement(list(p["devices"]), {k: int(v) for k, v in p["split"].items()}) for p in raw["placements"]]
        plan = self.searcher.assign(layers)
        if self.save_path:
            with open(self.save_path, "w") as f:
                json.dump({"strategy": type(self.searcher).__name__, "num_devices": self.n, "estimated_step_s": self.total_time(layers, plan),
                           "placements": [{"devices": list(p.devices), "split": dict(p.split)} for p in plan]}, f, indent=1)
        return plan
