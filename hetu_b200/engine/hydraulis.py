"""Hydraulis-style dynamic dispatch: every step the global batch of variable-length sequences is split over several
parallel strategies (e.g. a long-sequence strategy with large tp/cp and a short-sequence data-parallel one) so that the
estimated makespan is minimal, and inside a strategy over its data-parallel replicas.  The planner can run in its own
process and stream plans to the trainers through the KV store (`rpc.kv_store.ProducerConsumer`).
(ref: examples/hydraulis -- strategy/dynamic_scip.py solves the assignment as an ILP with pyscipopt over a profiled cost
model; here the same objective is solved with an exact DP over sorted sequences for <= 2 strategies and an LPT + local
search heuristic in general, which needs no solver package.)"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Sequence, Tuple


@dataclass
class StrategyCost:
    """time(seq) = a * seq + b * seq^2 + c per sequence on ONE data-parallel replica of the strategy; `max_seq` is the longest
    sequence that fits in memory; `dp` replicas work in parallel; switching to the strategy costs `switch_ms` once per step"""
    name: str
    dp: int
    a: float
    b: float
    c: float = 0.0
    max_seq: int = 1 << 30
    switch_ms: float = 0.0

    def seq_ms(self, n: int) -> float:
        return self.a * n + self.b * n * n + self.c

    @staticmethod
    def fit(name: str, dp: int, samples: Sequence[Tuple[int, float]], max_seq: int = 1 << 30, switch_ms: float = 0.0) -> "StrategyCost":
        """least-squares fit of (seq_len, ms) profile points to a*n + b*n^2 + c"""
        import numpy as np
        n = np.array([s for s, _ in samples], dtype=np.float64)
        t = np.array([m for _, m in samples], dtype=np.float64)
        A = np.stack([n, n * n, np.ones_like(n)], 1)
        coef, *_ = np.linalg.lstsq(A, t, rcond=None)
        return StrategyCost(name, dp, float(max(coef[0], 0)), float(max(coef[1], 0)), float(max(coef[2], 0)), max_seq, switch_ms)


def _replica_makespan(costs: List[float], replicas: int) -> Tuple[float, List[List[int]]]:
    """longest-processing-time-first assignment of per-sequence costs to `replicas` bins"""
    order = sorted(range(len(costs)), key=lambda i: -costs[i])
    loads = [0.0] * replicas
    bins: List[List[int]] = [[] for _ in range(replicas)]
    for i in order:
        r = min(range(replicas), key=loads.__getitem__)
        loads[r] += costs[i]
        bins[r].append(i)
    return (max(loads) if loads else 0.0), bins


def dispatch_batch(seq_lens: Sequence[int], strategies: Sequence[StrategyCost], sequential: bool = True) -> Dict:
    """-> {"assignment": [strategy index per sequence], "per_strategy": [{"indices", "replicas": [[...]], "ms"}], "makespan_ms"}.
    sequential=True: the strategies run one after another on the same devices (hot switching, HotSPa / Hydraulis) so the
    step time is the SUM of their makespans; False: they run concurrently on disjoint devices (MAX)."""
    n, S = len(seq_lens), len(strategies)
    feasible = [[j for j, st in enumerate(strategies) if seq_lens[i] <= st.max_seq] for i in range(n)]
    if any(not f for f in feasible):
        raise ValueError("a sequence fits no strategy")

    def evaluate(assign: List[int]):
        per, total = [], 0.0
        for j, st in enumerate(strategies):
            idx = [i for i in range(n) if assign[i] == j]
            ms, bins = _replica_makespan([st.seq_ms(seq_lens[i]) for i in idx], st.dp)
            ms += st.switch_ms if idx else 0.0
            per.append({"indices": idx, "replicas": [[idx[k] for k in b] for b in bins], "ms": ms})
            total = total + ms if sequential else max(total, ms)
        return total, per

    # start: every sequence on its cheapest feasible strategy (per-replica amortised cost)
    assign = [min(feasible[i], key=lambda j: strategies[j].seq_ms(seq_lens[i]) / strategies[j].dp) for i in range(n)]
    best, per = evaluate(assign)
    # threshold sweep (exact for two strategies ordered by length affinity): sequences longer than t go to the strategy that
    # supports the longest sequences
    if S >= 2:
        long_j = max(range(S), key=lambda j: strategies[j].max_seq)
        for t in sorted(set(seq_lens)) + [max(seq_lens) + 1]:
            cand = [long_j if (seq_lens[i] >= t or long_j not in feasible[i] and False) else
                    min((j for j in feasible[i] if j != long_j), default=long_j, key=lambda j: strategies[j].seq_ms(seq_lens[i]) / strategies[j].dp)
                    for i in range(n)]
            if any(c not in feasible[i] for i, c in enumerate(cand)):
                continue
            v, p = evaluate(cand)
            if v < best:
                best, per, assign = v, p, cand
    # local search: move single sequences while it helps
    improved = True
    while improved:
        improved = False
        for i in sorted(range(n), key=lambda k: -seq_lens[k]):
            for j in feasible[i]:
                if j == assign[i]:
                    continue
                cand = list(assign)
                cand[i] = j
                v, p = evaluate(cand)
                if v < best - 1e-9:
                    best, per, assign, improved = v, p, cand, True
    return {"assignment": assign, "per_strategy": per, "makespan_ms": best}


class HydraulisPlanner:
    """produces one dispatch plan per step; with a KV client the plans are published ahead of the trainer"""

    def __init__(self, strategies: Sequence[StrategyCost], producer=None):
        self.strategies, self.producer = list(strategies), producer
        self.step = 0

    def plan(self, seq_lens: Sequence[int]) -> Dict:
        p = dispatch_batch(seq_lens, self.strategies)
        p["step"] = self.step
        p["strategies"] = [s.name for s in self.strategies]
        if self.producer is not None:
            self.producer.produce({k: v for k, v in p.items()})
        self.step += 1
        return p


# ---------------------------------------------------------------------------------------------------------------------
# Exact formulations (mixed-integer programs solved with scipy's HiGHS interface; the reference uses pyscipopt / pulp)
# (ref: examples/hydraulis/strategy/dynamic_scip.py:9-85 dynamic_strategy, :88-170 solve_v_micro_batches / batching_strategy)
def dispatch_batch_ilp(seq_lens: Sequence[int], strategies: Sequence[StrategyCost], pipeline_stages: Sequence[int] = (),
                       time_limit: float = 5.0) -> Dict:
    """Assign every sequence to exactly one of the CONCURRENTLY running replicas `strategies` (heterogeneous data-parallel
    pipelines: one StrategyCost per pipeline, dp ignored) so that the slowest pipeline finishes as early as possible:

        min Z   s.t.  sum_j m[i,j] = 1,   m[i,j] = 0 when seq i does not fit pipeline j,
                      Z >= sum_i m[i,j] * cost_j(s_i) + (pp_j - 1) * cost_j(longest admissible seq of j)     for all j

    The second term is the 1F1B fill / drain bubble of a pp_j-stage pipeline, as in the reference.  Returns the same
    structure as `dispatch_batch` (+ "status")."""
    import numpy as np
    from scipy.optimize import Bounds, LinearConstraint, milp
    n, S = len(seq_lens), len(strategies)
    pps = list(pipeline_stages) or [1] * S
    pairs = [(i, j) for i in range(n) for j in range(S) if seq_lens[i] <= strategies[j].max_seq]
    if {i for i, _ in pairs} != set(range(n)):
        raise ValueError("a sequence fits no strategy")
    nv = len(pairs) + 1                                   # binaries + Z
    c = np.zeros(nv); c[-1] = 1.0
    A_eq = np.zeros((n, nv))
    A_le = np.zeros((S, nv))
    for k, (i, j) in enumerate(pairs):
        A_eq[i, k] = 1.0
        A_le[j, k] = strategies[j].seq_ms(seq_lens[i])
    A_le[:, -1] = -1.0
    bubble = np.array([(pps[j] - 1) * strategies[j].seq_ms(min(max(seq_lens), strategies[j].max_seq)) for j in range(S)])
    cons = [LinearConstraint(A_eq, 1.0, 1.0), LinearConstraint(A_le, -np.inf, -bubble)]
    integrality = np.ones(nv); integrality[-1] = 0
    res = milp(c, constraints=cons, integrality=integrality, bounds=Bounds(np.zeros(nv), np.append(np.ones(nv - 1), np.inf)),
               options={"time_limit": float(time_limit), "disp": False})
    if res.x is None:
        out = dispatch_batch(seq_lens, strategies, sequential=False)
        out["status"] = f"ilp failed ({res.message}); heuristic"
        return out
    assign = [-1] * n
    for k, (i, j) in enumerate(pairs):
        if res.x[k] > 0.5:
            assign[i] = j
    per = []
    for j, st in enumerate(strategies):
        idx = [i for i in range(n) if assign[i] == j]
        per.append({"indices": idx, "replicas": [idx], "ms": sum(st.seq_ms(seq_lens[i]) for i in idx) + float(bubble[j])})
    return {"assignment": assign, "per_strategy": per, "makespan_ms": max(p["ms"] for p in per), "status": res.message}


def solve_micro_batches_ilp(seqs: Sequence[int], costs: Sequence[float], max_tokens: int, min_tokens: int, v: int,
                            time_limit: float = 5.0):
    """pack the sequences of one pipeline into `v` micro-batches (packed rows of at most `max_tokens`, at least `min_tokens`
    tokens so the tensor cores stay utilised) minimising the slowest micro-batch -> (max cost, [micro-batch of each seq]) or
    (inf, None) when infeasible"""
    import numpy as np
    from scipy.optimize import Bounds, LinearConstraint, milp
    u = len(seqs)
    nv = u * v + 1
    c = np.zeros(nv); c[-1] = 1.0
    one = np.zeros((u, nv))
    cost_rows = np.zeros((v, nv))
    tok_rows = np.zeros((v, nv))
    for i in range(u):
        for j in range(v):
            one[i, i * v + j] = 1.0
            cost_rows[j, i * v + j] = costs[i]
            tok_rows[j, i * v + j] = seqs[i]
    cost_rows[:, -1] = -1.0
    cons = [LinearConstraint(one, 1.0, 1.0), LinearConstraint(cost_rows, -np.inf, 0.0), LinearConstraint(tok_rows, float(min_tokens), float(max_tokens))]
    integrality = np.ones(nv); integrality[-1] = 0
    res = milp(c, constraints=cons, integrality=integrality, bounds=Bounds(np.zeros(nv), np.append(np.ones(nv - 1), np.inf)),
               options={"time_limit": float(time_limit), "disp": False})
    if res.x is None:
        return float("inf"), None
    where = [int(max(range(v), key=lambda j: res.x[i * v + j])) for i in range(u)]
    return float(res.x[-1]), where


def batching_strategy_ilp(seqs: Sequence[int], strategy: StrategyCost, pp: int, max_tokens: int, min_tokens: int,
                          per_micro_batch_overhead_ms: float = 0.0, time_limit: float = 5.0) -> Dict:
    """choose the number of micro-batches v and the packing that minimise the pipeline's end-to-end time
    (max micro-batch cost + overhead) * (pp - 1 + v)   (ref: dynamic_scip.py:118 batching_strategy)"""
    total = sum(seqs)
    v_lo = max((total + max_tokens - 1) // max_tokens, 1)
    v_hi = max(total // max(min_tokens, 1), v_lo) if min_tokens > 0 else len(seqs)
    costs = [strategy.seq_ms(s) for s in seqs]
    best = None
    for v in range(v_lo, min(v_hi, len(seqs)) + 1):
        mx, where = solve_micro_batches_ilp(seqs, costs, max_tokens, min_tokens, v, time_limit)
        if where is None:
            continue
        e2e = (mx + per_micro_batch_overhead_ms) * (pp - 1 + v)
        if best is None or e2e < best["e2e_ms"]:
            best = {"num_micro_batches": v, "micro_batch_of_seq": where, "max_micro_batch_ms": mx, "e2e_ms": e2e,
                    "micro_batches": [[i for i, w in enumerate(where) if w == j] for j in range(v)]}
    if best is None:
        raise ValueError("no packing satisfies the token bounds (min_tokens too high for this batch?)")
    return best
