"""LobRA: multi-tenant LoRA fine-tuning over heterogeneous replicas.

Several fine-tuning tasks (one LoRA adapter each, `hetu_b200.peft` multi-LoRA) share one frozen base model.  Their
sequence-length distributions differ wildly, so instead of one (tp, pp) strategy for everything the job deploys a *mix* of
replicas -- small cheap ones for the many short sequences, large ones (more tensor/pipeline parallelism => more tokens per
micro-batch) for the few long ones -- and dispatches every step's sequences to the replicas so that all finish together.

* `LoraCostModel`        time of one micro-batch (mbs sequences of length s) on a (tp, pp) replica, fitted from profiles
* `optimized_scheme_pool`   prune dominated (tp, pp) candidates
* `GroupStaticPlanner` / `BalanceStaticPlanner` / `PruneStaticPlanner`   choose how many replicas of each scheme to deploy on
  N GPUs for the expected multi-task length distribution
* `GroupDynamicDispatcher` / `BalanceDynamicDispatcher`   per step: assign the sampled sequences to the deployed replicas
* `global_batch_scheduler` / `greedy_local_batch_scheduler` / `local_batch_pack_scheduler`   turn a dispatch into per-replica
  micro-batches (padded per bucket or packed up to max_tokens)

The reference formulates planning as a non-linear integer program for SCIP; here replica counts are enumerated (the search
space is tiny: compositions of N GPUs over <= a dozen schemes) and each candidate's dispatch is a small mixed-integer linear
program solved with scipy's HiGHS (`scipy.optimize.milp`), with an exact re-evaluation (micro-batch fragments, pipeline
bubbles) of the rounded solution.
(ref: examples/lobra/trainer/planner.py:10-1525, batch_scheduler.py:7-400, profiler/cost_model.py)
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np


# ---------------------------------------------------------------------------------------------------------------- schemes
@dataclass(frozen=True)
class Scheme:
    tp: int
    pp: int
    max_tokens: int                    # tokens one micro-batch may hold on this replica (memory bound)
    throughput_per_gpu: float = 0.0    # profiled tokens/s/GPU (only used to prune the pool)

    @property
    def ngpus(self) -> int:
        return self.tp * self.pp


def optimized_scheme_pool(candidates: Sequence[dict]) -> List[Scheme]:
    """keep, per distinct max_tokens, the highest-throughput scheme plus every cheaper (fewer GPUs) runner-up -- the other
    candidates are dominated (ref: planner.py:39-94)"""
    best: Dict[int, Scheme] = {}
    per_gpus: Dict[int, Dict[int, Scheme]] = {}
    for c in candidates:
        mt = c.get("max_tokens", c.get("mbs", 0) * c.get("seq_len", 0))
        if mt <= 0:
            continue
        s = Scheme(int(c["tp"]), int(c["pp"]), int(mt), float(c.get("throughput_per_gpu", 0.0)))
        g = per_gpus.setdefault(mt, {})
        if s.ngpus not in g or s.throughput_per_gpu > g[s.ngpus].throughput_per_gpu:
            g[s.ngpus] = s
        b = best.get(mt)
        if b is None or s.throughput_per_gpu > b.throughput_per_gpu or \
                (s.throughput_per_gpu == b.throughput_per_gpu and s.ngpus < b.ngpus):
            best[mt] = s
    pool = list(best.values())
    for mt, b in best.items():
        for n in sorted(per_gpus[mt]):
            if n < b.ngpus and per_gpus[mt][n] not in pool:
                pool.append(per_gpus[mt][n])
    return sorted(pool, key=lambda s: (s.max_tokens, s.ngpus))


# ------------------------------------------------------------------------------------------------------------- cost model
class LoraCostModel:
    """ms of one layer's forward+backward for a micro-batch of `mbs` sequences of length `s` on tensor-parallel degree tp:
        (c1*mbs + c2*f) s^2 + (c3*mbs + c4*f) s + (c5*mbs + c6*f),   f = 1 if the micro-batch is non-empty else 0
    (attention is quadratic, GEMMs linear, launch / collective latency constant).  A replica's micro-batch time is that times
    layers / pp.  (ref: planner.py:106-127, profiler/cost_model.py)"""

    def __init__(self, popt: Dict[int, Sequence[float]], max_tokens: Optional[Dict[Tuple[int, int], int]] = None):
        self.popt = {int(tp): tuple(float(c) for c in cs) for tp, cs in popt.items()}
        self.max_tokens = dict(max_tokens or {})

    @staticmethod
    def features(mbs: float, s: float, frag: float = 1.0) -> List[float]:
        return [mbs * s * s, frag * s * s, mbs * s, frag * s, mbs, frag]

    @classmethod
    def fit(cls, records: Sequence[Tuple[int, int, int, float]], max_tokens=None) -> "LoraCostModel":
        """records: (tp, mbs, seq_len, ms per layer) -> non-negative least squares per tp"""
        from scipy.optimize import nnls
        popt = {}
        for tp in sorted({r[0] for r in records}):
            rows = [r for r in records if r[0] == tp]
            A = np.array([cls.features(m, s) for _, m, s, _ in rows], dtype=np.float64)
            y = np.array([t for *_, t in rows], dtype=np.float64)
            scale = A.max(axis=0)
            scale[scale == 0] = 1.0
            coef, _ = nnls(A / scale, y)
            popt[tp] = tuple(coef / scale)
        return cls(popt, max_tokens)

    @classmethod
    def analytic(cls, hidden: int, ffn: int, tps: Sequence[int] = (1, 2, 4, 8), tflops: float = 900.0, launch_ms: float = 0.05,
                 comm_gbps: float = 600.0) -> "LoraCostModel":
        """roofline-style stand-in when no profile is available (fwd+bwd of a frozen-base LoRA layer ~ 2x forward GEMMs + attention)"""
        popt = {}
        for tp in tps:
            gemm = 2.0 * 2.0 * (4 * hidden * hidden + 3 * hidden * ffn) / tp / (tflops * 1e9)          # ms per token
            attn = 2.0 * 4.0 * hidden / tp / (tflops * 0.35 * 1e9)                                       # ms per token^2
            comm = 0.0 if tp == 1 else 4.0 * 2.0 * hidden * 2 * (tp - 1) / tp / (comm_gbps * 1e6)        # ms per token
            popt[tp] = (attn, 0.0, gemm + comm, 0.0, 0.0, launch_ms * (1 if tp == 1 else 2))
        return cls(popt)

    def layer_time(self, mbs: float, s: float, tp: int, frag: float = 1.0) -> float:
        if mbs <= 0:
            return 0.0
        return float(np.dot(self.popt[tp], self.features(mbs, s, frag)))

    def batch_time(self, mbs: float, s: float, scheme: Scheme, num_layers: int) -> float:
        return self.layer_time(mbs, s, scheme.tp) * num_layers / scheme.pp


# ------------------------------------------------------------------------------------------------------------------ plans
@dataclass
class LobraPlan:
    schemes: List[Scheme]
    dp: List[int]                                   # replicas per scheme
    dispatch: List[Dict[int, int]]                  # per scheme: seq_len bucket -> sequences (summed over its replicas)
    scheme_times: List[float]                       # estimated ms per scheme (max over its replicas)
    task_dispatch: Optional[List[List[Dict[int, int]]]] = None   # [task][scheme] bucket -> sequences

    @property
    def time(self) -> float:
        return max([t for t, d in zip(self.scheme_times, self.dp) if d > 0] or [0.0])

    @property
    def gpus(self) -> int:
        return sum(d * s.ngpus for d, s in zip(self.dp, self.schemes))

    def strategy(self) -> List[Tuple[int, int, int]]:
        """[(dp, tp, pp)] of the deployed replica groups"""
        return [(d, s.tp, s.pp) for d, s in zip(self.dp, self.schemes) if d > 0]

    def pipelines(self, num_layers: int, first_device: int = 0) -> List[dict]:
        """the heterogeneous pipeline list for generate_hetero_ds_parallel_config / engine.hetero.HeteroSession"""
        out, dev = [], first_device
        for d, s in zip(self.dp, self.schemes):
            for _ in range(d):
                base, rem = divmod(num_layers, s.pp)
                stages, lo = [], 0
                for st in range(s.pp):
                    n = base + (1 if st < rem else 0)
                    stages.append({"devices": list(range(dev, dev + s.tp)), "layers": [lo, lo + n - 1]})
                    dev += s.tp
                    lo += n
                out.append({"stages": stages})
        return out


class _PlannerCore:
    def __init__(self, cost_model: LoraCostModel, num_layers: int, schemes: Sequence[Scheme]):
        self.cm, self.num_layers = cost_model, num_layers
        self.schemes = list(schemes)

    # ---- exact evaluation of one scheme
    def replica_time(self, scheme: Scheme, per_replica: Dict[int, int]) -> float:
        """micro-batches of one replica: per bucket full micro-batches of mbs = max_tokens // s plus one fragment; 1F1B adds
        (pp - 1) x the slowest micro-batch"""
        total, worst = 0.0, 0.0
        for s, n in per_replica.items():
            if n <= 0:
                continue
            mbs = scheme.max_tokens // s
            if mbs <= 0:
                return math.inf
            m, r = divmod(n, mbs)
            tf = self.cm.batch_time(mbs, s, scheme, self.num_layers) if m else 0.0
            tr = self.cm.batch_time(r, s, scheme, self.num_layers) if r else 0.0
            total += m * tf + tr
            worst = max(worst, tf, tr)
        return total + (scheme.pp - 1) * worst

    def scheme_time(self, j: int, dp: int, counts: Dict[int, int]) -> float:
        if dp == 0:
            return 0.0 if not any(counts.values()) else math.inf
        return self.replica_time(self.schemes[j], {s: -(-n // dp) for s, n in counts.items()})

    # ---- dispatch for fixed replica counts
    def allowed(self, seq_lens: Sequence[int], mode: str) -> List[List[bool]]:
        """Balance: any scheme whose micro-batch holds the sequence.  Group: only the schemes with the SMALLEST sufficient
        max_tokens (every length bucket is owned by one group of identical-capacity replicas)."""
        out = []
        for s in seq_lens:
            fits = [sc.max_tokens >= s for sc in self.schemes]
            if mode == "group" and any(fits):
                smallest = min(sc.max_tokens for sc, f in zip(self.schemes, fits) if f)
                fits = [f and sc.max_tokens == smallest for sc, f in zip(self.schemes, fits)]
            out.append(fits)
        return out

    def dispatch(self, dp: Sequence[int], seq_counts: Dict[int, int], mode: str = "balance", use_milp: bool = True) -> Optional[LobraPlan]:
        seq_lens = sorted(s for s, n in seq_counts.items() if n > 0)
        J = len(self.schemes)
        ok = self.allowed(seq_lens, mode)
        live = [j for j in range(J) if dp[j] > 0]
        for i, s in enumerate(seq_lens):
            if not any(ok[i][j] for j in live):
                return None
        # per-sequence cost (ms) of bucket i on ONE replica of scheme j, full micro-batches
        t = np.full((len(seq_lens), J), np.inf)
        for i, s in enumerate(seq_lens):
            for j in live:
                if ok[i][j]:
                    mbs = self.schemes[j].max_tokens // s
                    t[i, j] = self.cm.batch_time(mbs, s, self.schemes[j], self.num_layers) / mbs
        n = self._solve(seq_lens, seq_counts, dp, live, t, use_milp)
        disp = [{s: int(n[i, j]) for i, s in enumerate(seq_lens) if n[i, j] > 0} for j in range(J)]
        times = [self.scheme_time(j, dp[j], disp[j]) for j in range(J)]
        plan = LobraPlan(self.schemes, list(dp), disp, times)
        self._polish(plan, seq_lens, ok)
        return plan

    def _solve(self, seq_lens, seq_counts, dp, live, t, use_milp):
        I, J = len(seq_lens), len(self.schemes)
        n = np.zeros((I, J), dtype=np.int64)
        pairs = [(i, j) for i in range(I) for j in live if np.isfinite(t[i, j])]
        solved = False
        if use_milp and pairs:
            try:
                from scipy.optimize import Bounds, LinearConstraint, milp
                nv = len(pairs) + 1                                 # + T
                c = np.zeros(nv)
                c[-1] = 1.0
                A, lb, ub = [], [], []
                for i in range(I):                                   # every sequence is dispatched
                    row = np.zeros(nv)
                    for k, (pi, _) in enumerate(pairs):
                        if pi == i:
                            row[k] = 1.0
                    A.append(row); lb.append(seq_counts[seq_lens[i]]); ub.append(seq_counts[seq_lens[i]])
                for j in live:                                       # per-replica load <= T
                    row = np.zeros(nv)
                    for k, (pi, pj) in enumerate(pairs):
                        if pj == j:
                            row[k] = t[pi, pj] / dp[j]
                    row[-1] = -1.0
                    A.append(row); lb.append(-np.inf); ub.append(0.0)
                integrality = np.ones(nv)
                integrality[-1] = 0
                res = milp(c, constraints=LinearConstraint(np.array(A), lb, ub), integrality=integrality,
                           bounds=Bounds(np.zeros(nv), np.full(nv, np.inf)), options={"time_limit": 5.0, "mip_rel_gap": 1e-3})
                if res.x is not None:
                    for k, (i, j) in enumerate(pairs):
                        n[i, j] = int(round(res.x[k]))
                    for i in range(I):                               # repair rounding drift
                        diff = seq_counts[seq_lens[i]] - int(n[i].sum())
                        if diff:
                            j = max(live, key=lambda q: n[i, q]) if diff < 0 else min((q for q in live if np.isfinite(t[i, q])), key=lambda q: t[i, q])
                            n[i, j] += diff
                    solved = True
            except Exception:                                        # solver unavailable / timed out -> greedy
                solved = False
        if not solved:
            load = np.zeros(J)
            for i in sorted(range(I), key=lambda i: -seq_lens[i]):   # longest first: least slack
                for _ in range(seq_counts[seq_lens[i]]):
                    j = min((q for q in live if np.isfinite(t[i, q])), key=lambda q: load[q] + t[i, q] / dp[q])
                    n[i, j] += 1
                    load[j] += t[i, j] / dp[j]
        return n

    def _polish(self, plan: LobraPlan, seq_lens, ok, rounds: int = 64):
        """local search on the exact objective: move one replica-row of sequences from the slowest scheme to the scheme where
        it hurts least, while that lowers the makespan"""
        J = len(self.schemes)
        for _ in range(rounds):
            j = max(range(J), key=lambda q: plan.scheme_times[q] if plan.dp[q] else -1.0)
            best = None
            for i, s in enumerate(seq_lens):
                have = plan.dispatch[j].get(s, 0)
                if not have:
                    continue
                step = min(have, plan.dp[j])
                for q in range(J):
                    if q == j or not plan.dp[q] or not ok[i][q]:
                        continue
                    dj = dict(plan.dispatch[j]); dq = dict(plan.dispatch[q])
                    dj[s] = have - step
                    dq[s] = dq.get(s, 0) + step
                    tj, tq = self.scheme_time(j, plan.dp[j], dj), self.scheme_time(q, plan.dp[q], dq)
                    others = max([plan.scheme_times[k] for k in range(J) if k not in (j, q) and plan.dp[k]] or [0.0])
                    new = max(tj, tq, others)
                    if new < plan.time - 1e-9 and (best is None or new < best[0]):
                        best = (new, s, q, dj, dq, tj, tq)
            if best is None:
                return
            _, s, q, dj, dq, tj, tq = best
            plan.dispatch[j], plan.dispatch[q] = {k: v for k, v in dj.items() if v > 0}, dq
            plan.scheme_times[j], plan.scheme_times[q] = tj, tq

    # ---- enumeration of replica counts
    def replica_counts(self, ngpus: int, exact: bool = True):
        sizes = [s.ngpus for s in self.schemes]

        def rec(j, left):
            if j == len(sizes):
                if left == 0 or not exact:
                    yield []
                return
            for d in range(left // sizes[j] + 1):
                for rest in rec(j + 1, left - d * sizes[j]):
                    yield [d] + rest
        yield from rec(0, ngpus)

    @staticmethod
    def merge_tasks(multi_task_seq_distribution: Sequence[Dict[int, int]]) -> Dict[int, int]:
        tot: Dict[int, int] = {}
        for d in multi_task_seq_distribution:
            for s, n in d.items():
                tot[int(s)] = tot.get(int(s), 0) + int(n)
        return tot

    @staticmethod
    def split_tasks(plan: LobraPlan, multi_task_seq_distribution: Sequence[Dict[int, int]]):
        """attribute each scheme's share of a bucket to the tasks (first fit in task order; the shares add up exactly)"""
        J = len(plan.schemes)
        out = [[{} for _ in range(J)] for _ in multi_task_seq_distribution]
        for s in {s for d in plan.dispatch for s in d}:
            left = [plan.dispatch[j].get(s, 0) for j in range(J)]
            for ti, dist in enumerate(multi_task_seq_distribution):
                need = int(dist.get(s, 0))
                for j in range(J):
                    take = min(need, left[j])
                    if take:
                        out[ti][j][s] = take
                        left[j] -= take
                        need -= take
        plan.task_dispatch = out
        return out


class _StaticPlanner(_PlannerCore):
    """(ref: planner.py BaseStaticPlanner: schedule(multi_task_seq_distribution) -> deployed replica mix + dispatch)"""
    mode = "balance"

    def __init__(self, cost_model: LoraCostModel, num_layers: int, train_task_num: int, global_batch_size_list: Sequence[int], ngpus: int,
                 scheme_candidates: Sequence, use_optimized_scheme_pool: bool = True, max_candidates: int = 20000):
        pool = [c for c in scheme_candidates if isinstance(c, Scheme)] or \
            (optimized_scheme_pool(scheme_candidates) if use_optimized_scheme_pool else
             [Scheme(c["tp"], c["pp"], c["max_tokens"], c.get("throughput_per_gpu", 0.0)) for c in scheme_candidates if c.get("max_tokens", 0) > 0])
        super().__init__(cost_model, num_layers, pool)
        self.train_task_num, self.global_batch_size_list, self.ngpus = train_task_num, list(global_batch_size_list), ngpus
        self.max_candidates = max_candidates
        self.evaluated = 0

    def candidates(self, counts):
        longest = max(counts)
        for k, dp in enumerate(self.replica_counts(self.ngpus)):
            if k >= self.max_candidates:
                return
            if any(d and s.max_tokens >= longest for d, s in zip(dp, self.schemes)):
                yield dp

    def schedule(self, multi_task_seq_distribution: Sequence[Dict[int, int]]) -> LobraPlan:
        counts = self.merge_tasks(multi_task_seq_distribution)
        best = None
        self.evaluated = 0
        for dp in self.candidates(counts):
            plan = self.dispatch(dp, counts, self.mode, use_milp=False)      # cheap screen ...
            self.evaluated += 1
            if plan is not None and (best is None or plan.time < best.time):
                best = plan
        assert best is not None, "no replica mix can hold the longest sequence"
        refined = self.dispatch(best.dp, counts, self.mode, use_milp=True)   # ... exact dispatch for the winner
        if refined is not None and refined.time < best.time:
            best = refined
        self.split_tasks(best, multi_task_seq_distribution)
        return best


class GroupStaticPlanner(_StaticPlanner):
    """every length bucket is served only by the replicas with the smallest sufficient capacity (ref: planner.py:211-416)"""
    mode = "group"


class BalanceStaticPlanner(_StaticPlanner):
    """a bucket may be spread over every replica that can hold it (ref: planner.py:418-627)"""
    mode = "balance"


class PruneStaticPlanner(_StaticPlanner):
    """Balance planning over a pruned candidate set: a candidate's *group* dispatch time upper-bounds its balanced time and
    the perfectly divisible load lower-bounds it; candidates whose lower bound exceeds the best upper bound seen are
    skipped before the (more expensive) balanced dispatch is solved (ref: planner.py:629-972)"""
    mode = "balance"

    def lower_bound(self, dp, counts) -> float:
        """fluid relaxation per capacity class: the sequences too long for every scheme below capacity v can only run on the
        replicas with max_tokens >= v -- at best at their cheapest per-sequence cost, perfectly divisible over those replicas"""
        caps = sorted({sc.max_tokens for sc in self.schemes})
        bound = 0.0
        for k, v in enumerate(caps):
            below = caps[k - 1] if k else 0
            live = [sc for d, sc in zip(dp, self.schemes) if d and sc.max_tokens >= v]
            replicas = sum(d for d, sc in zip(dp, self.schemes) if sc.max_tokens >= v)
            work = 0.0
            for s, n in counts.items():
                if s <= below or not n:
                    continue
                costs = [self.cm.batch_time(sc.max_tokens // s, s, sc, self.num_layers) / (sc.max_tokens // s) for sc in live if sc.max_tokens >= s]
                if not costs:
                    return math.inf
                work += n * min(costs)
            if work > 0:
                bound = max(bound, work / max(replicas, 1))
        return bound

    def schedule(self, multi_task_seq_distribution):
        counts = self.merge_tasks(multi_task_seq_distribution)
        cands = list(self.candidates(counts))
        bounds = sorted(((self.lower_bound(dp, counts), dp) for dp in cands), key=lambda x: x[0])
        best, self.evaluated, self.pruned = None, 0, 0
        for lb, dp in bounds:
            if best is not None and lb >= best.time:
                self.pruned += 1
                continue
            plan = self.dispatch(dp, counts, "group", use_milp=False)        # upper bound from the restricted problem
            bal = self.dispatch(dp, counts, "balance", use_milp=False)
            self.evaluated += 1
            for p in (plan, bal):
                if p is not None and (best is None or p.time < best.time):
                    best = p
        assert best is not None, "no replica mix can hold the longest sequence"
        refined = self.dispatch(best.dp, counts, "balance", use_milp=True)
        if refined is not None and refined.time < best.time:
            best = refined
        self.split_tasks(best, multi_task_seq_distribution)
        return best


class _DynamicDispatcher(_PlannerCore):
    """the replica mix is deployed; each step's sampled sequences are dispatched to it (ref: planner.py:974-1263)"""
    mode = "balance"

    def __init__(self, cost_model: LoraCostModel, num_layers: int, strategy: Sequence[Tuple[int, int, int]], max_tokens_list: Sequence[int],
                 train_task_num: int = 1):
        schemes = [Scheme(tp, pp, mt) for (_, tp, pp), mt in zip(strategy, max_tokens_list)]
        super().__init__(cost_model, num_layers, schemes)
        self.dp = [d for d, _, _ in strategy]
        self.train_task_num = train_task_num

    def schedule(self, multi_task_seq_distribution: Sequence[Dict[int, int]]) -> LobraPlan:
        counts = self.merge_tasks(multi_task_seq_distribution)
        plan = self.dispatch(self.dp, counts, self.mode, use_milp=True)
        assert plan is not None, "a sampled sequence does not fit any deployed replica"
        self.split_tasks(plan, multi_task_seq_distribution)
        return plan


class GroupDynamicDispatcher(_DynamicDispatcher):
    mode = "group"


class BalanceDynamicDispatcher(_DynamicDispatcher):
    mode = "balance"


# -------------------------------------------------------------------------------------------------------- batch schedulers
@dataclass
class MicroBatch:
    """(ref: batch_scheduler.py:7-34)"""
    batch_data: Optional[np.ndarray]
    batch_size: int
    seq_length: int
    batch_offset_list: List[int] = field(default_factory=list)     # per task: first row of the task inside the micro-batch
    batch_size_list: List[int] = field(default_factory=list)       # per task: rows of the task (multi-LoRA task routing)

    def token_num(self) -> int:
        return self.batch_size * self.seq_length

    def task_id(self) -> List[int]:
        return [i for i, b in enumerate(self.batch_size_list) if b > 0]


def bucket_of(seq_len: int, buckets: Sequence[int]) -> int:
    """smallest bucket boundary that holds the sequence"""
    for b in sorted(buckets):
        if seq_len <= b:
            return b
    raise ValueError(f"sequence of length {seq_len} exceeds the largest bucket {max(buckets)}")


def seq_distribution(task_batches: Sequence[Sequence[Sequence[int]]], buckets: Sequence[int]) -> List[Dict[int, int]]:
    """per task: bucket -> number of sequences in this step's global batch"""
    out = []
    for seqs in task_batches:
        d: Dict[int, int] = {}
        for s in seqs:
            b = bucket_of(len(s), buckets)
            d[b] = d.get(b, 0) + 1
        out.append(d)
    return out


def global_batch_scheduler(task_batches: Sequence[Sequence[Sequence[int]]], plan: LobraPlan, buckets: Sequence[int]):
    """split the step's multi-task global batch over the deployed replicas following `plan.task_dispatch`:
    -> [scheme][replica] list of (task, token list, bucket).  Sequences of a bucket go round-robin over a scheme's replicas
    (ref: batch_scheduler.py:239-400)"""
    assert plan.task_dispatch is not None
    J = len(plan.schemes)
    out = [[[] for _ in range(max(d, 0))] for d in plan.dp]
    rr = [0] * J
    for ti, seqs in enumerate(task_batches):
        quota = [dict(plan.task_dispatch[ti][j]) for j in range(J)]
        for s in sorted(seqs, key=len, reverse=True):
            b = bucket_of(len(s), buckets)
            j = next((q for q in range(J) if quota[q].get(b, 0) > 0), None)
            assert j is not None, f"dispatch plan has no room left for a task-{ti} sequence of bucket {b}"
            quota[j][b] -= 1
            out[j][rr[j] % plan.dp[j]].append((ti, list(s), b))
            rr[j] += 1
    return out


def greedy_local_batch_scheduler(rows: Sequence[Tuple[int, List[int], int]], max_tokens: int, train_task_num: int, pad_id: int = 0) -> List[MicroBatch]:
    """padded micro-batches of one replica: per bucket, rows sorted by task so a micro-batch holds contiguous task ranges
    (multi-LoRA applies adapter t to rows [offset_t, offset_t + size_t)) (ref: batch_scheduler.py:138-237)"""
    out: List[MicroBatch] = []
    for b in sorted({r[2] for r in rows}):
        mbs = max(max_tokens // b, 1)
        group = sorted((r for r in rows if r[2] == b), key=lambda r: r[0])
        for k in range(0, len(group), mbs):
            chunk = group[k:k + mbs]
            data = np.full((len(chunk), b), pad_id, dtype=np.int64)
            sizes, offs = [0] * train_task_num, [0] * train_task_num
            for r, (ti, toks, _) in enumerate(chunk):
                data[r, :len(toks)] = toks
                if sizes[ti] == 0:
                    offs[ti] = r
                sizes[ti] += 1
            out.append(MicroBatch(data, len(chunk), b, offs, sizes))
    return out


def local_batch_pack_scheduler(rows: Sequence[Tuple[int, List[int], int]], max_tokens: int, train_task_num: int, pad_id: int = 0):
    """packed micro-batches of one replica: first-fit-decreasing bins of <= max_tokens real tokens per task-contiguous row;
    returns (MicroBatch with one packed row, cu_seqlens) pairs (ref: batch_scheduler.py:84-136)"""
    bins: List[List[Tuple[int, List[int]]]] = []
    loads: List[int] = []
    for ti, toks, _ in sorted(rows, key=lambda r: -len(r[1])):
        assert len(toks) <= max_tokens, "sequence longer than the replica's micro-batch capacity"
        k = next((k for k in range(len(bins)) if loads[k] + len(toks) <= max_tokens), None)
        if k is None:
            bins.append([]); loads.append(0)
            k = len(bins) - 1
        bins[k].append((ti, toks))
        loads[k] += len(toks)
    out = []
    for items in bins:
        items.sort(key=lambda x: x[0])
        data = np.full((1, max_tokens), pad_id, dtype=np.int64)
        cu, pos = [0], 0
        sizes, offs = [0] * train_task_num, [0] * train_task_num
        for ti, toks in items:
            data[0, pos:pos + len(toks)] = toks
            if sizes[ti] == 0:
                offs[ti] = pos
            sizes[ti] += len(toks)
            pos += len(toks)
            cu.append(pos)
        out.append((MicroBatch(data, 1, max_tokens, offs, sizes), cu))
    return out
