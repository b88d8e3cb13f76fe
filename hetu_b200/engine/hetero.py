"""Execution of heterogeneous strategies (Malleus / Ampelos unions: pipelines with different tensor-parallel degrees, stage
counts and layer splits training ONE model on unequal shares of the global batch).

Design (B200-first, not the reference's union-aware executor): every rank instantiates the *member-local* graph of its own
pipeline -- an ordinary homogeneous (dp = 1, tp_m, pp_m) strategy, so all fused kernels / the 1F1B scheduler are reused
unchanged -- and the only cross-pipeline traffic is the parameter-gradient synchronisation, which the optimizer lowers to a
`grouped_all_reduce` per parameter: the finest common shard (lcm of the tp degrees) of a split parameter is all-reduced among
its holders in every pipeline; replicated parameters are all-reduced among one leader per pipeline and broadcast inside the
pipeline's tp group.  Per-pipeline gradient weights (tokens of the pipeline / tokens of the step) make the result equal to
the single-device gradient of the global mean loss.
(ref: hetu/graph/distributed_states.h DistributedStatesUnion + hetero_dim, hetu/graph/ops/Communication.cc SplitAllReduce,
 python/hetu/engine/strategy.py / strategy_ampelos.py for where such strategies come from)
"""
from __future__ import annotations

from typing import List, Optional, Sequence

from ..models.parallel_config import localize_hetero_config
from ..nn.parallel import HETERO_PARAMS, precreate_hetero_groups


def _pipelines_of(cfg: dict) -> List[List[List[int]]]:
    """pipelines -> ordered list of stage device groups, recovered from the per-layer unions"""
    blocks = sorted(cfg["blocks"].values(), key=lambda b: b["range"][0])
    n = len(cfg["input"]["device_group_union"])
    pipes: List[List[List[int]]] = [[] for _ in range(n)]
    for blk in blocks:
        for m, devs in enumerate(blk["layernorm1"]["device_group_union"]):
            if not pipes[m] or pipes[m][-1] != list(devs):
                pipes[m].append(list(devs))
    return pipes


class HeteroSession:
    """One rank's view of a heterogeneous strategy: the local homogeneous config to build the model with, the share of
    the global batch its pipeline trains on, the gradient weight and the loss reduction."""

    def __init__(self, hetero_cfg: dict, rank: Optional[int] = None, shares: Optional[Sequence[int]] = None):
        from .. import distributed
        self.cfg = hetero_cfg
        self.rank = distributed.rank() if rank is None else rank
        self.pipelines = _pipelines_of(hetero_cfg)
        self.num_pipelines = len(self.pipelines)
        # a rank the plan leaves without work (e.g. the other half of a tensor-parallel group that lost a device) is idle:
        # it owns no graph but still takes part in the collective creation of process groups and in world barriers
        self.pipeline = next((i for i, p in enumerate(self.pipelines) if any(self.rank in st for st in p)), None)
        self.idle = self.pipeline is None
        self.first_stage_ranks = [p[0][0] for p in self.pipelines]
        self.last_stage_ranks = [p[-1][0] for p in self.pipelines]
        self.shares = list(shares) if shares is not None else [1] * self.num_pipelines
        HETERO_PARAMS.clear()
        self.local_cfg = None if self.idle else localize_hetero_config(hetero_cfg, self.rank)

    # -- batch split ----------------------------------------------------------------------------------------------
    def split_batch(self, global_batch: int) -> List[int]:
        """sequences per pipeline, proportional to `shares` (largest remainder; every pipeline gets at least one)"""
        tot = float(sum(self.shares))
        raw = [global_batch * s / tot for s in self.shares]
        out = [max(1, int(r)) for r in raw]
        order = sorted(range(len(raw)), key=lambda i: -(raw[i] - int(raw[i])))
        i = 0
        while sum(out) < global_batch:
            out[order[i % len(out)]] += 1
            i += 1
        while sum(out) > global_batch:
            j = max(range(len(out)), key=lambda k: out[k])
            out[j] -= 1
        return out

    def batch_slice(self, global_batch: int) -> slice:
        per = self.split_batch(global_batch)
        if self.idle:
            return slice(0, 0)
        lo = sum(per[:self.pipeline])
        return slice(lo, lo + per[self.pipeline])

    def grad_scale(self, local_tokens: int, global_tokens: int) -> float:
        return float(local_tokens) / float(global_tokens)

    # -- collectives ----------------------------------------------------------------------------------------------
    def precreate_groups(self, extra=()):
        """collective over ALL ranks; call once after the model (and optimizer) graph has been built"""
        return precreate_hetero_groups(self.cfg, extra=[self.last_stage_ranks] + [list(e) for e in extra])

    def reduce_loss(self, local_loss, local_tokens: int, global_tokens: int):
        """global mean loss from the per-pipeline mean losses (valid on the last-stage leader of every pipeline)"""
        import torch
        from .. import _C
        if self.rank not in self.last_stage_ranks:
            return None
        t = torch.as_tensor(float(local_loss)).reshape(1) * (float(local_tokens) / float(global_tokens))
        if len(self.last_stage_ranks) > 1 and _C.comm_initialized():
            t = _C.comm_all_reduce(t, self.last_stage_ranks, "sum")
        return float(t[0])
