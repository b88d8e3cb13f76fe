"""Tensor/sequence-parallel modules driven by `ds_parallel_config` strategy JSON leaves.

One leaf per strategy (multi-strategy graphs for hot switching), each optionally heterogeneous:
    {"split": {"0": [tp]}, "dup": [dp], "device_group_union": [[0,1,2,3]], "type": "variable", "zero": true,
     "recompute": [false], "cpu_offload": [false]}
(ref: python/hetu/nn/modules/parallel_multi_ds.py:328-588, parallel_utils.py:83-116, parallel.py)
"""
from __future__ import annotations

from typing import List, Sequence

from .. import ops
from ..core import (DeviceGroup, DistributedStates, DistributedStatesUnion, normal_initializer, ones_initializer,
                    parallel_parameter, xavier_normal_initializer, zeros_initializer)
from ..distributed import global_device_group
from .module import Module

__all__ = ["config2ds", "get_multi_ds_parallel_config", "HtMultiColumnParallelLinear", "HtMultiRowParallelLinear",
           "HtMultiQKVColumnParallelLinear", "HtMultiParallelEmbedding", "HtMultiVocabParallelEmbedding",
           "HtMultiParallelLayerNorm", "HtMultiParallelRMSNorm", "ColumnParallelLinear", "RowParallelLinear",
           "ParallelEmbedding", "VocabParallelEmbedding", "ParallelLayerNorm", "single_config"]


def _device_group(ids: Sequence[int]) -> DeviceGroup:
    allg = global_device_group()
    kind = "cuda" if str(allg.get(0)).find("cuda") >= 0 else "cpu"
    return DeviceGroup([f"{kind}:{i}" for i in ids])


def config2ds(config: dict):
    """strategy leaf -> (DistributedStatesUnion, [DeviceGroup per hetero member])"""
    hetero_sum = len(config["device_group_union"])
    if config["type"] == "placeholder":
        hetero_dim = 0
    elif config["type"] == "variable":
        hetero_dim = -1
    else:
        raise RuntimeError(f"unsupported type {config['type']}")
    if hetero_sum == 1:
        hetero_dim = -3
    ds_list, dg_list = [], []
    for h in range(hetero_sum):
        n = len(config["device_group_union"][h])
        split = {int(k): (v[h] if isinstance(v, (list, tuple)) else v) for k, v in config["split"].items()}
        dup = config["dup"][h] if isinstance(config["dup"], (list, tuple)) else config["dup"]
        states = {-1: dup, **split}
        if config["type"] == "placeholder":
            order = sorted(split.keys()) + [-1]
            zero = False
        else:
            order = [-1] + sorted(split.keys())
            zero = bool(config.get("zero", False))
        ds_list.append(DistributedStates(n, states, order, zero))
        dg_list.append(_device_group(config["device_group_union"][h]))
    return DistributedStatesUnion(ds_list, hetero_dim), dg_list


def single_config(dp: int, tp: int, devices: Sequence[int], kind="variable", split_dim=0, zero=False) -> dict:
    """one-strategy leaf in the older flat form of the reference (tests/ci_test/ds_parallel_config)"""
    return {"split": {str(split_dim): [tp]} if tp > 1 else {}, "dup": [dp], "device_group_union": [list(devices)], "type": kind,
            "zero": zero, "recompute": [False], "cpu_offload": [False]}


def get_multi_ds_parallel_config(ds_parallel_configs: List[dict], module_name: str, _range: int = -1):
    """find `module_name` (optionally the instance whose 'range' covers block `_range`) in every strategy config"""
    out = []
    for cfg in ds_parallel_configs:
        stack = [cfg]
        found = None
        while stack and found is None:
            cur = stack.pop()
            if not isinstance(cur, dict):
                continue
            for k, v in cur.items():
                if not isinstance(v, dict):
                    continue
                if k == module_name or (module_name in ("layers", "blocks") and k.startswith(module_name)):
                    if _range >= 0 and "range" in v:
                        lo, hi = v["range"][0], v["range"][-1]
                        if not (lo <= _range <= hi):
                            stack.append(v)
                            continue
                    found = v
                    break
                stack.append(v)
        if found is None:
            raise KeyError(f"module {module_name} (range {_range}) not found in ds_parallel_config")
        out.append(found)
    return out


# name of a parameter -> (original heterogeneous leaf, member index, split dim or None): filled by the parallel modules when
# they are built from a localized heterogeneous config, consumed by hetero_grad_sync_spec()
HETERO_PARAMS: dict = {}


def _lcm(a, b):
    import math
    return a * b // math.gcd(a, b)


def hetero_grad_sync_spec(name: str, local_shape, rank: int):
    """slices / rank groups that synchronise the gradient of parameter `name` across the pipelines of a heterogeneous
    strategy -> dict(dim, offsets, lengths, groups, bcast_ranks) or None when nothing has to be exchanged"""
    if name not in HETERO_PARAMS:
        return None
    leaf, m, split_dim = HETERO_PARAMS[name]
    members = [list(d) for d in leaf["device_group_union"]]
    if len(members) < 2:
        return None
    me = members[m]
    if rank not in me:
        return None
    t = me.index(rank)
    if split_dim is None:
        leaders = [d[0] for d in members]
        spec = {"dim": 0, "offsets": [], "lengths": [], "groups": [], "bcast_ranks": me if len(me) > 1 else []}
        if t == 0:
            spec["offsets"], spec["lengths"], spec["groups"] = [0], [int(local_shape[0]) if local_shape else 1], [leaders]
        return spec
    tps = [len(d) for d in members]
    G = 1
    for v in tps:
        G = _lcm(G, v)
    n_local = int(local_shape[split_dim])
    per = G // tps[m]                      # finest pieces held by this rank
    f = n_local // per
    assert f * per == n_local, f"parameter {name}: local extent {n_local} is not divisible into {per} pieces"
    offs, lens, groups = [], [], []
    for j in range(per):
        k = t * per + j
        ranks = [members[q][k // (G // tps[q])] for q in range(len(members))]
        if groups and groups[-1] == ranks and offs[-1] + lens[-1] == j * f:
            lens[-1] += f
        else:
            offs.append(j * f); lens.append(f); groups.append(ranks)
    return {"dim": split_dim, "offsets": offs, "lengths": lens, "groups": groups, "bcast_ranks": []}


def precreate_hetero_groups(hetero_cfg: dict, extra=()):
    """Every rank of a heterogeneous job builds a different (member-local) graph, so the process groups cannot be created
    lazily: ALL ranks create, in one global order, the tensor-parallel group of every pipeline stage, the cross-pipeline
    gradient groups and the caller's `extra` groups."""
    from .. import _C
    groups = []

    def walk(node):
        if isinstance(node, dict):
            if "device_group_union" in node and "type" in node:
                for devs in node["device_group_union"]:
                    if len(devs) > 1 and list(devs) not in groups:
                        groups.append(list(devs))
            else:
                for k in sorted(node):
                    walk(node[k])

    walk(hetero_cfg)
    for r in all_hetero_groups() + [list(e) for e in extra]:
        if len(r) > 1 and r not in groups:
            groups.append(r)
    if _C.comm_initialized():
        for r in groups:
            _C.comm_create_group(r)
    return groups


def all_hetero_groups():
    """every rank group any rank needs for the heterogeneous gradient synchronisation (deterministic order)"""
    seen = []
    for name, (leaf, m, split_dim) in sorted(HETERO_PARAMS.items()):
        members = [list(d) for d in leaf["device_group_union"]]
        if len(members) < 2:
            continue
        if split_dim is None:
            cand = [[d[0] for d in members]] + [d for d in members if len(d) > 1]
        else:
            tps = [len(d) for d in members]
            G = 1
            for v in tps:
                G = _lcm(G, v)
            cand = [[members[q][k // (G // tps[q])] for q in range(len(members))] for k in range(G)]
        for c in cand:
            if c not in seen:
                seen.append(c)
    return seen


class _ParallelBase(Module):
    """common bookkeeping: per-strategy (dp, tp, device groups) + the handful of layouts every module needs"""

    def _hetero_register(self, param, split_dim):
        """remember how `param` is laid out in the other pipelines of a heterogeneous strategy"""
        leaf = self.ds_parallel_configs[0]
        if "_hetero_orig" in leaf:
            HETERO_PARAMS[param.name] = (leaf["_hetero_orig"], leaf["_hetero_member"], split_dim)
        return param

    def __init__(self, multi_ds_parallel_config: List[dict]):
        super().__init__()
        self.ds_parallel_configs = multi_ds_parallel_config
        self.device_group_unions = []
        self.dp, self.tp, self.zero = [], [], []
        for cfg in multi_ds_parallel_config:
            ds_union, dgs = config2ds(cfg)
            self.device_group_unions.append(dgs)
            ds0 = ds_union.get(0)
            split_dims = [k for k, v in ds0.states.items() if k >= 0 and v > 1]
            self.tp.append(ds0.get_dim(split_dims[0]) if split_dims else 1)
            self.dp.append(ds0.get_dim(-1))
            self.zero.append(bool(cfg.get("zero", False)))

    def _ds(self, states_fn, order, zero=False):
        """build a DS hierarchy: one union per strategy; states_fn(dp, tp) -> states dict"""
        out = []
        for i, dgs in enumerate(self.device_group_unions):
            n = dgs[0].num_devices
            st = {k: v for k, v in states_fn(self.dp[i], self.tp[i]).items() if v > 1}
            od = [o for o in order if o in st]
            out.append(DistributedStatesUnion([DistributedStates(n, st, od, zero and self.zero[i])]))
        return out

    # activations
    def ds_split0_dup(self):       # tokens split over dp, replicated over tp
        return self._ds(lambda d, t: {0: d, -1: t}, [0, -1])

    def ds_split0(self):           # sequence parallel: tokens split over dp*tp
        return self._ds(lambda d, t: {0: d * t}, [0])

    def ds_split01(self):          # column-parallel output
        return self._ds(lambda d, t: {0: d, 1: t}, [0, 1])

    def ds_dup(self):
        return self._ds(lambda d, t: {-1: d * t}, [-1])

    # weights
    def ds_dup_split0(self, zero=True):
        return self._ds(lambda d, t: {-1: d, 0: t}, [-1, 0], zero)

    def ds_dup_split1(self, zero=True):
        return self._ds(lambda d, t: {-1: d, 1: t}, [-1, 1], zero)

    def _all_split0(self):         # tokens split over every device of the group (norm leaves: dp * tp = dup)
        return self._ds(lambda d, t: {0: d * t}, [0])

    def ds_w_dup(self, zero=True):
        return self._ds(lambda d, t: {-1: d * t}, [-1], zero)

    def _adapt(self, x, target=None):
        """bring `x` to layout `target` on this module's devices: a layout change lowers to a collective, a change of
        device group with an unchanged layout to pipeline P2P send/recv"""
        src_group = x.device_group
        my_group = self.device_group_unions[0][0]
        moved = (not src_group.empty) and src_group != my_group
        if target is None:
            if not moved:
                return x
            target = list(x.ds_hierarchy)          # same layout, new devices: pipeline P2P
        if x.check_ds_hierarchy_equal(target) and not moved:
            return x
        return ops.comm(x, target, device_group_hierarchy=self.device_group_unions)


class HtMultiColumnParallelLinear(_ParallelBase):
    """Y = X A^T + b with A split along its output dim over the TP group.
    Input must be (or is converted to) split0_dup; output is split01 unless gather_output."""

    def __init__(self, in_features, out_features, multi_ds_parallel_config, bias=True, gather_output=True,
                 init_method="xavier_normal_", dtype="float32", name="colp", init_std=None):
        super().__init__(multi_ds_parallel_config)
        self.in_features, self.out_features, self.gather_output, self.name = in_features, out_features, gather_output, name
        init = normal_initializer(0.0, init_std) if init_std is not None else xavier_normal_initializer()
        self.weight = parallel_parameter(init, [out_features, in_features], self.ds_dup_split0(), dtype=dtype, requires_grad=True,
                                         device_group_hierarchy=self.device_group_unions, name=f"{name}_weight")
        self.bias = parallel_parameter(zeros_initializer(), [out_features], self.ds_dup_split0(), dtype=dtype, requires_grad=True,
                                       device_group_hierarchy=self.device_group_unions, name=f"{name}_bias") if bias else None
        self._hetero_register(self.weight, 0)
        if self.bias is not None:
            self._hetero_register(self.bias, 0)

    fp8 = False   # set per instance (or via model config) to run the GEMM in block-scaled e4m3

    def forward(self, x, act="none"):
        x = self._adapt(x, self.ds_split0_dup())   # sequence-parallel inputs are all-gathered here
        if self.fp8:
            y = ops.linear_fp8(x, self.weight, self.bias, act=act, device_group_hierarchy=self.device_group_unions,
                               name=f"linear_{self.name}")
        else:
            y = ops.linear(x, self.weight, self.bias, trans_b=True, act=act, device_group_hierarchy=self.device_group_unions,
                           name=f"linear_{self.name}")
        if self.gather_output:
            y = self._adapt(y, self.ds_split0_dup())
        return y


class HtMultiQKVColumnParallelLinear(HtMultiColumnParallelLinear):
    """Fused q/k/v projection: the output dim is laid out [q heads | k heads | v heads] per TP shard so that
    each rank's slice holds complete heads (GQA: num_kv_heads may differ from num_heads)."""

    def __init__(self, in_features, head_dim, num_heads, num_kv_heads, multi_ds_parallel_config, bias=True,
                 dtype="float32", name="qkv", init_std=None):
        out = (num_heads + 2 * num_kv_heads) * head_dim
        super().__init__(in_features, out, multi_ds_parallel_config, bias=bias, gather_output=False, dtype=dtype, name=name,
                         init_std=init_std)
        self.head_dim, self.num_heads, self.num_kv_heads = head_dim, num_heads, num_kv_heads


class HtMultiRowParallelLinear(_ParallelBase):
    """Y = X A^T + b with A split along its input dim; the partial outputs are all-reduced (or reduce-scattered
    onto the sequence dim when sequence_parallel)."""

    def __init__(self, in_features, out_features, multi_ds_parallel_config, sequence_parallel=False, bias=True,
                 init_method="xavier_normal_", dtype="float32", name="rowp", init_std=None):
        super().__init__(multi_ds_parallel_config)
        self.in_features, self.out_features, self.sequence_parallel, self.name = in_features, out_features, sequence_parallel, name
        init = normal_initializer(0.0, init_std) if init_std is not None else xavier_normal_initializer()
        self.weight = parallel_parameter(init, [out_features, in_features], self.ds_dup_split1(), dtype=dtype, requires_grad=True,
                                         device_group_hierarchy=self.device_group_unions, name=f"{name}_weight")
        self.bias = parallel_parameter(zeros_initializer(), [out_features], self.ds_w_dup(), dtype=dtype, requires_grad=True,
                                       device_group_hierarchy=self.device_group_unions, name=f"{name}_bias") if bias else None
        self._hetero_register(self.weight, 1)
        if self.bias is not None:
            self._hetero_register(self.bias, None)

    fp8 = False

    def forward(self, x, residual=None):
        x = self._adapt(x, self.ds_split01())
        tp_any = any(t > 1 for t in self.tp)
        if not tp_any:
            if self.fp8:
                y = ops.linear_fp8(x, self.weight, self.bias, device_group_hierarchy=self.device_group_unions, name=f"linear_{self.name}")
                return y if residual is None else y + residual
            return ops.linear(x, self.weight, self.bias, trans_b=True, residual=residual,
                              device_group_hierarchy=self.device_group_unions, name=f"linear_{self.name}")
        if self.fp8:
            y = ops.linear_fp8(x, self.weight, None, device_group_hierarchy=self.device_group_unions, name=f"linear_{self.name}")
        else:
            y = ops.linear(x, self.weight, None, trans_b=True, device_group_hierarchy=self.device_group_unions,
                           name=f"linear_{self.name}")                      # partial sums
        y = ops.comm(y, self.ds_split0() if self.sequence_parallel else self.ds_split0_dup())
        if self.bias is not None:
            y = y + self.bias
        if residual is not None:
            y = y + residual
        return y


class HtMultiParallelEmbedding(_ParallelBase):
    """replicated table, data-parallel lookups"""

    def __init__(self, num_embeddings, embedding_dim, multi_ds_parallel_config, init_method="xavier_normal_", dtype="float32",
                 name="embedding", init_std=None):
        super().__init__(multi_ds_parallel_config)
        init = normal_initializer(0.0, init_std) if init_std is not None else xavier_normal_initializer()
        self.embedding_table = parallel_parameter(init, [num_embeddings, embedding_dim], self.ds_w_dup(), dtype=dtype,
                                                  requires_grad=True, device_group_hierarchy=self.device_group_unions,
                                                  name=f"{name}_table")
        self._hetero_register(self.embedding_table, None)

    def forward(self, ids):
        return ops.embedding_lookup(self.embedding_table, ids, device_group_hierarchy=self.device_group_unions)


class HtMultiVocabParallelEmbedding(_ParallelBase):
    """table split along the vocabulary over TP: out-of-shard ids give zeros, partial results are all-reduced"""

    def __init__(self, num_embeddings, embedding_dim, multi_ds_parallel_config, init_method="xavier_normal_", dtype="float32",
                 name="vocab_embedding", init_std=None):
        super().__init__(multi_ds_parallel_config)
        self.num_embeddings = num_embeddings
        init = normal_initializer(0.0, init_std) if init_std is not None else xavier_normal_initializer()
        self.embedding_table = parallel_parameter(init, [num_embeddings, embedding_dim], self.ds_dup_split0(), dtype=dtype,
                                                  requires_grad=True, device_group_hierarchy=self.device_group_unions,
                                                  name=f"{name}_table")
        self._hetero_register(self.embedding_table, 0)

    def _vocab_offset(self, strategy=0):
        from ..distributed import rank
        dgs = self.device_group_unions[strategy]
        idx = dgs[0].indices().index(rank()) if rank() in dgs[0].indices() else 0
        ds = self.embedding_table.get_ds(strategy)
        shard = ds.map_device_to_state_index(idx).get(0, 0)
        return shard * (self.num_embeddings // max(self.tp[strategy], 1))

    def forward(self, ids, sequence_parallel=False):
        ids = self._adapt(ids, self.ds_split0_dup())
        offs = [self._vocab_offset(s) for s in range(len(self.device_group_unions))]
        y = ops.embedding_lookup(self.embedding_table, ids, vocab_offset=offs[0], vocab_offsets=offs,
                                 device_group_hierarchy=self.device_group_unions)
        if any(t > 1 for t in self.tp):
            y = ops.comm(y, self.ds_split0() if sequence_parallel else self.ds_split0_dup())
        return y


class HtMultiParallelLayerNorm(_ParallelBase):
    def __init__(self, normalized_shape, multi_ds_parallel_config, sequence_parallel=False, eps=1e-5, dtype="float32",
                 name="ln"):
        super().__init__(multi_ds_parallel_config)
        n = normalized_shape if isinstance(normalized_shape, int) else normalized_shape[-1]
        self.eps, self.sequence_parallel = eps, sequence_parallel
        self.weight = parallel_parameter(ones_initializer(), [n], self.ds_w_dup(), dtype=dtype, requires_grad=True,
                                         device_group_hierarchy=self.device_group_unions, name=f"{name}_weight")
        self.bias = parallel_parameter(zeros_initializer(), [n], self.ds_w_dup(), dtype=dtype, requires_grad=True,
                                       device_group_hierarchy=self.device_group_unions, name=f"{name}_bias")
        self._hetero_register(self.weight, None)
        self._hetero_register(self.bias, None)

    def forward(self, x):
        x = self._adapt(x, self._all_split0() if self.sequence_parallel else None)
        return ops.layer_norm(x, self.weight, self.bias, eps=self.eps, device_group_hierarchy=self.device_group_unions)


class HtMultiParallelRMSNorm(_ParallelBase):
    def __init__(self, normalized_shape, multi_ds_parallel_config, sequence_parallel=False, eps=1e-6, dtype="float32",
                 name="rmsnorm"):
        super().__init__(multi_ds_parallel_config)
        n = normalized_shape if isinstance(normalized_shape, int) else normalized_shape[-1]
        self.eps, self.sequence_parallel = eps, sequence_parallel
        self.weight = parallel_parameter(ones_initializer(), [n], self.ds_w_dup(), dtype=dtype, requires_grad=True,
                                         device_group_hierarchy=self.device_group_unions, name=f"{name}_weight")
        self._hetero_register(self.weight, None)

    def forward(self, x):
        x = self._adapt(x, self._all_split0() if self.sequence_parallel else None)
        return ops.rms_norm(x, self.weight, eps=self.eps, device_group_hierarchy=self.device_group_unions)


# ----------------------------------------------------------------------------- single-strategy API (hetu.nn.parallel)
def _one(device_group, dp, kind="variable", split_dim=0, zero=False):
    ids = device_group.indices() if hasattr(device_group, "indices") else list(device_group)
    tp = max(len(ids) // max(dp, 1), 1)
    return [single_config(dp, tp, ids, kind, split_dim, zero)]


class ColumnParallelLinear(HtMultiColumnParallelLinear):
    def __init__(self, in_features, out_features, device_group, dp=1, bias=True, gather_output=True, dtype="float32", name="colp"):
        super().__init__(in_features, out_features, _one(device_group, dp), bias=bias, gather_output=gather_output, dtype=dtype, name=name)


class RowParallelLinear(HtMultiRowParallelLinear):
    def __init__(self, in_features, out_features, device_group, dp=1, bias=True, dtype="float32", name="rowp"):
        super().__init__(in_features, out_features, _one(device_group, dp, split_dim=1), bias=bias, dtype=dtype, name=name)


class ParallelEmbedding(HtMultiParallelEmbedding):
    def __init__(self, num_embeddings, embedding_dim, device_group, dp=None, dtype="float32", name="embedding"):
        ids = device_group.indices()
        super().__init__(num_embeddings, embedding_dim, _one(device_group, dp or len(ids)), dtype=dtype, name=name)


class VocabParallelEmbedding(HtMultiVocabParallelEmbedding):
    def __init__(self, num_embeddings, embedding_dim, device_group, dp=1, dtype="float32", name="vocab_embedding"):
        super().__init__(num_embeddings, embedding_dim, _one(device_group, dp), dtype=dtype, name=name)


class ParallelLayerNorm(HtMultiParallelLayerNorm):
    def __init__(self, normalized_shape, device_group, dp=None, eps=1e-5, dtype="float32", name="ln"):
        ids = device_group.indices()
        super().__init__(normalized_shape, _one(device_group, dp or len(ids)), eps=eps, dtype=dtype, name=name)


# ----------------------------------------------------------------------------- single-config DS modules (hetu.nn.parallel_ds)
def _single(cls):
    """Ht<X> takes ONE ds_parallel_config dict where HtMulti<X> takes the list of all strategies
    (ref: python/hetu/nn/modules/parallel_ds.py)"""
    pos = list(cls.__init__.__code__.co_varnames[:cls.__init__.__code__.co_argcount])
    idx = pos.index("multi_ds_parallel_config") - 1

    class _One(cls):
        def __init__(self, *a, **k):
            if "ds_parallel_config" in k:
                k["multi_ds_parallel_config"] = [k.pop("ds_parallel_config")]
            elif len(a) > idx and isinstance(a[idx], dict):
                a = list(a)
                a[idx] = [a[idx]]
            super().__init__(*a, **k)
    _One.__name__ = _One.__qualname__ = cls.__name__.replace("HtMulti", "Ht")
    return _One


HtParallelRMSNorm = _single(HtMultiParallelRMSNorm)
HtParallelLayerNorm = _single(HtMultiParallelLayerNorm)
HtParallelEmbedding = _single(HtMultiParallelEmbedding)
HtVocabParallelEmbedding = _single(HtMultiVocabParallelEmbedding)
HtColumnParallelLinear = _single(HtMultiColumnParallelLinear)
HtRowParallelLinear = _single(HtMultiRowParallelLinear)

__all__ += ["HtParallelRMSNorm", "HtParallelLayerNorm", "HtParallelEmbedding", "HtVocabParallelEmbedding", "HtColumnParallelLinear",
            "HtRowParallelLinear", "HETERO_PARAMS", "hetero_grad_sync_spec", "precreate_hetero_groups", "all_hetero_groups"]
