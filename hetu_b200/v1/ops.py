"""The remaining `*_op` constructors of the v1 API (executor.py holds the most common ones): arithmetic and reductions, index /
order statistics, sampling, hashing, MoE layout transforms and gates' helpers, collective-communication nodes, quantisation.
Each maps onto an op of the graph library; gradients come from the graph's autodiff (the reference's explicit `*_gradient_op`
constructors live in grad_ops.py).  (ref: hetu/v1/python/hetu/gpu_ops/__init__.py and the per-op modules next to it)"""
from __future__ import annotations

import numpy as np

from .. import core, ops


# ---- arithmetic / elementwise
def addmm_op(c, a, b, alpha=1.0, beta=1.0): return ops.matmul(a, b) * float(alpha) + c * float(beta)          # noqa: E704
def baddbmm_op(c, a, b, alpha=1.0, beta=1.0): return ops.bmm(a, b) * float(alpha) + c * float(beta)          # noqa: E704
def matrix_dot_op(a, b, axes=0): return ops.mul(a, b)                                                        # noqa: E704
def mul_byconst_op(x, c): return x * float(c)                                                                 # noqa: E704
def minus_byconst_op(x, c): return float(c) - x                                                               # noqa: E704  (c - x, as in the reference)
def div_const_op(c, x): return float(c) / x                                                                   # noqa: E704
def div_handle_zero_op(a, b): return ops.where(ops.equal(b, 0.0), ops.zeros_like(a), a / (b + ops.equal(b, 0.0)))   # noqa: E704
def const_pow_op(x, c): return ops.exp(x * float(np.log(c)))                                                 # noqa: E704  (c ** x)
def power_op(x, p): return ops.pow(x, float(p))                                                               # noqa: E704
def cos_op(x): return ops.cos(x)                                                                              # noqa: E704
def sin_op(x): return ops.sin(x)                                                                              # noqa: E704
def floor_op(x): return ops.floor(x)                                                                          # noqa: E704
def sign_op(x): return ops.sign(x)                                                                            # noqa: E704
def bool_op(x): return ops.not_equal(x, 0.0)                                                                  # noqa: E704
def clamp_op(x, mmin=None, mmax=None, min=None, max=None):                                                   # noqa: A002
    lo = mmin if mmin is not None else min
    hi = mmax if mmax is not None else max
    return ops.clamp(x, -np.inf if lo is None else float(lo), np.inf if hi is None else float(hi))
def mask_op(x, mask): return ops.mul(x, mask)                                                                 # noqa: E704
def masked_fill_op(x, mask, val=0.0): return ops.masked_fill(x, mask, float(val))                            # noqa: E704
def where_const_op(cond, a, const): return ops.where(cond, a, ops.full_like(a, float(const)))                # noqa: E704
def stop_gradient_op(x): return ops.stop_gradient(x)                                                          # noqa: E704
def oneslike_op(x): return ops.ones_like(x)                                                                   # noqa: E704
def zeroslike_op(x): return ops.zeros_like(x)                                                                 # noqa: E704
def full_like_op(x, value): return ops.full_like(x, float(value))                                             # noqa: E704
def full_op(shape, value, dtype="float32"): return core.from_numpy(np.full(list(shape), value, dtype=np.dtype(dtype)))   # noqa: E704
def arange_op(start, end=None, step=1): return ops.arange(0 if end is None else start, start if end is None else end, step)   # noqa: E704
def log_softmax_op(x): return ops.log_softmax(x, -1)                                                          # noqa: E704
def dropout2d_op(x, keep_prob): return ops.dropout(x, 1.0 - float(keep_prob))                                 # noqa: E704
def binary_step_op(x): return ops.greater(x, 0.0)                                                             # noqa: E704
def datah2d_op(x, ctx=None): return x                                                                         # noqa: E704  (placement is the executor's job here)
def datad2h_op(x, ctx=None): return x                                                                         # noqa: E704


# ---- reductions / shape
def max_op(a, b): return ops.where(ops.greater(a - b, 0.0), a, b)                                             # noqa: E704
def min_op(a, b): return ops.where(ops.greater(a - b, 0.0), b, a)                                             # noqa: E704
def reduce_mul_op(x, axes, keepdims=False): return ops.prod(x, axes, keepdims)                                # noqa: E704
def reduce_norm1_op(x, axes, keepdims=False): return ops.sum(ops.abs(x), axes, keepdims)                      # noqa: E704
def reduce_norm2_op(x, axes, keepdims=False): return ops.sqrt(ops.sum(x * x, axes, keepdims))                 # noqa: E704
def reducesumaxiszero_op(x): return ops.sum(x, [0])                                                           # noqa: E704
def norm_op(x, axis=None, p=2, keepdims=False): return ops.norm(x, float(p), axis, keepdims)                  # noqa: E704
def broadcast_shape_op(x, shape, add_axes=()): return ops.broadcast(x, list(shape), add_axes)                 # noqa: E704
def reshape_to_op(x, like): return ops.reshape(x, list(like.shape))                                           # noqa: E704
def repeat_op(x, reps): return ops.repeat(x, list(reps))                                                      # noqa: E704
def tile_op(x, reps): return ops.repeat(x, list(reps))                                                        # noqa: E704
def roll_op(x, shift, axis=None): return ops.roll(x, [shift] if np.isscalar(shift) else list(shift), [0 if axis is None else axis] if np.isscalar(axis) or axis is None else list(axis))   # noqa: E704,E501
def interpolate_op(x, size=None, scale_factor=None, mode="bilinear", align_corners=False): return ops.interpolate(x, size, scale_factor, mode, align_corners)   # noqa: E704,E501
def slice_assign_op(x, begin, size, value):
    """x with x[begin : begin + size] = value (scalar)"""
    mask = np.zeros(list(x.shape), np.float32)
    mask[tuple(slice(b, b + s) for b, s in zip(begin, size))] = 1.0
    m = core.from_numpy(mask)
    return x * (1.0 - m) + m * float(value)


# ---- index / order statistics
def argmax_op(x, dim=-1): return ops.argmax(x, dim)                                                           # noqa: E704
def argsort_op(x, dim=-1, descending=False): return ops.argsort(x, dim, descending)                           # noqa: E704
def topk_val_op(x, k, dim=-1): return ops.topk(x, k, dim)[0]                                                  # noqa: E704
def topk_idx_op(x, k, dim=-1): return ops.topk(x, k, dim)[1]                                                  # noqa: E704
def gather_op(x, dim, index): return ops.gather(x, dim, index)                                                # noqa: E704
def scatter_op(x, dim, index, src): return ops.scatter(x, dim, index, src)                                    # noqa: E704
def scatter1d_op(x, index, src): return ops.scatter(x, 0, index, src)                                         # noqa: E704
def indexing_op(x, index): return ops.embedding_lookup(x, index)                                              # noqa: E704  (rows of x)
def cumsum_with_bias_op(x, bias=0.0, dim=0): return ops.cumsum(x, dim) + float(bias)                         # noqa: E704
def unique_indices_op(x): return ops.unique(x)                                                                # noqa: E704
def tril_lookup_op(x, offset=0):
    """the lower-triangular entries (row-major) of the last two dims as a vector"""
    n, m = x.shape[-2], x.shape[-1]
    rows, cols = np.tril_indices(n, offset, m)
    flat = ops.reshape(x, list(x.shape[:-2]) + [n * m])
    idx = core.from_numpy((rows * m + cols).astype(np.int64))
    lead = 1
    for s in x.shape[:-2]:
        lead *= s
    picked = ops.transpose(ops.embedding_lookup(ops.transpose(ops.reshape(flat, [lead, n * m]), [1, 0]), idx), [1, 0])
    return ops.reshape(picked, list(x.shape[:-2]) + [len(rows)])


def conv2d_broadcastto_op(bias, like): return ops.broadcast(ops.reshape(bias, [1, like.shape[1], 1, 1]), list(like.shape))                 # noqa: E704
def conv2d_reducesum_op(x): return ops.sum(x, [0, 2, 3])                                                     # noqa: E704
def sparse_embedding_lookup_op(table, ids): return ops.embedding_lookup(table, ids)                          # noqa: E704
def argmax_partial_op(x, full_mask, topk, dim=-1): return ops.argmax(x * full_mask + (full_mask - 1.0) * 1e30, dim)   # noqa: E704
def min_dist_op(query, codebook, indices=None, mode="eu"):
    """index of the nearest codebook row for every query row (euclidean, or largest inner product with mode='ip')"""
    dots = ops.matmul(query, codebook, trans_b=True)
    if mode == "ip":
        return ops.argmax(dots, -1)
    d2 = ops.sum(codebook * codebook, [1]) - dots * 2.0             # |q|^2 is constant per row
    return ops.argmax(d2 * -1.0, -1)
def csrmm_op(indices, values, shape, dense): return ops.spmm(indices, values, dense, int(shape[0]))          # noqa: E704,E302
def csrmv_op(indices, values, shape, vec): return ops.reshape(ops.spmm(indices, values, ops.reshape(vec, [vec.shape[0], 1]), int(shape[0])), [shape[0]])   # noqa: E704,E501
def mod_hash_negative_op(ids, nembed): return ops.remainder(ops.remainder(ids, int(nembed)) + int(nembed), int(nembed))   # noqa: E704
def robe_sign_op(ids, a=1000003, b=12345): return ops.hash_ids(ids, 2, a=int(a), b=int(b)) * 2 - 1          # noqa: E704  (+1 / -1 per id)
def reduceCommunicate_op(x, root=0, comm=None, ranks=None): return ops.all_reduce(x, ranks or _world())      # noqa: E704,N802  (every rank gets the sum)


# ---- losses
def crossentropy_op(probs, labels): return ops.sum(labels * ops.log(probs + 1e-12), [-1]) * -1.0             # noqa: E704
def crossentropy_sparse_op(probs, labels, ignored_index=-1): return ops.nll_loss(ops.log(probs + 1e-12), labels, reduction="none")   # noqa: E704
def nll_loss_op(logp, labels): return ops.nll_loss(logp, labels, reduction="none")                            # noqa: E704
def binarycrossentropywithlogits_op(logits, labels): return ops.binary_cross_entropy(ops.sigmoid(logits), labels, reduction="none")   # noqa: E704


# ---- sampling (values are drawn under the global seed when the node is created, as constants of the graph)
def _rng():
    return np.random.RandomState(core._global_seed[0] + _rng.calls)
_rng.calls = 0                                                                                                # noqa: E305


def _sample(fn, shape):
    _rng.calls += 1
    return core.from_numpy(fn(_rng(), list(shape)).astype(np.float32))
def rand_op(shape): return _sample(lambda r, s: r.rand(*s), shape)                                            # noqa: E704,E302
def uniform_sample_op(shape, low=0.0, high=1.0): return _sample(lambda r, s: r.uniform(low, high, s), shape)  # noqa: E704
def normal_sample_op(shape, mean=0.0, stddev=1.0): return _sample(lambda r, s: r.normal(mean, stddev, s), shape)   # noqa: E704
def truncated_normal_sample_op(shape, mean=0.0, stddev=1.0): return _sample(lambda r, s: np.clip(r.normal(mean, stddev, s), mean - 2 * stddev, mean + 2 * stddev), shape)   # noqa: E704,E501
def gumbel_sample_op(shape): return _sample(lambda r, s: r.gumbel(size=s), shape)                             # noqa: E704
def randint_sample_op(shape, low, high):
    _rng.calls += 1
    return core.from_numpy(_rng().randint(low, high, list(shape)).astype(np.int64))


# ---- hashing (embedding compression)
def mod_hash_op(ids, nembed): return ops.remainder(ids, int(nembed))                                          # noqa: E704
def div_hash_op(ids, nembed): return ops.floor_divide(ids, int(nembed))                                       # noqa: E704
def compo_hash_op(ids, ntable, nembed): return [ops.remainder(ops.floor_divide(ids, int(nembed) ** t), int(nembed)) for t in range(int(ntable))]   # noqa: E704,E501
def learn_hash_op(ids, a, b, prime, nbucket): return ops.remainder(ops.remainder(ids * int(a) + int(b), int(prime)), int(nbucket))   # noqa: E704
def robe_hash_op(ids, nbucket, a=1000003, b=12345): return ops.hash_ids(ids, int(nbucket), a=int(a), b=int(b))   # noqa: E704


# ---- MoE layout transforms and gate helpers
def _routing(t, dtype="int32"):
    """v1 passes routing data per choice: one [T] vector (top-1) or a list of k of them; the library ops take [T, k]"""
    if isinstance(t, (list, tuple)):
        cols = [ops.reshape(ops.cast(c, dtype), [-1, 1]) for c in t]
        return cols[0] if len(cols) == 1 else ops.concat(cols, 1)
    t = ops.cast(t, dtype)
    return ops.reshape(t, [-1, 1]) if len(t.shape) == 1 else t
def layout_transform_op(x, indices, locations, capacity, num_experts):                                       # noqa: E704
    """tokens [T, d] -> expert-major slots [E * capacity, d]"""
    y = ops.moe_dispatch(x, _routing(indices), _routing(locations), int(num_experts), int(capacity))
    return ops.reshape(y, [int(num_experts) * int(capacity), x.shape[-1]])
def _slots(y, capacity, num_experts):
    return y if len(y.shape) == 3 else ops.reshape(y, [int(num_experts), int(capacity), y.shape[-1]])
def reverse_layout_transform_op(y, indices, locations, gates, capacity, num_experts):                        # noqa: E704
    return ops.moe_combine(_slots(y, capacity, num_experts), _routing(indices), _routing(locations), _routing(gates, "float32"))
def reverse_layout_transform_no_gate_op(y, indices, locations, capacity, num_experts):                       # noqa: E704
    return ops.moe_combine(_slots(y, capacity, num_experts), _routing(indices), _routing(locations), None)
def balance_assignment_op(scores): return ops._op1("moe_balance_assign", [scores], {})                        # noqa: E704
def group_topk_idx_op(x, top1_group, topk=1, num_local_gpus=8):
    """top-k experts restricted to the expert group chosen by `top1_group` (SAM gate): scores outside the group are masked out"""
    e = x.shape[-1]
    group = ops.floor_divide(core.from_numpy(np.arange(e, dtype=np.int64).reshape(1, e)), e // int(num_local_gpus))
    inside = ops.equal(group - ops.reshape(top1_group, [x.shape[0], 1]), 0.0)
    return ops.topk(x * inside + (inside - 1.0) * 1e30, int(topk), -1)[1]
def sam_group_sum_op(gate, num_local_gpus): return ops.sum(ops.reshape(gate, [gate.shape[0], int(num_local_gpus), gate.shape[1] // int(num_local_gpus)]), [2])   # noqa: E704,E501,E302
def sam_max_op(gate, top1_group, topk_idx, num_local_gpus):
    """per token: max score outside its chosen group minus the chosen scores, clipped at 0 (the SAM gate's auxiliary term)"""
    e = gate.shape[-1]
    group = ops.floor_divide(core.from_numpy(np.arange(e, dtype=np.int64).reshape(1, e)), e // int(num_local_gpus))
    outside = ops.not_equal(group - ops.reshape(top1_group, [gate.shape[0], 1]), 0.0)
    mx = ops.max(gate * outside + (outside - 1.0) * 1e30, [1], True)
    return ops.relu(mx - ops.gather(gate, 1, topk_idx))


# ---- communication nodes (the executor lowers them onto the process groups / fused kernels)
def _world():
    from .. import distributed
    return list(range(max(distributed.world_size(), 1)))
def allreduceCommunicate_op(x, comm=None, ranks=None): return ops.all_reduce(x, ranks or _world())            # noqa: E704,N802,E302
def groupallreduceCommunicate_op(x, ranks): return ops.all_reduce(x, list(ranks))                             # noqa: E704,N802
def allgatherCommunicate_op(x, comm=None, ranks=None, dim=0): return ops.all_gather(x, ranks or _world(), dim)   # noqa: E704,N802
def reducescatterCommunicate_op(x, comm=None, ranks=None, dim=0): return ops.reduce_scatter(x, ranks or _world(), dim)   # noqa: E704,N802
def broadcastCommunicate_op(x, comm=None, root=0, ranks=None): return ops._op1("broadcast_comm", [x], {"ranks": ranks or _world(), "root": int(root)})   # noqa: E704,N802,E501
def alltoall_op(x, comm=None, ranks=None): return ops.all_to_all(x, ranks or _world())                        # noqa: E704
def halltoall_op(x, comm=None, ranks=None, gpus_per_node=8): return ops.hall_to_all(x, ranks or _world(), int(gpus_per_node))   # noqa: E704


# ---- quantisation
def quantize_op(x, digit=8, scale=None, minele=None): return ops.quantization(x, "int8" if int(digit) == 8 else "nf4", 64)   # noqa: E704
def dequantize_op(q, absmax, digit=8, shape=None, dtype="float32"): return ops.dequantization(q, absmax, dtype, 64, shape=shape, quant_type="int8" if int(digit) == 8 else "nf4")   # noqa: E704,E501
def prune_low_magnitude_op(x, rate):
    """zero the `rate` fraction of smallest-magnitude entries (threshold from the sorted magnitudes)"""
    n = 1
    for s in x.shape:
        n *= s
    k = max(int(n * float(rate)), 1)
    thr = ops.slice(ops.topk(ops.reshape(ops.abs(x), [n]), k, -1, largest=False)[0], [k - 1], [1])
    return x * ops.greater(ops.abs(x) - thr, 0.0)
def param_clip_op(x, min_value, max_value): return ops.clamp(x, float(min_value), float(max_value))          # noqa: E704,E302


__all__ = [n for n in dir() if n.endswith("_op")]
