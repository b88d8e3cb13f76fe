"""Python op surface: `hetu.<op>(...)` functions and `Tensor.<op>` methods.

The reference generates these from `_binding/codegen/ops.yml` (150 entries); here a declarative table drives
the generation (`OP_TABLE`) and the handful of ops with non-trivial argument handling are written out.
Every function accepts the OP_META kwargs of the reference: name, device_group_hierarchy, stream_index, extra_deps.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _C
from .core import (DistributedStates, IntSymbol, Tensor, _normalize_dsh, cur_graph, dtype_name, from_numpy, make_op)

_META_KEYS = ("name", "device_group_hierarchy", "stream_index", "extra_deps", "is_cpu")


def _meta(kw):
    return {k: kw.pop(k) for k in list(kw) if k in _META_KEYS}


def _t(x, like: Optional[Tensor] = None):
    """python scalars / numpy arrays become constants"""
    if isinstance(x, Tensor):
        return x
    arr = torch.as_tensor(np.asarray(x))
    if like is not None and arr.is_floating_point():
        from .core import to_torch_dtype
        arr = arr.to(to_torch_dtype(like.dtype))
    return from_numpy(arr)


def _op1(op_type, inputs, attrs=None, **kw):
    return make_op(op_type, inputs, attrs or {}, **_meta(kw))[0]


# ----------------------------------------------------------------------------- arithmetic
def _binary(op_type):
    def f(a, b, **kw):
        if isinstance(a, Tensor) and isinstance(b, Tensor):
            return _op1(op_type, [a, b], **kw)
        if isinstance(a, Tensor):
            if op_type == "div":
                return _op1("div", [a], {"value": float(b)}, **kw)
            return _op1(op_type, [a], {"value": float(b)}, **kw)
        # const (op) tensor
        if op_type in ("add", "mul"):
            return _op1(op_type, [b], {"value": float(a)}, **kw)
        return _op1(op_type, [b], {"value": float(a), "from_const": True}, **kw)
    f.__name__ = op_type
    return f


add = _binary("add")
sub = _binary("sub")
mul = _binary("mul")
div = _binary("div")


def neg(x, **kw):
    return _op1("neg", [x], **kw)


def pow(x, exponent, **kw):  # noqa: A001
    return _op1("pow", [x], {"exponent": float(exponent)}, **kw)


def _unary(op_type, hot_kind=None):
    def f(x, **kw):
        if hot_kind is not None:
            return _op1("unary_act", [x], {"kind": hot_kind}, **kw)
        return _op1(op_type, [x], **kw)
    f.__name__ = op_type
    return f


abs = _unary("abs")  # noqa: A001
ceil = _unary("ceil")
floor = _unary("floor")
round = _unary("round")  # noqa: A001
exp = _unary("exp")
log = _unary("log")
sqrt = _unary("sqrt")
rsqrt = _unary("rsqrt")
sin = _unary("sin")
cos = _unary("cos")
reciprocal = _unary("reciprocal")
sigmoid = _unary("sigmoid")
tanh = _unary("tanh")
hardsigmoid = _unary("hardsigmoid")
hardswish = _unary("hardswish")
logsigmoid = _unary("logsigmoid")
mish = _unary("mish")
relu = _unary("relu", "relu")
gelu = _unary("gelu", "gelu")
silu = _unary("silu", "silu")
contiguous = _unary("contiguous")
checknumeric = _unary("checknumeric")


def leakyrelu(x, alpha=0.01, **kw):
    return _op1("leakyrelu", [x], {"alpha": float(alpha)}, **kw)


def elu(x, alpha=1.0, scale=1.0, **kw):
    return _op1("elu", [x], {"alpha": float(alpha), "scale": float(scale)}, **kw)


def hardshrink(x, lambda_=0.5, **kw):
    return _op1("hardshrink", [x], {"lambda": float(lambda_)}, **kw)


def softshrink(x, lambda_=0.5, **kw):
    return _op1("softshrink", [x], {"lambda": float(lambda_)}, **kw)


def hardtanh(x, min_val=-1.0, max_val=1.0, **kw):
    return _op1("hardtanh", [x], {"min_val": float(min_val), "max_val": float(max_val)}, **kw)


def softplus(x, beta=1.0, threshold=20.0, **kw):
    return _op1("softplus", [x], {"beta": float(beta), "threshold": float(threshold)}, **kw)


def softmax(x, dim=-1, **kw):
    return _op1("softmax", [x], {"dim": int(dim)}, **kw)


def log_softmax(x, dim=-1, **kw):
    return _op1("log_softmax", [x], {"dim": int(dim)}, **kw)


def swiglu(x, interleaved=False, **kw):
    """silu(gate) * up of a fused gate/up projection; interleaved: columns are (gate_0, up_0, gate_1, up_1, ...)"""
    return _op1("swiglu", [x], {"interleaved": bool(interleaved)}, **kw)


def clamp(x, min=-math.inf, max=math.inf, **kw):  # noqa: A002
    return _op1("clamp", [x], {"min": float(min), "max": float(max)}, **kw)


def where(cond, a, b, **kw):
    return _op1("where", [cond, _t(a), _t(b)], **kw)


def masked_fill(x, mask, value, **kw):
    return _op1("masked_fill", [x, mask], {"value": float(value)}, **kw)


def triu(x, lower=False, diagonal=0, **kw):
    return _op1("triu", [x], {"lower": bool(lower), "diagonal": int(diagonal)}, **kw)


def onehot(x, num_classes, **kw):
    return _op1("onehot", [x], {"num_classes": int(num_classes)}, **kw)


def range_mask(x, min, max, **kw):  # noqa: A002
    return _op1("range_mask", [x], {"min": int(min), "max": int(max)}, **kw)


def arange(start, end, step=1, dtype="int64", **kw):
    return _op1("arange", [], {"start": float(start), "end": float(end), "step": float(step), "dtype": dtype_name(dtype)}, **kw)


def ones_like(x, **kw):
    return _op1("ones_like", [x], **kw)


def zeros_like(x, **kw):
    return _op1("zeros_like", [x], **kw)


def full_like(x, value, **kw):
    return _op1("full_like", [x], {"value": float(value)}, **kw)


def data_transfer(x, dtype, **kw):
    return _op1("data_transfer", [x], {"dtype": dtype_name(dtype)}, **kw)


# ----------------------------------------------------------------------------- reductions
def reduce(x, mode="sum", axes=None, keepdims=False, **kw):  # noqa: A001
    axes = [] if axes is None else ([int(axes)] if np.isscalar(axes) else [int(a) for a in axes])
    if isinstance(keepdims, (list, tuple)):
        keepdims = bool(keepdims[0])
    return _op1("reduce", [x], {"mode": str(mode), "axes": axes, "keepdims": bool(keepdims)}, **kw)


def sum(x, axes=None, keepdims=False, **kw):  # noqa: A001
    if isinstance(x, (list, tuple)):
        return _op1("sum_n", list(x), **kw)
    return reduce(x, "sum", axes, keepdims, **kw)


def mean(x, axes=None, keepdims=False, **kw):
    return reduce(x, "mean", axes, keepdims, **kw)


def max(x, axes=None, keepdims=False, **kw):  # noqa: A001
    return reduce(x, "max", axes, keepdims, **kw)


def min(x, axes=None, keepdims=False, **kw):  # noqa: A001
    return reduce(x, "min", axes, keepdims, **kw)


def prod(x, axes=None, keepdims=False, **kw):
    return reduce(x, "prod", axes, keepdims, **kw)


def norm(x, p=2.0, dim=None, keepdim=False, **kw):
    axes = [] if dim is None else ([int(dim)] if np.isscalar(dim) else [int(a) for a in dim])
    return _op1("norm", [x], {"p": float(p), "axes": axes, "keepdims": bool(keepdim)}, **kw)


# ----------------------------------------------------------------------------- shape ops
def _shape_args(shape):
    """(static ints, symbolic list) -- a shape may mix ints and IntSymbols"""
    if any(isinstance(s, IntSymbol) for s in shape):
        return [], [s if isinstance(s, IntSymbol) else IntSymbol(int(s)) for s in shape]
    return [int(s) for s in shape], []


def reshape(x, shape, **kw):
    st, sy = _shape_args(list(shape))
    meta = _meta(kw)
    if sy:
        vals = [s.get_data() for s in sy]
        return make_op("reshape", [x], {"shape": vals}, sy_shape=sy, **meta)[0]
    return make_op("reshape", [x], {"shape": st}, **meta)[0]


def transpose(x, perm=None, **kw):
    return _op1("transpose", [x], {"perm": [int(p) for p in (perm or [])]}, **kw)


def slice(x, begin, size, **kw):  # noqa: A001
    meta = _meta(kw)
    st, sy = _shape_args(list(size))
    attrs = {"begin": [int(b) for b in begin], "size": st if st else [s.get_data() for s in sy]}
    return make_op("slice", [x], attrs, sy_shape=sy, **meta)[0]


def split(x, num_chunks_or_sections, dim=0, **kw) -> List[Tensor]:
    meta = _meta(kw)
    if isinstance(num_chunks_or_sections, (list, tuple)):
        attrs = {"sections": [int(s) for s in num_chunks_or_sections], "dim": int(dim)}
    else:
        attrs = {"num_chunks": int(num_chunks_or_sections), "dim": int(dim)}
    return make_op("split", [x], attrs, **meta)


def concat(tensors: Sequence[Tensor], axis=0, **kw):
    return _op1("concat", list(tensors), {"dim": int(axis)}, **kw)


concatenate = concat


def dynamic_concat(tensors, axis=0, **kw):
    return _op1("dynamic_concat", list(tensors), {"dim": int(axis)}, **kw)


def broadcast(x, shape, add_axes=(), **kw):
    return _op1("broadcast", [x], {"shape": [int(s) for s in shape], "add_axes": [int(a) for a in add_axes]}, **kw)


def repeat(x, repeats, **kw):
    return _op1("repeat", [x], {"repeats": [int(r) for r in repeats]}, **kw)


def roll(x, shifts, dims, **kw):
    return _op1("roll", [x], {"shifts": [int(s) for s in shifts], "dims": [int(d) for d in dims]}, **kw)


def pad(x, paddings, mode="constant", constant=0.0, **kw):
    return _op1("pad", [x], {"paddings": [int(p) for p in paddings], "value": float(constant)}, **kw)


def gather(x, dim, index, **kw):
    return _op1("gather", [x, index], {"dim": int(dim)}, **kw)


def argmax(x, dim=-1, keepdims=False, **kw):
    return _op1("argmax", [x], {"dim": int(dim), "keepdims": bool(keepdims)}, **kw)


def argsort(x, dim=-1, descending=False, **kw):
    return _op1("argsort", [x], {"dim": int(dim), "descending": bool(descending)}, **kw)


def topk(x, k, dim=-1, largest=True, **kw):
    """-> (values, indices)"""
    outs = make_op("topk", [x], {"k": int(k), "dim": int(dim), "largest": bool(largest)}, **_meta(kw))
    return outs[0], outs[1]


def cumsum(x, dim=-1, **kw):
    return _op1("cumsum", [x], {"dim": int(dim)}, **kw)


def sign(x, **kw):
    return _op1("sign", [x], **kw)


def scatter(x, dim, index, src, **kw):
    return _op1("scatter", [x, index, src], {"dim": int(dim)}, **kw)


def unique(x, **kw):
    """sorted unique values of a 1-D tensor padded to the input length, and every element's index among them -> (values, inverse)"""
    outs = make_op("unique_consecutive_count", [x], {}, **_meta(kw))
    return outs[0], outs[1]


def index_add_(x, index, src, dim=0, **kw):
    return _op1("index_add", [x, index, src], {"dim": int(dim)}, **kw)


def diagonal(x, offset=0, dim1=0, dim2=1, **kw):
    return _op1("diagonal", [x], {"offset": int(offset), "dim1": int(dim1), "dim2": int(dim2)}, **kw)


def as_strided(x, shape, stride, storage_offset=0, **kw):
    return _op1("as_strided", [x], {"shape": [int(s) for s in shape], "stride": [int(s) for s in stride],
                                   "storage_offset": int(storage_offset)}, **kw)


def interpolate(x, size, mode="bilinear", align_corners=False, **kw):
    return _op1("interpolate", [x], {"size": [int(s) for s in size], "mode": mode, "align_corners": bool(align_corners)}, **kw)


# ----------------------------------------------------------------------------- linear algebra
def matmul(a, b, trans_a=False, trans_b=False, **kw):
    return _op1("matmul", [a, b], {"trans_a": bool(trans_a), "trans_b": bool(trans_b)}, **kw)


def linear(x, w, bias=None, trans_a=False, trans_b=True, act="none", residual=None, **kw):
    """y = act(x @ w^T + bias) + residual.  With `act` the op has two outputs (y, pre-activation); y is returned."""
    assert not trans_a, "linear does not transpose its activation operand"
    ins = [x, w]
    attrs = {"trans_b": bool(trans_b), "has_bias": bias is not None, "has_residual": residual is not None, "act": str(act)}
    if bias is not None:
        ins.append(bias)
    if residual is not None:
        ins.append(residual)
    return make_op("linear", ins, attrs, **_meta(kw))[0]


def linear_fp8(x, w, bias=None, act="none", **kw):
    """y = act(x @ w^T + bias) with both operands quantised to e4m3 with one fp32 scale per row of x and per output row of w
    (i.e. per 1 x K slice -- coarser than 1 x 128 / 128 x 128 block scaling; the scales are applied in the tcgen05 GEMM epilogue);
    the input-gradient GEMM also runs in fp8, the weight gradient in bf16.  w is [out, in]."""
    ins = [x, w] + ([bias] if bias is not None else [])
    attrs = {"trans_b": True, "has_bias": bias is not None, "has_residual": False, "act": str(act)}
    return make_op("linear_fp8", ins, attrs, **_meta(kw))[0]


def bmm(a, b, **kw):
    return _op1("bmm", [a, b], **kw)


def dot(a, b, **kw):
    return _op1("dot", [a, b], **kw)


def outer(a, b, **kw):
    return _op1("outer", [a, b], **kw)


def einsum(equation, *tensors, **kw):
    if len(tensors) == 1 and isinstance(tensors[0], (list, tuple)):
        tensors = tuple(tensors[0])
    return _op1("einsum", list(tensors), {"equation": equation}, **kw)


# ----------------------------------------------------------------------------- nn
def layer_norm(x, weight, bias, normalized_shape=None, eps=1e-5, **kw):
    return make_op("fused_norm", [x, weight, bias], {"rms": False, "eps": float(eps)}, **_meta(kw))[0]


fused_layernorm = layer_norm


def rms_norm(x, weight, eps=1e-6, **kw):
    return make_op("fused_norm", [x, weight], {"rms": True, "eps": float(eps)}, **_meta(kw))[0]


def fused_rmsnorm(x, weight, normalized_shape=None, eps=1e-6, **kw):
    return rms_norm(x, weight, eps, **kw)


def embedding_lookup(table, ids, vocab_offset=0, vocab_offsets=(), **kw):
    """vocab_offset: first row of this rank's shard of a vocab-parallel table; vocab_offsets: the same per strategy"""
    return _op1("embedding_lookup", [table, ids], {"vocab_offset": int(vocab_offset), "vocab_offsets": [int(v) for v in vocab_offsets]}, **kw)


def dropout(x, p=0.5, inplace=False, **kw):
    if p <= 0:
        return x
    return _op1("dropout", [x], {"p": float(p)}, **kw)


dropout_ = dropout
dropout2d = dropout
dropout2d_ = dropout


def rotary(x, positions=None, base=10000.0, rot_dim=0, pos_offset=0, **kw):
    ins = [x] if positions is None else [x, positions]
    return _op1("rotary", ins, {"base": float(base), "rot_dim": int(rot_dim), "pos_offset": int(pos_offset)}, **kw)


def _attn_with_dropout(q, k, v, p_dropout, softmax_scale, is_causal):
    """attention with dropout on the softmax probabilities, composed from graph ops (scores -> mask -> softmax -> dropout -> PV).
    The fused flash kernels have no in-kernel dropout, so p_dropout > 0 takes this path: same semantics as the reference's
    flash-attention dropout (inverted dropout on P, mask replayed in backward from the op's Philox counters), O(S^2) memory."""
    b, s, h, d = q.shape
    sk, hkv = k.shape[1], k.shape[2]
    g = h // hkv
    scale = float(softmax_scale) if softmax_scale and softmax_scale > 0 else 1.0 / float(d) ** 0.5
    qt = reshape(transpose(q, [0, 2, 1, 3]), [b * h, s, d])

    def heads(t):                                     # [B, Sk, Hkv, D] -> [B * H, Sk, D] with every kv head repeated g times
        t = transpose(t, [0, 2, 1, 3])
        if g > 1:
            t = broadcast(reshape(t, [b, hkv, 1, sk, d]), [b, hkv, g, sk, d])
        return reshape(t, [b * h, sk, d])
    kt, vt = heads(k), heads(v)
    scores = bmm(qt, transpose(kt, [0, 2, 1])) * scale
    if is_causal:
        future = triu(ones_like(scores), False, 1 + (sk - s))      # 1 above the (bottom-right aligned) diagonal
        scores = masked_fill(scores, future, -1e30)
    probs = dropout(softmax(scores, -1), float(p_dropout))
    out = bmm(probs, vt)
    return transpose(reshape(out, [b, h, s, d]), [0, 2, 1, 3]), probs


def attn(q, k, v, p_dropout=0.0, softmax_scale=-1.0, is_causal=True, return_softmax=False, **kw):
    """q [B,S,H,D], k/v [B,S,Hkv,D] -> [B,S,H,D] (flash attention; tcgen05 kernels on B200).  p_dropout > 0 (training-time
    attention dropout) runs the composed path `_attn_with_dropout`."""
    if p_dropout and float(p_dropout) > 0.0:
        out, probs = _attn_with_dropout(q, k, v, float(p_dropout), softmax_scale, is_causal)
        return (out, probs) if return_softmax else out
    outs = make_op("attn", [q, k, v], {"causal": bool(is_causal), "softmax_scale": float(softmax_scale if softmax_scale > 0 else 0.0)},
                   **_meta(kw))
    return outs if return_softmax else outs[0]


flash_attn = attn


def attn_qkvpacked(qkv, num_heads, num_kv_heads=None, head_dim=None, is_causal=True, softmax_scale=-1.0, **kw):
    """qkv [B, S, (H + 2*Hkv) * D] packed projection output."""
    num_kv_heads = num_kv_heads or num_heads
    b, s, w = qkv.shape
    d = head_dim or w // (num_heads + 2 * num_kv_heads)
    q, k, v = split(qkv, [num_heads * d, num_kv_heads * d, num_kv_heads * d], dim=2)
    q = reshape(q, [b, s, num_heads, d])
    k = reshape(k, [b, s, num_kv_heads, d])
    v = reshape(v, [b, s, num_kv_heads, d])
    return attn(q, k, v, is_causal=is_causal, softmax_scale=softmax_scale, **kw)


def attn_varlen_qkvpacked(qkv, cu_seqlens, max_seqlen=None, num_heads=None, num_kv_heads=None, head_dim=None, p_dropout=0.0,
                          softmax_scale=-1.0, is_causal=True, **kw):
    """variable-length self attention over a packed batch: qkv [T, (H + 2*Hkv) * D] (or [T, 3, H, D]), cu_seqlens int32 [n + 1]
    cumulative document boundaries; every document attends only to itself -> [T, H * D]
    (ref: hetu.attn_varlen_qkvpacked / AttentionVarlenOp)"""
    from .ops_extra import attn_packed
    if p_dropout and float(p_dropout) > 0.0:
        raise NotImplementedError("attention dropout on packed variable-length rows is not implemented (the fused kernels have no in-kernel "
                                  "dropout); use p_dropout=0 here, or `attn(..., p_dropout=p)` on padded batches")
    if len(qkv.shape) == 4:                        # [T, 3, H, D]
        t, _, h, d = qkv.shape
        num_heads, num_kv_heads, head_dim = h, h, d
        qkv = reshape(qkv, [t, 3 * h * d])
    assert num_heads is not None, "num_heads is required for a 2-D packed qkv"
    return attn_packed(qkv, int(qkv.shape[0]), num_heads, num_kv_heads or num_heads, head_dim, is_causal=is_causal,
                       softmax_scale=softmax_scale, layout="qkv", cu_seqlens=cu_seqlens, **kw)


def attn_varlen(q, k, v, cu_seqlens_q, cu_seqlens_k=None, max_seqlen_q=None, max_seqlen_k=None, p_dropout=0.0, softmax_scale=-1.0,
                is_causal=True, **kw):
    """q [T, H, D], k / v [T, Hkv, D] packed along tokens, cu_seqlens_* the document boundaries (self attention: both lists
    describe the same packing) -> [T, H, D]  (ref: hetu.attn_varlen)"""
    t, h, d = q.shape
    hkv = k.shape[1]
    packed = concat([reshape(q, [t, h * d]), reshape(k, [t, hkv * d]), reshape(v, [t, hkv * d])], axis=1)
    o = attn_varlen_qkvpacked(packed, cu_seqlens_q, max_seqlen_q, num_heads=h, num_kv_heads=hkv, head_dim=d, softmax_scale=softmax_scale,
                              is_causal=is_causal, **kw)
    return reshape(o, [t, h, d])


def _slice_group_attrs(dim, offsets, lengths, groups):
    return {"dim": int(dim), "offsets": [int(o) for o in offsets], "lengths": [int(n) for n in lengths],
            "group_sizes": [len(g) for g in groups], "ranks_flat": [int(r) for g in groups for r in g]}


def grouped_all_reduce(x, dim, offsets, lengths, groups, bcast_ranks=(), **kw):
    """slice i = x[offsets[i] : offsets[i] + lengths[i]] along `dim` is all-reduced over rank group groups[i]; optionally the
    result is broadcast inside `bcast_ranks` (heterogeneous data parallelism, ref: SplitAllReduceOp)"""
    a = _slice_group_attrs(dim, offsets, lengths, groups)
    a["bcast_ranks"] = [int(r) for r in bcast_ranks]
    return _op1("grouped_all_reduce", [x], a, **kw)


def grouped_reduce_scatter(x, dim, offsets, lengths, groups, **kw):
    """slice-wise reduce-scatter: every member of groups[i] keeps 1 / |group| of reduced slice i (ref: SplitReduceScatterOp)"""
    return _op1("grouped_reduce_scatter", [x], _slice_group_attrs(dim, offsets, lengths, groups), **kw)


def grouped_all_gather(x, dim, offsets, lengths, groups, **kw):
    """inverse of grouped_reduce_scatter: parts are all-gathered slice by slice into the full tensor (ref: SplitAllGatherOp)"""
    return _op1("grouped_all_gather", [x], _slice_group_attrs(dim, offsets, lengths, groups), **kw)


def parallel_attn(q, k, v, ranks, is_causal=True, softmax_scale=-1.0, split_pattern="SYM", cu_seqlens=None, **kw):
    """Context-parallel attention over the ring `ranks` (ref: hetu.parallel_attn / ParallelAttentionOp).
    cu_seqlens (int tensor [n + 1], boundaries in the coordinates of the WHOLE row = local length x ring size) makes it
    variable-length attention over packed rows: every document attends only to itself, across chunk and rank borders."""
    outs = make_op("parallel_attn", [q, k, v] if cu_seqlens is None else [q, k, v, cu_seqlens],
                   {"causal": bool(is_causal), "softmax_scale": float(softmax_scale if softmax_scale > 0 else 0.0),
                    "ranks": [int(r) for r in ranks], "split_pattern": split_pattern}, **_meta(kw))
    return outs[0]


def conv2d(x, w, bias=None, padding=0, stride=1, **kw):
    ins = [x, w] + ([bias] if bias is not None else [])
    return _op1("conv2d", ins, {"padding": int(padding), "stride": int(stride)}, **kw)


def avgpool(x, kernel_H, kernel_W, padding=0, stride=1, **kw):
    return _op1("avgpool", [x], {"kernel_H": int(kernel_H), "kernel_W": int(kernel_W), "padding": int(padding), "stride": int(stride)}, **kw)


def maxpool(x, kernel_H, kernel_W, padding=0, stride=1, **kw):
    return _op1("maxpool", [x], {"kernel_H": int(kernel_H), "kernel_W": int(kernel_W), "padding": int(padding), "stride": int(stride)}, **kw)


def batch_norm(x, scale, bias, running_mean, running_var, momentum=0.1, eps=1e-5, **kw):
    return _op1("batch_norm", [x, scale, bias, running_mean, running_var], {"momentum": float(momentum), "eps": float(eps)}, **kw)


def instance_norm(x, eps=1e-7, **kw):
    return _op1("instance_norm", [x], {"eps": float(eps)}, **kw)


# ----------------------------------------------------------------------------- losses
def softmax_cross_entropy(logits, labels, reduction="mean", **kw):
    return _op1("softmax_cross_entropy", [logits, labels], {"reduction": reduction}, **kw)


def softmax_cross_entropy_sparse(logits, labels, ignored_index=-1, reduction="mean", **kw):
    return make_op("softmax_cross_entropy_sparse", [logits, labels], {"ignore_index": int(ignored_index), "reduction": reduction},
                   **_meta(kw))[0]


def vocab_parallel_cross_entropy(logits, labels, ignored_index=-1, reduction="mean", ranks=(), **kw):
    attrs = {"ignore_index": int(ignored_index), "reduction": reduction}
    if len(ranks):
        attrs["ranks"] = [int(r) for r in ranks]
    return make_op("vocab_parallel_cross_entropy", [logits, labels], attrs, **_meta(kw))[0]


def mse_loss(pred, target, reduction="mean", **kw):
    return _op1("mse_loss", [pred, target], {"reduction": reduction}, **kw)


def binary_cross_entropy(pred, target, reduction="mean", **kw):
    return _op1("binary_cross_entropy", [pred, target], {"reduction": reduction}, **kw)


def nll_loss(pred, target, reduction="mean", **kw):
    return _op1("nll_loss", [pred, target], {"reduction": reduction}, **kw)


def kl_div(pred, target, reduction="mean", **kw):
    return _op1("kl_div", [pred, target], {"reduction": reduction}, **kw)


# ----------------------------------------------------------------------------- communication
def comm(x, dst_ds_hierarchy, **kw):
    """Declare the target layout of `x`; the executor lowers it to all-reduce / all-gather / reduce-scatter /
    scatter / P2P / batched send-recv (ref: hetu.comm, hetu/graph/ops/Communication.cc)."""
    return make_op("comm", [x], {}, dst_ds=_normalize_dsh(dst_ds_hierarchy), **_meta(kw))[0]


def all_reduce(x, ranks, reduction="sum", **kw):
    return _op1("all_reduce", [x], {"ranks": [int(r) for r in ranks], "reduction": reduction}, **kw)


def all_gather(x, ranks, dim=0, **kw):
    return _op1("all_gather", [x], {"ranks": [int(r) for r in ranks], "dim": int(dim)}, **kw)


def reduce_scatter(x, ranks, dim=0, **kw):
    return _op1("reduce_scatter", [x], {"ranks": [int(r) for r in ranks], "dim": int(dim)}, **kw)


def dropout_add_norm(x, gamma, beta=None, residual=None, p=0.0, eps=1e-5, rms=False, **kw):
    """fused  z = residual + dropout(x);  y = norm(z)  ->  (y, z)   (z is the next block's residual stream)
    (ref: hetu/impl/kernel/RMSNorm.cu:90,257 DropoutAddLn{Fwd,Bwd}Cuda; hetu.fused_layernorm / fused_rmsnorm with residual)"""
    ins = [x] + ([residual] if residual is not None else []) + [gamma] + ([beta] if (beta is not None and not rms) else [])
    outs = make_op("dropout_add_norm", ins, {"rms": bool(rms), "eps": float(eps), "p": float(p), "has_residual": residual is not None},
                   **_meta(kw))
    return outs[0], outs[1]


def dropout_add_layer_norm(x, residual, gamma, beta, p=0.0, eps=1e-5, **kw):
    return dropout_add_norm(x, gamma, beta, residual, p, eps, rms=False, **kw)


def dropout_add_rms_norm(x, residual, gamma, p=0.0, eps=1e-5, **kw):
    return dropout_add_norm(x, gamma, None, residual, p, eps, rms=True, **kw)


def hall_to_all(x, ranks, gpus_per_node, **kw):
    """hierarchical (intra-node, then inter-node) all-to-all of the dim-0 chunks; same result as all_to_all
    (ref: hetu/v1 halltoall_op)"""
    return _op1("hall_to_all", [x], {"ranks": [int(r) for r in ranks], "gpus_per_node": int(gpus_per_node)}, **kw)


halltoall = hall_to_all


def all_to_all(x, ranks, split_dim=0, concat_dim=0, **kw):
    return _op1("all_to_all", [x], {"ranks": [int(r) for r in ranks], "split_dim": int(split_dim), "concat_dim": int(concat_dim)}, **kw)


def group(tensors, **kw):
    return _op1("group", list(tensors), **kw)


# ----------------------------------------------------------------------------- MoE
def moe_gate(logits, k=1, capacity=None, capacity_factor=1.0, **kw):
    tokens, experts = logits.shape
    if capacity is None:
        capacity = int(math.ceil(k * tokens / experts * capacity_factor))
    return make_op("moe_gate", [logits], {"k": int(k), "capacity": int(capacity)}, **_meta(kw))


def moe_gate_values(logits, idx, loc, **kw):
    return _op1("moe_gate_values", [logits, idx, loc], **kw)


def moe_dispatch(x, idx, loc, experts, capacity, scale=None, ep_ranks=(), **kw):
    """layout transform [T, H] -> [E, C, H]; with `ep_ranks` (expert parallel group) the dispatch all-to-all is fused in:
    rows are stored straight into the expert's rank over NVLink and the result is [E / ep, ep * C, H]"""
    ins = [x, idx, loc] + ([scale] if scale is not None else [])
    return _op1("moe_dispatch", ins, {"experts": int(experts), "capacity": int(capacity), "ep_ranks": [int(r) for r in ep_ranks]}, **kw)


def moe_combine(expert_out, idx, loc, gates=None, ep_ranks=(), **kw):
    """reverse layout transform (+ gate weighting); with `ep_ranks` the combine all-to-all is fused in (peer loads)"""
    ins = [expert_out, idx, loc] + ([gates] if gates is not None else [])
    return _op1("moe_combine", ins, {"ep_ranks": [int(r) for r in ep_ranks]}, **kw)


def compare(x, mode, value, dtype="float32", **kw):
    """0/1 mask of x <mode> value, mode in lt | le | gt | ge | eq | ne"""
    return _op1("compare_scalar", [x], {"mode": str(mode), "value": float(value), "dtype": dtype_name(dtype)}, **kw)


def less(x, v, **kw): return compare(x, "lt", v, **kw)              # noqa: E704
def less_equal(x, v, **kw): return compare(x, "le", v, **kw)        # noqa: E704
def greater(x, v, **kw): return compare(x, "gt", v, **kw)           # noqa: E704
def greater_equal(x, v, **kw): return compare(x, "ge", v, **kw)     # noqa: E704
def equal(x, v, **kw): return compare(x, "eq", v, **kw)             # noqa: E704
def not_equal(x, v, **kw): return compare(x, "ne", v, **kw)         # noqa: E704


def scaled_masked_softmax(x, mask=None, scale=1.0, **kw):
    """softmax(scale * x) over the last dim of [B, H, Sq, Sk] with an optional boolean mask [B, 1, Sq, Sk] (true = masked)"""
    return _op1("scaled_masked_softmax", [x] + ([mask] if mask is not None else []), {"scale": float(scale), "causal": False}, **kw)


def scaled_upper_triang_masked_softmax(x, scale=1.0, **kw):
    """causal variant: entries above the (bottom-right aligned) diagonal are masked"""
    return _op1("scaled_masked_softmax", [x], {"scale": float(scale), "causal": True}, **kw)


def stop_gradient(x, **kw):
    return _op1("stop_gradient", [x], **kw)


detach = stop_gradient


def cast(x, dtype, **kw):
    return data_transfer(x, dtype, **kw)


def permute(x, dims, **kw):
    return transpose(x, list(dims), **kw)


# ----------------------------------------------------------------------------- id arithmetic / sparse
def remainder(x, divisor, **kw):
    return _op1("remainder", [x], {"divisor": int(divisor)}, **kw)


def floor_divide(x, divisor, **kw):
    return _op1("floor_divide", [x], {"divisor": int(divisor)}, **kw)


def hash_ids(ids, buckets, a=1000003, b=12345, p=2147483647, **kw):
    """universal hash ((a * id + b) mod p) mod buckets of an integer id tensor"""
    return _op1("hash_ids", [ids], {"buckets": int(buckets), "a": int(a), "b": int(b), "p": int(p)}, **kw)


def spmm(indices, values, dense, rows, **kw):
    """sparse (COO: indices [2, nnz], values [nnz], `rows` x dense.shape[0]) times dense"""
    return _op1("spmm", [indices, values, dense], {"rows": int(rows)}, **kw)


csrmm = spmm


# ----------------------------------------------------------------------------- quantization (blockwise absmax)
def quantization(x, dtype="int8", blocksize=64, **kw):
    from .utils.quant import quantize_blockwise_op
    return quantize_blockwise_op(x, dtype, blocksize, **kw)


def dequantization(q, absmax, dtype="float32", blocksize=64, shape=None, quant_type=None, **kw):
    from .utils.quant import dequantize_blockwise_op
    return dequantize_blockwise_op(q, absmax, dtype, blocksize, shape, quant_type, **kw)


def matmul4bit(x, w_q, absmax, blocksize=64, quant_type="nf4", weight_shape=None, **kw):
    from .utils.quant import matmul4bit_op
    return matmul4bit_op(x, w_q, absmax, blocksize, quant_type, weight_shape, **kw)


# ----------------------------------------------------------------------------- in-place aliases (graph semantics: functional)
abs_ = abs
add_ = add
sub_ = sub
mul_ = mul
div_ = div
neg_ = neg
exp_ = exp
log_ = log
sqrt_ = sqrt
rsqrt_ = rsqrt
sin_ = sin
ceil_ = ceil
floor_ = floor
round_ = round
pow_ = pow
relu_ = relu
sigmoid_ = sigmoid
tanh_ = tanh
leakyrelu_ = leakyrelu
reciprocal_ = reciprocal
where_ = where

OP_NAMES = [n for n, v in list(globals().items()) if callable(v) and not n.startswith("_") and n not in (
    "List", "Optional", "Sequence", "make_op", "from_numpy", "cur_graph", "dtype_name", "DistributedStates", "IntSymbol", "Tensor")]


# ----------------------------------------------------------------------------- Tensor methods
def _install_tensor_methods():
    T = Tensor
    self_ops = ["abs", "ceil", "floor", "round", "exp", "log", "sqrt", "rsqrt", "sin", "cos", "reciprocal", "sigmoid", "tanh", "relu",
                "gelu", "silu", "neg", "pow", "reshape", "transpose", "slice", "split", "broadcast", "repeat", "roll", "pad", "gather",
                "softmax", "sum", "mean", "reduce", "norm", "contiguous", "clamp", "masked_fill", "triu", "where", "dropout",
                "leakyrelu", "elu", "swiglu", "data_transfer", "diagonal", "as_strided", "onehot"]
    g = globals()
    for n in self_ops:
        setattr(T, n, (lambda fn: lambda self, *a, **k: fn(self, *a, **k))(g[n]))
    # every remaining public op is reachable as a method too (the tensor is the op's first argument), as in the reference's
    # generated bindings (ref: python/hetu/_binding/codegen/ops.yml `self` entries)
    for n in OP_NAMES:
        if not hasattr(T, n) and not n.endswith("_initializer") and n not in ("placeholder", "parallel_placeholder", "parameter", "parallel_parameter"):
            setattr(T, n, (lambda fn: lambda self, *a, **k: fn(self, *a, **k))(g[n]))
    T.__add__ = lambda a, b: add(a, b)
    T.__radd__ = lambda a, b: add(b, a)
    T.__sub__ = lambda a, b: sub(a, b)
    T.__rsub__ = lambda a, b: sub(b, a)
    T.__mul__ = lambda a, b: mul(a, b)
    T.__rmul__ = lambda a, b: mul(b, a)
    T.__truediv__ = lambda a, b: div(a, b)
    T.__rtruediv__ = lambda a, b: div(b, a)
    T.__neg__ = lambda a: neg(a)
    T.__matmul__ = lambda a, b: matmul(a, b)
    T.to = lambda self, dtype=None, **k: data_transfer(self, dtype) if dtype is not None else self

    def numpy(self, force=False):
        from .core import _graphs_by_id
        d = self.eager_data()
        if d is None or not d.defined() if hasattr(d, "defined") else d is None:
            gr = _graphs_by_id.get(self.graph_id)
            # define-by-run: evaluate (only) the part of the ancestry that has no cached value yet
            d = gr.materialize(self) if gr.kind == _C.GraphKind.DEFINE_BY_RUN else gr.get_param(self)
        d = d.detach().cpu()
        if d.dtype == torch.bfloat16:
            d = d.float()
        return d.numpy()

    def get_data(self):
        from .core import NDArray, _graphs_by_id
        d = self.eager_data()
        if d is None:
            gr = _graphs_by_id[self.graph_id]
            d = gr.materialize(self) if gr.kind == _C.GraphKind.DEFINE_BY_RUN else gr.get_param(self)
        return NDArray(d)

    def reset_data(self, value):
        from .core import NDArray, _graphs_by_id, _to_torch
        _graphs_by_id[self.graph_id].set_param(self, _to_torch(value.t if isinstance(value, NDArray) else value))

    def backward(self, grad=None):
        from .core import _graphs_by_id
        _graphs_by_id[self.graph_id].backward(self, grad)

    def reset_data_from_splits(self, provided_datas):
        """the local shard given as its ordered pieces (split-checkpoint loader): concatenated along the tensor's split
        dimension (ref: ResetVariableDataFromSplits, hetu/graph/graph.h:1227)"""
        from .core import NDArray, _to_torch
        parts = [_to_torch(p.t if isinstance(p, NDArray) else p) for p in provided_datas]
        ds = self.distributed_states
        dims = [d for d, n in (dict(ds.states).items() if ds is not None else []) if d >= 0 and n > 1]
        dim = dims[0] if len(dims) == 1 else 0
        full = parts[0] if len(parts) == 1 else torch.cat(parts, dim=dim)
        reset_data(self, full.reshape(list(self.shape)))

    def timecost(self, micro_batch_id=0):
        """ms the producer of this tensor took in micro-batch `micro_batch_id` of the last profiled run
        (`with hetu.profiler(): graph.run(...)`; ref: Tensor.timecost / OpDef::TimeCost)"""
        from .core import _graphs_by_id
        key = self.producer_name
        hits = [ms for name, ms in _graphs_by_id[self.graph_id].op_times() if name == key]
        if not hits:
            raise RuntimeError(f"no timing recorded for {key}: run the graph inside `with hetu.profiler():`")
        return hits[micro_batch_id if micro_batch_id < len(hits) else len(hits) - 1]

    def _device(self):
        from .distributed import local_device
        dg = self.device_group
        dev = local_device()
        return dev if (dg.empty or dg.contains(dev)) else None

    T.numpy = numpy
    T.reset_data_from_splits = reset_data_from_splits
    T.timecost = timecost
    T.device = property(_device)
    T.graph = property(lambda self: __import__("hetu_b200.core", fromlist=["_graphs_by_id"])._graphs_by_id.get(self.graph_id))
    T.get_data = get_data
    T.reset_data = reset_data
    T.backward = backward
    T.symbolic = lambda self: self.symbolic_shape
    T.get_device_group_union = lambda self: self.device_group
    T.check_ds_hierarchy_equal = lambda self, other: all(
        a.check_equal(b) for a, b in zip(self.ds_hierarchy, _normalize_dsh(other)))


_install_tensor_methods()
