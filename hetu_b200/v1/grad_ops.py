"""The v1 API's explicit gradient-node constructors (`relu_gradient_op(x, dy)`, `conv2d_gradient_of_data_op(w, dy, x)`, ...) and its
remaining sparse / quantised-embedding / pipeline nodes.  The graph's autodiff builds gradients on its own, so nothing in this
framework needs these to train; they exist because v1 user code and v1's own layers call them directly.  Simple cases are closed
forms over library ops; the structured ones (convolution, pooling, batch-norm, interpolation, pad / slice / concat / split) are the
vector-Jacobian product of the forward node, taken from the graph with the caller's gradient as the seed.
(ref: hetu/v1/python/hetu/gpu_ops/{Relu,Conv2d,MaxPool,AvgPool,BatchNorm,Pad,Slice,Concat,Split,QuantizeEmbedding,QuantizeALPTEmb,
PipelineSend,PipelineReceive,ParameterServerCommunicate}.py)"""
from __future__ import annotations

import numpy as np

from .. import core, ops
from ..graph_api import gradients
from . import executor as _ex
from . import ops as _v1


def _vjp(y, x, dy):
    """d<dy, y>/dx"""
    return gradients([y], [x], [dy])[0]


# ---- elementwise
def relu_gradient_op(node_A, node_B, ctx=None): return ops.mul(node_B, ops.greater(node_A, 0.0))              # noqa: E704
def leaky_relu_gradient_op(node_A, node_B, alpha, ctx=None):                                                # noqa: E704
    pos = ops.greater(node_A, 0.0)
    return node_B * (pos + (1.0 - pos) * float(alpha))
def gelu_gradient_op(node_A, node_B, ctx=None): return _vjp(ops.gelu(node_A), node_A, node_B)                 # noqa: E704
def tanh_gradient_op(forward_node, output_grad, ctx=None): return output_grad * (1.0 - forward_node * forward_node)   # noqa: E704
def abs_gradient_op(node_A, node_B, ctx=None): return ops.mul(node_A, ops.sign(node_B))                       # noqa: E704  (grad, input)
def binary_step_gradient_op(node, ctx=None): return ops.zeros_like(node)                                      # noqa: E704
def log_grad_op(output_grad, input, eps=1e-7, ctx=None): return output_grad / (input + float(eps))           # noqa: E704,A002
def pow_gradient_op(node_A, node_B, eps, ctx=None):                                                         # noqa: E704
    """d(x ** eps) = eps * x ** (eps - 1) * dy   (node_A = x, node_B = dy)"""
    return node_B * (ops.pow(node_A, float(eps) - 1.0) * float(eps))
def const_pow_gradient_op(input_node, grad_node, val, ctx=None):                                            # noqa: E704
    """d(val ** x) = ln(val) * val ** x * dy"""
    return grad_node * (_v1.const_pow_op(input_node, val) * float(np.log(val)))
def softmax_gradient_op(node_y, grad, ctx=None):                                                            # noqa: E704
    return node_y * (grad - ops.sum(node_y * grad, [-1], True))
def log_softmax_gradient_op(node_y, grad, ctx=None):                                                        # noqa: E704
    """node_y = log_softmax(x):  dx = dy - softmax(x) * sum(dy)"""
    return grad - ops.exp(node_y) * ops.sum(grad, [-1], True)
def dropout_gradient_op(node_in, keep_prob, forward_node, ctx=None):                                        # noqa: E704
    """the forward's own mask, recovered from its output: dx = dy / keep_prob where the forward kept the element"""
    return node_in * ops.not_equal(forward_node, 0.0) * (1.0 / float(keep_prob))
dropout2d_gradient_op = dropout_gradient_op
def binarycrossentropywithlogits_gradient_op(node_A, node_B, node_C, ctx=None):                             # noqa: E704
    """(logits, labels, dy) -> dy * (sigmoid(logits) - labels)"""
    return node_C * (ops.sigmoid(node_A) - node_B)
def nll_loss_grad_op(output_grad, target, cols, ctx=None):                                                  # noqa: E704
    """d(-mean_i logp[i, target_i]) / dlogp: -dy / N at the target column of every row"""
    n = target.shape[0]
    return ops.onehot(target, int(cols)) * (output_grad * (-1.0 / n))
def norm_gradient_op(node, node_y, grad_y, axis, p, ctx=None):                                              # noqa: E704
    return _vjp(ops.norm(node, float(p), axis, True), node, ops.reshape(grad_y, list(ops.norm(node, float(p), axis, True).shape)))
def addmm_gradient_op(node_input, node_grad, beta=1.0, ctx=None):                                           # noqa: E704
    """gradient of the additive term: beta * dy reduced to the term's (broadcast) shape"""
    g = node_grad * float(beta)
    extra = len(g.shape) - len(node_input.shape)
    if extra > 0:
        g = ops.sum(g, list(range(extra)))
    axes = [i for i, (a, b) in enumerate(zip(node_input.shape, g.shape)) if a == 1 and b != 1]
    return ops.sum(g, axes, True) if axes else g


# ---- shape
def array_reshape_gradient_op(node_in, node_out, ctx=None): return ops.reshape(node_out, list(node_in.shape))   # noqa: E704
def repeat_gradient_op(node_input, node_grad, ctx=None):                                                    # noqa: E704
    """sum the tiles back: grad has shape reps * input.shape (leading dims added when reps is longer)"""
    gs, xs = list(node_grad.shape), list(node_input.shape)
    xs_full = [1] * (len(gs) - len(xs)) + xs
    split, axes = [], []
    for i, (g, x) in enumerate(zip(gs, xs_full)):
        split += [g // x, x]
        axes.append(2 * i)
    return ops.reshape(ops.sum(ops.reshape(node_grad, split), axes), xs)
def pad_gradient_op(node_A, paddings, mode="CONSTANT", ctx=None):                                           # noqa: E704
    begin = [int(p[0]) for p in paddings]
    size = [int(s) - int(p[0]) - int(p[1]) for s, p in zip(node_A.shape, paddings)]
    return ops.slice(node_A, begin, size)
def slice_gradient_op(node, begin, size=None, ctx=None):                                                    # noqa: E704
    """scatter the slice's gradient back into zeros of the input shape (`size` = the input's shape)"""
    assert size is not None, "slice_gradient_op needs the shape of the sliced input"
    pads = [[int(b), int(s) - int(b) - int(g)] for b, s, g in zip(begin, size, node.shape)]
    return _ex.pad_op(node, pads)
def concat_gradient_op(grad_node, input_node, axis, idx, ctx=None):                                         # noqa: E704
    """the part of the concatenation's gradient that belongs to operand idx (0: first, 1: second)"""
    shape = list(input_node.shape)
    begin = [0] * len(shape)
    if idx == 1:
        begin[axis] = grad_node.shape[axis] - shape[axis]
    return ops.slice(grad_node, begin, shape)
def concatenate_gradient_op(grad_node, input_node, axis, ctx=None, offset=None):                           # noqa: E704
    """`offset` = where input_node starts along axis (the reference sets it on the node after construction)"""
    shape = list(input_node.shape)
    begin = [0] * len(shape)
    begin[axis] = int(offset or 0)
    return ops.slice(grad_node, begin, shape)
def split_gradient_op(node, axes, indices, splits, ctx=None):                                               # noqa: E704
    """zeros everywhere except the part that the split kept"""
    pads = [[0, 0] for _ in node.shape]
    for ax, ind, sp in zip(axes, indices, splits):
        part = node.shape[ax]
        pads[ax] = [int(ind) * part, (int(sp) - int(ind) - 1) * part]
    return _ex.pad_op(node, pads)
def gather_gradient_op(input, grad, dim, index, ctx=None):                                                  # noqa: E704,A002
    return _vjp(ops.gather(input, dim, index), input, grad)
def scatter1d_grad_op(output_grad_mat, index_mat, ctx=None): return ops.embedding_lookup(output_grad_mat, index_mat)   # noqa: E704
def tril_lookup_gradient_op(array, offset=0, ctx=None):                                                     # noqa: E704
    """vector of lower-triangular entries -> [.., n, n] matrix with them in place (n from the vector length)"""
    k = array.shape[-1]
    n = next(n for n in range(1, 4096) if len(np.tril_indices(n, offset)[0]) == k)
    rows, cols = np.tril_indices(n, offset)
    place = np.zeros((k, n * n), np.float32)
    place[np.arange(k), rows * n + cols] = 1.0
    return ops.reshape(ops.matmul(ops.reshape(array, [-1, k]), core.from_numpy(place)), list(array.shape[:-1]) + [n, n])
def interpolate_grad_op(grad, input, mode="bicubic", align_corners=False, ctx=None):                       # noqa: E704,A002
    y = ops.interpolate(input, [int(grad.shape[-2]), int(grad.shape[-1])], mode, align_corners)
    return _vjp(y, input, grad)


# ---- convolution / pooling / normalisation
def conv2d_gradient_of_data_op(node_A, node_B, node_C, padding=0, stride=1, ctx=None):                      # noqa: E704
    """(filter, dy, x) -> dx"""
    return _vjp(ops.conv2d(node_C, node_A, None, padding=padding, stride=stride), node_C, node_B)
def conv2d_gradient_of_filter_op(input_X, gradient_Y, input_filter, padding=0, stride=1, ctx=None):         # noqa: E704
    return _vjp(ops.conv2d(input_X, input_filter, None, padding=padding, stride=stride), input_filter, gradient_Y)
def max_pool2d_gradient_op(node_out, node_out_gradient, node_in, kernel_H, kernel_W, padding, stride, ctx=None):   # noqa: E704
    return _vjp(ops.maxpool(node_in, kernel_H, kernel_W, padding, stride), node_in, node_out_gradient)
def avg_pool2d_gradient_op(node_out, node_out_gradient, node_in, kernel_H, kernel_W, padding, stride, ctx=None):   # noqa: E704
    return _vjp(ops.avgpool(node_in, kernel_H, kernel_W, padding, stride), node_in, node_out_gradient)


class _BNGrad:
    """the three gradients of one batch-norm node, built once and handed out by the `_of_data / _of_scale / _of_bias` constructors"""

    def __init__(self, dy, x, scale, eps):
        mean = ops.mean(x, [0, 2, 3], True) if len(x.shape) == 4 else ops.mean(x, [0], True)
        cen = x - mean
        axes = [0, 2, 3] if len(x.shape) == 4 else [0]
        var = ops.mean(cen * cen, axes, True)
        xhat = cen * ops.rsqrt(var + float(eps))
        bshape = [1, x.shape[1], 1, 1] if len(x.shape) == 4 else [1, x.shape[1]]
        y = xhat * ops.reshape(scale, bshape)
        self.dx = _vjp(y, x, dy)
        self.dscale = ops.sum(dy * xhat, axes)
        self.dbias = ops.sum(dy, axes)
        self.shape = list(x.shape)


def batch_normalization_gradient_op(out_gradient, in_node, bn_scale, forward_node=None, eps=1e-5, ctx=None):   # noqa: E704
    return _BNGrad(out_gradient, in_node, bn_scale, eps)
def batch_normalization_gradient_of_data_op(bn_gradient, in_arr=None, ctx=None): return bn_gradient.dx       # noqa: E704
def batch_normalization_gradient_of_scale_op(bn_gradient, in_scale=None, ctx=None): return bn_gradient.dscale   # noqa: E704
def batch_normalization_gradient_of_bias_op(bn_gradient, in_bias=None, ctx=None): return bn_gradient.dbias   # noqa: E704


# ---- MoE layout transforms
def layout_transform_gradient_op(input, indice, location, capacity, ctx=None):                              # noqa: E704,A002
    """dispatch's gradient: every token takes back the rows its copies were written to ([E * capacity, d] -> [tokens, d])"""
    experts = input.shape[0] // int(capacity) if len(input.shape) == 2 else input.shape[0]
    return _v1.reverse_layout_transform_no_gate_op(input, indice, location, capacity, experts)
def reverse_layout_transform_gradient_data_op(input, indices, locations, gates, capacity, num_experts, ctx=None):   # noqa: E704,A002
    """combine's gradient towards the expert outputs: dy rows, weighted by their gates, back at their expert slots"""
    y = ops.moe_dispatch(input, _v1._routing(indices), _v1._routing(locations), int(num_experts), int(capacity), scale=_v1._routing(gates, "float32"))
    return ops.reshape(y, [int(num_experts) * int(capacity), input.shape[-1]])
def reverse_layout_transform_no_gate_gradient_op(input, indices, locations, capacity, num_experts, ctx=None):   # noqa: E704,A002
    return _v1.layout_transform_op(input, indices, locations, capacity, num_experts)
def reverse_layout_transform_gradient_gate_op(combined_output, expert_output, indices, locations, capacity, ctx=None):   # noqa: E704
    """combine's gradient towards the gates: <dy_token, expert_output[slot of the token's j-th choice]> -> [T, k]"""
    idx, loc = _v1._routing(indices), _v1._routing(locations)
    experts = expert_output.shape[0] // int(capacity) if len(expert_output.shape) == 2 else expert_output.shape[0]
    eo = _v1._slots(expert_output, capacity, experts)
    return ops._op1("moe_combine_gate_grad", [combined_output, eo, idx, loc], {"ep_ranks": []})


# ---- sparse rows / de-duplication
def slice_by_matrix_op(node_A, index1, index2, ctx=None):                                                   # noqa: E704
    """x[index1[i], index2[i], :] for every i"""
    d0, d1 = node_A.shape[0], node_A.shape[1]
    flat = ops.reshape(node_A, [d0 * d1] + list(node_A.shape[2:]))
    return ops.embedding_lookup(flat, _lin_index(index1, index2, d1))
def _lin_index(i1, i2, d1):
    return ops.cast(ops.cast(i1, "float32") * float(d1) + ops.cast(i2, "float32"), "int64")
def slice_by_matrix_gradient_op(input, grad, index1, index2, ctx=None):                                     # noqa: E704,A002
    return _vjp(slice_by_matrix_op(input, index1, index2), input, grad)
def slice_assign_matrix_op(node_A, node_B, begin_A, size_A, begin_B, size_B, ctx=None):                     # noqa: E704
    """A with A[begin_A : +size_A] = B[begin_B : +size_B]"""
    piece = ops.slice(node_B, list(begin_B), list(size_B))
    pads = [[int(b), int(s) - int(b) - int(g)] for b, s, g in zip(begin_A, node_A.shape, size_A)]
    mask = np.ones(list(node_A.shape), np.float32)
    mask[tuple(slice(b, b + s) for b, s in zip(begin_A, size_A))] = 0.0
    return node_A * core.from_numpy(mask) + _ex.pad_op(piece, pads)
def sparse_set_op(table, ind, data, ctx=None):                                                              # noqa: E704
    """table with rows `ind` replaced by `data` (negative ids are skipped)"""
    return _rows_assign(table, ind, data)
def assign_with_indexedslices_op(embed, unique, newparam, ctx=None): return _rows_assign(embed, unique, newparam)   # noqa: E704
def _rows_assign(table, ids, rows):
    n = table.shape[0]
    keep = ops.cast(ops.greater_equal(ids, 0), "float32")
    safe = ops.cast(ops.cast(ids, "float32") * keep, "int64")
    hit = ops.sum(ops.onehot(safe, n) * ops.reshape(keep, [-1, 1]), [0])                # [n]: rows being replaced
    hit = ops.greater(hit, 0.0)
    scattered = ops.matmul(ops.onehot(safe, n) * ops.reshape(keep, [-1, 1]), ops.cast(rows, "float32"), trans_a=True)
    return ops.cast(ops.cast(table, "float32") * (1.0 - ops.reshape(hit, [-1, 1])) + scattered, table.dtype_name if hasattr(table, "dtype_name") else "float32")
def unique_indices_offsets_op(unique, ctx=None):                                                            # noqa: E704
    """`unique` = the (values, inverse, counts) triple of `unique_indices_op`; the id-offset structure groups the positions of
    every distinct id: (inverse, counts) is that grouping in this framework"""
    return (unique[1], unique[2]) if isinstance(unique, (list, tuple)) else unique
def deduplicate_lookup_op(lookup, idoffsets, ctx=None):                                                     # noqa: E704
    """one row per distinct id (the first occurrence's row): [n, d] -> [n_unique, d], padded with zeros to n rows"""
    inverse = idoffsets[0] if isinstance(idoffsets, (list, tuple)) else idoffsets
    n = lookup.shape[0]
    sel = ops.onehot(inverse, n)                                                        # [n, n_unique<=n]
    counts = ops.sum(sel, [0])
    return ops.matmul(sel, lookup, trans_a=True) / ops.reshape(counts + ops.equal(counts, 0.0), [-1, 1])
def deduplicate_grad_op(grad, idoffsets, ctx=None):                                                         # noqa: E704
    """the summed gradient of every distinct id: [n, d] -> [n (first n_unique used), d]"""
    inverse = idoffsets[0] if isinstance(idoffsets, (list, tuple)) else idoffsets
    return ops.matmul(ops.onehot(inverse, grad.shape[0]), grad, trans_a=True)
def sum_sparse_gradient_op(dense_shape, *pairs_or_denses, dtype=np.float32, ctx=None):                      # noqa: E704
    """sum of dense gradients and (indices, rows) sparse gradients into one dense [dense_shape] tensor"""
    total = core.from_numpy(np.zeros(list(dense_shape), dtype))
    for item in pairs_or_denses:
        if isinstance(item, (list, tuple)):
            ids, rows = item
            flat_ids = ops.reshape(ids, [-1])
            total = total + ops.matmul(ops.onehot(flat_ids, int(dense_shape[0])), ops.reshape(rows, [-1, int(dense_shape[-1])]), trans_a=True)
        else:
            total = total + item
    return total


# ---- quantised embedding tables (ref: src/ops/QuantizeEmbedding.cu, SignedQuantize.cu)
def _limits(digit, signed):
    return (-(2 ** (digit - 1)), 2 ** (digit - 1) - 1) if signed else (0, 2 ** digit - 1)
def quantized_embedding_lookup_op(embed, indices, qparams, digit, ctx=None):                                # noqa: E704
    """rows of an unsigned `digit`-bit table with per-row (scale, zero point) pairs: q * scale + zero_point"""
    q = ops.cast(ops.embedding_lookup(embed, indices), "float32")
    qp = ops.embedding_lookup(qparams, indices)
    last = len(qp.shape) - 1
    scale = ops.slice(qp, [0] * last + [0], list(qp.shape[:-1]) + [1])
    zero = ops.slice(qp, [0] * last + [1], list(qp.shape[:-1]) + [1])
    return q * scale + zero
def unified_quantized_embedding_lookup_op(embed, indices, scale, zero_point, digit, ctx=None):              # noqa: E704
    """one (scale, zero point) for the whole table; out-of-range ids give zero rows"""
    n = embed.shape[0]
    ok = ops.cast(ops.greater_equal(indices, 0), "float32") * ops.cast(ops.less(indices, n), "float32")
    safe = ops.cast(ops.cast(indices, "float32") * ok, "int64")
    rows = ops.cast(ops.embedding_lookup(embed, safe), "float32") * float(scale) + float(zero_point)
    return rows * ops.reshape(ok, list(ok.shape) + [1])
def alpt_embedding_lookup_op(embed, indices, scale, zero_point, digit, ctx=None):                           # noqa: E704
    """signed integer table with a LEARNED per-row step size (ALPT): q[id] * scale[id] + zero_point"""
    q = ops.cast(ops.embedding_lookup(embed, indices), "float32")
    return q * ops.embedding_lookup(scale, indices) + float(zero_point)
def alpt_rounding_op(lookup, scale, middle, digit, ctx=None):                                               # noqa: E704
    """LSQ fake quantisation of looked-up rows (already divided by their step): clamp, round half up, rescale; the rounding passes
    gradients straight through to `lookup` and the step size gets the LSQ gradient (`alpt_scale_gradient_op`)"""
    lo, hi = _limits(int(digit), True)
    r = ops.floor(ops.clamp(lookup, float(lo), float(hi)) + 0.5)
    hard = r * scale + float(middle)
    soft = lookup * ops.stop_gradient(scale) + ops.stop_gradient(alpt_scale_gradient_op(lookup, digit)) * scale
    return soft + ops.stop_gradient(hard - soft)
def alpt_scale_gradient_op(lookup, digit, ctx=None):                                                        # noqa: E704
    """d(quantised value)/d(step): the bound outside the range, round(v) - v inside"""
    lo, hi = _limits(int(digit), True)
    inside = ops.floor(lookup + 0.5) - lookup
    return ops.where(ops.greater_equal(lookup, float(hi)), ops.full_like(lookup, float(hi)),
                     ops.where(ops.less_equal(lookup, float(lo)), ops.full_like(lookup, float(lo)), inside))
def assign_quantized_embedding_op(embed, unique, newparam, digit, scale=None, minele=None, middle=None, qparam=None, ctx=None):   # noqa: E704,E501
    """write float rows back into a quantised table: unified (scale, minele), ALPT (per-row `scale` tensor + middle) or per-row
    qparams (re-derived from each new row, returned second)"""
    x = ops.cast(newparam, "float32")
    if qparam is not None:
        lo, hi = _limits(int(digit), False)
        mn, mx = ops.min(x, [-1], True), ops.max(x, [-1], True)
        step = (mx - mn) / float(hi)
        q = ops.clamp(ops.floor((x - mn) / (step + ops.equal(step, 0.0)) + 0.5), float(lo), float(hi))
        return _rows_assign(embed, unique, q), _rows_assign(qparam, unique, ops.concat([step, mn], -1))
    if middle is not None:
        lo, hi = _limits(int(digit), True)
        step = ops.embedding_lookup(scale, unique) if not np.isscalar(scale) else float(scale)
        q = ops.clamp(ops.floor((x - float(middle)) / step + 0.5), float(lo), float(hi))
    else:
        lo, hi = _limits(int(digit), False)
        q = ops.clamp(ops.floor((x - float(minele)) / float(scale) + 0.5), float(lo), float(hi))
    return _rows_assign(embed, unique, q)


# ---- pipeline / parameter-server nodes
def pipeline_send_op(node, destination, comm=None, ctx=None, channel=0):                                    # noqa: E704
    """activation hand-off to rank `destination`; yields a 1-element token to fetch / depend on"""
    return ops._op1("pipeline_send", [node], {"dst": int(destination), "channel": int(channel)})
def pipeline_receive_op(source, comm=None, use_indexed_slices=False, ctx=None, shape=None, dtype="float32", channel=0):   # noqa: E704
    """the tensor rank `source` sends; `shape` / `dtype` describe it (the reference infers them at run time from the sender)"""
    assert shape is not None, "pipeline_receive_op needs the shape of the incoming tensor"
    _ex._g()                                               # a node without inputs: make sure the v1 graph is the current one
    return ops._op1("pipeline_recv", [], {"src": int(source), "shape": [int(v) for v in shape], "dtype": str(dtype), "channel": int(channel)})
def allreduceCommunicatep2p_op(node, comm=None): return ops.all_reduce(node)                                  # noqa: E704,N802
def parameterServerCommunicate_op(node, parameter, optimizer):                                             # noqa: N802
    """push the gradient `node` of `parameter` to the server, which applies `optimizer` (the executor's comm_mode='PS' builds this
    itself; explicit use marks one parameter for the server path)"""
    server_opt = optimizer._server_opt() if hasattr(optimizer, "_server_opt") else ("sgd", float(getattr(optimizer, "learning_rate", 0.01)))
    return _ex.annotate(node, "ps_target", (parameter, server_opt))
def parameterServerSparsePull_op(parameter, deps_node):                                                    # noqa: N802
    """the rows of `parameter` named by `deps_node`, fetched from the server at run time"""
    out = ops.embedding_lookup(parameter, deps_node)
    return _ex.annotate(out, "ps_sparse_pull", parameter)
def distgcn_15d_op(node_A, node_B, node_C, node_Count_Self=None, node_Count_All=None, size=1, replication=1, device_id=0, comm=None,   # noqa: E704,N803
                   comm_groups=(None, None), need_W=True):
    """one 1.5-D GCN layer A @ (H @ W) on the caller's row block (`models.gnn.DistGCN15D` is the partitioned trainer); on one rank
    this is the plain product"""
    hw = ops.matmul(node_B, node_C) if need_W else node_B
    return ops.spmm(node_A, hw) if getattr(node_A, "is_sparse", False) else ops.matmul(node_A, hw)
