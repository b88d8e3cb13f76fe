"""v1 layer library: small callable objects that create their Variables and emit `*_op` nodes
(ref: hetu/v1/python/hetu/layers/{base,linear,conv,normalization,embedding,dropout,relu,gelu,mish,identity,reshape,
sequence,concatenate,sum,slice,pooling,attention,loss}.py; the MoE layers / gates live in hetu_b200.models.moe)."""
from __future__ import annotations

import math
from typing import Callable, Sequence

from .. import ops
from . import initializers as init


class BaseLayer:
    def __call__(self, *a, **k):
        raise NotImplementedError

    def make_dataloader_func(self):
        return None


class OpLayer(BaseLayer):
    """wraps any `*_op` function as a layer"""

    def __init__(self, op: Callable, *args, **kwargs):
        self.op, self.args, self.kwargs = op, args, kwargs

    def __call__(self, *x):
        return self.op(*x, *self.args, **self.kwargs)


class Linear(BaseLayer):
    def __init__(self, in_features, out_features, initializer=None, bias=True, activation=None, weight_transpose=False, name="linear"):
        shape = (out_features, in_features) if weight_transpose else (in_features, out_features)
        self.weight_transpose, self.activation = weight_transpose, activation
        self.weight_var = (initializer or init.GenXavierUniform())(shape, name=f"{name}_weight")
        self.bias_var = init.zeros((out_features,), name=f"{name}_bias") if bias else None

    def __call__(self, x):
        y = ops.linear(x, self.weight_var, self.bias_var, trans_b=self.weight_transpose)
        return _act(self.activation, y)


def _act(name, y):
    if name is None:
        return y
    if callable(name):
        return name(y)
    return {"relu": ops.relu, "gelu": ops.gelu, "sigmoid": ops.sigmoid, "tanh": ops.tanh, "mish": ops.mish, "silu": ops.silu}[name](y)


class Conv2d(BaseLayer):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, initializer=None, bias=True, activation=None, name="conv2d"):
        k = (kernel_size, kernel_size) if isinstance(kernel_size, int) else tuple(kernel_size)
        self.stride, self.padding, self.activation = stride, padding, activation
        self.weight_var = (initializer or init.GenHeUniform())((out_channels, in_channels) + k, name=f"{name}_weight")
        self.bias_var = init.zeros((out_channels,), name=f"{name}_bias") if bias else None

    def __call__(self, x):
        return _act(self.activation, ops.conv2d(x, self.weight_var, self.bias_var, padding=self.padding, stride=self.stride))


class MaxPool2d(BaseLayer):
    def __init__(self, kernel_size, stride=None, padding=0):
        self.k = (kernel_size, kernel_size) if isinstance(kernel_size, int) else tuple(kernel_size)
        self.stride, self.padding = stride or self.k[0], padding

    def __call__(self, x):
        return ops.maxpool(x, self.k[0], self.k[1], padding=self.padding, stride=self.stride)


class AvgPool2d(MaxPool2d):
    def __call__(self, x):
        return ops.avgpool(x, self.k[0], self.k[1], padding=self.padding, stride=self.stride)


class BatchNorm(BaseLayer):
    def __init__(self, num_channels, momentum=0.1, eps=1e-5, name="batchnorm"):
        self.scale, self.bias = init.ones((num_channels,), name=f"{name}_scale"), init.zeros((num_channels,), name=f"{name}_bias")
        self.mean = init.zeros((num_channels,), name=f"{name}_running_mean", trainable=False)
        self.var = init.ones((num_channels,), name=f"{name}_running_var", trainable=False)
        self.momentum, self.eps = momentum, eps

    def __call__(self, x):
        return ops.batch_norm(x, self.scale, self.bias, self.mean, self.var, momentum=self.momentum, eps=self.eps)


class LayerNorm(BaseLayer):
    def __init__(self, num_channels, eps=1e-5, name="layernorm"):
        self.scale, self.bias, self.eps = init.ones((num_channels,), name=f"{name}_scale"), init.zeros((num_channels,), name=f"{name}_bias"), eps

    def __call__(self, x):
        return ops.layer_norm(x, self.scale, self.bias, eps=self.eps)


class InstanceNorm2d(BaseLayer):
    def __init__(self, eps=1e-7):
        self.eps = eps

    def __call__(self, x):
        return ops.instance_norm(x, eps=self.eps)


class Embedding(BaseLayer):
    def __init__(self, num_embeddings, embedding_dim, initializer=None, name="embedding"):
        self.num_embeddings, self.embedding_dim = num_embeddings, embedding_dim
        self.embedding_table = (initializer or init.GenXavierNormal())((num_embeddings, embedding_dim), name=f"{name}_table")

    def __call__(self, ids):
        return ops.embedding_lookup(self.embedding_table, ids)


class DropOut(BaseLayer):
    def __init__(self, p=0.5):
        self.keep_prob = 1.0 - p

    def __call__(self, x):
        return ops.dropout(x, 1.0 - self.keep_prob)


class Relu(OpLayer):
    def __init__(self): super().__init__(ops.relu)                                 # noqa: E704


class Gelu(OpLayer):
    def __init__(self): super().__init__(ops.gelu)                                 # noqa: E704


class Mish(OpLayer):
    def __init__(self): super().__init__(ops.mish)                                 # noqa: E704


class Sigmoid(OpLayer):
    def __init__(self): super().__init__(ops.sigmoid)                              # noqa: E704


class Tanh(OpLayer):
    def __init__(self): super().__init__(ops.tanh)                                 # noqa: E704


class Identity(BaseLayer):
    def __call__(self, x):
        return x


class Reshape(BaseLayer):
    def __init__(self, shape):
        self.shape = list(shape)

    def __call__(self, x):
        return ops.reshape(x, self.shape)


class Slice(BaseLayer):
    def __init__(self, begin, size):
        self.begin, self.size = list(begin), list(size)

    def __call__(self, x):
        return ops.slice(x, self.begin, self.size)


class Sequence(BaseLayer):
    def __init__(self, *layers):
        self.layers = list(layers)

    def __call__(self, x):
        for layer in self.layers:
            x = layer(x)
        return x


class ConcatenateLayers(BaseLayer):
    """apply every branch to the input and concatenate the results"""

    def __init__(self, layers: Sequence, axis=0):
        self.layers, self.axis = list(layers), axis

    def __call__(self, x):
        return ops.concat([layer(x) for layer in self.layers], self.axis)


class SumLayers(BaseLayer):
    def __init__(self, layers: Sequence):
        self.layers = list(layers)

    def __call__(self, x):
        outs = [layer(x) for layer in self.layers]
        y = outs[0]
        for o in outs[1:]:
            y = ops.add(y, o)
        return y


class MultiHeadAttention(BaseLayer):
    """[batch * seq, hidden] token-major self attention with fused qkv projection"""

    def __init__(self, hidden_size, num_heads, seq_len, batch_size, dropout=0.0, causal=False, name="attn"):
        assert hidden_size % num_heads == 0
        self.h, self.nh, self.s, self.b, self.causal = hidden_size, num_heads, seq_len, batch_size, causal
        self.qkv = Linear(hidden_size, 3 * hidden_size, name=f"{name}_qkv")
        self.out = Linear(hidden_size, hidden_size, name=f"{name}_out")
        self.drop = DropOut(dropout) if dropout > 0 else None

    def __call__(self, x):
        d = self.h // self.nh
        qkv = ops.reshape(self.qkv(x), [self.b, self.s, 3 * self.h])
        q, k, v = ops.split(qkv, 3, dim=2)
        q, k, v = (ops.reshape(t, [self.b, self.s, self.nh, d]) for t in (q, k, v))
        o = ops.attn(q, k, v, is_causal=self.causal, softmax_scale=1.0 / math.sqrt(d))
        y = self.out(ops.reshape(o, [self.b * self.s, self.h]))
        return self.drop(y) if self.drop is not None else y


class SoftmaxCrossEntropyLoss(BaseLayer):
    def __init__(self, sparse=False, ignored_index=-1, reduce_mean=True):
        self.sparse, self.ignored_index, self.reduce_mean = sparse, ignored_index, reduce_mean

    def __call__(self, logits, labels):
        red = "mean" if self.reduce_mean else "none"
        if self.sparse:
            return ops.softmax_cross_entropy_sparse(logits, labels, ignored_index=self.ignored_index, reduction=red)
        return ops.softmax_cross_entropy(logits, labels, reduction=red)


class BCELoss(BaseLayer):
    def __call__(self, p, y):
        return ops.mean(ops.binary_cross_entropy(p, y, reduction="none"))


class MSELoss(BaseLayer):
    def __call__(self, p, y):
        return ops.mean(ops.mse_loss(p, y, reduction="none"))


# ------------------------------------------------------------------ remaining v1 layer classes
class _Reduced(BaseLayer):
    """loss layers share the reduction switch (ref: layers/loss.py BaseLossLayer)"""

    def __init__(self, reduction="mean"):
        assert reduction in ("mean", "sum", "none", None)
        self.reduction = reduction

    def reduce(self, loss):
        if self.reduction == "mean":
            return ops.mean(loss, [0])
        if self.reduction == "sum":
            return ops.sum(loss, [0])
        return loss


class MAELoss(_Reduced):
    def __call__(self, inputs, targets):
        return self.reduce(ops.abs(inputs - targets))


class BCEWithLogitsLoss(_Reduced):
    def __call__(self, inputs, targets):
        return self.reduce(ops.binary_cross_entropy(ops.sigmoid(inputs), targets, reduction="none"))


class Concatenate(BaseLayer):
    """concatenate the call's arguments along `axis` (ref: layers/concatenate.py)"""

    def __init__(self, axis):
        self.axis = axis

    def __call__(self, *args):
        return args[0] if len(args) == 1 else ops.concat(list(args), self.axis)


class BatchSplitOnlyLayer(BaseLayer):
    """a block that the v1 planners may only split along the batch dimension (ref: layers/batch_split_layer.py).  The wrapped
    sequence runs unchanged; `split_dims` is what strategy searches read."""
    split_dims = (0,)

    def __init__(self, sequence, ctx=None):
        self.sequence, self.ctx = sequence, ctx

    def __call__(self, x):
        y = self.sequence(x)
        from .executor import annotate
        return annotate(y, "layer_constraint", type(self).__name__)


class ReserveSplitLayer(BatchSplitOnlyLayer):
    """as above, and a split of the hidden dimension survives through the block's reshapes (attention-style head splits)"""
    split_dims = (0, 1)


def _moe():
    from ..models import moe
    return moe


class Expert(BaseLayer):
    """one feed-forward expert [*, d] -> [*, d] (ref: layers/moe_layer.py:7 -- two matmuls, optional bias, relu / gelu, dropout)"""

    def __init__(self, embed_dim, ffn_dim, dropout_rate=0.0, initializer=None, bias=False, activation=None, name="expert"):
        self.embed_dim, self.drop, self.activation = embed_dim, dropout_rate, activation
        self.fc1 = Linear(embed_dim, ffn_dim, initializer=initializer, bias=bias, name=f"{name}_fc1") if initializer is not None \
            else Linear(embed_dim, ffn_dim, bias=bias, name=f"{name}_fc1")
        self.fc2 = Linear(ffn_dim, embed_dim, initializer=initializer, bias=bias, name=f"{name}_fc2") if initializer is not None \
            else Linear(ffn_dim, embed_dim, bias=bias, name=f"{name}_fc2")

    def __call__(self, x):
        h = self.fc1(ops.reshape(x, [-1, self.embed_dim]))
        h = _act(self.activation, h) if self.activation else h
        if self.drop and self.drop > 0:
            h = ops.dropout(h, float(self.drop))
        return self.fc2(h)


class _GatedMoE(BaseLayer):
    """v1's MoE layers differ only in their gate; tokens [.., d] -> [.., d] through `models.moe.MoELayer`'s dispatch -> grouped
    experts -> combine (all-to-alls fused in when `ep_ranks` spans ranks).  `l_aux` holds the gate's balance loss after a call."""
    gate_type = "topk"

    def __init__(self, embed_dim, ffn_dim, num_experts, top=1, capacity_factor=1.0, ep_ranks=(), activation="gelu", name=None, **_ignored):
        self.embed_dim = embed_dim
        self.inner = _moe().MoELayer(embed_dim, ffn_dim, num_experts, k=top, capacity_factor=capacity_factor, gate_type=self.gate_type,
                                     ep_ranks=ep_ranks, act=activation, name=name or type(self).__name__.lower())
        self.l_aux = None

    def __call__(self, x):
        shape = list(x.shape)
        y = self.inner(ops.reshape(x, [-1, self.embed_dim]))
        self.l_aux = self.inner.l_aux
        return ops.reshape(y, shape)

    def parameters(self):
        return list(self.inner.parameters())


class MoELayer(_GatedMoE): gate_type = "topk"          # noqa: E701
class HashLayer(_GatedMoE): gate_type = "hash"         # noqa: E701
class KTop1Layer(_GatedMoE): gate_type = "ktop1"       # noqa: E701
class SAMLayer(_GatedMoE): gate_type = "sam"           # noqa: E701


def __getattr__(name):
    # the gates live with the MoE model family; v1 code reaches them as ht.layers.TopKGate etc.
    alias = {"TopKGate": "TopKGate", "KTop1Gate": "KTop1Gate", "HashGate": "HashGate", "SAMGate": "SAMGate",
             "BalanceAssignmentGate": "BalanceGate", "BalanceGate": "BalanceGate"}
    if name in alias:
        return getattr(_moe(), alias[name])
    raise AttributeError(name)
