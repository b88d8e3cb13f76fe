"""Plain (non-parallel) layers (ref: python/hetu/nn/modules/{linear,normalization,activation,loss,sparse,dropout,conv}.py)."""
from __future__ import annotations

import math

from .. import ops
from ..core import (he_uniform_initializer, normal_initializer, ones_initializer, parallel_parameter, uniform_initializer,
                    xavier_normal_initializer, zeros_initializer)
from .module import Module

__all__ = ["Linear", "Embedding", "LayerNorm", "RMSNorm", "Dropout", "Dropout2d", "ReLU", "GELU", "SiLU", "Sigmoid", "Tanh",
           "LeakyReLU", "Softmax", "Identity", "Conv2d", "MaxPool2d", "AvgPool2d", "BatchNorm", "InstanceNorm", "MSELoss",
           "BCELoss", "NLLLoss", "KLDivLoss", "CrossEntropyLoss", "SoftmaxCrossEntropySparse"]


class Identity(Module):
    def forward(self, x):
        return x


class Linear(Module):
    def __init__(self, in_features, out_features, bias=True, dtype="float32", name="linear"):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        bound = 1.0 / math.sqrt(in_features)
        self.weight = parallel_parameter(he_uniform_initializer(gain=math.sqrt(1.0 / 3.0)), [out_features, in_features], None,
                                         dtype=dtype, requires_grad=True, name=f"{name}_weight")
        self.bias = parallel_parameter(uniform_initializer(-bound, bound), [out_features], None, dtype=dtype, requires_grad=True,
                                       name=f"{name}_bias") if bias else None

    def forward(self, x, act="none", residual=None):
        return ops.linear(x, self.weight, self.bias, trans_b=True, act=act, residual=residual)

    def extra_repr(self):
        return f"{self.in_features}, {self.out_features}, bias={self.bias is not None}"


class Embedding(Module):
    def __init__(self, num_embeddings, embedding_dim, dtype="float32", name="embedding"):
        super().__init__()
        self.num_embeddings, self.embedding_dim = num_embeddings, embedding_dim
        self.weight = parallel_parameter(xavier_normal_initializer(), [num_embeddings, embedding_dim], None, dtype=dtype,
                                         requires_grad=True, name=f"{name}_weight")

    def forward(self, ids):
        return ops.embedding_lookup(self.weight, ids)


class LayerNorm(Module):
    def __init__(self, normalized_shape, eps=1e-5, dtype="float32", name="ln"):
        super().__init__()
        n = normalized_shape if isinstance(normalized_shape, int) else normalized_shape[-1]
        self.eps = eps
        self.weight = parallel_parameter(ones_initializer(), [n], None, dtype=dtype, requires_grad=True, name=f"{name}_weight")
        self.bias = parallel_parameter(zeros_initializer(), [n], None, dtype=dtype, requires_grad=True, name=f"{name}_bias")

    def forward(self, x):
        return ops.layer_norm(x, self.weight, self.bias, eps=self.eps)


class RMSNorm(Module):
    def __init__(self, normalized_shape, eps=1e-6, dtype="float32", name="rmsnorm"):
        super().__init__()
        n = normalized_shape if isinstance(normalized_shape, int) else normalized_shape[-1]
        self.eps = eps
        self.weight = parallel_parameter(ones_initializer(), [n], None, dtype=dtype, requires_grad=True, name=f"{name}_weight")

    def forward(self, x):
        return ops.rms_norm(x, self.weight, eps=self.eps)


class Dropout(Module):
    def __init__(self, p=0.5, inplace=False):
        super().__init__()
        self.p = p

    def forward(self, x):
        return ops.dropout(x, self.p) if self.training and self.p > 0 else x


Dropout2d = Dropout


class _Act(Module):
    fn = None

    def forward(self, x):
        return type(self).fn(x)


class ReLU(_Act):
    fn = staticmethod(ops.relu)


class GELU(_Act):
    fn = staticmethod(ops.gelu)


class SiLU(_Act):
    fn = staticmethod(ops.silu)


class Sigmoid(_Act):
    fn = staticmethod(ops.sigmoid)


class Tanh(_Act):
    fn = staticmethod(ops.tanh)


class LeakyReLU(Module):
    def __init__(self, negative_slope=0.01):
        super().__init__()
        self.alpha = negative_slope

    def forward(self, x):
        return ops.leakyrelu(x, self.alpha)


class Softmax(Module):
    def __init__(self, dim=-1):
        super().__init__()
        self.dim = dim

    def forward(self, x):
        return ops.softmax(x, self.dim)


class Conv2d(Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, bias=True, name="conv"):
        super().__init__()
        k = kernel_size if isinstance(kernel_size, int) else kernel_size[0]
        self.stride, self.padding = stride, padding
        self.weight = parallel_parameter(he_uniform_initializer(), [out_channels, in_channels, k, k], None, requires_grad=True,
                                         name=f"{name}_weight")
        self.bias = parallel_parameter(zeros_initializer(), [out_channels], None, requires_grad=True, name=f"{name}_bias") if bias else None

    def forward(self, x):
        return ops.conv2d(x, self.weight, self.bias, padding=self.padding, stride=self.stride)


class MaxPool2d(Module):
    def __init__(self, kernel_size, stride=None, padding=0):
        super().__init__()
        self.k, self.s, self.p = kernel_size, stride or kernel_size, padding

    def forward(self, x):
        return ops.maxpool(x, self.k, self.k, self.p, self.s)


class AvgPool2d(MaxPool2d):
    def forward(self, x):
        return ops.avgpool(x, self.k, self.k, self.p, self.s)


class BatchNorm(Module):
    def __init__(self, num_features, eps=1e-5, momentum=0.1, name="bn"):
        super().__init__()
        self.eps, self.momentum = eps, momentum
        self.weight = parallel_parameter(ones_initializer(), [num_features], None, requires_grad=True, name=f"{name}_weight")
        self.bias = parallel_parameter(zeros_initializer(), [num_features], None, requires_grad=True, name=f"{name}_bias")
        self.register_buffer("running_mean", parallel_parameter(zeros_initializer(), [num_features], None, name=f"{name}_rm"))
        self.register_buffer("running_var", parallel_parameter(ones_initializer(), [num_features], None, name=f"{name}_rv"))

    def forward(self, x):
        return ops.batch_norm(x, self.weight, self.bias, self.running_mean, self.running_var, self.momentum, self.eps)


class InstanceNorm(Module):
    def __init__(self, num_features=None, eps=1e-7):
        super().__init__()
        self.eps = eps

    def forward(self, x):
        return ops.instance_norm(x, self.eps)


class _Loss(Module):
    def __init__(self, reduction="mean"):
        super().__init__()
        self.reduction = reduction


class MSELoss(_Loss):
    def forward(self, pred, target):
        return ops.mse_loss(pred, target, self.reduction)


class BCELoss(_Loss):
    def forward(self, pred, target):
        return ops.binary_cross_entropy(pred, target, self.reduction)


class NLLLoss(_Loss):
    def forward(self, pred, target):
        return ops.nll_loss(pred, target, self.reduction)


class KLDivLoss(_Loss):
    def forward(self, pred, target):
        return ops.kl_div(pred, target, self.reduction)


class CrossEntropyLoss(_Loss):
    def forward(self, logits, onehot_labels):
        return ops.softmax_cross_entropy(logits, onehot_labels, self.reduction)


class SoftmaxCrossEntropySparse(_Loss):
    def __init__(self, reduction="mean", ignored_index=-1):
        super().__init__(reduction)
        self.ignored_index = ignored_index

    def forward(self, logits, labels):
        return ops.softmax_cross_entropy_sparse(logits, labels, self.ignored_index, self.reduction)


class NewGeLU(Module):
    """tanh approximation used by GPT-2 (ref: nn/modules/activation.py NewGeLU)"""

    def forward(self, x):
        c = 0.7978845608028654   # sqrt(2 / pi)
        return x * 0.5 * (ops.tanh((x + ops.pow(x, 3.0) * 0.044715) * c) + 1.0)


class ConstantPad2d(Module):
    """pad the last two dimensions: padding = int or (left, right, top, bottom) (ref: nn/modules/padding.py)"""

    def __init__(self, padding, value: float = 0.0):
        super().__init__()
        self.padding = [int(padding)] * 4 if isinstance(padding, int) else [int(p) for p in padding]
        assert len(self.padding) == 4
        self.value = float(value)

    def forward(self, x):
        return ops.pad(x, self.padding, "constant", self.value)


class ZeroPad2d(ConstantPad2d):
    def __init__(self, padding):
        super().__init__(padding, 0.0)


__all__ += ["NewGeLU", "ConstantPad2d", "ZeroPad2d"]


def _nlist(x, n: int):
    """an int (or a 1-element sequence) repeated n times, an n-sequence as a list (ref: python/hetu/nn/modules/utils.py)"""
    if isinstance(x, (list, tuple)):
        x = list(x)
        return x * n if len(x) == 1 else x
    return [x] * n
