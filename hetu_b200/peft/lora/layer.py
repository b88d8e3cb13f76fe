"""LoRA adapters around the tensor-parallel linears: y = W x + (alpha / r) * B A x with A, B laid out so that no extra
full-size collective is needed (column-parallel: A replicated, B split like W; row-parallel: A split like W's input dim,
the tiny [tokens, r] partial product is all-reduced, B replicated).
(ref: python/hetu/peft/lora/layer.py -- HtLoRAMultiColumnParallelLinear / HtLoRAMultiRowParallelLinear and the
multi-task variants)"""
from __future__ import annotations

from typing import List, Optional

from ... import ops
from ...core import normal_initializer, parallel_parameter, zeros_initializer
from ...nn.module import Module
from ...nn.parallel import HtMultiColumnParallelLinear, HtMultiRowParallelLinear
from .config import LoraConfig


def _freeze(m: Module):
    for _, p in m.named_parameters():
        p.requires_grad = False


class LoraColumnParallelLinear(Module):
    def __init__(self, base: HtMultiColumnParallelLinear, config: LoraConfig, name: Optional[str] = None):
        super().__init__()
        self.base, self.config = base, config
        _freeze(base)
        name = name or base.name
        dt = base.weight.dtype
        r = config.rank
        self.lora_A = parallel_parameter(normal_initializer(0.0, config.init_std), [r, base.in_features], base.ds_dup(), dtype=dt,
                                         requires_grad=True, device_group_hierarchy=base.device_group_unions, name=f"{name}_lora_A")
        self.lora_B = parallel_parameter(zeros_initializer(), [base.out_features, r], base.ds_dup_split0(), dtype=dt,
                                         requires_grad=True, device_group_hierarchy=base.device_group_unions, name=f"{name}_lora_B")

    # attribute passthrough so model code written against the base module keeps working
    def __getattr__(self, k):
        try:
            return super().__getattr__(k)
        except AttributeError:
            return getattr(self.__dict__["_modules"]["base"], k)

    def forward(self, x, act="none"):
        b = self.base
        x = b._adapt(x, b.ds_split0_dup())
        h = x if self.config.lora_dropout <= 0 else ops.dropout(x, self.config.lora_dropout)
        t = ops.linear(h, self.lora_A, None, trans_b=True, device_group_hierarchy=b.device_group_unions)       # [T, r] dup over tp
        delta = ops.linear(t, self.lora_B, None, trans_b=True, device_group_hierarchy=b.device_group_unions)   # [T, out/tp]
        if act != "none":
            # activation applies to W x + delta: fold the delta in as the residual of a bias-only pass is not possible,
            # so add first and activate explicitly
            y = ops.linear(x, b.weight, b.bias, trans_b=True, device_group_hierarchy=b.device_group_unions) + delta * self.config.scaling
            y = getattr(ops, act)(y)
        else:
            y = ops.linear(x, b.weight, b.bias, trans_b=True, residual=None, device_group_hierarchy=b.device_group_unions)
            y = y + delta * self.config.scaling
        if b.gather_output:
            y = b._adapt(y, b.ds_split0_dup())
        return y


class LoraRowParallelLinear(Module):
    def __init__(self, base: HtMultiRowParallelLinear, config: LoraConfig, name: Optional[str] = None):
        super().__init__()
        self.base, self.config = base, config
        _freeze(base)
        name = name or base.name
        dt = base.weight.dtype
        r = config.rank
        self.lora_A = parallel_parameter(normal_initializer(0.0, config.init_std), [r, base.in_features], base.ds_dup_split1(), dtype=dt,
                                         requires_grad=True, device_group_hierarchy=base.device_group_unions, name=f"{name}_lora_A")
        self.lora_B = parallel_parameter(zeros_initializer(), [base.out_features, r], base.ds_w_dup(), dtype=dt,
                                         requires_grad=True, device_group_hierarchy=base.device_group_unions, name=f"{name}_lora_B")

    def __getattr__(self, k):
        try:
            return super().__getattr__(k)
        except AttributeError:
            return getattr(self.__dict__["_modules"]["base"], k)

    def forward(self, x, residual=None):
        b = self.base
        x = b._adapt(x, b.ds_split01())
        h = x if self.config.lora_dropout <= 0 else ops.dropout(x, self.config.lora_dropout)
        t = ops.linear(h, self.lora_A, None, trans_b=True, device_group_hierarchy=b.device_group_unions)   # partial over tp
        out_ds = b.ds_split0() if b.sequence_parallel else b.ds_split0_dup()
        if any(tp > 1 for tp in b.tp):
            t = ops.comm(t, out_ds)                      # [T, r]: r << hidden, a tiny collective
        delta = ops.linear(t, self.lora_B, None, trans_b=True, device_group_hierarchy=b.device_group_unions)
        y = b(x, residual=residual)
        return y + delta * self.config.scaling


class _MultiMixin:
    """multi-task LoRA: `task_mask` [tokens, num_tasks] (one-hot float) routes every token to its task's adapter"""

    def set_task_mask(self, mask):
        self._task_mask = mask

    def _route(self, deltas: List):
        # one [tokens, 1] column per task (a split follows the fed token count; a static slice would pin it)
        cols = ops.split(self._task_mask, len(deltas), dim=1) if len(deltas) > 1 else [self._task_mask]
        out = None
        for d, m in zip(deltas, cols):
            out = d * m if out is None else out + d * m
        return out


class MultiLoraColumnParallelLinear(Module, _MultiMixin):
    def __init__(self, base: HtMultiColumnParallelLinear, config: LoraConfig, name: Optional[str] = None):
        super().__init__()
        self.base, self.config = base, config
        _freeze(base)
        name = name or base.name
        dt, r = base.weight.dtype, config.rank
        self.lora_As, self.lora_Bs = [], []
        for t in range(config.num_tasks):
            a = parallel_parameter(normal_initializer(0.0, config.init_std), [r, base.in_features], base.ds_dup(), dtype=dt,
                                   requires_grad=True, device_group_hierarchy=base.device_group_unions, name=f"{name}_lora_A_task{t}")
            b = parallel_parameter(zeros_initializer(), [base.out_features, r], base.ds_dup_split0(), dtype=dt, requires_grad=True,
                                   device_group_hierarchy=base.device_group_unions, name=f"{name}_lora_B_task{t}")
            self.register_parameter(f"lora_A_task{t}", a)
            self.register_parameter(f"lora_B_task{t}", b)
            self.lora_As.append(a)
            self.lora_Bs.append(b)
        self._task_mask = None

    def __getattr__(self, k):
        try:
            return super().__getattr__(k)
        except AttributeError:
            return getattr(self.__dict__["_modules"]["base"], k)

    def forward(self, x, act="none"):
        b = self.base
        x = b._adapt(x, b.ds_split0_dup())
        y = ops.linear(x, b.weight, b.bias, trans_b=True, device_group_hierarchy=b.device_group_unions)
        deltas = [ops.linear(ops.linear(x, a, None, trans_b=True), bb, None, trans_b=True) for a, bb in zip(self.lora_As, self.lora_Bs)]
        y = y + self._route(deltas) * self.config.scaling
        if act != "none":
            y = getattr(ops, act)(y)
        if b.gather_output:
            y = b._adapt(y, b.ds_split0_dup())
        return y


class MultiLoraRowParallelLinear(Module, _MultiMixin):
    def __init__(self, base: HtMultiRowParallelLinear, config: LoraConfig, name: Optional[str] = None):
        super().__init__()
        self.base, self.config = base, config
        _freeze(base)
        name = name or base.name
        dt, r = base.weight.dtype, config.rank
        self.lora_As, self.lora_Bs = [], []
        for t in range(config.num_tasks):
            a = parallel_parameter(normal_initializer(0.0, config.init_std), [r, base.in_features], base.ds_dup_split1(), dtype=dt,
                                   requires_grad=True, device_group_hierarchy=base.device_group_unions, name=f"{name}_lora_A_task{t}")
            b = parallel_parameter(zeros_initializer(), [base.out_features, r], base.ds_w_dup(), dtype=dt, requires_grad=True,
                                   device_group_hierarchy=base.device_group_unions, name=f"{name}_lora_B_task{t}")
            self.register_parameter(f"lora_A_task{t}", a)
            self.register_parameter(f"lora_B_task{t}", b)
            self.lora_As.append(a)
            self.lora_Bs.append(b)
        self._task_mask = None

    def __getattr__(self, k):
        try:
            return super().__getattr__(k)
        except AttributeError:
            return getattr(self.__dict__["_modules"]["base"], k)

    def forward(self, x, residual=None):
        b = self.base
        x = b._adapt(x, b.ds_split01())
        out_ds = b.ds_split0() if b.sequence_parallel else b.ds_split0_dup()
        deltas = []
        for a, bb in zip(self.lora_As, self.lora_Bs):
            t = ops.linear(x, a, None, trans_b=True)
            if any(tp > 1 for tp in b.tp):
                t = ops.comm(t, out_ds)
            deltas.append(ops.linear(t, bb, None, trans_b=True))
        return b(x, residual=residual) + self._route(deltas) * self.config.scaling
