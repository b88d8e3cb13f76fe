"""Optimizers (ref: python/hetu/optim, hetu/graph/optim/optimizer.{h,cc}, optimizerParamScheduler.h).

`minimize(loss)` = compute gradients -> (ZeRO) re-layout of optimizer states -> one update op per parameter,
grouped by a `group` op.  Gradients whose layout differs from the parameter's (partial over the data-parallel
axis, ...) go through a `comm` op that the executor defers until all micro-batches are accumulated.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch

from .. import _C
from ..core import (DistributedStates, DistributedStatesUnion, Tensor, _graphs_by_id, constant_initializer, cur_graph,
                    parallel_parameter, zeros_initializer)
from ..ops import comm, group
from ..core import make_op


class OptimizerParamScheduler:
    """lr warmup + decay (constant / linear / cosine / inverse-square-root) and weight-decay increment schedule."""

    def __init__(self, init_lr=0.0, max_lr=1e-3, min_lr=0.0, lr_warmup_steps=0, lr_decay_steps=0, lr_decay_style="constant",
                 start_wd=0.0, end_wd=0.0, wd_incr_steps=0, wd_incr_style="constant"):
        self.init_lr, self.max_lr, self.min_lr = init_lr, max_lr, min_lr
        self.lr_warmup_steps, self.lr_decay_steps, self.lr_decay_style = lr_warmup_steps, lr_decay_steps, lr_decay_style
        self.start_wd, self.end_wd, self.wd_incr_steps, self.wd_incr_style = start_wd, end_wd, wd_incr_steps, wd_incr_style

    def get_lr(self, step: int) -> float:
        if self.lr_warmup_steps > 0 and step <= self.lr_warmup_steps:
            return self.init_lr + (self.max_lr - self.init_lr) * step / self.lr_warmup_steps
        if self.lr_decay_style == "constant" or self.lr_decay_steps <= 0:
            return self.max_lr
        if step > self.lr_decay_steps:
            return self.min_lr
        if self.lr_decay_style == "inverse-square-root":
            w = max(self.lr_warmup_steps, 1)
            return max(self.min_lr, self.max_lr * math.sqrt(w) / math.sqrt(max(step, 1)))
        ratio = (step - self.lr_warmup_steps) / max(self.lr_decay_steps - self.lr_warmup_steps, 1)
        if self.lr_decay_style == "linear":
            coeff = 1.0 - ratio
        elif self.lr_decay_style == "cosine":
            coeff = 0.5 * (math.cos(math.pi * ratio) + 1.0)
        else:
            raise ValueError(f"unknown decay style {self.lr_decay_style}")
        return self.min_lr + coeff * (self.max_lr - self.min_lr)

    def get_wd(self, step: int) -> float:
        if self.wd_incr_steps <= 0 or self.wd_incr_style == "constant":
            return self.end_wd
        ratio = min(step / self.wd_incr_steps, 1.0)
        if self.wd_incr_style == "linear":
            coeff = ratio
        elif self.wd_incr_style == "cosine":
            coeff = 0.5 * (math.cos(math.pi * (1 - ratio)) + 1.0)
        else:
            raise ValueError(f"unknown wd style {self.wd_incr_style}")
        return self.start_wd + coeff * (self.end_wd - self.start_wd)


def _zero_ds(ds: DistributedStates) -> DistributedStates:
    """ApplyZero: fold the duplicate axis into a dim-0 split (optimizer states are sharded over DP)."""
    if ds.get_dim(-1) <= 1:
        return ds
    st = dict(ds.combine_states([-1], 0))
    order = ds.combine_order([-1], 0)
    st = {k: v for k, v in st.items() if v > 1}
    return DistributedStates(ds.device_num, st, order, True)


class Optimizer:
    update_type = "sgd_update"

    def __init__(self):
        self.params: List[Tensor] = []
        self.states: Dict[int, Dict[str, Tensor]] = {}
        self.update_ops: List[Tensor] = []
        self.step_count = 0
        self._zero_geom: Dict[int, tuple] = {}   # param id -> (interval, count) of its reduce-scatter group

    # -- to be provided by subclasses
    def _make_states(self, param: Tensor, ds_hier, dgh) -> List[Tensor]:
        return []

    def _attrs(self) -> dict:
        return {}

    @staticmethod
    def _hetero_sync(p, gr, g):
        """heterogeneous strategies (one homogeneous graph per pipeline): synchronise the gradient across the pipelines
        that hold the same parameter with a different tensor-parallel degree (grouped all-reduce, see ops_comm.cc)"""
        from ..nn.parallel import HETERO_PARAMS, hetero_grad_sync_spec
        if not HETERO_PARAMS or p.name not in HETERO_PARAMS:
            return gr
        from ..distributed import rank
        spec = hetero_grad_sync_spec(p.name, list(p.shape), rank())
        if spec is None:
            return gr
        attrs = {"dim": spec["dim"], "offsets": spec["offsets"], "lengths": spec["lengths"],
                 "group_sizes": [len(r) for r in spec["groups"]], "ranks_flat": [x for r in spec["groups"] for x in r],
                 "bcast_ranks": list(spec["bcast_ranks"])}
        return make_op("grouped_all_reduce", [gr], attrs,
                       device_group_hierarchy=[[p.device_group]] if not p.device_group.empty else None, graph=g)[0]

    def minimize(self, loss: Tensor, var_list: Optional[List[Tensor]] = None, grad_loss: Optional[Tensor] = None) -> Tensor:
        g = _graphs_by_id.get(loss.graph_id) or cur_graph()
        params = list(var_list) if var_list is not None else g.parameters()
        grads = g.gradients([loss], params, [grad_loss] if grad_loss is not None else [])
        updates = []
        for p, gr in zip(params, grads):
            if gr is None:
                continue
            ds_h = list(p.ds_hierarchy)
            state_ds = ds_h
            # gradient layout -> parameter layout (all-reduce) or ZeRO layout (reduce-scatter over the replicas that hold
            # partial gradients; optimizer states are sharded the same way)
            if ds_h and any(u.size() > 0 for u in ds_h):
                target = []
                for s, u in enumerate(ds_h):
                    dsl = []
                    for d in u.ds_list:
                        gds = gr.get_ds(s)
                        if d.zero and gds is not None and gds.get_dim(-2) > 1:
                            st = {k: v for k, v in dict(gds.combine_states([-2], 0)).items() if v > 1}
                            dsl.append(DistributedStates(gds.device_num, st, gds.combine_order([-2], 0), True))
                            order = list(gds.order)
                            interval = 1
                            for o in order[order.index(-2) + 1:]:
                                interval *= gds.get_dim(o)
                            self._zero_geom[p.id] = (int(interval), int(gds.get_dim(-2)))
                        elif d.zero and (gds is None or gds.get_dim(-2) <= 1):
                            dsl.append(DistributedStates(d.device_num, {k: v for k, v in d.states.items() if v > 1}, d.order, False))
                        else:
                            dsl.append(d)
                    target.append(DistributedStatesUnion(dsl, u.hetero_dim))
                need = False
                for s, u in enumerate(target):
                    gds = gr.get_ds(s)
                    if gds is not None and not gds.check_equal(u.get(0)):
                        need = True
                if need:
                    gr = comm(gr, target, device_group_hierarchy=[[p.device_group]] if not p.device_group.empty else None)
                state_ds = target
            gr = self._hetero_sync(p, gr, g)
            states = self._make_states(p, state_ds, p.device_group)
            self.params.append(p)
            out = make_op(self.update_type, [p, gr] + states, self._attrs(),
                          device_group_hierarchy=[[p.device_group]] if not p.device_group.empty else None, graph=g)[0]
            updates.append(out)
        self.update_ops = updates
        with _in_graph(g):
            return group(updates) if updates else None

    def get_states(self, param: Tensor) -> Dict[str, Tensor]:
        return self.states.get(param.id, {})

    def set_states(self, param: Tensor, name: str, value):
        st = self.states[param.id][name]
        _graphs_by_id[st.graph_id].set_param(st, torch.as_tensor(value))

    def step_lr(self):
        """advance the lr / weight-decay schedule by one optimizer step and push the new values into the update ops (their
        bodies read `lr` / `weight_decay` at execution time); returns the lr the next run() will use
        (ref: OptimizerParamScheduler, hetu/graph/optim/optimizerParamScheduler.h)"""
        self.step_count += 1
        sched = getattr(self, "scheduler", None)
        if sched is not None:
            nxt = self.step_count + 1                       # the step about to be executed
            self.learning_rate = float(sched.get_lr(nxt))
            if hasattr(sched, "get_wd"):
                self.weight_decay = float(sched.get_wd(nxt))
        self.apply_hyper_parameters()
        return self.learning_rate

    def set_step(self, n: int):
        """position the schedule after `n` completed optimizer steps (checkpoint resume, Trainer.rebuild): the next run()
        uses the scheduled lr / weight decay of step n + 1 instead of the freshly constructed optimizer's warm-up / peak value"""
        self.step_count = int(n)
        sched = getattr(self, "scheduler", None)
        if sched is not None:
            self.learning_rate = float(sched.get_lr(self.step_count + 1))
            if hasattr(sched, "get_wd"):
                self.weight_decay = float(sched.get_wd(self.step_count + 1))
        self.apply_hyper_parameters()
        return self.learning_rate

    def set_learning_rate(self, lr: float):
        """external schedulers (e.g. the v1 lr_scheduler classes) drive the rate directly"""
        self.learning_rate = float(lr)
        self.apply_hyper_parameters()

    def apply_hyper_parameters(self):
        attrs = {"lr": float(self.learning_rate)}
        if hasattr(self, "weight_decay"):
            attrs["weight_decay"] = float(self.weight_decay)
        for u in self.update_ops:
            _graphs_by_id[u.graph_id].set_op_attrs(u.producer_id, attrs)


class _in_graph:
    def __init__(self, g):
        self.g = g

    def __enter__(self):
        from ..core import _state, _register_graph
        _state.graph_stack.append(_register_graph(self.g))

    def __exit__(self, *a):
        from ..core import _state
        _state.graph_stack.pop()
        return False


def _state_var(param: Tensor, suffix: str, ds_h, dtype="float32", shape=None, init=None):
    """optimizer-state variable with the (possibly ZeRO-sharded) layout `ds_h`"""
    g = _graphs_by_id[param.graph_id]
    with _in_graph(g):
        gshape = shape if shape is not None else (param.global_shape if ds_h else param.shape)
        dg = param.device_group
        return parallel_parameter(init or zeros_initializer(), gshape, list(ds_h) if (ds_h and shape is None) else None, dtype=dtype,
                                  requires_grad=False, device_group_hierarchy=[[dg]] if not dg.empty else None,
                                  name=f"{param.name}_{suffix}")


class SGDOptimizer(Optimizer):
    update_type = "sgd_update"

    def __init__(self, lr=0.01, momentum=0.0, nesterov=False, weight_decay=0.0, **kw):
        super().__init__()
        self.learning_rate, self.momentum, self.nesterov, self.weight_decay = lr, momentum, nesterov, weight_decay

    def _attrs(self):
        return {"lr": float(self.learning_rate), "momentum": float(self.momentum), "nesterov": bool(self.nesterov),
                "weight_decay": float(self.weight_decay)}

    def _make_states(self, param, ds_h, dgh):
        if self.momentum == 0.0:
            return []
        vel = _state_var(param, "velocity", ds_h, dtype=param.dtype)
        self.states[param.id] = {"velocity": vel}
        return [vel]


class AdamOptimizer(Optimizer):
    """Adam / AdamW with fp32 master weights when parameters are kept in bf16 (autocast)."""
    update_type = "adam_update"

    def __init__(self, lr=None, init_lr=None, max_lr=None, min_lr=0.0, lr_warmup_steps=0, lr_decay_steps=0, lr_decay_style="constant",
                 start_wd=0.0, end_wd=None, wd_incr_steps=0, wd_incr_style="constant", beta1=0.9, beta2=0.999, eps=1e-8,
                 weight_decay=0.0, **kw):
        super().__init__()
        peak = lr if lr is not None else (max_lr if max_lr is not None else 1e-3)
        self.scheduler = OptimizerParamScheduler(init_lr if init_lr is not None else (0.0 if lr_warmup_steps else peak), peak, min_lr,
                                                 lr_warmup_steps, lr_decay_steps, lr_decay_style, start_wd,
                                                 end_wd if end_wd is not None else weight_decay, wd_incr_steps, wd_incr_style)
        self.learning_rate = self.scheduler.get_lr(1) if lr_warmup_steps else peak
        self.beta1, self.beta2, self.eps = beta1, beta2, eps
        self.weight_decay = self.scheduler.get_wd(1)

    def _attrs(self):
        return {"lr": float(self.learning_rate), "beta1": float(self.beta1), "beta2": float(self.beta2), "eps": float(self.eps),
                "weight_decay": float(self.weight_decay)}

    def _make_states(self, param, ds_h, dgh):
        m = _state_var(param, "mean", ds_h)
        v = _state_var(param, "variance", ds_h)
        step = _state_var(param, "step", [], dtype="int64", shape=[1])
        st = {"mean": m, "variance": v, "step": step}
        ins = [m, v, step]
        if param.dtype != "float32":
            from ..core import Initializer
            master = _state_var(param, "master", ds_h, init=_CopyOf(param, self._zero_geom.get(param.id)))
            st["master"] = master
            ins.append(master)
        self.states[param.id] = st
        return ins


class _RuleOptimizer(Optimizer):
    """optimizers served by the `rule_update` op (csrc/graph/ops_optim.cc): AdaGrad, AMSGrad, LAMB"""
    update_type = "rule_update"
    rule = "adagrad"

    def __init__(self, lr=0.01, beta1=0.9, beta2=0.999, eps=1e-7, weight_decay=0.0, l2reg=0.0, **kw):
        super().__init__()
        self.learning_rate, self.beta1, self.beta2, self.eps, self.weight_decay, self.l2reg = lr, beta1, beta2, eps, weight_decay, l2reg

    def _attrs(self):
        return {"rule": self.rule, "lr": float(self.learning_rate), "beta1": float(self.beta1), "beta2": float(self.beta2),
                "eps": float(self.eps), "weight_decay": float(self.weight_decay), "l2": float(self.l2reg)}

    def _named_states(self, param, ds_h, names, with_step=True):
        st = {n: _state_var(param, n, ds_h) for n in names}
        if with_step:
            st["step"] = _state_var(param, "step", [], dtype="int64", shape=[1])
        self.states[param.id] = st
        return list(st.values())


class AdaGradOptimizer(_RuleOptimizer):
    """p -= lr * g / (sqrt(sum g^2) + eps)   (ref: hetu/v1/python/hetu/optimizer.py:418)"""
    rule = "adagrad"

    def __init__(self, lr=0.01, initial_accumulator_value=0.0, eps=1e-7, l2reg=0.0, **kw):
        super().__init__(lr=lr, eps=eps, l2reg=l2reg)
        self.initial_accumulator_value = float(initial_accumulator_value)

    def _make_states(self, param, ds_h, dgh):
        from ..core import constant_initializer
        acc = _state_var(param, "accumulator", ds_h, init=constant_initializer(self.initial_accumulator_value)) \
            if self.initial_accumulator_value else _state_var(param, "accumulator", ds_h)
        self.states[param.id] = {"accumulator": acc}
        return [acc]


class AMSGradOptimizer(_RuleOptimizer):
    """Adam with the running maximum of the second moment in the denominator (ref: optimizer.py:624)"""
    rule = "amsgrad"

    def _make_states(self, param, ds_h, dgh):
        return self._named_states(param, ds_h, ["mean", "variance", "max_variance"])


class LambOptimizer(_RuleOptimizer):
    """layer-wise adaptive moments: the Adam(W) direction rescaled per tensor by |p| / |update| (ref: optimizer.py:730)"""
    rule = "lamb"

    def _make_states(self, param, ds_h, dgh):
        return self._named_states(param, ds_h, ["mean", "variance"])


class AdamWOptimizer(AdamOptimizer):
    """Adam with decoupled weight decay -- what `adam_update` does with a non-zero weight_decay (ref: optimizer.py:671)"""


class _CopyOf:
    """initializer that mirrors another parameter's (bf16-rounded) initial value: fp32 master weights.
    `zero_geom` = (interval, count) of the reduce-scatter group so the shard matches the collective's chunk order."""

    def __init__(self, param: Tensor, zero_geom=None):
        self._op_id = param.producer_id
        self._geom = zero_geom
        self.data = None

    def attrs(self):
        a = {"init": "copy_of", "copy_of_op": int(self._op_id)}
        if self._geom is not None:
            a["zero_interval"], a["zero_count"] = self._geom
        return a


class GradScaler:
    """Dynamic loss scaling (ref: hetu/graph/autocast/gradscaler.h, optimizer_update.cc SGDUpdateWithGradScaler).

    `minimize(optimizer, loss)` multiplies the loss by a scale *variable* inside the graph and registers the scaler with the
    graph's executor, whose update phase un-scales the accumulated gradients, checks them for inf/nan on every rank, skips
    the optimizer step when one is found (backing the scale off) and grows the scale after `growth_interval` clean steps.
    `scale()/update()` remain for hand-written loops (eager graphs)."""

    def __init__(self, init_scale=65536.0, growth_factor=2.0, backoff_factor=0.5, growth_interval=2000, enabled=True):
        self.scale_value = float(init_scale)
        self.growth_factor, self.backoff_factor, self.growth_interval = growth_factor, backoff_factor, growth_interval
        self.enabled = enabled
        self._tracker = 0
        self._graph = None

    def scale(self, loss: Tensor) -> Tensor:
        return loss * self.scale_value if self.enabled else loss

    def get_scale(self):
        if self._graph is not None:
            self.scale_value = float(self._graph.loss_scale())
        return self.scale_value

    def found_inf(self) -> bool:
        """whether the last executed step was skipped because of a non-finite gradient"""
        return bool(self._graph.loss_scaler_state()[4]) if self._graph is not None else False

    def skipped_steps(self) -> int:
        return int(self._graph.loss_scaler_state()[3]) if self._graph is not None else 0

    def update(self, found_inf: bool):
        if not self.enabled:
            return
        if found_inf:
            self.scale_value *= self.backoff_factor
            self._tracker = 0
        else:
            self._tracker += 1
            if self._tracker >= self.growth_interval:
                self.scale_value *= self.growth_factor
                self._tracker = 0
        if self._graph is not None:
            self._graph.set_loss_scale(self.scale_value)

    def minimize(self, optimizer: Optimizer, loss: Tensor, var_list=None):
        if not self.enabled:
            return optimizer.minimize(loss, var_list)
        g = _graphs_by_id.get(loss.graph_id) or cur_graph()
        with _in_graph(g):
            scale_var = parallel_parameter(constant_initializer(self.scale_value), [1], None, dtype="float32", requires_grad=False,
                                           name="loss_scale")
            scaled = loss * scale_var
        train_op = optimizer.minimize(scaled, var_list)
        g.set_loss_scaler(scale_var, self.scale_value, float(self.growth_factor), float(self.backoff_factor), int(self.growth_interval))
        self._graph = g
        return train_op


SGD = SGDOptimizer      # ref: python/hetu/optim/sgd.py
