"""v1 static-dataflow API: nodes are created with `*_op` functions, an `Executor` evaluates a list of nodes for a feed
dict; training nodes come from `optimizer.minimize(loss)`.  comm_mode: None (single device), 'AllReduce' (data parallel
over the process group), 'PS' (all parameters on the parameter server), 'Hybrid' (dense by all-reduce, embeddings on the
PS with the HET cache).  (ref: hetu/v1/python/hetu/gpu_ops/executor.py Executor/SubExecutor/HetuConfig)"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

from .. import core, ops
from ..core import Tensor


class _Ctx:
    def __init__(self, kind, index=0, host="localhost"):
        self.kind, self.index, self.host = kind, index, host

    # the reference's DLContext spellings
    device_id = property(lambda self: self.index)
    hostname = property(lambda self: self.host)
    local = property(lambda self: self.host in ("localhost", "127.0.0.1"))

    def __eq__(self, other):
        return isinstance(other, _Ctx) and (self.kind, self.index, self.host) == (other.kind, other.index, other.host)

    def __hash__(self):
        return hash((self.kind, self.index, self.host))

    def __repr__(self):
        return f"{self.host}:{self.kind}:{self.index}"


def cpu(i=0): return _Ctx("cpu", i)            # noqa: E704
def gpu(i=0): return _Ctx("gpu", i)            # noqa: E704
def rcpu(host, i=0): return _Ctx("cpu", i, host)   # noqa: E704
def rgpu(host, i=0): return _Ctx("gpu", i, host)   # noqa: E704


_EMBED_IDS = set()
_graph = None
_graph_ctx = None      # the context manager must stay alive, otherwise the graph is popped again


def _g():
    """all v1 nodes of a process live in one define-and-run graph"""
    global _graph, _graph_ctx
    if _graph is None:
        _graph_ctx = core.graph("define_and_run", create_new=True, prefix="v1")
        _graph = _graph_ctx.__enter__()
    return _graph


_UPDATE_TYPES = ("adam_update", "sgd_update", "rule_update", "group")
_notes: Dict = {}


def annotate(t, key, value):
    """side table for per-node facts the v1 API hangs on nodes (graph tensors are native objects without a __dict__)"""
    _notes[(t.graph_id, t.id, key)] = value
    return t


def annotation(t, key, default=None):
    return _notes.get((getattr(t, "graph_id", None), getattr(t, "id", None), key), default)


def reset_graph():
    global _graph, _graph_ctx
    _notes.clear()
    if _graph_ctx is not None:
        _graph_ctx.__exit__(None, None, None)
    _graph = _graph_ctx = None


def Variable(name, value=None, initializer=None, trainable=True, shape=None, dtype="float32", is_embed=False):
    _g()
    if value is not None:
        arr = np.asarray(value, dtype=np.float32 if dtype == "float32" else dtype)
        t = core.parameter(core.provided_initializer(arr), list(arr.shape), dtype=dtype, requires_grad=trainable, name=name)
    else:
        init = initializer or core.xavier_uniform_initializer()
        t = core.parameter(init, list(shape), dtype=dtype, requires_grad=trainable, name=name)
    if is_embed:
        _EMBED_IDS.add(t.id)         # Hybrid mode keeps these tables on the parameter server (sparse push / pull)
    return t


def placeholder_op(name, shape=None, dtype="float32", trainable=False):
    _g()
    return core.placeholder(dtype, list(shape) if shape is not None else [1], name=name)


# ---- op constructors (v1 naming)
def matmul_op(a, b, trans_A=False, trans_B=False): return ops.matmul(a, b, trans_a=trans_A, trans_b=trans_B)     # noqa: E704
def linear_op(x, w, b=None, trans_B=False): return ops.linear(x, w, b, trans_b=trans_B)                         # noqa: E704
def batch_matmul_op(a, b, trans_A=False, trans_B=False): return ops.bmm(ops.transpose(a, [0, 2, 1]) if trans_A else a, ops.transpose(b, [0, 2, 1]) if trans_B else b)   # noqa: E704,E501
def relu_op(x): return ops.relu(x)                                     # noqa: E704
def sigmoid_op(x): return ops.sigmoid(x)                               # noqa: E704
def tanh_op(x): return ops.tanh(x)                                     # noqa: E704
def gelu_op(x): return ops.gelu(x)                                     # noqa: E704
def sqrt_op(x): return ops.sqrt(x)                                     # noqa: E704
def exp_op(x): return ops.exp(x)                                       # noqa: E704
def log_op(x): return ops.log(x)                                       # noqa: E704
def softmax_op(x): return ops.softmax(x, -1)                           # noqa: E704
def add_op(a, b): return ops.add(a, b)                                 # noqa: E704
def mul_op(a, b): return ops.mul(a, b)                                 # noqa: E704
def addbyconst_op(x, c): return x + float(c)                           # noqa: E704
def mulbyconst_op(x, c): return x * float(c)                           # noqa: E704
def reduce_mean_op(x, axes, keepdims=False): return ops.mean(x, axes, keepdims)    # noqa: E704
def reduce_sum_op(x, axes, keepdims=False): return ops.sum(x, axes, keepdims)      # noqa: E704
def array_reshape_op(x, shape): return ops.reshape(x, list(shape))     # noqa: E704
def transpose_op(x, perm=None): return ops.transpose(x, list(perm) if perm is not None else list(range(len(x.shape)))[::-1])   # noqa: E704
def broadcastto_op(x, y): return ops.broadcast(x, list(y.shape))       # noqa: E704
def concat_op(a, b, axis=0): return ops.concat([a, b], axis)           # noqa: E704
def slice_op(x, begin, size): return ops.slice(x, list(begin), list(size))   # noqa: E704
def dropout_op(x, keep_prob): return ops.dropout(x, 1.0 - keep_prob)   # noqa: E704
def embedding_lookup_op(table, ids): return ops.embedding_lookup(table, ids)   # noqa: E704
def layer_normalization_op(x, scale, bias, eps=1e-5): return ops.layer_norm(x, scale, bias, eps=eps)   # noqa: E704
def softmaxcrossentropy_op(logits, labels): return ops.softmax_cross_entropy(logits, labels, reduction="none")   # noqa: E704
def softmaxcrossentropy_sparse_op(logits, labels, ignored_index=-1): return ops.softmax_cross_entropy_sparse(logits, labels, ignored_index=ignored_index, reduction="none")   # noqa: E704,E501
def binarycrossentropy_op(p, y): return ops.binary_cross_entropy(p, y, reduction="none")   # noqa: E704
def mse_op(p, y): return ops.mse_loss(p, y, reduction="none")          # noqa: E704


def conv2d_op(x, w, padding=0, stride=1): return ops.conv2d(x, w, None, padding=padding, stride=stride)                    # noqa: E704
def conv2d_add_bias_op(x, w, b, padding=0, stride=1): return ops.conv2d(x, w, b, padding=padding, stride=stride)            # noqa: E704
def max_pool2d_op(x, kernel_H, kernel_W, padding=0, stride=1): return ops.maxpool(x, kernel_H, kernel_W, padding, stride)   # noqa: E704
def avg_pool2d_op(x, kernel_H, kernel_W, padding=0, stride=1): return ops.avgpool(x, kernel_H, kernel_W, padding, stride)   # noqa: E704
def batch_normalization_op(x, scale, bias, mean, var, momentum=0.1, eps=1e-5): return ops.batch_norm(x, scale, bias, mean, var, momentum, eps)   # noqa: E704,E501
def instance_normalization2d_op(x, eps=1e-7): return ops.instance_norm(x, eps)   # noqa: E704
def pad_op(x, paddings, mode="constant", constant_values=0.0):
    """paddings: [[before, after] per dimension] (numpy style) -> the flat last-dimension-first list of the pad op"""
    flat = [int(v) for pair in reversed([list(p) for p in paddings]) for v in pair]
    return ops.pad(x, flat, mode, constant_values)
def div_op(a, b): return ops.div(a, b)                                  # noqa: E704
def minus_op(a, b): return ops.sub(a, b)                                # noqa: E704
def opposite_op(x): return ops.neg(x)                                   # noqa: E704
def abs_op(x): return ops.abs(x)                                        # noqa: E704
def pow_op(x, p): return ops.pow(x, p)                                  # noqa: E704
def rsqrt_op(x): return ops.rsqrt(x)                                    # noqa: E704
def leaky_relu_op(x, alpha=0.1): return ops.leakyrelu(x, alpha)         # noqa: E704
def mish_op(x): return ops.mish(x)                                      # noqa: E704
def silu_op(x): return ops.silu(x)                                      # noqa: E704
def where_op(c, a, b): return ops.where(c, a, b)                        # noqa: E704
def one_hot_op(x, num_classes): return ops.onehot(x, num_classes)       # noqa: E704
def reduce_max_op(x, axes, keepdims=False): return ops.reduce(x, "max", axes, keepdims)   # noqa: E704
def reduce_min_op(x, axes, keepdims=False): return ops.reduce(x, "min", axes, keepdims)   # noqa: E704
def concatenate_op(nodes, axis=0): return ops.concat(list(nodes), axis)   # noqa: E704
def split_op(x, axes, indices, splits): return ops.split(x, splits[0], dim=axes[0])[indices[0]]   # noqa: E704
def sum_op(nodes):                                                      # noqa: E704
    y = nodes[0]
    for n in nodes[1:]:
        y = ops.add(y, n)
    return y


def gradients(loss, nodes):
    from ..graph_api import gradients as g
    return g(loss, list(nodes))


class HetuConfig:
    def __init__(self, eval_node_list, ctx=None, comm_mode=None, seed=None, bsp=-1, cstable_policy=None, cache_bound=100, **kw):
        self.eval_node_list, self.context, self.comm_mode, self.seed = eval_node_list, ctx, comm_mode, seed
        self.bsp, self.cstable_policy, self.cache_bound = bsp, cstable_policy, cache_bound
        self.extra = kw


class Executor:
    """Executor({name: [nodes]} | [nodes], ctx=..., comm_mode=...).run(name?, feed_dict) -> list of numpy arrays"""

    def __init__(self, eval_node_dict, ctx=None, comm_mode=None, seed=None, **kw):
        if not isinstance(eval_node_dict, dict):
            eval_node_dict = {"default": list(eval_node_dict)}
        self.eval_node_dict = {k: list(v) for k, v in eval_node_dict.items()}
        self.config = HetuConfig(self.eval_node_dict, ctx, comm_mode, seed, **kw)
        self.graph = _g()
        self.comm_mode = comm_mode
        if seed is not None:
            core.set_seed(seed)
        self.step = 0
        self.ps = kw.get("ps")                    # a PSContext / ShardedPSContext; default: ps.connect() on first use
        self._ps_plans = {}                       # train node id -> (params on the server, their gradient nodes, server optimizer)

    # ---- parameter-server modes -----------------------------------------------------------------------------------
    # 'PS': every trainable variable lives on the server(s); a step fetches the gradients instead of running the local update ops,
    # pushes them, and installs the values the server answers with (the server applies the optimizer: BSP through the barrier).
    # 'Hybrid': only the embedding tables (Variable(..., is_embed=True)) are on the server, as sparse tables -- a step pushes the
    # rows that received gradient and pulls them back; dense variables keep the local optimizer behind the gradient all-reduce.
    # (ref: hetu/v1/python/hetu/gpu_ops/executor.py comm_mode handling, ParameterServerCommunicate / ParameterServerSparsePull ops)
    def _ps_context(self):
        if self.ps is None:
            from . import ps as _ps
            from .runtime_api import get_worker_communicate
            self.ps = get_worker_communicate() or _ps.connect()      # worker_init() already registered this process
        return self.ps

    def _ps_plan(self, opt):
        key = opt.v1_train_node.id
        if key in self._ps_plans:
            return self._ps_plans[key]
        from ..graph_api import gradients as _grads
        params = list(opt.v1_var_list) if opt.v1_var_list is not None else [p for p in self.graph.parameters() if p.requires_grad]
        if self.comm_mode == "Hybrid":
            params = [p for p in params if p.id in _EMBED_IDS]
        grads = _grads(opt.v1_loss, params) if params else []
        pairs = [(p, g) for p, g in zip(params, grads) if g is not None]
        ps = self._ps_context()
        kind, lr = opt.v1_server_opt
        if getattr(ps, "worker_id", 0) == 0:
            for p, _ in pairs:
                value = self.graph.get_param(p).float().cpu().numpy()
                if self.comm_mode == "Hybrid":
                    ps.init_sparse(p.name, value.reshape(value.shape[0], -1), opt=kind, lr=lr)
                else:
                    ps.init_dense(p.name, value, opt=kind, lr=lr)
        ps.barrier()
        for p, _ in pairs:                         # every worker starts from the server's copy
            if self.comm_mode == "Hybrid":
                continue
            if hasattr(ps, "_dense_len"):
                ps._dense_len[p.name] = int(np.prod(p.shape))
            self.graph.set_param(p, torch.as_tensor(ps.pull(p.name, list(p.shape))))
        self._ps_plans[key] = (pairs, kind, lr)
        return self._ps_plans[key]

    def _ps_explicit(self, node, grad, param, server_opt):
        ps = self._ps_context()
        kind, lr = server_opt
        done = self.__dict__.setdefault("_ps_explicit_init", set())
        if param.id not in done:
            if getattr(ps, "worker_id", 0) == 0:
                ps.init_dense(param.name, self.graph.get_param(param).float().cpu().numpy(), opt=kind, lr=lr)
            if hasattr(ps, "_dense_len"):
                ps._dense_len[param.name] = int(np.prod(param.shape))
            ps.barrier()
            done.add(param.id)
        nw = max(int(getattr(ps, "num_workers", 1)), 1)
        ps.push(param.name, grad.float().cpu().numpy().reshape(list(param.shape)) / nw)
        ps.barrier()
        self.graph.set_param(param, torch.as_tensor(ps.pull(param.name, list(param.shape))))
        ps.barrier()

    def _ps_step(self, opt, grad_values):
        ps = self._ps_context()
        pairs, _, _ = self._ps_plan(opt)
        nw = max(int(getattr(ps, "num_workers", 1)), 1)
        if self.comm_mode == "PS":
            for (p, _), g in zip(pairs, grad_values):
                ps.push(p.name, g.float().cpu().numpy() / nw)           # the server sums the workers' shares
            ps.barrier()                                                 # BSP: all pushes of the step are in
            for p, _ in pairs:
                self.graph.set_param(p, torch.as_tensor(ps.pull(p.name, list(p.shape))))
            ps.barrier()                                                 # nobody pushes step t+1 before everyone pulled step t
            return
        for (p, _), g in zip(pairs, grad_values):                       # Hybrid: sparse rows of the embedding tables
            g2 = g.float().cpu().numpy().reshape(p.shape[0], -1)
            rows = np.nonzero(np.abs(g2).sum(1))[0]
            if rows.size:
                ps.sparse_push(p.name, rows.tolist(), g2[rows] / nw)
        ps.barrier()
        for (p, _), g in zip(pairs, grad_values):
            g2 = g.float().cpu().numpy().reshape(p.shape[0], -1)
            rows = np.nonzero(np.abs(g2).sum(1))[0]
            if rows.size:
                table = self.graph.get_param(p).float().cpu().clone()
                table.reshape(p.shape[0], -1)[torch.as_tensor(rows)] = torch.as_tensor(ps.sparse_pull(p.name, rows.tolist(), g2.shape[1]))
                self.graph.set_param(p, table)
        ps.barrier()

    def run(self, name="default", eval_node_list=None, feed_dict: Optional[Dict] = None, convert_to_numpy_ret_vals=False, **kw):
        if isinstance(name, dict) and feed_dict is None:
            name, feed_dict = "default", name
        nodes = list(eval_node_list) if eval_node_list else self.eval_node_dict[name]
        feed = {k: torch.as_tensor(np.asarray(v)) if not isinstance(v, torch.Tensor) else v for k, v in (feed_dict or {}).items()}
        # dataloader_op nodes: the loader registered under this executor pass (`name`) provides the batch
        from . import dataloader as _dl
        fed = {k.id for k in feed}
        for op in _dl.loaders_of(nodes):
            if op.node.id not in fed and name in op.dataloaders and op.node.graph_id == self.graph.id:
                feed[op.node] = torch.as_tensor(op.get_arr(name))
        for op in _dl._GNN_REGISTRY.values():         # graph feeders: the handler turns the current graph into this node's array
            if op.node.id not in fed and op.node.graph_id == self.graph.id and type(op).graph is not None:
                feed[op.node] = torch.as_tensor(np.asarray(op.get_arr(name)))
        # an lr scheduler object as the optimizer's learning rate: apply the current value, advance after the step
        for opt in self._optimizers():
            sched = getattr(opt, "v1_scheduler", None)
            if sched is not None:
                opt.set_learning_rate(sched.get())
        loss = next((n for n in nodes if isinstance(n, Tensor) and n.producer_type not in _UPDATE_TYPES), None)
        dp = 1
        if self.comm_mode in ("AllReduce", "Hybrid"):
            from .. import distributed
            dp = max(distributed.world_size(), 1)
        training = any(isinstance(n, Tensor) and n.producer_type in _UPDATE_TYPES for n in nodes)
        ps_opts = []
        if training and self.comm_mode in ("PS", "Hybrid"):
            ps_opts = [o for o in self._optimizers() if any(n is o.v1_train_node or getattr(n, "id", None) == o.v1_train_node.id for n in nodes)]
        if ps_opts:
            # fetch the gradients of the server-held variables next to the user's nodes; in 'PS' mode the local update ops are
            # dropped from the fetch list (the server is the optimizer), in 'Hybrid' they still run for the dense variables
            plans = [self._ps_plan(o) for o in ps_opts]
            extra = [g for pairs, _, _ in plans for _, g in pairs]
            train_ids = {o.v1_train_node.id for o in ps_opts}
            fetch = [n for n in nodes if not (self.comm_mode == "PS" and getattr(n, "id", None) in train_ids)] + extra
            got = self.graph.run(loss, fetch, feed, grad_scale=1.0 / dp)
            n_user = len(fetch) - len(extra)
            user_vals, grad_vals = got[:n_user], got[n_user:]
            at = 0
            for o, (pairs, _, _) in zip(ps_opts, plans):
                self._ps_step(o, grad_vals[at:at + len(pairs)])
                at += len(pairs)
            it = iter(user_vals)
            outs = [None if (self.comm_mode == "PS" and getattr(n, "id", None) in train_ids) else next(it) for n in nodes]
        else:
            outs = self.graph.run(loss if training else None, nodes, feed, grad_scale=1.0 / dp)
        # explicit parameterServerCommunicate_op nodes: the fetched value is a gradient that goes to the server, which applies the
        # node's optimizer; the fresh parameter comes back before the next step
        for n, o in zip(nodes, outs):
            target = annotation(n, "ps_target") if isinstance(n, Tensor) else None
            if target is not None and o is not None:
                self._ps_explicit(n, o, *target)
        self.step += 1
        if any(isinstance(n, Tensor) and n.producer_type in _UPDATE_TYPES for n in nodes):
            for opt in self._optimizers():
                if getattr(opt, "v1_scheduler", None) is not None:
                    opt.v1_scheduler.step()
        res = []
        for o in outs:
            if o is None:
                res.append(None)
            else:
                res.append(o.float().cpu().numpy() if convert_to_numpy_ret_vals else _ND(o))
        return res

    def _optimizers(self):
        from .optimizer import _LIVE
        return [o for o in _LIVE if o.update_ops and o.update_ops[0].graph_id == self.graph.id]

    def get_batch_num(self, name="default"):
        """batches per epoch of the loaders registered under `name` (ref: Executor.get_batch_num)"""
        from . import dataloader as _dl
        nums = [op.get_batch_num(name) for op in _dl.loaders_of([]) if name in op.dataloaders and op.node.graph_id == self.graph.id]
        return min(nums) if nums else None

    def save(self, file_path, file_name="checkpoint.pt"):
        import os
        os.makedirs(file_path, exist_ok=True)
        state = {name: t.float().cpu() for name, t in self.graph.named_parameters()} if hasattr(self.graph, "named_parameters") else {}
        torch.save(state, os.path.join(file_path, file_name))

    def load(self, file_path, file_name="checkpoint.pt"):
        import os
        state = torch.load(os.path.join(file_path, file_name))
        for name, t in state.items():
            self.graph.set_param_by_name(name, t) if hasattr(self.graph, "set_param_by_name") else None


class _ND:
    """minimal NDArray-like wrapper (asnumpy) returned by Executor.run"""

    def __init__(self, t):
        self.t = t

    def asnumpy(self):
        return self.t.float().cpu().numpy()

    @property
    def shape(self):
        return tuple(self.t.shape)
