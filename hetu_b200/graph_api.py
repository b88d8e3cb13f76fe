"""`graph.run(...)` with the reference's signature and `hetu.gradients`."""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import numpy as np
import torch

from . import _C
from .core import Graph, IntSymbol, NDArray, Tensor, _graphs_by_id, _to_torch, cur_graph, cur_run_level


def gradients(ys, xs, grad_ys=None):
    ys = [ys] if isinstance(ys, Tensor) else list(ys)
    xs = list(xs)
    g = _graphs_by_id.get(ys[0].graph_id) or cur_graph()
    return g.gradients(ys, xs, list(grad_ys) if grad_ys else [])


def _feed_value(v):
    if isinstance(v, (list, tuple)) and len(v) > 0 and not np.isscalar(v[0]):
        return [_to_torch(e.t if isinstance(e, NDArray) else e) for e in v]
    return [_to_torch(v.t if isinstance(v, NDArray) else v)]


def run_graph(g: Graph, loss: Optional[Tensor], fetches: Sequence[Tensor], feed_dict: Optional[Dict] = None,
              int_symbol_dict: Optional[Dict] = None, num_micro_batches: int = 1, compute_strategy_id: int = 0,
              optimize_strategy_id: int = 0, run_level=None, save_checkpoint: bool = False, grad_scale: float = 1.0,
              run_dict=None, cur_strategy_id: Optional[int] = None):
    """graph.run(loss, fetches, feed_dict, int_symbol_dict, num_micro_batches, cur_strategy_id, run_level, ...)
    (ref: python/hetu/_binding/graph/graph.cc:110-121).  Returns one torch tensor (or None) per fetch."""
    if cur_strategy_id is not None:
        compute_strategy_id = cur_strategy_id
    symbols = []
    if int_symbol_dict:
        # {IntSymbol: value | [value per micro-batch]}: the executor sets every symbol before each micro-batch task and
        # re-infers static shapes when a micro-batch differs from the previous one
        for sym, vals in int_symbol_dict.items():
            vs = [int(v) for v in vals] if isinstance(vals, (list, tuple)) else [int(vals)]
            sym.set_data(vs[0])
            symbols.append((sym, vs))
    feed = {}
    for t, v in (feed_dict or {}).items():
        vals = _feed_value(v)
        want = torch.float32
        from .core import to_torch_dtype
        want = to_torch_dtype(t.dtype)
        feed[t] = [x.to(want) if x.dtype != want else x for x in vals]
    lvl = cur_run_level() if run_level is None else (run_level if isinstance(run_level, int) else
                                                     {"update": 0, "grad": 1, "compute_only": 2, "alloc": 3, "topo": 4}[run_level])
    fetches = list(fetches)
    for i, f in enumerate(fetches):
        if not isinstance(f, Tensor):
            raise TypeError(f"fetch {i} is {type(f).__name__}, not a graph tensor (a gradient that does not exist comes back as None)")
    res = g.run_native(loss, fetches, feed, int(num_micro_batches), int(compute_strategy_id), int(lvl),
                       float(grad_scale), bool(save_checkpoint), symbols)
    # the executor also evaluates `loss` (it drives the backward pass); the caller gets exactly one value per fetch
    return res[:len(fetches)] if (loss is not None and fetches and len(res) > len(fetches)) else res


def _graph_run(self, loss, fetches=None, feed_dict=None, *args, **kwargs):
    # accepted overloads: run(loss, fetches, feed_dict, ...) and run(fetches, feed_dict)
    if fetches is None or isinstance(fetches, dict):
        feed_dict, fetches, loss = fetches, (loss if isinstance(loss, (list, tuple)) else [loss]), None
    names = ["int_symbol_dict", "num_micro_batches", "compute_strategy_id", "optimize_strategy_id", "run_level",
             "save_checkpoint", "grad_scale", "run_dict"]
    for n, a in zip(names, args):
        kwargs.setdefault(n, a)
    return run_graph(self, loss, fetches, feed_dict, **kwargs)


Graph.run_native = Graph.run
Graph.run = _graph_run
