"""hetu.nn.Module: torch-like module tree over graph tensors.

Every `__call__` opens a named sub-graph so the executor / profiler can attribute ops to modules
(ref: python/hetu/nn/modules/module.py, hetu/graph/subgraph.h).
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Iterator, List, Optional, Tuple

import torch

from ..core import Tensor, _graphs_by_id, subgraph


class Module:
    def __init__(self):
        object.__setattr__(self, "_parameters", OrderedDict())
        object.__setattr__(self, "_buffers", OrderedDict())
        object.__setattr__(self, "_modules", OrderedDict())
        object.__setattr__(self, "training", True)
        object.__setattr__(self, "_module_name", type(self).__name__)

    # ------------------------------------------------------------------ registration
    def __setattr__(self, name, value):
        if isinstance(value, Module):
            self._modules[name] = value
            object.__setattr__(value, "_module_name", name)
        elif isinstance(value, Tensor) and value.producer_type == "variable":
            # only graph variables are parameters: an activation kept on the module (e.g. an auxiliary loss) is not
            self._parameters[name] = value
        object.__setattr__(self, name, value)

    def register_parameter(self, name: str, param: Optional[Tensor]):
        if param is not None:
            self._parameters[name] = param
        object.__setattr__(self, name, param)

    def register_buffer(self, name: str, tensor: Optional[Tensor]):
        if tensor is not None:
            self._buffers[name] = tensor
        object.__setattr__(self, name, tensor)

    def add_module(self, name: str, module: "Module"):
        self._modules[name] = module
        object.__setattr__(module, "_module_name", name)
        object.__setattr__(self, name, module)

    # ------------------------------------------------------------------ traversal
    def named_modules(self, prefix: str = "") -> Iterator[Tuple[str, "Module"]]:
        yield prefix, self
        for n, m in self._modules.items():
            yield from m.named_modules(prefix + ("." if prefix else "") + n)

    def modules(self):
        for _, m in self.named_modules():
            yield m

    def children(self):
        return iter(self._modules.values())

    def named_parameters(self, prefix: str = "", recurse: bool = True) -> Iterator[Tuple[str, Tensor]]:
        seen = set()
        for mn, m in (self.named_modules(prefix) if recurse else [(prefix, self)]):
            for n, p in m._parameters.items():
                if p is None or p.id in seen:
                    continue
                seen.add(p.id)
                yield mn + ("." if mn else "") + n, p

    def parameters(self, recurse: bool = True):
        for _, p in self.named_parameters(recurse=recurse):
            yield p

    def named_buffers(self, prefix: str = ""):
        for mn, m in self.named_modules(prefix):
            for n, b in m._buffers.items():
                yield mn + ("." if mn else "") + n, b

    # ------------------------------------------------------------------ state
    def state_dict(self, format: str = "torch") -> Dict[str, torch.Tensor]:
        out = OrderedDict()
        for n, p in list(self.named_parameters()) + list(self.named_buffers()):
            g = _graphs_by_id.get(p.graph_id)
            d = p.eager_data()
            if d is None and g is not None:
                d = g.get_param(p)
            out[n] = d if format == "hetu" else d.detach().clone()
        return out

    def load_state_dict(self, state: Dict[str, torch.Tensor], strict: bool = True):
        own = dict(list(self.named_parameters()) + list(self.named_buffers()))
        missing = [k for k in own if k not in state]
        unexpected = [k for k in state if k not in own]
        if strict and (missing or unexpected):
            raise KeyError(f"state_dict mismatch: missing {missing}, unexpected {unexpected}")
        for k, v in state.items():
            if k in own:
                p = own[k]
                g = _graphs_by_id[p.graph_id]
                g.set_param(p, torch.as_tensor(v))
                if p.eager_data() is not None:
                    p.set_eager_data(g.get_param(p))
        return missing, unexpected

    def buffers(self):
        for _, b in self.named_buffers():
            yield b

    def named_children(self) -> Iterator[Tuple[str, "Module"]]:
        for n, m in self._modules.items():
            if m is not None:
                yield n, m

    def apply(self, fn):
        """fn(module) on every sub-module (children first), then on self"""
        for m in self.children():
            m.apply(fn)
        fn(self)
        return self

    def to(self, dtype=None, **kw):
        """cast the floating-point parameters / buffers (device placement is decided by the graph's device groups)"""
        if dtype is None:
            return self
        from ..core import to_torch_dtype
        td = to_torch_dtype(dtype) if not isinstance(dtype, torch.dtype) else dtype
        for _, p in list(self.named_parameters()) + list(self.named_buffers()):
            g = _graphs_by_id.get(p.graph_id)
            if g is None or not g.has_param(p):
                continue
            d = g.get_param(p)
            if d.is_floating_point() and d.dtype != td:
                g.set_param(p, d.to(td))
        return self

    def train(self, mode: bool = True):
        for m in self.modules():
            object.__setattr__(m, "training", mode)
        return self

    def eval(self):
        return self.train(False)

    def zero_grad(self):
        pass  # gradients are graph tensors; the executor owns accumulation buffers

    # ------------------------------------------------------------------ call
    def forward(self, *args, **kwargs):
        raise NotImplementedError

    def __call__(self, *args, **kwargs):
        with subgraph(self._module_name, "MODULE"):
            return self.forward(*args, **kwargs)

    def extra_repr(self):
        return ""

    def __repr__(self):
        lines = [f"{type(self).__name__}({self.extra_repr()}"]
        for n, m in self._modules.items():
            sub = repr(m).replace("\n", "\n  ")
            lines.append(f"  ({n}): {sub}")
        return "\n".join(lines) + ("\n)" if self._modules else ")")


class ModuleList(Module):
    def __init__(self, modules=()):
        super().__init__()
        self._list: List[Module] = []
        for m in modules:
            self.append(m)

    def append(self, m: Module):
        self.add_module(str(len(self._list)), m)
        self._list.append(m)
        return self

    def extend(self, ms):
        for m in ms:
            self.append(m)
        return self

    def __getitem__(self, i):
        return self._list[i]

    def __len__(self):
        return len(self._list)

    def __iter__(self):
        return iter(self._list)


class ModuleDict(Module):
    def __init__(self, modules=None):
        super().__init__()
        for k, v in (modules or {}).items():
            self.add_module(k, v)

    def __getitem__(self, k):
        return self._modules[k]

    def __setitem__(self, k, v):
        self.add_module(k, v)

    def keys(self):
        return self._modules.keys()

    def items(self):
        return self._modules.items()

    def values(self):
        return self._modules.values()


class Sequential(Module):
    def __init__(self, *modules):
        super().__init__()
        if len(modules) == 1 and isinstance(modules[0], (OrderedDict, dict)):
            for k, v in modules[0].items():
                self.add_module(k, v)
        else:
            for i, m in enumerate(modules):
                self.add_module(str(i), m)

    def forward(self, x):
        for m in self._modules.values():
            x = m(x)
        return x

    def __getitem__(self, i):
        return list(self._modules.values())[i]

    def __len__(self):
        return len(self._modules)
