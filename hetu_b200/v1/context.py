"""v1 placement vocabulary: `DeviceGroup` (the workers / servers a node runs on, model-parallel workers as tuples) and `NodeStatus`
(how a tensor is laid out over a device group: {dim: parts} splits + duplicate + partial counts and the device-axis order) -- the
1.x ancestor of DistributedStates, to which it converts.  (ref: hetu/v1/python/hetu/context.py DeviceGroup :28, NodeStatus :248)"""
from __future__ import annotations

from typing import Dict, Iterable, Optional, Sequence, Tuple

import numpy as np

from .executor import _Ctx, cpu, gpu, rcpu, rgpu


class DeviceGroup:
    """`DeviceGroup([gpu(0), gpu(1)])` data parallel; `DeviceGroup([(gpu(0), gpu(1)), (gpu(2), gpu(3))])` two model-parallel workers;
    strings like 'gpu:0', 'node2:gpu:1', 'cpu:0' are parsed; CPU contexts are servers"""

    def __init__(self, ctxs):
        self._contexts = self.parse_contexts(ctxs)
        self._servers, self._workers = [], []
        for c in self._contexts:
            if isinstance(c, tuple) or c.kind == "gpu":
                self._workers.append(c)
            else:
                self._servers.append(c)
        self._is_mp = any(isinstance(c, tuple) and len(c) > 1 for c in self._workers)
        self._mp_dev_num = max([len(c) for c in self._workers if isinstance(c, tuple)] + [1]) if self._workers else None
        self._index = None

    @classmethod
    def parse_contexts(cls, ctxs):
        if isinstance(ctxs, DeviceGroup):
            return list(ctxs._contexts)
        if isinstance(ctxs, (str, _Ctx)):
            ctxs = [ctxs]
        out = []
        for c in ctxs:
            if isinstance(c, (tuple, list)):
                out.append(tuple(cls.str2ctx(x) for x in c))
            else:
                out.append(cls.str2ctx(c))
        return out

    @staticmethod
    def str2ctx(c):
        if isinstance(c, _Ctx):
            return c
        parts = str(c).lower().split(":")
        host = "localhost" if len(parts) == 2 else ":".join(parts[:-2])
        kind, idx = parts[-2], int(parts[-1])
        assert kind in ("cpu", "gpu"), f"cannot parse device '{c}'"
        return _Ctx(kind, idx, host)

    def index(self, ctx):
        for i, c in enumerate(self._contexts):
            if c == ctx or (isinstance(c, tuple) and ctx in c):
                return i
        raise ValueError(f"{ctx} is not in {self}")

    def __getitem__(self, i): return self._contexts[i]                 # noqa: E704
    def __iter__(self): return iter(self._contexts)                     # noqa: E704
    def __len__(self): return len(self._contexts)                       # noqa: E704
    is_mp = property(lambda self: self._is_mp)
    mp_dev_num = property(lambda self: self._mp_dev_num)
    worker_num = property(lambda self: len(self._workers))
    server_num = property(lambda self: len(self._servers))
    workers = property(lambda self: tuple(self._workers))
    servers = property(lambda self: tuple(self._servers))

    def check_mp_num(self, mp_dev_num: int):
        assert all((len(c) if isinstance(c, tuple) else 1) == mp_dev_num for c in self._workers), f"every worker must span {mp_dev_num} devices"

    def flat_workers(self):
        return [d for w in self._workers for d in (w if isinstance(w, tuple) else (w,))]

    def get_sorted(self):
        key = lambda c: min((d.host, d.index) for d in (c if isinstance(c, tuple) else (c,)))    # noqa: E731
        return DeviceGroup(sorted(self._contexts, key=key))

    def get_only(self):
        assert len(self._contexts) == 1 and not isinstance(self._contexts[0], tuple)
        return self._contexts[0]

    def __repr__(self): return "DeviceGroup(" + ", ".join(map(str, self._contexts)) + ")"     # noqa: E704
    full_repr = __repr__
    def __hash__(self): return hash(tuple(self._contexts))                                      # noqa: E704
    def __eq__(self, other): return isinstance(other, DeviceGroup) and self._contexts == other._contexts   # noqa: E704


class ContextStack:
    def __init__(self): self._stack = []                                 # noqa: E704
    def peek(self): return self._stack[-1] if self._stack else None      # noqa: E704
    def push(self, ctx): self._stack.append(ctx)                         # noqa: E704
    def pop(self): return self._stack.pop()                              # noqa: E704


class NodeStatus:
    """layout of one tensor over `dev_num` devices.  `state` {dim: parts}; `duplicate` replicas; `partial` partial-sum copies;
    `order` the nesting of the device axes (-1 duplicate, -2 partial, d >= 0 a split dim), outermost first."""

    def __init__(self, state: Optional[Dict[int, int]] = None, dev_num: Optional[int] = None, partial_or_node=None,
                 duplicate: Optional[int] = None, partial: Optional[int] = None, order: Optional[Sequence[int]] = None):
        if isinstance(state, (tuple, list)):
            state = {i: int(v) for i, v in enumerate(state)}
        self._state = {int(k): int(v) for k, v in (state or {}).items() if int(v) > 1} if state is not None else None
        self._dev_num = dev_num
        self._duplicate, self._partial = duplicate, partial
        self._order = tuple(order) if order is not None else None
        self.try_get_duplicate()

    # ---- reading
    def get(self): return self._state, self._duplicate                               # noqa: E704
    def get_all(self): return self._state, self._duplicate, self._order              # noqa: E704
    state = property(lambda self: self._state)
    duplicate = property(lambda self: self._duplicate)
    partial = property(lambda self: self._partial)
    order = property(lambda self: self._order)
    dev_num = property(lambda self: self._dev_num)

    def is_dist(self) -> bool:
        return bool(self._dev_num and self._dev_num > 1)

    def get_dim(self, ind: int) -> int:
        if ind == -1:
            return self._duplicate or 1
        if ind == -2:
            return self._partial or 1
        return (self._state or {}).get(ind, 1)

    # ---- writing
    def try_get_duplicate(self):
        """fill in whichever of duplicate / partial / dev_num follows from the others"""
        if self._state is None:
            return
        split = int(np.prod(list(self._state.values()))) if self._state else 1
        if self._dev_num is not None:
            rest = self._dev_num // split
            if self._duplicate is None and self._partial is not None:
                self._duplicate = rest // self._partial
            elif self._partial is None and self._duplicate is not None:
                self._partial = rest // self._duplicate
            elif self._duplicate is None and self._partial is None:
                self._duplicate, self._partial = rest, 1
        elif self._duplicate is not None:
            self._partial = self._partial or 1
            self._dev_num = split * self._duplicate * self._partial

    def set_state(self, state=None, duplicate=None, partial=None):
        if state is not None:
            self._state = {int(k): int(v) for k, v in state.items() if int(v) > 1}
        if duplicate is not None:
            self._duplicate = int(duplicate)
        if partial is not None:
            self._partial = int(partial)
        self.try_get_duplicate()

    def set_duplicate(self, duplicate=None): self.set_state(duplicate=duplicate)       # noqa: E704
    def set_partial(self, partial=None): self.set_state(partial=partial)               # noqa: E704
    def set_order(self, order=None): self._order = tuple(order) if order is not None else self._order   # noqa: E704
    def set_one(self): self._state, self._duplicate, self._partial, self._order, self._dev_num = {}, 1, 1, (), 1   # noqa: E704

    def copy_state_from(self, other): self._state, self._duplicate, self._partial, self._dev_num = dict(other._state or {}), other._duplicate, other._partial, other._dev_num   # noqa: E704,E501
    def copy_order_from(self, other): self._order = other._order                       # noqa: E704
    def copy_from(self, other, copy_order=True):                                       # noqa: E704
        self.copy_state_from(other)
        if copy_order:
            self.copy_order_from(other)

    def get_default_order(self):
        if self._order is None:
            axes = ([-2] if (self._partial or 1) > 1 else []) + ([-1] if (self._duplicate or 1) > 1 else []) + sorted(self._state or {})
            self._order = tuple(axes)
        return self._order

    # ---- validity
    def valid_state(self) -> bool:
        return self._state is not None and self._duplicate is not None

    def valid_all(self) -> bool:
        if not self.valid_state() or self._order is None:
            return False
        need = {d for d in self._state} | ({-1} if self._duplicate > 1 else set()) | ({-2} if (self._partial or 1) > 1 else set())
        return need <= set(self._order) and (self._dev_num is None or self._dev_num == self._total())

    def valid(self, include_order: bool) -> bool:
        return self.valid_all() if include_order else self.valid_state()

    def check_state(self, max_dim: int, check_order: bool):
        assert all(d < max_dim for d in (self._state or {})), f"split dims {sorted(self._state)} exceed {max_dim}"
        if check_order and self._order is not None:
            assert all(d < max_dim for d in self._order)

    def _total(self):
        return int(np.prod(list((self._state or {}).values()) or [1])) * (self._duplicate or 1) * (self._partial or 1)

    # ---- device <-> shard coordinates
    def get_loop_sizes(self) -> Tuple[int, ...]:
        """stride of every device axis in `order` (how many consecutive devices share one coordinate of that axis)"""
        order = self.get_default_order()
        sizes, acc = [], 1
        for d in reversed(order):
            sizes.append(acc)
            acc *= self.get_dim(d)
        return tuple(reversed(sizes))

    def map_dev_to_index(self, global_index: int, containing_duplicate: bool = False) -> Dict[int, int]:
        order, out = self.get_default_order(), {}
        for d, stride in zip(order, self.get_loop_sizes()):
            coord = (global_index // stride) % self.get_dim(d)
            if d >= 0 or containing_duplicate:
                out[d] = coord
        return out

    def get_devices_by_dim(self, dim: int, index: int, devices: Optional[Sequence] = None):
        """the devices whose coordinate along `dim` equals `index`"""
        devs = list(devices) if devices is not None else list(range(self._total()))
        return [dv for i, dv in enumerate(devs) if self.map_dev_to_index(i, True).get(dim, 0) == index]

    # ---- algebra
    def combine_state(self, *src2dst):
        """fold device axes: (src, dst) moves the `src` axis (or list of axes) into `dst` -- e.g. (-2, -1) partial -> duplicate is what
        an all-reduce does, (0, -1) is an all-gather of dim 0.  -> (state, duplicate, partial)"""
        state, dup, par = dict(self._state or {}), self._duplicate or 1, self._partial or 1
        for srcs, dst in src2dst:
            for s in ([srcs] if isinstance(srcs, int) else list(srcs)):
                if s == dst:
                    continue
                n = dup if s == -1 else par if s == -2 else state.get(s, 1)
                if s == -1:
                    dup = 1
                elif s == -2:
                    par = 1
                else:
                    state.pop(s, None)
                if dst == -1:
                    dup *= n
                elif dst == -2:
                    par *= n
                else:
                    state[dst] = state.get(dst, 1) * n
        return state, dup, par

    def combine_order(self, *src2dst):
        order = list(self.get_default_order())
        for srcs, dst in src2dst:
            for s in ([srcs] if isinstance(srcs, int) else list(srcs)):
                if s == dst or s not in order:
                    continue
                if dst in order:
                    order.remove(s)
                else:
                    order[order.index(s)] = dst
        return tuple(order)

    def get_combine_from(self, other: "NodeStatus", deduce_order: bool, *src2dst):
        if deduce_order:
            self._order = other.combine_order(*src2dst)
        else:
            st, dup, par = other.combine_state(*src2dst)
            self._state, self._duplicate, self._partial, self._dev_num = st, dup, par, other._dev_num

    def reduce_state(self, dim: int):
        return self.combine_state((dim, -2))

    def reduce_order(self, dim: int):
        return self.combine_order((dim, -2))

    def remove_partial(self) -> "NodeStatus":
        st, dup, par = self.combine_state((-2, -1))
        return NodeStatus(st, dev_num=self._dev_num, duplicate=dup, partial=par, order=self.combine_order((-2, -1)))

    def exchange_state(self, n1: int, n2: int):
        st = dict(self._state or {})
        a, b = st.pop(n1, 1), st.pop(n2, 1)
        if b > 1:
            st[n1] = b
        if a > 1:
            st[n2] = a
        return st, self._duplicate, self._partial

    def exchange_order(self, n1: int, n2: int):
        return tuple(n2 if d == n1 else n1 if d == n2 else d for d in self.get_default_order())

    # ---- which collective turns self into other
    def check_combine(self, other: "NodeStatus", *src2dst) -> bool:
        st, dup, par = self.combine_state(*src2dst)
        return (st, dup, par) == (other._state or {}, other._duplicate or 1, other._partial or 1) and \
            tuple(self.combine_order(*src2dst)) == tuple(other.get_default_order())

    def check_allreduce(self, other): return (self._partial or 1) > 1 and self.check_combine(other, (-2, -1))          # noqa: E704
    def check_allgather(self, other): return 0 in (self._state or {}) and self.check_combine(other, (0, -1))           # noqa: E704
    def check_reducescatter(self, other): return (self._partial or 1) > 1 and self.check_combine(other, (-2, 0))       # noqa: E704
    def check_broadcast(self, other): return other.is_dist() and not self.is_dist() and not (other._state or {})       # noqa: E704
    def check_reduce(self, other): return (self._partial or 1) > 1 and not other.is_dist()                             # noqa: E704
    def check_reduce_dim(self, other, dim): return self.check_combine(other, (dim, -2))                                # noqa: E704

    # ---- identity
    def effect_equal(self, other) -> bool:
        return other is not None and (self._state or {}) == (other._state or {}) and (self._duplicate or 1) == (other._duplicate or 1) and \
            (self._partial or 1) == (other._partial or 1)

    def value_equal(self, state, duplicate, partial, order) -> bool:
        return (self._state or {}) == {k: v for k, v in state.items() if v > 1} and self._duplicate == duplicate and (self._partial or 1) == partial and \
            tuple(self.get_default_order()) == tuple(order)

    def __eq__(self, other):
        return isinstance(other, NodeStatus) and self.effect_equal(other) and tuple(self.get_default_order()) == tuple(other.get_default_order())

    def content_hash(self):
        return hash((tuple(sorted((self._state or {}).items())), self._duplicate, self._partial, tuple(self.get_default_order())))

    __hash__ = content_hash

    def __repr__(self):
        return f"NodeStatus(state={self._state}, duplicate={self._duplicate}, partial={self._partial}, order={self._order}, dev_num={self._dev_num})"

    # ---- bridge to the 2.x layout type
    def to_distributed_states(self):
        from ..core import DistributedStates
        states = dict(self._state or {})
        if (self._duplicate or 1) > 1:
            states[-1] = self._duplicate
        if (self._partial or 1) > 1:
            states[-2] = self._partial
        return DistributedStates(self._total(), states, [d for d in self.get_default_order() if states.get(d, 1) > 1])

    @classmethod
    def from_distributed_states(cls, ds):
        st = dict(ds.states)
        return cls({k: v for k, v in st.items() if k >= 0}, dev_num=ds.device_num, duplicate=st.get(-1, 1), partial=st.get(-2, 1), order=list(ds.order))


__all__ = ["DeviceGroup", "ContextStack", "NodeStatus", "cpu", "gpu", "rcpu", "rgpu"]
