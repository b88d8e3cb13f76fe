"""Core Python front-end: graphs, context managers, tensor constructors, initializers.

API parity with the reference's `hetu` package (python/hetu/context.py, _binding/graph/tensor_ctor.cc,
_binding/graph/init/initializer.cc): `hetu.graph(...)`, `autocast`, `recompute`, `cpu_offload`, `subgraph`,
`context`, `control_dependencies`, `run_level`, `parallel_placeholder`, `parallel_parameter`, `from_numpy`, ...
"""
from __future__ import annotations

import contextlib
import threading
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _C

Tensor = _C.Tensor
Graph = _C.Graph
DistributedStates = _C.DistributedStates
DistributedStatesUnion = _C.DistributedStatesUnion
DeviceGroup = _C.DeviceGroup
DeviceGroupUnion = _C.DeviceGroupUnion
IntSymbol = _C.IntSymbol
device = _C.device
HetuError = _C.HetuError

# ----------------------------------------------------------------------------- dtypes
float32 = "float32"
float16 = "float16"
bfloat16 = "bfloat16"
float64 = "float64"
int8 = "int8"
uint8 = "uint8"
int16 = "int16"
int32 = "int32"
int64 = "int64"
bool_ = "bool"
float4 = "float4"
nfloat4 = "nfloat4"
float8_e4m3 = "float8_e4m3"
float8_e5m2 = "float8_e5m2"

_TORCH_DTYPES = {
    "float32": torch.float32, "float16": torch.float16, "bfloat16": torch.bfloat16, "float64": torch.float64,
    "int8": torch.int8, "uint8": torch.uint8, "int16": torch.int16, "int32": torch.int32, "int64": torch.int64,
    "bool": torch.bool, "float8_e4m3": torch.float8_e4m3fn, "float8_e5m2": torch.float8_e5m2,
}
_FROM_TORCH = {v: k for k, v in _TORCH_DTYPES.items()}


def to_torch_dtype(dt) -> torch.dtype:
    if isinstance(dt, torch.dtype):
        return dt
    return _TORCH_DTYPES[str(dt)]


def dtype_name(dt) -> str:
    if isinstance(dt, torch.dtype):
        return _FROM_TORCH[dt]
    if isinstance(dt, np.dtype) or (isinstance(dt, type) and issubclass(dt, np.generic)):
        return str(np.dtype(dt))
    return str(dt)


# ----------------------------------------------------------------------------- graph contexts
_GRAPH_KINDS = {
    "eager": _C.GraphKind.EAGER,
    "define_by_run": _C.GraphKind.DEFINE_BY_RUN,
    "define_and_run": _C.GraphKind.DEFINE_AND_RUN,
    "executable": _C.GraphKind.EXECUTABLE,
}


class _State(threading.local):
    def __init__(self):
        self.graph_stack: List[Graph] = []
        self.named_graphs: Dict[str, Graph] = {}
        self.autocast_stack: List[Optional[str]] = []
        self.run_level_stack: List[int] = []
        self.counter = 0


_state = _State()


def cur_graph() -> Graph:
    if _state.graph_stack:
        return _state.graph_stack[-1]
    return _register_graph(Graph.default_eager())


def get_default_eager_graph() -> Graph:
    return Graph.default_eager()


@contextlib.contextmanager
def graph(kind_or_graph="define_and_run", create_new: bool = False, prefix: str = "default", num_strategy: int = -1,
          tmp: bool = False):
    """`with hetu.graph("define_and_run", num_strategy=2): ...` (ref: python/hetu/context.py:50-110)."""
    if isinstance(kind_or_graph, Graph):
        g = kind_or_graph
    else:
        kind = str(kind_or_graph)
        if kind not in _GRAPH_KINDS:
            raise ValueError(f"unknown graph kind {kind}")
        if kind == "eager" and not create_new:
            g = Graph.default_eager()
        else:
            key = f"{prefix}_{kind}"
            if create_new or tmp or key not in _state.named_graphs:
                _state.counter += 1
                g = Graph(_GRAPH_KINDS[kind], f"{key}_{_state.counter}", max(num_strategy, 1))
                if not tmp:
                    _state.named_graphs[key] = g
            else:
                g = _state.named_graphs[key]
    if num_strategy > 0 and g.num_strategy < num_strategy:
        g.num_strategy = num_strategy
    _state.graph_stack.append(_register_graph(g))
    try:
        yield g
    finally:
        _state.graph_stack.pop()


@contextlib.contextmanager
def autocast(dtype=None):
    """Mixed-precision region: parameters created inside get a `dtype` compute copy and matmul-class ops run in it."""
    _state.autocast_stack.append(None if dtype is None else dtype_name(dtype))
    try:
        yield
    finally:
        _state.autocast_stack.pop()


def autocast_dtype() -> Optional[str]:
    return _state.autocast_stack[-1] if _state.autocast_stack else None


@contextlib.contextmanager
def context(eager_device=None, device_group_hierarchy=None, stream_index: int = -1, extra_deps: Sequence[Tensor] = ()):
    g = cur_graph()
    g.push_ctx(device_group_hierarchy=_normalize_dgh(device_group_hierarchy), stream_index=stream_index,
               extra_deps=list(extra_deps))
    try:
        yield
    finally:
        g.pop_ctx()


@contextlib.contextmanager
def control_dependencies(deps: Sequence[Tensor]):
    with context(extra_deps=deps):
        yield


@contextlib.contextmanager
def recompute(multi_recompute):
    """multi_recompute: per strategy list of booleans (hetero lists are flattened to any())."""
    flags = [bool(np.any(x)) for x in multi_recompute] if isinstance(multi_recompute, (list, tuple)) else [bool(multi_recompute)]
    g = cur_graph()
    g.push_ctx(recompute=flags)
    try:
        yield
    finally:
        g.pop_ctx()


@contextlib.contextmanager
def cpu_offload(multi_cpu_offload):
    flags = [bool(np.any(x)) for x in multi_cpu_offload] if isinstance(multi_cpu_offload, (list, tuple)) else [bool(multi_cpu_offload)]
    g = cur_graph()
    g.push_ctx(cpu_offload=flags)
    try:
        yield
    finally:
        g.pop_ctx()


@contextlib.contextmanager
def subgraph(name: str, module_type: str = "MODULE"):
    g = cur_graph()
    g.push_subgraph(name, module_type)
    try:
        yield
    finally:
        g.pop_subgraph()


_RUN_LEVELS = {"update": 0, "grad": 1, "compute_only": 2, "alloc": 3, "topo": 4}


@contextlib.contextmanager
def run_level(name="update"):
    _state.run_level_stack.append(_RUN_LEVELS[name] if isinstance(name, str) else int(name))
    try:
        yield
    finally:
        _state.run_level_stack.pop()


def cur_run_level() -> int:
    return _state.run_level_stack[-1] if _state.run_level_stack else 0


@contextlib.contextmanager
def merge_strategy(target_graph: Graph = None, num_strategy: int = -1):
    g = target_graph or cur_graph()
    if num_strategy > 0:
        g.num_strategy = max(g.num_strategy, num_strategy)
    yield g


class profiler:
    """`with hetu.profiler(enabled=True) as prof: ...; prof.summary()` -- per-op timing of executor runs."""

    def __init__(self, enabled=True, use_cpu=False, use_cuda=True, record_shapes=False, profile_memory=False, graph=None):
        self.enabled = enabled
        self.graph = graph

    def __enter__(self):
        if self.graph is None:
            self.graph = cur_graph()
        if self.enabled:
            self.graph.set_profile(True)
        return self

    def __exit__(self, *exc):
        if self.enabled:
            self.graph.set_profile(False)
        return False

    def records(self):
        return self.graph.op_times() if self.graph is not None else []

    def chrome_trace(self, pid: int = 0) -> dict:
        """the recorded ops as a Chrome / Perfetto trace (`chrome://tracing`): complete events laid end to end in execution
        order on one track per category (ref: the torch-profiler bridge of engine/trainer.py:713-730 writes one per device)"""
        events, ts = [], 0.0
        tracks = {"compute": 0, "attention": 1, "comm": 2, "optimizer": 3}
        for name, ms in self.records():
            ty = name.split(":")[0]
            cat = "attention" if "attn" in ty else "comm" if ty in ("comm", "all_reduce", "all_gather", "reduce_scatter", "all_to_all",
                                                                     "grouped_all_reduce") else "optimizer" if "update" in ty else "compute"
            events.append({"name": name, "cat": cat, "ph": "X", "ts": ts * 1e3, "dur": max(ms, 0.0) * 1e3, "pid": pid, "tid": tracks[cat]})
            ts += max(ms, 0.0)
        meta = [{"name": "thread_name", "ph": "M", "pid": pid, "tid": t, "args": {"name": n}} for n, t in tracks.items()]
        return {"traceEvents": meta + events, "displayTimeUnit": "ms"}

    def _op_index(self) -> Dict[str, tuple]:
        """record name ("type:op name") -> (module path, first output shape)"""
        idx = {}
        if self.graph is None:
            return idx
        for i in range(self.graph.num_ops):
            try:
                info = self.graph.op_info(i)
            except Exception:      # noqa: BLE001 -- pruned ops
                continue
            shape = tuple(info["outputs"][0].shape) if info["outputs"] else ()
            idx[f'{info["type"]}:{info["name"]}'] = (info.get("subgraph") or "", shape)
        return idx

    def summary(self, group_by="optype"):
        """rows (key, total ms, calls) sorted by time.  group_by: "optype" | "op" (every op instance) | "optype_shape" (type + output
        shape: which GEMM sizes cost what) | "subgraph" (module view: time per `hetu.subgraph` / nn.Module scope, children included in
        their parents)  (ref: impl/profiler op / optype / optype+shape / graph views, Graph::SubGraphProfiling)"""
        agg: Dict[str, List[float]] = {}
        index = self._op_index() if group_by in ("optype_shape", "subgraph") else {}
        for name, ms in self.records():
            if group_by == "optype":
                keys = [name.split(":")[0]]
            elif group_by == "optype_shape":
                keys = [f'{name.split(":")[0]} {list(index.get(name, ("", ()))[1])}']
            elif group_by == "subgraph":
                path = index.get(name, ("", ()))[0]
                parts = [p for p in path.split(".") if p]
                keys = [".".join(parts[:i]) for i in range(1, len(parts) + 1)] or ["(top level)"]
            else:
                keys = [name]
            for key in keys:
                agg.setdefault(key, []).append(ms)
        rows = sorted(((k, sum(v), len(v)) for k, v in agg.items()), key=lambda r: -r[1])
        return {"by_" + group_by: rows, "breakdown": self.graph.step_breakdown() if self.graph is not None else {}}


# ----------------------------------------------------------------------------- helpers for DS / device-group arguments
def _normalize_dgh(dgh):
    """Accepts None | DeviceGroup | [DeviceGroup per strategy] | [[DeviceGroup per hetero member] per strategy]."""
    if dgh is None:
        return None
    if isinstance(dgh, DeviceGroup):
        return [[dgh]]
    out = []
    for u in dgh:
        if isinstance(u, DeviceGroup):
            out.append([u])
        elif isinstance(u, DeviceGroupUnion):
            out.append(list(u.raw()))
        else:
            out.append(list(u))
    return out


def _normalize_dsh(dsh):
    if dsh is None:
        return None
    if isinstance(dsh, DistributedStates):
        return [DistributedStatesUnion([dsh])]
    out = []
    for u in dsh:
        if isinstance(u, DistributedStates):
            out.append(DistributedStatesUnion([u]))
        elif isinstance(u, DistributedStatesUnion):
            out.append(u)
        else:
            out.append(DistributedStatesUnion(list(u)))
    return out


def _next_name(prefix):
    _state.counter += 1
    return f"{prefix}_{_state.counter}"


def make_op(op_type: str, inputs: Sequence[Tensor], attrs: Optional[dict] = None, *, name: str = "",
            device_group_hierarchy=None, dst_ds=None, sy_shape=(), const_data=None, stream_index: int = -1,
            extra_deps: Sequence[Tensor] = (), graph: Optional[Graph] = None, **_ignored) -> List[Tensor]:
    g = graph or _graph_of(inputs) or cur_graph()
    ac = autocast_dtype()
    if ac is not None and op_type in _AUTOCAST_OPS:
        # autocast: tensor-core ops run in the context's dtype -- fp32 operands (e.g. fed activations) are cast on entry,
        # tensors already in a 16-bit type are left alone (ref: hetu/graph/autocast, DataTransferOp insertion)
        cast = None
        new = []
        for t in inputs:
            if t.dtype == "float32":
                if cast is None:
                    from .ops import data_transfer as cast
                t = cast(t, ac)
            new.append(t)
        inputs = new
    return g.make_op(op_type, list(inputs), attrs or {}, name, _normalize_dgh(device_group_hierarchy), _normalize_dsh(dst_ds),
                     list(sy_shape), const_data, stream_index, list(extra_deps))


_AUTOCAST_OPS = {"linear", "matmul", "bmm", "conv2d", "attn", "attn_packed", "einsum"}


_graphs_by_id: Dict[int, Graph] = {}


def _register_graph(g: Graph):
    _graphs_by_id[g.id] = g
    return g


def _graph_of(inputs):
    # ops go to the graph that owns their inputs (so Tensor methods work outside a `with graph` block)
    for t in inputs:
        g = _graphs_by_id.get(t.graph_id)
        if g is not None:
            return g
    return None


# ----------------------------------------------------------------------------- initializers
class Initializer:
    def __init__(self, kind: str, **params):
        self.kind = kind
        self.params = params

    def attrs(self):
        d = {"init": self.kind}
        d.update(self.params)
        return d


def voidified_initializer():
    return Initializer("zeros")


def provided_initializer(data):
    t = torch.as_tensor(np.asarray(data)) if not isinstance(data, torch.Tensor) else data
    init = Initializer("provided")
    init.data = t
    return init


def zeros_initializer():
    return Initializer("zeros")


def ones_initializer():
    return Initializer("ones")


def constant_initializer(value):
    return Initializer("constant", value=float(value))


def uniform_initializer(lb=-1.0, ub=1.0):
    return Initializer("uniform", lb=float(lb), ub=float(ub))


def normal_initializer(mean=0.0, stddev=1.0):
    return Initializer("normal", mean=float(mean), stddev=float(stddev))


def truncated_normal_initializer(mean=0.0, stddev=1.0, lb=-2.0, ub=2.0):
    return Initializer("truncated_normal", mean=float(mean), stddev=float(stddev), lb=float(lb), ub=float(ub))


def xavier_uniform_initializer(gain=1.0):
    return Initializer("xavier_uniform", gain=float(gain))


def xavier_normal_initializer(gain=1.0):
    return Initializer("xavier_normal", gain=float(gain))


def he_uniform_initializer(mode="fan_in", gain=1.0):
    return Initializer("he_uniform", mode=mode, gain=float(gain))


def he_normal_initializer(mode="fan_in", gain=1.0):
    return Initializer("he_normal", mode=mode, gain=float(gain))


def lecun_uniform_initializer(gain=1.0):
    return Initializer("lecun_uniform", gain=float(gain))


def lecun_normal_initializer(gain=1.0):
    return Initializer("lecun_normal", gain=float(gain))


# ----------------------------------------------------------------------------- tensor constructors
_global_seed = [0]


def set_seed(seed: int):
    _global_seed[0] = int(seed)
    torch.manual_seed(seed)


def placeholder(dtype, shape, ds_hierarchy=None, name: str = "", device_group_hierarchy=None, **kw) -> Tensor:
    return parallel_placeholder(dtype, list(shape), ds_hierarchy, device_group_hierarchy=device_group_hierarchy, name=name, **kw)


def parallel_placeholder(dtype, global_shape, ds_hierarchy=None, device_group_hierarchy=None, name: str = "", symbolic_shape=(),
                         **kw) -> Tensor:
    attrs = {"dtype": dtype_name(dtype), "global_shape": [int(s) for s in global_shape]}
    g = cur_graph()
    t = g.make_op("placeholder", [], attrs, name or _next_name("placeholder"), _normalize_dgh(device_group_hierarchy),
                  _normalize_dsh(ds_hierarchy), list(symbolic_shape), None, -1, [])[0]
    return t


def parallel_parameter(init: Initializer, global_shape, ds_hierarchy=None, local_idx=(-1,), dtype=None, requires_grad: bool = False,
                       parameter_dict=None, device_group_hierarchy=None, name: str = "", **kw) -> Tensor:
    """Sharded parameter whose every shard is a slice of the same seeded global tensor
    (ref: _binding/graph/tensor_ctor.cc:185, hetu/graph/ops/variable.cc)."""
    dt = dtype_name(dtype) if dtype is not None else "float32"
    ac = autocast_dtype()
    if ac is not None and dt == "float32" and requires_grad:
        dt = ac  # bf16 compute copy; the fp32 master lives in the optimizer state
    attrs = init.attrs()
    attrs.update({"dtype": dt, "global_shape": [int(s) for s in global_shape], "requires_grad": bool(requires_grad),
                  "seed": int(kw.get("seed", _global_seed[0]))})
    g = cur_graph()
    const = getattr(init, "data", None)
    t = g.make_op("variable", [], attrs, name or _next_name("parameter"), _normalize_dgh(device_group_hierarchy),
                  _normalize_dsh(ds_hierarchy), [], const, -1, [])[0]
    if g.kind == _C.GraphKind.EAGER:
        t.set_eager_data(g.get_param(t))
    return t


def parameter(init: Initializer, shape, dtype=None, requires_grad: bool = True, name: str = "", **kw) -> Tensor:
    return parallel_parameter(init, shape, None, dtype=dtype, requires_grad=requires_grad, name=name, **kw)


def _to_torch(data, dtype=None) -> torch.Tensor:
    if isinstance(data, torch.Tensor):
        t = data
    else:
        t = torch.as_tensor(np.asarray(data))
    if dtype is not None:
        t = t.to(to_torch_dtype(dtype))
    elif t.dtype == torch.float64:
        t = t.to(torch.float32)
    return t


def from_numpy(arr, requires_grad: bool = False, name: str = "", dtype=None) -> Tensor:
    """Constant / leaf tensor holding `arr` (eager graphs compute with it immediately)."""
    t = _to_torch(arr, dtype)
    g = cur_graph()
    if g.kind != _C.GraphKind.EAGER and requires_grad:
        return parallel_parameter(provided_initializer(t), list(t.shape), None, dtype=dtype_name(t.dtype), requires_grad=True, name=name)
    out = g.make_op("const", [], {"requires_grad": bool(requires_grad)}, name or _next_name("const"), None, None, [], t, -1, [])[0]
    return out


def from_numpy_parallel(arr, ds_hierarchy, device_group_hierarchy=None, requires_grad: bool = False, name: str = "") -> Tensor:
    t = _to_torch(arr)
    return parallel_parameter(provided_initializer(t), list(t.shape), ds_hierarchy, dtype=dtype_name(t.dtype),
                              requires_grad=requires_grad, device_group_hierarchy=device_group_hierarchy, name=name)


def tensor(data, dtype=None, requires_grad: bool = False, trainable: bool = False, name: str = "", **kw) -> Tensor:
    return from_numpy(data, requires_grad=requires_grad or trainable, name=name, dtype=dtype)


def ones(shape, dtype="float32", requires_grad=False, **kw):
    return from_numpy(torch.ones(list(shape), dtype=to_torch_dtype(dtype)), requires_grad)


def zeros(shape, dtype="float32", requires_grad=False, **kw):
    return from_numpy(torch.zeros(list(shape), dtype=to_torch_dtype(dtype)), requires_grad)


def full(shape, value, dtype="float32", requires_grad=False, **kw):
    return from_numpy(torch.full(list(shape), value, dtype=to_torch_dtype(dtype)), requires_grad)


def empty(shape, dtype="float32", requires_grad=False, **kw):
    return zeros(shape, dtype, requires_grad)


def rand(shape, lb=0.0, ub=1.0, dtype="float32", requires_grad=False, **kw):
    return from_numpy(torch.empty(list(shape)).uniform_(lb, ub).to(to_torch_dtype(dtype)), requires_grad)


def randn(shape, mean=0.0, stddev=1.0, dtype="float32", requires_grad=False, **kw):
    return from_numpy((torch.randn(list(shape)) * stddev + mean).to(to_torch_dtype(dtype)), requires_grad)


def randint(shape, low, high, dtype="int64", **kw):
    return from_numpy(torch.randint(low, high, list(shape)).to(to_torch_dtype(dtype)))


# ----------------------------------------------------------------------------- NDArray (host-side value type)
class NDArray:
    """Thin value type over torch.Tensor (the reference's hetu.NDArray surface: numpy(), to, slice, copy, ...)."""

    def __init__(self, t: torch.Tensor):
        self.t = t

    @property
    def shape(self):
        return list(self.t.shape)

    @property
    def dtype(self):
        return dtype_name(self.t.dtype)

    def numpy(self, force=False):
        t = self.t.detach().cpu()
        if t.dtype == torch.bfloat16:
            t = t.float()
        return t.numpy()

    def to(self, *a, **k):
        return NDArray(self.t.to(*a, **k))

    def copy(self):
        return NDArray(self.t.clone())

    def contiguous(self):
        return NDArray(self.t.contiguous())

    def transpose(self, perm=None):
        return NDArray(self.t.permute(*perm) if perm else self.t.T)

    def view(self, shape):
        return NDArray(self.t.reshape(shape))

    def slice(self, begin, size):
        t = self.t
        for d, (b, s) in enumerate(zip(begin, size)):
            t = t.narrow(d, b, s)
        return NDArray(t)

    def __repr__(self):
        return f"NDArray({self.t})"


def numpy_to_NDArray(arr, dtype=None):
    return NDArray(_to_torch(arr, dtype))


def buffer_to_NDArray(buf, dtype="float32", shape=None):
    t = torch.frombuffer(bytearray(buf), dtype=to_torch_dtype(dtype))
    return NDArray(t.reshape(shape) if shape is not None else t)
