"""v1's measurement tools for its strategy searches: `HetuProfiler` times graph nodes on this device, `NCCLProfiler` times
collectives over the live communicator, `HetuSimulator` answers "how long would this node / this transfer take" from measured
cases (cached on disk) and an alpha-beta link model.  (ref: hetu/v1/python/hetu/profiler.py -- HetuProfiler :55, NCCLProfiler :390,
HetuSimulator :609)"""
from __future__ import annotations

import enum
import json
import os
import time
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from .. import core


class NCCLOP(enum.Enum):
    AllReduce = 0
    AllGather = 1
    ReduceScatter = 2
    Reduce = 3
    Broadcast = 4


def _sync():
    if torch.cuda.is_available():
        torch.cuda.synchronize()


def _timed(fn, iters: int, warmup: int = 2) -> float:
    """ms per call: CUDA events on the current stream on a GPU, wall clock on the CPU"""
    for _ in range(warmup):
        fn()
    _sync()
    if torch.cuda.is_available():
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        b.synchronize()
        return a.elapsed_time(b) / iters
    t0 = time.perf_counter()
    for _ in range(iters):
        fn()
    return (time.perf_counter() - t0) * 1e3 / iters


class BaseProfiler:
    def __init__(self):
        self.idx = 0
        self.ctx = None

    def set_ctx(self, ctx):
        self.ctx = ctx

    @staticmethod
    def init_memory(shape, seed=0, dtype=np.float32):
        return np.random.RandomState(seed).normal(0.0, 0.1, size=list(shape)).astype(dtype)

    def renew_instances(self, *a, **k):
        raise NotImplementedError


class HetuProfiler(BaseProfiler):
    """time a set of graph nodes under random feeds.  `feed_shapes`: {placeholder node: shape}; `node_to_arr_map` (optional) gives
    explicit arrays instead.  `profile()` -> ms per run of all nodes; `profile_all()` -> {op record: ms} from the per-op profiler."""

    def __init__(self, computing_nodes, feed_shapes: Dict, node_to_arr_map: Optional[Dict] = None, ctx=None):
        super().__init__()
        self.ctx = ctx
        self.renew_nodes(computing_nodes, feed_shapes, node_to_arr_map)

    def renew_nodes(self, computing_nodes, feed_shapes, node_to_arr_map=None):
        from .executor import Executor
        self.computing_nodes = list(computing_nodes) if isinstance(computing_nodes, (list, tuple)) else [computing_nodes]
        self.feed_shapes = dict(feed_shapes or {})
        self.node_to_arr_map = dict(node_to_arr_map or {})
        self.executor = Executor(self.computing_nodes)
        self.renew_instances()

    def renew_instances(self, num_instances: int = 5):
        """`num_instances` random feeds per placeholder; runs rotate through them so a cache-resident input is not re-read"""
        self.instances = []
        for i in range(num_instances):
            feed = {}
            for node, shape in self.feed_shapes.items():
                if node in self.node_to_arr_map and self.node_to_arr_map[node] is not None:
                    feed[node] = np.asarray(self.node_to_arr_map[node])
                elif str(getattr(node, "dtype", "float32")).startswith("int"):
                    feed[node] = np.random.RandomState(i).randint(0, max(int(self.lookup_vocab), 1), size=list(shape)).astype(np.int64)
                else:
                    feed[node] = self.init_memory(shape, seed=i)
            self.instances.append(feed)
        self.idx = 0

    lookup_vocab = 2          # upper bound for random integer feeds (set it to the vocabulary size for embedding inputs)

    def get_lookup_sampler(self, vocab_size: int, ignore_rate: float = 0.0, zipf: bool = False):
        """integer-id sampler for embedding inputs: uniform or Zipf-distributed ids, a fraction `ignore_rate` set to -1"""
        def sampler(shape, seed=0):
            r = np.random.RandomState(seed)
            ids = (np.minimum(r.zipf(1.2, size=list(shape)), vocab_size) - 1) if zipf else r.randint(0, vocab_size, size=list(shape))
            if ignore_rate > 0:
                ids = np.where(r.rand(*shape) < ignore_rate, -1, ids)
            return ids.astype(np.int64)
        return sampler

    def step_idx(self):
        self.idx = (self.idx + 1) % max(len(self.instances), 1)

    def pure_compute(self):
        self.executor.run(feed_dict=self.instances[self.idx] if self.instances else {})
        self.step_idx()

    def profile(self, num_iterations: int = 100, profiler: str = "gpu") -> float:
        return _timed(self.pure_compute, int(num_iterations))

    def profile_all(self, num_iterations: int = 100, profiler: str = "gpu") -> Dict[str, float]:
        g = self.executor.graph
        self.pure_compute()
        with core.profiler(graph=g) as prof:
            for _ in range(int(num_iterations)):
                self.pure_compute()
            _sync()
            rows = prof.summary(group_by="op")["by_op"]
        return {k: ms / max(calls, 1) for k, ms, calls in rows}

    def profile_n_log(self, log_file: str, profiler: str = "cpu", num_iterations: int = 20):
        res = self.profile_all(num_iterations, profiler)
        with open(log_file, "w") as f:
            for k, v in sorted(res.items(), key=lambda kv: -kv[1]):
                f.write(f"{k}\t{v:.6f} ms\n")
        return res

    def free_mem(self):
        self.instances = []

    clean = free_mem


class NCCLProfiler(BaseProfiler):
    """time collectives / point-to-point transfers over the process's communicator (every participating rank makes the same call)"""

    def __init__(self):
        super().__init__()
        from .runtime_api import wrapped_mpi_nccl_init
        self.comm = wrapped_mpi_nccl_init()
        self.rank, self.nrank = self.comm.rank, self.comm.nrank

    def _buffer(self, size):
        n = int(size)
        dev = "cuda" if torch.cuda.is_available() and not os.environ.get("HETU_B200_FORCE_CPU") else "cpu"
        return torch.ones(max(n, 1), dtype=torch.float32, device=dev)

    def profile_allreduce(self, size, devices: Optional[Sequence[int]] = None, num_iterations: int = 10, primitive=NCCLOP.AllReduce) -> float:
        """ms per collective of `size` fp32 elements among `devices` (ranks); ranks outside the group return 0"""
        from .runtime_api import Communicator
        ranks = sorted(int(getattr(d, "device_id", d)) for d in devices) if devices is not None else list(range(self.nrank))
        comm = Communicator(ranks)                     # collective creation: every rank reaches this line
        if len(ranks) < 2 or not self.comm._live or self.comm._C.comm_rank() not in ranks:
            return 0.0
        x = self._buffer(size if primitive != NCCLOP.ReduceScatter else (int(size) // len(ranks)) * len(ranks))
        fn = {NCCLOP.AllReduce: lambda: comm.all_reduce(x), NCCLOP.AllGather: lambda: comm.all_gather(x),
              NCCLOP.ReduceScatter: lambda: comm.reduce_scatter(x), NCCLOP.Reduce: lambda: comm.reduce(x, 0),
              NCCLOP.Broadcast: lambda: comm.broadcast(x, 0)}[primitive]
        return _timed(fn, int(num_iterations), warmup=1)

    def profile_sendrecv(self, size, devices: Sequence[int], num_iterations: int = 10) -> float:
        """ms per transfer of `size` fp32 elements from devices[0] to devices[1]"""
        src, dst = (int(getattr(d, "device_id", d)) for d in devices)
        me = self.comm._C.comm_rank() if self.comm._live else 0
        if me not in (src, dst) or src == dst or not self.comm._live:
            return 0.0
        x = self._buffer(size)
        C = self.comm._C

        def once():
            if me == src:
                C.comm_send(x, dst, 0)
            else:
                C.comm_recv([x.numel()], "float32", src, 0)
        return _timed(once, int(num_iterations), warmup=1)


class HetuSimulator:
    """Execution-time oracle of the v1 searches.  Node times are measured once per (op type, input shapes) on this device and kept in
    `cache_path`; transfers come from an alpha-beta model of the node's links (NVLink inside a host, the NIC across hosts) unless a
    communicator is live and `measure_comm` is set, in which case they are measured through NCCLProfiler and cached too."""

    # B200 HGX defaults; `link_model` overrides ({"nvlink_GBps", "nic_GBps", "latency_us"})
    LINKS = {"nvlink_GBps": 900.0, "nic_GBps": 50.0, "latency_us": 8.0, "hbm_GBps": 7700.0}

    def __init__(self, feed_shapes: Optional[Dict] = None, ctx=None, mpi_comm=None, num_ctxs: int = 1, pix: bool = True,
                 cache_path: str = "/tmp/hetu_cached_exetime.json", link_model: Optional[dict] = None, measure_comm: bool = False):
        self.feed_shapes, self.ctx, self.mpi_comm, self.num_ctxs, self.pix = dict(feed_shapes or {}), ctx, mpi_comm, int(num_ctxs), pix
        self.cache_path = cache_path
        self.links = dict(self.LINKS, **(link_model or {}))
        self.measure_comm = bool(measure_comm)
        self.cached_exetime: Dict[str, float] = {}
        if cache_path and os.path.exists(cache_path):
            try:
                self.cached_exetime = json.load(open(cache_path))
            except Exception:      # noqa: BLE001 -- a stale / foreign cache is ignored
                self.cached_exetime = {}
        self.nccl_profiler = None

    # ---- compute
    @staticmethod
    def _key(kind, *parts):
        return kind + "|" + "|".join(json.dumps(p, sort_keys=True) if not isinstance(p, str) else p for p in parts)

    def profile_new_case(self, builder, input_shapes: Sequence[Sequence[int]], num_iterations: int = 10) -> float:
        """measure `builder(*placeholders)` on random inputs of `input_shapes`"""
        from . import executor as v1ex
        saved = (v1ex._graph, v1ex._graph_ctx, dict(v1ex._notes))
        v1ex._graph = v1ex._graph_ctx = None
        try:
            phs = [v1ex.placeholder_op(f"sim_in{i}", list(s)) for i, s in enumerate(input_shapes)]
            out = builder(*phs)
            prof = HetuProfiler([out], {p: list(s) for p, s in zip(phs, input_shapes)})
            return prof.profile(num_iterations)
        finally:
            v1ex.reset_graph()
            v1ex._graph, v1ex._graph_ctx = saved[0], saved[1]
            v1ex._notes.update(saved[2])

    def get_node_time(self, node, input_shapes, output_shape=None, builder=None) -> float:
        """ms for one execution of `node`'s op on inputs of `input_shapes`.  `node`: a graph tensor (its producer type names the op), an
        op-type string, or anything with `.producer_type`; `builder` rebuilds the op for measurement (defaults exist for the common
        types; unknown types fall back to a memory-bound estimate from the bytes touched)"""
        ty = node if isinstance(node, str) else getattr(node, "producer_type", type(node).__name__)
        key = self._key("node", ty, [list(s) for s in input_shapes])
        if key in self.cached_exetime:
            return self.cached_exetime[key]
        builder = builder or _DEFAULT_BUILDERS.get(ty)
        if builder is not None:
            t = self.profile_new_case(builder, input_shapes)
        else:
            nbytes = 4 * (sum(int(np.prod(s)) for s in input_shapes) + (int(np.prod(output_shape)) if output_shape else 0))
            t = nbytes / (self.links["hbm_GBps"] * 1e9) * 1e3 + 0.005
        self.cached_exetime[key] = float(t)
        return float(t)

    def get_split_time(self, input_shape, axes, inds, splits, return_shape: bool = False):
        shape = self.get_split_shape({a: s for a, s in zip(axes, splits)}, input_shape)
        t = self._copy_time(int(np.prod(shape)))
        return (t, shape) if return_shape else t

    def get_concatenate_time(self, input_shapes, axis, return_shape: bool = False):
        shape = self.get_concatenate_shape(input_shapes, axis)
        t = self._copy_time(int(np.prod(shape)))
        return (t, shape) if return_shape else t

    def get_sum_time(self, input_shapes, return_shape: bool = False):
        n = int(np.prod(input_shapes[0]))
        t = 4.0 * n * (len(input_shapes) + 1) / (self.links["hbm_GBps"] * 1e9) * 1e3 + 0.005
        return (t, list(input_shapes[0])) if return_shape else t

    def get_update_time(self, shape, sparse_shape=None, states: int = 2) -> float:
        """optimizer step on a parameter of `shape` (reads grad + param + states, writes param + states); sparse: only the touched rows"""
        n = int(np.prod(sparse_shape if sparse_shape is not None else shape))
        return 4.0 * n * (2 + 2 * states + 1) / (self.links["hbm_GBps"] * 1e9) * 1e3 + 0.005

    def _copy_time(self, numel: int) -> float:
        return 8.0 * numel / (self.links["hbm_GBps"] * 1e9) * 1e3 + 0.005

    # ---- communication
    def get_dev_distance(self, from_device, to_device) -> int:
        """0 same device, 1 same host (NVLink / NVSwitch), 2 different hosts"""
        fh, th = getattr(from_device, "hostname", "localhost"), getattr(to_device, "hostname", "localhost")
        fi, ti = getattr(from_device, "device_id", from_device), getattr(to_device, "device_id", to_device)
        if fh != th:
            return 2
        return 0 if fi == ti else 1

    def _bw(self, distance: int) -> float:
        return (self.links["nic_GBps"] if distance == 2 else self.links["nvlink_GBps"]) * 1e9

    def get_comm_time(self, from_device, to_device, shape) -> float:
        d = self.get_dev_distance(from_device, to_device)
        if d == 0:
            return 0.0
        nbytes = 4 * int(np.prod(shape))
        key = self._key("p2p", d, nbytes)
        if key not in self.cached_exetime:
            self.cached_exetime[key] = self.links["latency_us"] * 1e-3 + nbytes / self._bw(d) * 1e3
        return self.cached_exetime[key]

    def get_group_comm_time(self, dev_n_shape) -> float:
        """a batch of transfers [(from, to, shape)] issued together: per-device send / receive volumes serialise, devices overlap"""
        send, recv = {}, {}
        for f, t, shape in dev_n_shape:
            c = self.get_comm_time(f, t, shape)
            send[f] = send.get(f, 0.0) + c
            recv[t] = recv.get(t, 0.0) + c
        return max(list(send.values()) + list(recv.values()) + [0.0])

    def get_allreduce_time(self, shape, device_group, primitive=NCCLOP.AllReduce) -> float:
        devs = list(getattr(device_group, "workers", None) or getattr(device_group, "devices", None) or device_group)
        n = len(devs)
        if n <= 1:
            return 0.0
        nbytes = 4 * int(np.prod(shape))
        span = max(self.get_dev_distance(devs[0], d) for d in devs[1:])
        key = self._key("coll", primitive.name, n, span, nbytes)
        if key in self.cached_exetime:
            return self.cached_exetime[key]
        if self.measure_comm:
            self.nccl_profiler = self.nccl_profiler or NCCLProfiler()
            t = self.nccl_profiler.profile_allreduce(nbytes // 4, [getattr(d, "device_id", d) for d in devs], primitive=primitive)
        else:
            # ring / switch volume per rank: 2 (n-1)/n for all-reduce, (n-1)/n for gather / scatter / broadcast-like
            vol = (2.0 if primitive == NCCLOP.AllReduce else 1.0) * (n - 1) / n * nbytes
            steps = (2 if primitive == NCCLOP.AllReduce else 1) * (n - 1)
            t = steps * self.links["latency_us"] * 1e-3 + vol / self._bw(span) * 1e3
        self.cached_exetime[key] = float(t)
        return float(t)

    def wrapped_get_allreduce_time(self, shape, device_group, status=None, primitive=NCCLOP.AllReduce, dim=None) -> float:
        """all-reduce of a tensor that `status` splits (state {dim: parts}): each group moves its own shard"""
        local = list(shape)
        for d, parts in (getattr(status, "state", None) or {}).items():
            if 0 <= int(d) < len(local):
                local[int(d)] = max(local[int(d)] // int(parts), 1)
        return self.get_allreduce_time(local, device_group, primitive)

    def get_allgather_time(self, indices_shape, value_shape, device_group) -> float:
        """sparse gradients travel as (indices, rows) pairs gathered from every rank"""
        return self.get_allreduce_time(indices_shape, device_group, NCCLOP.AllGather) + self.get_allreduce_time(value_shape, device_group, NCCLOP.AllGather)

    def wrapped_get_allgather_time(self, indices_shape, value_shape, device_group, status=None) -> float:
        return self.get_allgather_time(indices_shape, value_shape, device_group)

    def get_general_comm_time(self, pre_status, tar_status, pre_rawctx, tar_rawctx, shape, use_nccl_collectives: bool = True) -> float:
        """re-sharding between two layouts ({dim: parts} states on device lists): every target shard pulls the overlap it lacks from the
        source device holding it; transfers into one device serialise"""
        pre_devs = list(getattr(pre_rawctx, "workers", None) or pre_rawctx)
        tar_devs = list(getattr(tar_rawctx, "workers", None) or tar_rawctx)
        ps, ts = dict(getattr(pre_status, "state", pre_status) or {}), dict(getattr(tar_status, "state", tar_status) or {})

        def boxes(state, n):
            dims = sorted(k for k in state if k >= 0)
            grid = [int(state[d]) for d in dims]
            out = []
            for i in range(n):
                rem, box = i % max(int(np.prod(grid)) if grid else 1, 1), [(0, s) for s in shape]
                for d, g in zip(reversed(dims), reversed(grid)):
                    idx, rem = rem % g, rem // g
                    w = shape[d] // g
                    box[d] = (idx * w, (idx + 1) * w)
                out.append(box)
            return out
        pb, tb = boxes(ps, len(pre_devs)), boxes(ts, len(tar_devs))
        per_dev = {}
        for j, (tdev, tbox) in enumerate(zip(tar_devs, tb)):
            need = 0.0
            covered = []
            for i, (pdev, pbox) in enumerate(zip(pre_devs, pb)):
                inter = [(max(a0, b0), min(a1, b1)) for (a0, a1), (b0, b1) in zip(tbox, pbox)]
                if any(lo >= hi for lo, hi in inter) or inter in covered:
                    continue
                covered.append(inter)
                need += self.get_comm_time(pdev, tdev, [hi - lo for lo, hi in inter])
            per_dev[j] = need
        return max(per_dev.values()) if per_dev else 0.0

    # ---- shapes
    @staticmethod
    def get_split_shape(parts, shape):
        shape = list(shape)
        for d, p in (parts.items() if isinstance(parts, dict) else enumerate(parts)):
            shape[int(d)] = shape[int(d)] // int(p)
        return shape

    @staticmethod
    def get_concatenate_shape(input_shapes, dim):
        shape = list(input_shapes[0])
        shape[dim] = sum(int(s[dim]) for s in input_shapes)
        return shape

    def profile_allreduce(self, *args, **kw):
        self.nccl_profiler = self.nccl_profiler or NCCLProfiler()
        return self.nccl_profiler.profile_allreduce(*args, **kw)

    def profile_sendrecv(self, *args, **kw):
        self.nccl_profiler = self.nccl_profiler or NCCLProfiler()
        return self.nccl_profiler.profile_sendrecv(*args, **kw)

    def write_cache(self):
        if self.cache_path:
            tmp = self.cache_path + ".tmp"
            with open(tmp, "w") as f:
                json.dump(self.cached_exetime, f)
            os.replace(tmp, self.cache_path)


def _builders():
    from . import executor as e
    return {
        "matmul": lambda a, b: e.matmul_op(a, b), "relu": e.relu_op, "gelu": e.gelu_op, "tanh": e.tanh_op, "sigmoid": e.sigmoid_op,
        "softmax": e.softmax_op, "add": lambda a, b: e.add_op(a, b), "mul": lambda a, b: e.mul_op(a, b),
        "conv2d": lambda x, w: e.conv2d_op(x, w, padding=1, stride=1), "layer_norm": lambda x, s, b: e.layer_normalization_op(x, s, b),
        "maxpool": lambda x: e.max_pool2d_op(x, 2, 2, 0, 2), "avgpool": lambda x: e.avg_pool2d_op(x, 2, 2, 0, 2),
    }


class _Lazy(dict):
    def get(self, k, default=None):
        if not self:
            self.update(_builders())
        return super().get(k, default)


_DEFAULT_BUILDERS = _Lazy()
