"""The process-level entry points of the v1 API: collective communicator handles (`wrapped_mpi_nccl_init`, `new_group_comm`), the
ps-lite style role functions (`scheduler_init`, `server_init`, `worker_init` and their `_finish` twins, driven by the DMLC_*
environment), `context` / `get_current_context` / `DistConfig`, `dispatch`, the numpy `softmax_func`, and `hetu.random`.
(ref: hetu/v1/python/hetu/gpu_ops/executor.py:65-137, context.py, communicator/mpi_nccl_comm.py, random.py, gpu_ops/Dispatch.py)"""
from __future__ import annotations

import contextlib
import os
import threading
from typing import Dict, List, Optional, Sequence

import numpy as np

from .. import core, ops


# ------------------------------------------------------------------ collective communicators
class Communicator:
    """A group of ranks and the collectives over it (NCCL on GPUs, gloo on CPU) -- what `ncclInit()` / `ncclGroupInit(devices)`
    hand back in v1.  Tensors are torch tensors, numpy arrays or v1 NDArrays; results come back in the type that went in."""

    def __init__(self, ranks: Optional[Sequence[int]] = None):
        from .. import _C
        self._C = _C
        if not _C.comm_initialized() and "RANK" in os.environ:
            from .. import distributed
            distributed.init_comm_group()
        self._live = _C.comm_initialized()            # a lone process (no launcher environment) is a one-rank world
        self.ranks = [int(r) for r in ranks] if ranks is not None else list(range(_C.comm_world() if self._live else 1))
        if len(self.ranks) > 1 and self._live:
            _C.comm_create_group(self.ranks)

    # -- identity
    @property
    def rank(self) -> int:
        me = self._C.comm_rank() if self._live else 0
        return self.ranks.index(me) if me in self.ranks else -1
    @property
    def nrank(self) -> int: return len(self.ranks)                                                                        # noqa: E704
    @property
    def local_rank(self) -> int: return int(os.environ.get("LOCAL_RANK", self._C.comm_rank() if self._live else 0))        # noqa: E704
    @property
    def dev_id(self) -> int: return self.local_rank                                                                       # noqa: E704
    def getRank(self): return self.rank                                                                                   # noqa: E704,N802
    def getNRanks(self): return self.nrank                                                                                # noqa: E704,N802

    # -- data plumbing
    @staticmethod
    def _in(x):
        import torch
        if isinstance(x, np.ndarray):
            return torch.from_numpy(np.ascontiguousarray(x)), "np"
        if hasattr(x, "t") and not isinstance(x, torch.Tensor):        # v1 NDArray wrapper
            return x.t, "nd"
        return x, "t"

    @staticmethod
    def _out(y, kind, like=None):
        if kind == "np":
            return y.cpu().numpy()
        if kind == "nd":
            like.t = y
            return like
        return y

    def _run(self, fn, x, output=None):
        t, kind = self._in(x)
        y = fn(t.contiguous()) if (len(self.ranks) > 1 and self._live) else t
        if output is not None:                                          # v1 style: write into the caller's output array
            ot, okind = self._in(output)
            ot.copy_(y.reshape(ot.shape))
            return output
        return self._out(y, kind, x)

    # -- collectives (pythonic names + the v1 spellings)
    def all_reduce(self, x, op="sum", output=None): return self._run(lambda t: self._C.comm_all_reduce(t, self.ranks, op), x, output)   # noqa: E704
    def all_gather(self, x, dim=0, output=None): return self._run(lambda t: self._C.comm_all_gather(t, self.ranks, dim), x, output)      # noqa: E704
    def reduce_scatter(self, x, dim=0, output=None): return self._run(lambda t: self._C.comm_reduce_scatter(t, self.ranks, dim), x, output)   # noqa: E704,E501
    def broadcast(self, x, root=0, output=None): return self._run(lambda t: self._C.comm_broadcast(t, self.ranks, self.ranks[root]), x, output)   # noqa: E704,E501
    def reduce(self, x, root=0, op="sum", output=None): return self._run(lambda t: self._C.comm_reduce(t, self.ranks, self.ranks[root], op), x, output)   # noqa: E704,E501
    def all_to_all(self, x, output=None): return self._run(lambda t: self._C.comm_all_to_all(t, self.ranks, 0, 0), x, output)           # noqa: E704
    def send(self, x, dst, channel=0): self._C.comm_send(self._in(x)[0], self.ranks[dst], channel)                                        # noqa: E704
    def recv(self, shape, src, dtype="float32", channel=0): return self._C.comm_recv(list(shape), dtype, self.ranks[src], channel)        # noqa: E704
    def barrier(self): self._live and self._C.comm_barrier()                                                                                             # noqa: E704

    def dlarrayNcclAllReduce(self, input_arr, output_arr, dtype=None, reduceop="sum", stream=None):                       # noqa: N802
        return self.all_reduce(input_arr, _op_name(reduceop), output_arr)
    def dlarrayAllGather(self, input_arr, output_arr, dtype=None, stream=None): return self.all_gather(input_arr, 0, output_arr)          # noqa: E704,N802
    def dlarrayReduceScatter(self, input_arr, output_arr, dtype=None, reduceop="sum", stream=None): return self.reduce_scatter(input_arr, 0, output_arr)   # noqa: E704,N802,E501
    def dlarrayBroadcast(self, input_arr, output_arr, dtype=None, root=0, stream=None): return self.broadcast(input_arr, root, output_arr)   # noqa: E704,N802,E501
    def dlarrayNcclReduce(self, input_arr, output_arr, root, dtype=None, reduceop="sum", stream=None): return self.reduce(input_arr, root, _op_name(reduceop), output_arr)   # noqa: E704,N802,E501
    def dlarrayAllToAll(self, input_arr, output_arr, dtype=None, stream=None): return self.all_to_all(input_arr, output_arr)              # noqa: E704,N802
    def dlarraySend(self, arr, dtype=None, target=0, stream=None): return self.send(arr, target)                                          # noqa: E704,N802
    def dlarrayRecv(self, arr, dtype=None, src=0, stream=None):                                                           # noqa: N802
        t, _ = self._in(arr)
        t.copy_(self.recv(list(t.shape), src, str(t.dtype).replace("torch.", "")))
        return arr


def _op_name(op):
    if isinstance(op, str):
        return op.lower()
    return {0: "sum", 1: "prod", 2: "max", 3: "min"}.get(int(getattr(op, "value", op)), "sum")


_world_comm: List[Optional[Communicator]] = [None]


def wrapped_mpi_nccl_init(init_nccl: bool = True, devices: Optional[List[int]] = None) -> Optional[Communicator]:
    """the process's world communicator (created on first use; rank / size come from the torchrun or heturun environment)"""
    if _world_comm[0] is None:
        _world_comm[0] = Communicator()
    return _world_comm[0] if init_nccl else None


def new_group_comm(devices_context=None) -> Communicator:
    """a communicator over a subset of the ranks: a `DeviceGroup`, a list of device contexts, or plain rank numbers.  Every rank of
    the world must make the same sequence of calls (group creation is collective)."""
    wrapped_mpi_nccl_init()
    if devices_context is None:
        return Communicator()
    items = getattr(devices_context, "workers", None) or getattr(devices_context, "devices", None) or devices_context
    ranks = sorted(int(getattr(d, "device_id", getattr(d, "index", d))) for d in items)
    return Communicator(ranks)


def get_mpi_communicate() -> Optional[Communicator]: return _world_comm[0]                                               # noqa: E704
def get_nccl_communicate() -> Optional[Communicator]: return _world_comm[0]                                              # noqa: E704


# ------------------------------------------------------------------ parameter-server roles
_ps: Dict[str, object] = {}


def _root():
    return f"{os.environ.get('DMLC_PS_ROOT_URI', '127.0.0.1')}:{int(os.environ.get('DMLC_PS_ROOT_PORT', 0))}"


def scheduler_init() -> None:
    """start this process's scheduler role (DMLC_ROLE=scheduler): servers and workers register at DMLC_PS_ROOT_URI:PORT"""
    from .. import _C
    n_server, n_worker = int(os.environ.get("DMLC_NUM_SERVER", 1)), int(os.environ.get("DMLC_NUM_WORKER", 1))
    _ps["scheduler"] = _C.PsScheduler(n_server, n_worker, int(os.environ.get("DMLC_PS_ROOT_PORT", 0)), "0.0.0.0")


def scheduler_finish(timeout_s: float = 3600.0) -> None:
    """wait for every node to check out, then stop"""
    s = _ps.pop("scheduler", None)
    if s is not None:
        s.wait_finalized(float(timeout_s))
        s.stop()


def server_init() -> None:
    """start a server role: registers at the scheduler and serves its key range until server_finish()"""
    from .ps import ShardedPSContext
    n_worker = int(os.environ.get("DMLC_NUM_WORKER", 1))
    box: Dict[str, object] = {}
    t = threading.Thread(target=lambda: box.__setitem__("srv", ShardedPSContext.serve(_root(), num_workers=n_worker, heartbeat_s=1.0)), daemon=True)
    t.start()                                   # registration completes once all nodes are up
    _ps["server_thread"], _ps["server_box"] = t, box


def server_finish(timeout_s: float = 3600.0) -> None:
    t, box = _ps.pop("server_thread", None), _ps.pop("server_box", {})
    if t is not None:
        t.join(float(timeout_s))
    srv = box.get("srv")
    if srv is not None:
        net, client = srv
        client.barrier()                        # ps-lite's Finalize: nobody leaves before every node is done
        client.finalize()
        net.stop()


def worker_init() -> None:
    """connect this worker to the parameter servers (through the scheduler when there is one)"""
    from .ps import ShardedPSContext, connect
    if "HETU_PS_SCHEDULER" not in os.environ and "HETU_PS_ADDRESS" not in os.environ and os.environ.get("DMLC_PS_ROOT_PORT"):
        _ps["worker"] = ShardedPSContext(_root(), heartbeat_s=1.0)
    else:
        _ps["worker"] = connect()


def worker_finish() -> None:
    w = _ps.pop("worker", None)
    if w is not None and hasattr(w, "sched"):
        w.sched.barrier()
        w.finalize()


def get_worker_communicate():
    """the worker's parameter-server handle (push / pull / sparse_push / sparse_pull / barrier ...)"""
    return _ps.get("worker")


# ------------------------------------------------------------------ device contexts
class DistConfig:
    """cluster description: a yaml / dict with `nodes: [{host, servers, workers, chief}]` (ref: context.py DistConfig)"""

    def __init__(self, file: Optional[str] = None, num_local_servers: int = 0, num_local_workers: int = 1, settings: Optional[dict] = None):
        if file is not None:
            import yaml
            settings = yaml.safe_load(open(file))
        if settings is None:
            settings = {"nodes": [{"host": "localhost", "servers": num_local_servers, "workers": num_local_workers, "chief": True}]}
        self.settings = settings
        nodes = settings["nodes"]
        self.hosts = [n["host"] for n in nodes]
        self.servers = {n["host"]: int(n.get("servers", 0)) for n in nodes}
        self.workers = {n["host"]: int(n.get("workers", 0)) for n in nodes}
        chiefs = [n["host"] for n in nodes if n.get("chief")]
        assert len(chiefs) == 1, "exactly one node must be the chief"
        self.chief = chiefs[0]
        self.num_servers, self.num_workers = sum(self.servers.values()), sum(self.workers.values())
        self.enable_PS = self.num_servers > 0
        self.chief_address = settings.get("chief_address", "127.0.0.1")

    def make_ps_config(self, port: int = 13100) -> dict:
        return {"DMLC_PS_ROOT_URI": self.chief_address, "DMLC_PS_ROOT_PORT": port, "DMLC_NUM_WORKER": self.num_workers,
                "DMLC_NUM_SERVER": self.num_servers, "DMLC_PS_VAN_TYPE": "p3"}

    def __str__(self):
        return "\n".join(f"{h}: servers={self.servers[h]} workers={self.workers[h]}{' (chief)' if h == self.chief else ''}" for h in self.hosts)

    def save(self, path: str):
        import yaml
        yaml.safe_dump(self.settings, open(path, "w"))


_ctx_stack: List[object] = []


@contextlib.contextmanager
def context(ctx):
    """`with ht.context(ht.gpu(0)):` -- nodes built inside are placed on `ctx` (a device, a tuple of devices for data parallelism,
    or a DeviceGroup); the executor reads it through get_current_context()"""
    _ctx_stack.append(ctx)
    try:
        yield ctx
    finally:
        _ctx_stack.pop()


def get_current_context():
    return _ctx_stack[-1] if _ctx_stack else None


# ------------------------------------------------------------------ model-parallel annotation
def dispatch(node, parts=None):
    """Split `node` into parts[d] pieces along every dimension d for the nodes that follow (model parallelism): {dim: n} or a tuple
    of per-dimension counts.  On one process this is the identity; on several the node gets that layout through a `comm` op over
    the first prod(parts) ranks."""
    from .. import _C
    if not parts:
        return node
    split = {int(d): int(n) for d, n in (parts.items() if isinstance(parts, dict) else enumerate(parts)) if int(n) > 1}
    total = int(np.prod(list(split.values()))) if split else 1
    from .executor import annotate
    annotate(node, "dispatch_parts", dict(split))
    if total <= 1 or not _C.comm_initialized() or _C.comm_world() < total:
        return node
    ds = core.DistributedStates(total, split, sorted(split))
    out = ops.comm(node, [ds])
    return annotate(out, "dispatch_parts", dict(split))


def softmax_func(y):
    """numerically stable softmax of a numpy array along the last axis"""
    y = np.asarray(y)
    e = np.exp(y - y.max(-1, keepdims=True))
    return e / e.sum(-1, keepdims=True)


# ------------------------------------------------------------------ hetu.random
class _Random:
    """seed bookkeeping of v1 (`ht.random.set_random_seed`, `get_seed`, `get_seed_seqnum`, `step_seqnum`): one global seed plus a
    sequence number that every random node advances, so runs with the same seed draw the same streams"""

    def __init__(self):
        self.seed, self.seqnum = 0, 0

    def set_random_seed(self, seed: int):
        self.seed, self.seqnum = int(seed), 0
        core.set_seed(int(seed))
        np.random.seed(int(seed) % (2 ** 32))

    def reset_seed_seqnum(self): self.seqnum = 0                                         # noqa: E704
    def get_seed(self): return self.seed                                                 # noqa: E704
    def get_seed_seqnum(self): return self.seqnum                                        # noqa: E704
    def get_seed_status(self): return self.seed, self.seqnum                             # noqa: E704
    def step_seqnum(self, step: int = 1): self.seqnum += int(step)                       # noqa: E704
    def get_np_rand(self, step: int = 1):
        r = np.random.RandomState((self.seed + self.seqnum) % (2 ** 32))
        self.step_seqnum(step)
        return r
    # numpy-style draws under the managed seed
    def normal(self, loc=0.0, scale=1.0, size=None): return self.get_np_rand().normal(loc, scale, size)          # noqa: E704
    def uniform(self, low=0.0, high=1.0, size=None): return self.get_np_rand().uniform(low, high, size)          # noqa: E704
    def randint(self, low, high=None, size=None): return self.get_np_rand().randint(low, high, size)             # noqa: E704
    def permutation(self, x): return self.get_np_rand().permutation(x)                                           # noqa: E704


random = _Random()
