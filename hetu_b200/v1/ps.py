"""Parameter-server context and the cache-enabled sparse table (HET) for the v1 PS / Hybrid modes.
Workers are threads or processes: in-process they share one native `ParameterServer`; across processes the server role
(`PSContext.serve`, or `python -m hetu.v1.launcher`) exposes it through the native TCP transport (`csrc/v1/ps_net.cc`) and
workers connect with `PSContext(address="host:port")` (or the HETU_PS_ADDRESS environment variable).  (ref: hetu/v1/python/hetu/communicator + ps-lite worker API, hetu/cstable.py)"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import numpy as np
import torch

from .. import _C

_OPT = {"none": _C.PsOptimizer.NONE, "sgd": _C.PsOptimizer.SGD, "momentum": _C.PsOptimizer.MOMENTUM, "adagrad": _C.PsOptimizer.ADAGRAD,
        "adam": _C.PsOptimizer.ADAM}


class PSContext:
    _shared: Dict[str, "_C.ParameterServer"] = {}

    def __init__(self, num_workers: int = 1, worker_id: Optional[int] = None, name: str = "default", address: Optional[str] = None):
        import os
        address = address or os.environ.get("HETU_PS_ADDRESS")
        if worker_id is None:
            worker_id = int(os.environ.get("HETU_PS_WORKER_ID", os.environ.get("WORKER_ID", "0")))
        if address:
            host, port = address.rsplit(":", 1)
            self.server = _C.PsNetClient(host, int(port))
            num_workers = self.server.num_workers()
        else:
            if name not in PSContext._shared:
                PSContext._shared[name] = _C.ParameterServer(num_workers)
            self.server = PSContext._shared[name]
        self.worker_id, self.num_workers = worker_id, num_workers
        self._keys: Dict[str, int] = {}

    @staticmethod
    def serve(num_workers: int, port: int = 0, bind_addr: str = "0.0.0.0", name: str = "default"):
        """server role: host the parameter store for `num_workers` remote workers -> the running `_C.PsNetServer`
        (`.port`, `.requests`, `.stop()`)"""
        ps = PSContext._shared.setdefault(name, _C.ParameterServer(num_workers))
        return _C.PsNetServer(ps, int(port), bind_addr)

    def key(self, name: str) -> int:
        # stable across processes (Python's str hash is salted per process)
        import zlib
        return self._keys.setdefault(name, (zlib.crc32(name.encode()) | (len(name) << 32)) & ((1 << 40) - 1))

    # ---- dense
    def init_dense(self, name, value: np.ndarray, opt="sgd", lr=0.01):
        self.server.init_dense(self.key(name), np.asarray(value, np.float32).reshape(-1).tolist(), _OPT[opt], lr)

    def push(self, name, grad: np.ndarray):
        self.server.push_dense(self.key(name), np.asarray(grad, np.float32).reshape(-1).tolist())

    def pull(self, name, shape=None) -> np.ndarray:
        v = np.asarray(self.server.pull_dense(self.key(name)), np.float32)
        return v.reshape(shape) if shape is not None else v

    def push_pull(self, name, grad, shape=None):
        v = np.asarray(self.server.push_pull_dense(self.key(name), np.asarray(grad, np.float32).reshape(-1).tolist()), np.float32)
        return v.reshape(shape) if shape is not None else v

    # ---- sparse
    def init_sparse(self, name, value: np.ndarray, opt="sgd", lr=0.01):
        value = np.asarray(value, np.float32)
        self.server.init_sparse(self.key(name), value.shape[0], value.shape[1], value.reshape(-1).tolist(), _OPT[opt], lr)

    def sparse_pull(self, name, rows: Sequence[int], width: int) -> np.ndarray:
        return np.asarray(self.server.pull_sparse(self.key(name), [int(r) for r in rows]), np.float32).reshape(len(rows), width)

    def sparse_push(self, name, rows: Sequence[int], grads: np.ndarray):
        self.server.push_sparse(self.key(name), [int(r) for r in rows], np.asarray(grads, np.float32).reshape(-1).tolist())

    # ---- consistency
    def barrier(self):
        self.server.barrier(self.worker_id)

    def ssp_init(self, staleness: int):
        self.server.ssp_init(staleness)

    def ssp_sync(self, clock: int):
        self.server.ssp_sync(self.worker_id, clock)

    def preduce(self, name, value: np.ndarray, min_workers=2, wait_ms=50):
        out, partners = self.server.preduce(self.worker_id, self.key(name), np.asarray(value, np.float32).reshape(-1).tolist(), min_workers, wait_ms)
        return np.asarray(out, np.float32).reshape(np.asarray(value).shape), partners


class CacheSparseTable:
    """client-side embedding cache with bounded staleness in front of a PS sparse table (HET, VLDB'22)"""

    def __init__(self, ps: PSContext, name: str, rows: int, width: int, limit: int, policy: str = "LRU", bound: int = 100, lr: float = 0.01):
        self.ps, self.name, self.rows, self.width, self.bound, self.lr = ps, name, rows, width, bound, lr
        pol = {"LRU": _C.CachePolicy.LRU, "LFU": _C.CachePolicy.LFU, "LFUOPT": _C.CachePolicy.LFUOPT, "LFUOpt": _C.CachePolicy.LFUOPT}[policy]
        self.cache = _C.EmbeddingCache(limit, width, pol, bound, bound)

    def embedding_lookup(self, ids) -> torch.Tensor:
        keys = [int(i) for i in np.asarray(ids).reshape(-1)]
        uniq = sorted(set(keys))
        vers = self.ps.server.row_versions(self.ps.key(self.name), uniq)
        out, miss = self.cache.lookup(uniq, vers)
        if miss:
            mk = [uniq[i] for i in miss]
            rows = self.ps.sparse_pull(self.name, mk, self.width)
            ek, eg = self.cache.insert(mk, torch.as_tensor(rows), [vers[i] for i in miss])
            if len(ek):
                self.ps.sparse_push(self.name, ek, eg.numpy())
            out[torch.as_tensor(miss)] = torch.as_tensor(rows)
        pos = {k: i for i, k in enumerate(uniq)}
        return out[torch.as_tensor([pos[k] for k in keys])].reshape(*np.asarray(ids).shape, self.width)

    def embedding_update(self, ids, grads):
        keys = [int(i) for i in np.asarray(ids).reshape(-1)]
        g = torch.as_tensor(np.asarray(grads, np.float32)).reshape(len(keys), self.width)
        uniq, inv = np.unique(np.asarray(keys), return_inverse=True)
        acc = torch.zeros(len(uniq), self.width).index_add_(0, torch.as_tensor(inv), g)
        pk, pg = self.cache.update([int(u) for u in uniq], acc, self.lr)
        if len(pk):
            self.ps.sparse_push(self.name, pk, pg.numpy())

    def stats(self):
        return self.cache.stats()


class ShardedPSContext:
    """Worker-side view of a multi-server deployment managed by the native scheduler (`_C.PsScheduler`): the worker registers,
    learns the server table, and partitions parameters over the servers the way ps-lite's key ranges do -- a dense parameter
    is cut into `num_servers` contiguous slices (slice s lives on server s), a sparse table is split by `row % num_servers`.
    Barriers go through the scheduler (worker group); every server sees `num_workers` workers.
    (ref: hetu/v1/ps-lite/src/postoffice.cc GetServerKeyRanges, kv_app.h DefaultSlicer)"""

    SERVER, WORKER = 0, 1

    def __init__(self, scheduler: str, my_host: str = "127.0.0.1", heartbeat_s: float = 0.0):
        host, port = scheduler.rsplit(":", 1)
        self.sched = _C.PsSchedulerClient(host, int(port), self.WORKER, my_host, 0)
        self.worker_id, self.num_workers, self.num_servers = self.sched.rank, self.sched.num_workers, self.sched.num_servers
        self.servers = [_C.PsNetClient(s.host, s.port) for s in self.sched.servers]
        self._keys: Dict[str, int] = {}
        self._dense_len: Dict[str, int] = {}
        if heartbeat_s > 0:
            self.sched.start_heartbeat(heartbeat_s)

    @staticmethod
    def serve(scheduler: str, my_host: str = "127.0.0.1", port: int = 0, num_workers: Optional[int] = None, heartbeat_s: float = 0.0):
        """server role: start the native store + its TCP transport, then register the address with the scheduler (blocks until
        the whole deployment has registered).  The worker count is a property of the deployment (DMLC_NUM_WORKER /
        HETU_PS_NUM_WORKERS, as in ps-lite) unless given.  -> (PsNetServer, PsSchedulerClient)"""
        import os
        if num_workers is None:
            num_workers = int(os.environ.get("DMLC_NUM_WORKER", os.environ.get("HETU_PS_NUM_WORKERS", "1")))
        host, sport = scheduler.rsplit(":", 1)
        net = _C.PsNetServer(_C.ParameterServer(int(num_workers)), int(port), "0.0.0.0")
        client = _C.PsSchedulerClient(host, int(sport), ShardedPSContext.SERVER, my_host, net.port)
        if heartbeat_s > 0:
            client.start_heartbeat(heartbeat_s)
        return net, client

    def key(self, name: str) -> int:
        import zlib
        return self._keys.setdefault(name, (zlib.crc32(name.encode()) | (len(name) << 32)) & ((1 << 40) - 1))

    def _ranges(self, n: int):
        b = self.sched.key_ranges(int(n))
        return [(int(b[i]), int(b[i + 1])) for i in range(self.num_servers)]

    # ---- dense: slice s of the flattened parameter lives on server s
    def init_dense(self, name, value: np.ndarray, opt="sgd", lr=0.01):
        v = np.asarray(value, np.float32).reshape(-1)
        self._dense_len[name] = v.size
        for s, (lo, hi) in enumerate(self._ranges(v.size)):
            if hi > lo:
                self.servers[s].init_dense(self.key(name), v[lo:hi].tolist(), _OPT[opt], lr)

    def push(self, name, grad: np.ndarray):
        g = np.asarray(grad, np.float32).reshape(-1)
        for s, (lo, hi) in enumerate(self._ranges(g.size)):
            if hi > lo:
                self.servers[s].push_dense(self.key(name), g[lo:hi].tolist())

    def pull(self, name, shape=None) -> np.ndarray:
        n = self._dense_len.get(name) or (int(np.prod(shape)) if shape is not None else None)
        assert n is not None, f"unknown size of dense parameter {name}: pass shape"
        out = np.empty(n, np.float32)
        for s, (lo, hi) in enumerate(self._ranges(n)):
            if hi > lo:
                out[lo:hi] = self.servers[s].pull_dense(self.key(name))
        return out.reshape(shape) if shape is not None else out

    def push_pull(self, name, grad, shape=None):
        g = np.asarray(grad, np.float32).reshape(-1)
        out = np.empty_like(g)
        for s, (lo, hi) in enumerate(self._ranges(g.size)):
            if hi > lo:
                out[lo:hi] = self.servers[s].push_pull_dense(self.key(name), g[lo:hi].tolist())
        return out.reshape(shape) if shape is not None else out

    # ---- sparse: row r lives on server r % S as local row r // S
    def init_sparse(self, name, value: np.ndarray, opt="sgd", lr=0.01):
        value = np.asarray(value, np.float32)
        for s in range(self.num_servers):
            part = value[s::self.num_servers]
            if part.shape[0]:
                self.servers[s].init_sparse(self.key(name), part.shape[0], part.shape[1], part.reshape(-1).tolist(), _OPT[opt], lr)

    def sparse_pull(self, name, rows: Sequence[int], width: int) -> np.ndarray:
        rows = np.asarray(rows, np.int64).reshape(-1)
        out = np.empty((rows.size, width), np.float32)
        for s in range(self.num_servers):
            sel = np.nonzero(rows % self.num_servers == s)[0]
            if sel.size:
                got = self.servers[s].pull_sparse(self.key(name), (rows[sel] // self.num_servers).tolist())
                out[sel] = np.asarray(got, np.float32).reshape(sel.size, width)
        return out

    def sparse_push(self, name, rows: Sequence[int], grads: np.ndarray):
        rows = np.asarray(rows, np.int64).reshape(-1)
        grads = np.asarray(grads, np.float32).reshape(rows.size, -1)
        for s in range(self.num_servers):
            sel = np.nonzero(rows % self.num_servers == s)[0]
            if sel.size:
                self.servers[s].push_sparse(self.key(name), (rows[sel] // self.num_servers).tolist(), grads[sel].reshape(-1).tolist())

    # ---- control plane
    def barrier(self):
        self.sched.barrier(2)          # worker group

    def dead_nodes(self, timeout_s: float):
        return self.sched.dead_nodes(float(timeout_s))

    def finalize(self):
        self.sched.finalize()


def connect(**kw):
    """the parameter-server context of this worker process: sharded over several servers when the launcher started a scheduler
    (HETU_PS_SCHEDULER), a single server otherwise (HETU_PS_ADDRESS or in-process)"""
    import os
    sched = os.environ.get("HETU_PS_SCHEDULER")
    if sched:
        return ShardedPSContext(sched, heartbeat_s=float(kw.pop("heartbeat_s", 1.0)))
    return PSContext(**kw)
