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
