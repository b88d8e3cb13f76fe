"""v1 metric loggers: values are buffered per step under unique names, averaged over the data-parallel ranks when a communicator is
given, and flushed by `step()` (rank 0 only) to stdout / a JSON-lines file; `WandbLogger` forwards to wandb when it is importable
and otherwise keeps the same records on disk.  (ref: hetu/v1/python/hetu/logger.py)"""
from __future__ import annotations

import json
import time
from collections.abc import Iterable
from typing import Optional

import numpy as np


class HetuLogger:
    def __init__(self, rank=None, nrank=None, ctx=None, comm=None, handle=None, file: Optional[str] = None, echo: bool = True):
        self._rank, self._nrank, self._ctx, self._comm, self._handle = rank, nrank, ctx, comm, handle
        self._buffer, self.history, self._step = {}, [], 0
        self._file, self._echo, self.config = file, echo, {}

    rank = property(lambda self: self._rank)
    nrank = property(lambda self: self._nrank)
    need_log = property(lambda self: self._nrank is None or self._rank == 0)

    @staticmethod
    def item(value):
        if hasattr(value, "asnumpy"):
            value = value.asnumpy()
        if hasattr(value, "detach"):
            value = value.detach().cpu().numpy()
        if isinstance(value, np.ndarray):
            return value.item()
        while isinstance(value, Iterable) and not isinstance(value, (str, bytes)):
            assert len(value) == 1, "only single values can be logged"
            value = value[0]
        return value

    def log(self, name, value):
        assert name not in self._buffer, f"{name} already exists in log buffer!"
        self._buffer[name] = self.item(value)

    def dist_log(self, name, value):
        """mean of `value` over the ranks (a reduce to rank 0), logged there"""
        v = np.array([float(self.item(value))], np.float32)
        if self._comm is not None and (self._nrank or 1) > 1:
            v = np.asarray(self._comm.reduce(v, 0, "sum"))
        if self.need_log:
            self.log(name, float(v[0]) / float(self._nrank or 1))

    def wrapped_log(self, name, value):
        self.log(name, value) if self._nrank is None else self.dist_log(name, value)

    def step(self):
        if self._buffer and self.need_log:
            rec = dict(self._buffer, _step=self._step, _time=time.time())
            self.history.append(rec)
            self._emit(rec)
        self._buffer.clear()
        self._step += 1

    def _emit(self, rec):
        if self._echo:
            print(" ".join(f"{k}={v:.6g}" if isinstance(v, float) else f"{k}={v}" for k, v in rec.items() if not k.startswith("_time")), flush=True)
        if self._file:
            with open(self._file, "a") as f:
                f.write(json.dumps(rec) + "\n")

    def set_config(self, attrs):
        self.config.update(dict(attrs))
        if self.need_log and self._echo:
            print(attrs, flush=True)

    def __del__(self):
        try:
            if self._buffer:
                self.step()
        except Exception:      # noqa: BLE001 -- interpreter shutdown
            pass


class WandbLogger(HetuLogger):
    def __init__(self, project, name, id=None, rank=None, nrank=None, ctx=None, comm=None, handle=None, file: Optional[str] = None):   # noqa: A002
        super().__init__(rank, nrank, ctx, comm, handle, file=file or f"{project}-{name}.jsonl", echo=False)
        self._name, self._run = name, None
        if self.need_log:
            try:
                import wandb
                self._run = wandb.init(project=project, name=name, id=id, resume="allow")
            except Exception:      # noqa: BLE001 -- not installed / offline: the JSON-lines file is the record
                self._run = None

    name = property(lambda self: self._name)

    def _emit(self, rec):
        if self._run is not None:
            self._run.log({k: v for k, v in rec.items() if not k.startswith("_")}, step=rec["_step"])
        super()._emit(rec)

    def set_config(self, attrs):
        super().set_config(attrs)
        if self._run is not None:
            self._run.config.update(dict(attrs), allow_val_change=True)
