"""v1 stream / event handles over CUDA streams (no-ops on a CPU-only process) and the parameter-server completion events.
(ref: hetu/v1/python/hetu/stream.py)"""
from __future__ import annotations

import time

import torch


class Stream:
    def __init__(self, ctx=None):
        self.ctx = ctx
        self.handle = torch.cuda.Stream(device=int(getattr(ctx, "index", 0))) if (torch.cuda.is_available() and getattr(ctx, "kind", "cpu") == "gpu") else None

    def sync(self):
        if self.handle is not None:
            self.handle.synchronize()


def create_stream_handle(ctx): return Stream(ctx)          # noqa: E704


class Event:
    def __init__(self, ctx=None):
        self.ctx = ctx
        self.handle = torch.cuda.Event(enable_timing=True) if (torch.cuda.is_available() and getattr(ctx, "kind", "cpu") == "gpu") else None
        self._t = None

    def record(self, stream_handle=None):
        if self.handle is not None:
            self.handle.record(getattr(stream_handle, "handle", None) or torch.cuda.current_stream())
        self._t = time.perf_counter()

    def sync(self):
        if self.handle is not None:
            self.handle.synchronize()

    def time_since(self, event) -> float:
        """ms from `event` to this event"""
        if self.handle is not None and event.handle is not None:
            return event.handle.elapsed_time(self.handle)
        return (self._t - event._t) * 1e3


def create_event_handle(ctx): return Event(ctx)            # noqa: E704


class PSEvent:
    """completion of the parameter-server requests issued for one node: the worker's client calls are synchronous here, so the event
    is complete as soon as it has been updated"""

    def __init__(self, comm, nid):
        self.comm, self.nid, self.updated = comm, nid, False

    def update(self): self.updated = True                  # noqa: E704
    def sync(self): self.updated = False                   # noqa: E704


class CSEvent(PSEvent):
    def __init__(self, comm, nid):
        super().__init__(comm, nid)
        self.ts = None

    def update_ts(self, ts):
        self.ts, self.updated = ts, True
