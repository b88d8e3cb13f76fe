"""Partial reduce: instead of waiting for every data-parallel worker, the workers that are ready within a short window average
their gradients among themselves -- a straggler joins a later round (or reduces alone).  The scheduler does the matchmaking
(`preduce_partners`), the reduction runs over point-to-point transfers between exactly the partners, so no communicator has to be
created by ranks outside the group.  (ref: hetu/v1/python/hetu/preduce.py, ps-lite preduce_get_partner)"""
from __future__ import annotations

from typing import Sequence, Tuple

import numpy as np
import torch

from .runtime_api import get_worker_communicate, wrapped_mpi_nccl_init


class PartialReduce:
    def __init__(self, reduce_key: int = 0):
        """reduce_key: workers of different pipeline stages match under different keys"""
        self._reduce_key = int(reduce_key)
        self.ps_comm = get_worker_communicate()
        assert self.ps_comm is not None and hasattr(self.ps_comm, "sched"), "PartialReduce needs worker_init() against a scheduler"
        self.comm = wrapped_mpi_nccl_init()
        self.rank, self.nrank = self.comm.rank, self.comm.nrank
        self.rounds = 0

    def get_partner(self, max_worker: int = -1, wait_time: float = 1.0) -> Tuple[int, ...]:
        """the group for this step: returns as soon as `max_worker` workers asked, or `wait_time` ms after the first one did"""
        if max_worker < 0:
            max_worker = self.nrank
        return tuple(self.ps_comm.sched.preduce_partners(self._reduce_key, self.rank, int(max_worker), float(wait_time)))

    def preduce(self, array, partner: Sequence[int], stream=None):
        """in-place mean of `array` over `partner` (numpy array, torch tensor or v1 NDArray)"""
        partner = tuple(int(p) for p in partner)
        self.rounds += 1
        if len(partner) <= 1:
            return array
        t = array.t if hasattr(array, "t") and not isinstance(array, torch.Tensor) else array
        is_np = isinstance(t, np.ndarray)
        x = torch.from_numpy(t) if is_np else t
        C = self.comm._C
        leader, me = partner[0], self.comm._C.comm_rank()
        dt = str(x.dtype).replace("torch.", "")
        if me == leader:
            total = x.clone()
            for p in partner[1:]:
                total += C.comm_recv(list(x.shape), dt, p, 0).to(x.device)
            total /= float(len(partner))
            for p in partner[1:]:
                C.comm_send(total, p, 1)
            x.copy_(total)
        else:
            C.comm_send(x.contiguous(), leader, 0)
            x.copy_(C.comm_recv(list(x.shape), dt, leader, 1).to(x.device))
        return array
