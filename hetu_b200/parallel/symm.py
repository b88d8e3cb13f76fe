"""Symmetric memory over NVLink: a buffer that every rank of the node maps from every other rank (CUDA IPC), plus the
device-side collectives / fused GEMM->reduce-scatter built on it (csrc/runtime/symm_mem.*, csrc/kernels/symm_comm.cu).

    buf = SymmetricBuffer("tp_stage", nbytes)          # collective: all ranks, same order
    x = buf.tensor([T, H], "bfloat16")                 # view of this rank's copy
    buf.all_reduce_(x)                                  # in place, in-kernel peer loads (no NCCL)
"""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.distributed as dist

from .. import _C


def symm_available() -> bool:
    return torch.cuda.is_available() and dist.is_initialized() and dist.get_world_size() > 1 and dist.get_world_size() <= 8


def multicast_available() -> bool:
    """NVLink-SHARP multicast objects (cuMulticastCreate) are supported by this rank's device / driver"""
    return symm_available() and bool(_C.symm_multicast_supported())


class SymmetricBuffer:
    """`multicast=True` allocates through the CUDA VMM API and adds an NVLS multicast mapping (csrc/runtime/symm_vmm.cc): the
    `mc_*` collectives then let the NVSwitch do the reduction (multimem.ld_reduce) and the replication (multimem.st)."""

    def __init__(self, name: str, nbytes: int, multicast: bool = False):
        assert dist.is_initialized(), "init_comm_group first"
        self.name, self.nbytes = name, int(nbytes)
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        handles: List[bytes] = [None] * self.world
        if multicast:
            dist.all_gather_object(handles, _C.symm_alloc_vmm(name, self.nbytes, self.rank, self.world))
            _C.symm_open_vmm(name, handles)
            dist.barrier()                       # every device has joined the multicast object before memory is bound
            _C.symm_bind_multicast(name)
            torch.cuda.synchronize()
            dist.barrier()
            self.has_multicast = bool(_C.symm_has_multicast(name))
            return
        self.has_multicast = False
        handle = _C.symm_alloc(name, self.nbytes, self.rank, self.world)
        dist.all_gather_object(handles, handle)
        _C.symm_open(name, handles)
        dist.barrier()

    # ---- NVLS collectives (need has_multicast) ----
    def mc_all_reduce_(self, elems: int, bf16: bool = True, byte_offset: int = 0):
        """in place on every rank's copy: one switch-side load-reduce + one multicast store per 16 bytes"""
        _C.symm_mc_all_reduce(self.name, int(byte_offset), int(elems), bool(bf16))

    def mc_reduce_scatter(self, elems_per_rank: int, out: torch.Tensor, byte_offset: int = 0):
        _C.symm_mc_reduce_scatter(self.name, int(byte_offset), out, int(elems_per_rank))
        return out

    def mc_all_gather(self, src: torch.Tensor, byte_offset: int = 0):
        """every rank's buffer[byte_offset + r * src.nbytes : ...] = rank r's `src`"""
        _C.symm_mc_all_gather(self.name, src, int(byte_offset))

    def tensor(self, shape: Sequence[int], dtype="bfloat16", byte_offset: int = 0) -> torch.Tensor:
        return _C.symm_tensor(self.name, int(byte_offset), [int(s) for s in shape], str(dtype))

    def barrier(self):
        _C.symm_barrier(self.name)

    def all_gather(self, nbytes_per_rank: int, out: torch.Tensor, byte_offset: int = 0):
        _C.symm_all_gather(self.name, int(byte_offset), out, int(nbytes_per_rank))
        return out

    def reduce_scatter(self, elems_per_rank: int, out: torch.Tensor, byte_offset: int = 0):
        _C.symm_reduce_scatter(self.name, int(byte_offset), out, int(elems_per_rank))
        return out

    def all_reduce_(self, elems: int, bf16: bool = True, byte_offset: int = 0):
        """in place on this rank's copy; the buffer must hold 2x the payload (scratch for the reduced chunks)"""
        _C.symm_all_reduce(self.name, int(byte_offset), int(elems), bool(bf16))

    def all_to_all(self, nbytes_per_chunk: int, out: torch.Tensor, byte_offset: int = 0):
        _C.symm_all_to_all(self.name, int(byte_offset), out, int(nbytes_per_chunk))
        return out

    def gemm_reduce_scatter(self, x: torch.Tensor, w: torch.Tensor, bias=None, residual=None) -> torch.Tensor:
        return _C.gemm_reduce_scatter(x, w, self.name, bias, residual)
