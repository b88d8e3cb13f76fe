"""Process bootstrap: one process per GPU, rendezvous through torch.distributed (NCCL on GPUs, gloo on CPU).

API parity: hetu.init_comm_group / local_device / global_device_group / global_comm_barrier_* and
hetu.utils.parallel.distributed.distributed_init (ref: python/hetu/_binding/distributed/comm_group.cc:55-60,
hetu/impl/communication/rpc_comm.cc:36-266).  The reference's gRPC DeviceController is replaced by the
torch.distributed TCPStore (same role: rank assignment, id exchange, barrier, KV) -- see hetu_b200.rpc for the
stand-alone KV / heartbeat / elastic servers.
"""
from __future__ import annotations

import os
from typing import List, Optional

import torch
import torch.distributed as dist

from . import _C
from .core import DeviceGroup, device

_local_device = None
_global_group = None
_rpc_client = None


def _factory(ranks: List[int]):
    # collective over the whole world; ranks outside the group get a sentinel instead of a ProcessGroup
    pg = dist.new_group(ranks=list(ranks))
    return pg if isinstance(pg, dist.ProcessGroup) else None


def init_comm_group(device_num: Optional[int] = None, device_idxs=(), server_address: str = "127.0.0.1:23457", backend: Optional[str] = None):
    """Join the job.  World size / rank come from the launcher environment (RANK, WORLD_SIZE, LOCAL_RANK, MASTER_*)."""
    global _local_device, _global_group, _rpc_client
    _rpc_store = None
    if "RANK" not in os.environ and os.environ.get("HETU_RENDEZVOUS", "") == "rpc":
        # reference-style bootstrap: workers started by the pssh launcher know only the controller's address; the
        # DeviceController hands out rank / local device / world size (native client, csrc/runtime/rpc_client.cc) and
        # carries the address of the torch.distributed store chosen by rank 0
        import socket

        from .rpc import NativeDeviceClient
        server_address = os.environ.get("HETU_RPC_SERVER", server_address)
        _rpc_client = NativeDeviceClient(server_address, hostname=os.environ.get("HETU_LOCAL_HOSTNAME"))
        r, local, w = _rpc_client.connect()
        import datetime
        st_timeout = datetime.timedelta(seconds=float(os.environ.get("HETU_PG_TIMEOUT_S", "1800")))
        if r == 0:
            # rank 0 owns the torch.distributed store: it binds FIRST (retrying on another port if the candidate was taken in
            # the meantime) and only then publishes the address, so the other ranks can never race a half-chosen port
            host = os.environ.get("HETU_MASTER_HOST", server_address.rsplit(":", 1)[0])
            last = None
            for _ in range(16):
                sk = socket.socket()
                sk.bind(("", 0))
                port = sk.getsockname()[1]
                sk.close()
                try:
                    _rpc_store = dist.TCPStore(host, port, w, True, st_timeout, wait_for_workers=False)
                    break
                except (RuntimeError, OSError) as e:      # address in use: try the next free port
                    last = e
            else:
                raise RuntimeError(f"could not bind a store port: {last}")
            _rpc_client.put_string("torch_master", f"{host}:{port}")
        master = _rpc_client.get_string("torch_master")
        if r != 0:
            mh, mp = master.rsplit(":", 1)
            _rpc_store = dist.TCPStore(mh, int(mp), w, False, st_timeout)
        os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = master.rsplit(":", 1)
        os.environ["RANK"], os.environ["WORLD_SIZE"], os.environ["LOCAL_RANK"] = str(r), str(w), str(local)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", str(device_num or 1)))
    local_rank = int(os.environ.get("LOCAL_RANK", str(rank)))
    use_cuda = torch.cuda.is_available() and os.environ.get("HETU_B200_FORCE_CPU", "0") == "0"
    if use_cuda:
        torch.cuda.set_device(local_rank % torch.cuda.device_count())
    if world > 1 or "RANK" in os.environ:
        if not dist.is_initialized():
            if "MASTER_ADDR" not in os.environ:
                host, port = server_address.split(":")
                os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = host, port
            be = backend or ("nccl" if use_cuda else "gloo")
            kw = {}
            # lazy communicator initialisation (no device_id): point-to-point sends / receives between pipeline stages then get
            # their own 2-rank communicators.  With eager initialisation c10d runs unbatched p2p ops on the parent communicator
            # "as independent collective ops, serialized with all other ops" of the group, which interlocks 1F1B schedules of
            # tensor-parallel stages (observed as a hang of tp2 x pp2 on 4 B200s).  HETU_NCCL_EAGER_INIT=1 restores it.
            if use_cuda and os.environ.get("HETU_NCCL_EAGER_INIT", "0") == "1":
                kw["device_id"] = torch.device("cuda", local_rank % torch.cuda.device_count())
            if os.environ.get("HETU_PG_TIMEOUT_S"):       # collectives that cannot complete fail after this long (tests)
                import datetime
                kw["timeout"] = datetime.timedelta(seconds=float(os.environ["HETU_PG_TIMEOUT_S"]))
            if _rpc_store is not None:
                kw["store"] = _rpc_store
            dist.init_process_group(backend=be, rank=rank, world_size=world, **kw)
        pg = dist.distributed_c10d._get_default_group()
        _C.init_comm(rank, world, pg, _factory)
    kind = "cuda" if use_cuda else "cpu"
    _local_device = device(f"{kind}:{rank}")
    _global_group = DeviceGroup([f"{kind}:{i}" for i in range(world)])
    os.environ.setdefault("HETU_LOCAL_HOSTNAME", os.uname().nodename)
    return _local_device


def local_device():
    return _local_device if _local_device is not None else device("cuda:0" if torch.cuda.is_available() else "cpu:0")


def global_device_group():
    return _global_group if _global_group is not None else DeviceGroup([str(local_device())])


def global_comm_barrier_rpc():
    """barrier through the DeviceController when the job was bootstrapped by it, else through the process group"""
    if _rpc_client is not None:
        _rpc_client.barrier(tag="global_comm_barrier")
    elif dist.is_initialized():
        dist.barrier()


def all_ranks_ok(ok: bool) -> bool:
    """logical AND of a per-rank flag over the whole job (checkpoint publication: every rank's files must be on disk)"""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return bool(ok)
    import torch
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return bool(int(t.item()))


def global_comm_barrier_mpi():
    if dist.is_initialized():
        dist.barrier()


def rpc_client():
    """the rendezvous client of this worker (None unless HETU_RENDEZVOUS=rpc)"""
    return _rpc_client


def world_size() -> int:
    return dist.get_world_size() if dist.is_initialized() else 1


def rank() -> int:
    return dist.get_rank() if dist.is_initialized() else 0


def map_to_local_data(ds, device_index: int):
    """{split dim -> shard index} of a device under a DistributedStates (ref: hetu.map_to_local_data)."""
    return {k: v for k, v in ds.map_device_to_state_index(device_index).items() if k >= 0}
