"""worker started WITHOUT RANK / WORLD_SIZE: rank, world size and the torch.distributed store address all come from the
DeviceController through the native client (reference-style pssh + gRPC bootstrap)"""
import sys

import torch

import hetu_b200 as ht
from hetu_b200 import distributed

ht.init_comm_group(int(sys.argv[2]), server_address=sys.argv[1])
c = distributed.rpc_client()
t = ht._C.comm_all_reduce(torch.ones(4) * (c.rank + 1), list(range(c.world_size)), "sum")
distributed.global_comm_barrier_rpc()
print(f"BOOT rank={c.rank} local={c.local_device} world={c.world_size} sum={float(t[0])}")
c.exit()
