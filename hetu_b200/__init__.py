"""hetu_b200: a B200-native distributed deep-learning framework with the capabilities and Python API of
PKU-DAIR/Hetu (define-and-run graphs, DistributedStates placement algebra, DP/ZeRO/TP/SP/PP/CP/EP parallelism,
hot switching, Galvatron planning) over hand-written sm_100a kernels.  `import hetu` is an alias of this package.
"""
import os as _os

from . import _C  # noqa: F401  native core (must be built in-tree: `python build.py`)
from . import _refpaths as _refpaths

_refpaths.install()      # reference module paths that are table entries instead of files

if _os.environ.get("HETU_NATIVE_ALLOCATOR", "0") == "1":
    # put the framework's own caching memory pool under every CUDA tensor (must happen before the first CUDA allocation)
    import torch as _torch
    if _torch.cuda.is_available():
        _C.use_native_allocator()
from .core import *  # noqa: F401,F403
from .core import (DeviceGroup, DeviceGroupUnion, DistributedStates, DistributedStatesUnion, Graph, HetuError, IntSymbol, NDArray,
                   Tensor, autocast, context, control_dependencies, cpu_offload, cur_graph, device, graph, merge_strategy,
                   profiler, recompute, run_level, subgraph)
from .ops import *  # noqa: F401,F403
from . import ops  # noqa: F401
from .graph_api import gradients, run_graph  # noqa: F401
from .optim import (AdamOptimizer, SGDOptimizer, GradScaler, AdaGradOptimizer, AMSGradOptimizer, AdamWOptimizer,  # noqa: F401
                    LambOptimizer)
from . import nn  # noqa: F401
from . import logger  # noqa: F401  (module: hetu.logger.info(...), hetu.logger.get_logger(name))
from .distributed import (init_comm_group, local_device, global_device_group, global_comm_barrier_rpc,  # noqa: F401
                          global_comm_barrier_mpi, map_to_local_data)



def memory_pool_summary(device: str = "cuda:0") -> str:
    """statistics of the native caching pool of `device` (reserved / allocated / peak, splits, merges, cache hits)"""
    if _os.environ.get("HETU_NATIVE_ALLOCATOR", "0") == "1" and device.startswith("cuda"):
        return _C.tensor_allocator(device).summary()      # caching / bfc / stream_ordered (HETU_MEMORY_POOL)
    return _C.get_memory_pool(device).summary()


Dataloader = _C.Dataloader       # native prefetching loader (ref: hetu.Dataloader, hetu/graph/data/dataloader.h)

__version__ = "0.1.0"
