"""hetu_b200: a B200-native distributed deep-learning framework with the capabilities and Python API of
PKU-DAIR/Hetu (define-and-run graphs, DistributedStates placement algebra, DP/ZeRO/TP/SP/PP/CP/EP parallelism,
hot switching, Galvatron planning) over hand-written sm_100a kernels.  `import hetu` is an alias of this package.
"""
from . import _C  # noqa: F401  native core (must be built in-tree: `python build.py`)
from .core import *  # noqa: F401,F403
from .core import (DeviceGroup, DeviceGroupUnion, DistributedStates, DistributedStatesUnion, Graph, HetuError, IntSymbol, NDArray,
                   Tensor, autocast, context, control_dependencies, cpu_offload, cur_graph, device, graph, merge_strategy,
                   profiler, recompute, run_level, subgraph)
from .ops import *  # noqa: F401,F403
from . import ops  # noqa: F401
from .graph_api import gradients, run_graph  # noqa: F401
from .optim import AdamOptimizer, SGDOptimizer, GradScaler  # noqa: F401
from . import nn  # noqa: F401
from . import logger  # noqa: F401  (module: hetu.logger.info(...), hetu.logger.get_logger(name))
from .distributed import (init_comm_group, local_device, global_device_group, global_comm_barrier_rpc,  # noqa: F401
                          global_comm_barrier_mpi, map_to_local_data)

Dataloader = _C.Dataloader       # native prefetching loader (ref: hetu.Dataloader, hetu/graph/data/dataloader.h)

__version__ = "0.1.0"
