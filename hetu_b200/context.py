"""context managers of the Python API (ref: python/hetu/context.py:8-299)"""
from .core import (autocast, context, control_dependencies, cpu_offload, graph, merge_strategy, profiler, recompute, run_level,  # noqa: F401
                   subgraph)
