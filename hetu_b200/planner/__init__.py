"""Galvatron automatic parallelism planner (ref: tools/Galvatron -- csrc/dp_core.cpp, galvatron/core/{profiler,
search_engine,cost_model}.py, profile_hardware/*): profile -> cost model -> layer-wise dynamic programming (C++ core) ->
plan JSON -> ds_parallel_config consumed by the executor."""
from .cost_model import LayerProfile, HardwareProfile, MemoryCostModel, TimeCostModel, Strategy  # noqa: F401
from .search_engine import GalvatronSearchEngine, galvatron_plan_to_ds_parallel_config  # noqa: F401
from .profiler import ModelProfiler, HardwareProfiler, profile_overlap_coefficient  # noqa: F401
from .runtime import load_plan, gen_comm_groups, build_hybrid_parallel_model, GalvatronRuntime  # noqa: F401
