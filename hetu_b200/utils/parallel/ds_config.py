"""(ref: python/hetu/utils/parallel/ds_config.py)"""
from . import RecomputeConfig, StrategyConfig, convert_strategy, generate_recompute_config  # noqa: F401
