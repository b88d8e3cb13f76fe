"""(ref: python/hetu/utils/parallel/generate_ds.py)"""
from ...models.parallel_config import generate_ds_parallel_config, generate_hetero_ds_parallel_config, save_ds_parallel_config  # noqa: F401
from . import convert_strategy  # noqa: F401
