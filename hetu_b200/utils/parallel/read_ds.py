"""(ref: python/hetu/utils/parallel/read_ds.py)"""
from ...models.parallel_config import read_ds_parallel_config  # noqa: F401
from ...nn.parallel import config2ds, get_multi_ds_parallel_config  # noqa: F401
from . import parse_multi_ds_parallel_config  # noqa: F401
