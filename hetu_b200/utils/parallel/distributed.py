"""hetu.utils.parallel.distributed (ref: python/hetu/utils/parallel/distributed.py)"""
from . import distributed_init, get_device_index, get_local_index, get_dg_from_union  # noqa: F401
