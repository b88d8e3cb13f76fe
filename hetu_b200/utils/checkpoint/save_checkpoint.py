"""(ref: python/hetu/utils/checkpoint/save_checkpoint.py)"""
from .legacy import save_checkpoint  # noqa: F401
