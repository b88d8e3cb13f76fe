"""(ref: python/hetu/data/messages/utils.py)"""
from . import build_chat_sample  # noqa: F401
