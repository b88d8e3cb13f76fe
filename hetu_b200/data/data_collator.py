"""(ref: python/hetu/data/data_collator.py)"""
from ..engine.data_collator import DataCollatorForLanguageModel  # noqa: F401
