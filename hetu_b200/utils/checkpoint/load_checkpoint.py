"""(ref: python/hetu/utils/checkpoint/load_checkpoint.py)"""
from .legacy import convert_llama_hf_to_ht, load_checkpoint, load_checkpoint_from_megatron  # noqa: F401
