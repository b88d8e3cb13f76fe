"""(ref: python/hetu/models/gpt/gpt_tokenizer.py)"""
from ...data.tokenizers.gpt2_tokenizer import GPT2BPETokenizer  # noqa: F401
