"""tokenizer factory (ref: python/hetu/data/tokenizers/tokenizer.py `build_tokenizer`)"""
from . import ByteTokenizer, build_tokenizer  # noqa: F401
