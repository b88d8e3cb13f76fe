"""tiktoken encodings (ref: python/hetu/data/tokenizers/tiktoken_tokenizer.py)"""
from . import build_tokenizer


def TikTokenizer(name="cl100k_base"):
    return build_tokenizer("tiktoken", name_or_path=name)
