"""HuggingFace tokenizer from a local directory (ref: python/hetu/data/tokenizers/hf_tokenizer.py)"""
from . import build_tokenizer


def HFTokenizer(name_or_path):
    return build_tokenizer("hf", name_or_path=name_or_path)
