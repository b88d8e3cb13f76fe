"""GPT-2 byte-pair tokenizer from local vocab / merges files (ref: python/hetu/data/tokenizers/gpt2_tokenizer.py)"""
from . import build_tokenizer


def GPT2BPETokenizer(vocab_file, merge_file):
    return build_tokenizer("gpt2", vocab_file=vocab_file, merge_file=merge_file)
