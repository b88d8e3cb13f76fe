"""sentencepiece model file tokenizer (ref: python/hetu/data/tokenizers/sentencepiece_tokenizer.py)"""
from . import build_tokenizer


def SentencePieceTokenizer(model_file):
    return build_tokenizer("sentencepiece", vocab_file=model_file)
