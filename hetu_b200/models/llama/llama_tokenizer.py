"""(ref: python/hetu/models/llama/llama_tokenizer.py)"""
from ...data.tokenizers.sentencepiece_tokenizer import SentencePieceTokenizer as LlamaTokenizer  # noqa: F401
