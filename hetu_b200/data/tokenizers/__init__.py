"""Tokenizer factory (ref: python/hetu/data/tokenizers/*): GPT-2 BPE / HuggingFace / sentencepiece / tiktoken when their
vocabulary files are available locally, and a dependency-free byte-level tokenizer otherwise (no network in CI)."""
from __future__ import annotations

from typing import List, Optional


class ByteTokenizer:
    """UTF-8 bytes + 3 specials; always available"""
    pad_id, bos_id, eos_id = 256, 257, 258
    vocab_size = 259

    def encode(self, text: str, add_special_tokens: bool = True) -> List[int]:
        ids = list(text.encode("utf-8"))
        return [self.bos_id] + ids + [self.eos_id] if add_special_tokens else ids

    def decode(self, ids) -> str:
        return bytes(i for i in ids if i < 256).decode("utf-8", errors="replace")

    @property
    def pad(self):
        return self.pad_id

    @property
    def eod(self):
        return self.eos_id


class _HFWrapper:
    def __init__(self, tok):
        self.tok = tok
        self.vocab_size = len(tok)
        self.pad_id = tok.pad_token_id if tok.pad_token_id is not None else (tok.eos_token_id or 0)
        self.eos_id = tok.eos_token_id or 0
        self.bos_id = tok.bos_token_id or self.eos_id

    def encode(self, text, add_special_tokens=True):
        return self.tok.encode(text, add_special_tokens=add_special_tokens)

    def decode(self, ids):
        return self.tok.decode(ids)

    pad = property(lambda self: self.pad_id)
    eod = property(lambda self: self.eos_id)


class _SPWrapper:
    def __init__(self, model_file):
        import sentencepiece as spm
        self.sp = spm.SentencePieceProcessor(model_file=model_file)
        self.vocab_size = self.sp.get_piece_size()
        self.pad_id = self.sp.pad_id() if self.sp.pad_id() >= 0 else 0
        self.eos_id, self.bos_id = self.sp.eos_id(), self.sp.bos_id()

    def encode(self, text, add_special_tokens=True):
        ids = self.sp.encode(text)
        return [self.bos_id] + ids + [self.eos_id] if add_special_tokens else ids

    def decode(self, ids):
        return self.sp.decode(list(ids))

    pad = property(lambda self: self.pad_id)
    eod = property(lambda self: self.eos_id)


def build_tokenizer(tokenizer_type: str = "byte", vocab_file: Optional[str] = None, merge_file: Optional[str] = None,
                    name_or_path: Optional[str] = None):
    t = tokenizer_type.lower()
    if t in ("byte", "bytes"):
        return ByteTokenizer()
    if t in ("gpt2bpetokenizer", "gpt2", "gpt2bpe"):
        from transformers import GPT2TokenizerFast
        return _HFWrapper(GPT2TokenizerFast(vocab_file=vocab_file, merges_file=merge_file))
    if t in ("hf", "huggingface", "hftokenizer"):
        from transformers import AutoTokenizer
        return _HFWrapper(AutoTokenizer.from_pretrained(name_or_path, local_files_only=True))
    if t in ("sentencepiece", "sentencepiecetokenizer", "llama"):
        return _SPWrapper(vocab_file or name_or_path)
    if t in ("tiktoken", "tiktokentokenizer"):
        import tiktoken
        enc = tiktoken.get_encoding(name_or_path or "cl100k_base")

        class _T:
            vocab_size, pad_id, eos_id, bos_id = enc.n_vocab, 0, enc.eot_token, enc.eot_token
            encode = staticmethod(lambda text, add_special_tokens=True: enc.encode(text))
            decode = staticmethod(enc.decode)
        return _T()
    raise ValueError(f"unknown tokenizer type {tokenizer_type}")


# class-style constructors of the reference's per-tokenizer modules (ref: python/hetu/data/tokenizers/{gpt2,hf,sentencepiece,
# tiktoken}_tokenizer.py, models/llama/llama_tokenizer.py)
def GPT2BPETokenizer(vocab_file, merge_file):
    return build_tokenizer("gpt2", vocab_file=vocab_file, merge_file=merge_file)


def HFTokenizer(name_or_path):
    return build_tokenizer("hf", name_or_path=name_or_path)


def SentencePieceTokenizer(model_file):
    return build_tokenizer("sentencepiece", vocab_file=model_file)


LlamaTokenizer = SentencePieceTokenizer


def TikTokenizer(name="cl100k_base"):
    return build_tokenizer("tiktoken", name_or_path=name)
