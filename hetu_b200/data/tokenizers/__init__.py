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
        from .bpe import GPT2BPE                      # own byte-level BPE (no dependency on the transformers tokenizers)
        return GPT2BPE(vocab_file, merge_file)
    if t in ("gpt2-hf", "gpt2_hf"):
        from transformers import GPT2TokenizerFast
        return _HFWrapper(GPT2TokenizerFast(vocab_file=vocab_file, merges_file=merge_file))
    if t in ("hf", "huggingface", "hftokenizer"):
        from transformers import AutoTokenizer
        return _HFWrapper(AutoTokenizer.from_pretrained(name_or_path, local_files_only=True))
    if t in ("sentencepiece", "sentencepiecetokenizer", "llama"):
        return _SPWrapper(vocab_file or name_or_path)
    if t in ("bert", "wordpiece", "berttokenizer"):
        from .wordpiece import BertTokenizer
        return BertTokenizer(vocab_file or name_or_path)
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


# ---- base classes of the reference's tokenizer package (ref: python/hetu/data/tokenizers/{utils.py, pretrained_tokenizer.py})
class SpecialToken:
    """names of the special tokens a tokenizer may define"""
    PAD, BOS, EOS, UNK, SEP, CLS, MASK = "pad", "bos", "eos", "unk", "sep", "cls", "mask"

    def __init__(self, content: str, token_id: Optional[int] = None, kind: str = ""):
        self.content, self.token_id, self.kind = content, token_id, kind

    def __repr__(self):
        return f"SpecialToken({self.content!r}, id={self.token_id}, kind={self.kind!r})"


class BaseTokenizer:
    """interface every tokenizer of this package satisfies: encode / decode, vocab_size, pad / bos / eos ids, batch helpers"""
    vocab_size: int = 0
    pad_id: int = 0
    bos_id: int = 0
    eos_id: int = 0

    def encode(self, text: str, add_special_tokens: bool = True) -> List[int]:
        raise NotImplementedError

    def decode(self, ids) -> str:
        raise NotImplementedError

    def batch_encode(self, texts, max_length: Optional[int] = None, padding: bool = False, add_special_tokens: bool = True):
        rows = [self.encode(t, add_special_tokens)[:max_length] if max_length else self.encode(t, add_special_tokens) for t in texts]
        if padding:
            width = max(len(r) for r in rows)
            rows = [r + [self.pad_id] * (width - len(r)) for r in rows]
        return rows

    def batch_decode(self, rows):
        return [self.decode([i for i in r if i != self.pad_id]) for r in rows]

    @property
    def pad(self):
        return self.pad_id

    @property
    def eod(self):
        return self.eos_id


class PreTrainedTokenizer(BaseTokenizer):
    """a tokenizer restored from a local directory (`from_pretrained(dir)`: HuggingFace files, a sentencepiece model or a GPT-2
    vocab/merges pair are recognised) behind the BaseTokenizer interface"""

    def __init__(self, inner):
        self.inner = inner
        for k in ("vocab_size", "pad_id", "bos_id", "eos_id"):
            setattr(self, k, getattr(inner, k, 0))

    @classmethod
    def from_pretrained(cls, path: str):
        import os
        files = set(os.listdir(path)) if os.path.isdir(path) else set()
        if {"vocab.json", "merges.txt"} <= files and "tokenizer.json" not in files and "tokenizer_config.json" not in files:
            return cls(build_tokenizer("gpt2", vocab_file=os.path.join(path, "vocab.json"), merge_file=os.path.join(path, "merges.txt")))
        sp = next((f for f in files if f.endswith(".model")), None)
        if sp and "tokenizer.json" not in files:
            return cls(build_tokenizer("sentencepiece", vocab_file=os.path.join(path, sp)))
        return cls(build_tokenizer("hf", name_or_path=path))

    def encode(self, text, add_special_tokens: bool = True):
        return list(self.inner.encode(text, add_special_tokens=add_special_tokens))

    def decode(self, ids):
        return self.inner.decode(list(ids))
