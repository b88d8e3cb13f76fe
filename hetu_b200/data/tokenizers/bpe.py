"""GPT-2 byte-level BPE tokenizer from `vocab.json` + `merges.txt`: text is split by the GPT-2 pre-tokenisation pattern, every piece is
mapped to printable byte symbols, and the ranked merges are applied until no pair applies.  No third-party tokenizer dependency
(the `regex` module provides the Unicode classes of the pattern).
(ref: python/hetu/data/tokenizers/gpt2_tokenizer.py, python/hetu/models/gpt/gpt_tokenizer.py)"""
from __future__ import annotations

import json
from functools import lru_cache
from typing import Dict, List, Tuple

import regex as re

_PATTERN = r"""'s|'t|'re|'ve|'m|'ll|'d| ?\p{L}+| ?\p{N}+| ?[^\s\p{L}\p{N}]+|\s+(?!\S)|\s+"""


@lru_cache()
def bytes_to_unicode() -> Dict[int, str]:
    """a reversible byte <-> printable character table (control / whitespace bytes are shifted above 255)"""
    keep = list(range(ord("!"), ord("~") + 1)) + list(range(ord("¡"), ord("¬") + 1)) + list(range(ord("®"), ord("ÿ") + 1))
    chars, n = keep[:], 0
    for b in range(256):
        if b not in keep:
            keep.append(b)
            chars.append(256 + n)
            n += 1
    return dict(zip(keep, (chr(c) for c in chars)))


class GPT2BPE:
    def __init__(self, vocab_file: str, merges_file: str, eos_token: str = "<|endoftext|>"):
        with open(vocab_file, encoding="utf-8") as f:
            self.encoder: Dict[str, int] = json.load(f)
        self.decoder = {i: t for t, i in self.encoder.items()}
        with open(merges_file, encoding="utf-8") as f:
            lines = [l.rstrip("\n") for l in f if l.strip() and not l.startswith("#version")]
        self.ranks: Dict[Tuple[str, str], int] = {tuple(l.split()): i for i, l in enumerate(lines)}
        self.byte_enc = bytes_to_unicode()
        self.byte_dec = {c: b for b, c in self.byte_enc.items()}
        self.pat = re.compile(_PATTERN)
        self.cache: Dict[str, List[str]] = {}
        self.eos_token = eos_token
        self.eos_id = self.encoder.get(eos_token, len(self.encoder) - 1)
        self.bos_id = self.pad_id = self.eos_id
        self.vocab_size = len(self.encoder)

    def bpe(self, token: str) -> List[str]:
        if token in self.cache:
            return self.cache[token]
        word = list(token)
        while len(word) > 1:
            pairs = {(a, b) for a, b in zip(word, word[1:])}
            best = min(pairs, key=lambda p: self.ranks.get(p, float("inf")))
            if best not in self.ranks:
                break
            first, second = best
            merged, i = [], 0
            while i < len(word):
                if i < len(word) - 1 and word[i] == first and word[i + 1] == second:
                    merged.append(first + second)
                    i += 2
                else:
                    merged.append(word[i])
                    i += 1
            word = merged
        self.cache[token] = word
        return word

    def tokenize(self, text: str) -> List[str]:
        out: List[str] = []
        for piece in self.pat.findall(text):
            out.extend(self.bpe("".join(self.byte_enc[b] for b in piece.encode("utf-8"))))
        return out

    def encode(self, text: str, add_special_tokens: bool = False) -> List[int]:
        unk = self.encoder.get("<|unk|>", self.eos_id)
        ids = [self.encoder.get(t, unk) for t in self.tokenize(text)]
        return ids + [self.eos_id] if add_special_tokens else ids

    def decode(self, ids) -> str:
        text = "".join(self.decoder.get(int(i), "") for i in ids)
        return bytearray(self.byte_dec[c] for c in text if c in self.byte_dec).decode("utf-8", errors="replace")

    @property
    def pad(self):
        return self.pad_id

    @property
    def eod(self):
        return self.eos_id
