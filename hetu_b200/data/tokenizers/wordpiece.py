"""BERT WordPiece tokenizer from a `vocab.txt`: basic tokenisation (text cleaning, CJK characters as single tokens, lower-casing and
accent stripping, punctuation splitting) followed by greedy longest-match-first sub-word splitting with `##` continuation pieces.
No third-party dependency.  (ref: hetu/v1/python/hetu/tokenizers/bert_tokenizer.py, hetu/v1/examples/nlp/bert/tokenization.py)"""
from __future__ import annotations

import unicodedata
from typing import Dict, List, Optional, Sequence, Tuple


def _is_whitespace(ch: str) -> bool:
    return ch in " \t\n\r" or unicodedata.category(ch) == "Zs"


def _is_control(ch: str) -> bool:
    return ch not in "\t\n\r" and unicodedata.category(ch).startswith("C")


def _is_punctuation(ch: str) -> bool:
    cp = ord(ch)
    if 33 <= cp <= 47 or 58 <= cp <= 64 or 91 <= cp <= 96 or 123 <= cp <= 126:
        return True
    return unicodedata.category(ch).startswith("P")


def _is_cjk(cp: int) -> bool:
    return (0x4E00 <= cp <= 0x9FFF or 0x3400 <= cp <= 0x4DBF or 0x20000 <= cp <= 0x2A6DF or 0x2A700 <= cp <= 0x2B73F or 0x2B740 <= cp <= 0x2B81F
            or 0x2B820 <= cp <= 0x2CEAF or 0xF900 <= cp <= 0xFAFF or 0x2F800 <= cp <= 0x2FA1F)


class BasicTokenizer:
    def __init__(self, do_lower_case: bool = True, never_split: Sequence[str] = ()):
        self.do_lower_case, self.never_split = do_lower_case, set(never_split)

    def tokenize(self, text: str) -> List[str]:
        cleaned = []
        for ch in text:
            cp = ord(ch)
            if cp == 0 or cp == 0xFFFD or _is_control(ch):
                continue
            if _is_cjk(cp):
                cleaned.append(f" {ch} ")
            else:
                cleaned.append(" " if _is_whitespace(ch) else ch)
        out: List[str] = []
        for tok in "".join(cleaned).split():
            if tok in self.never_split:
                out.append(tok)
                continue
            if self.do_lower_case:
                tok = "".join(c for c in unicodedata.normalize("NFD", tok.lower()) if unicodedata.category(c) != "Mn")
            cur = ""
            for ch in tok:
                if _is_punctuation(ch):
                    if cur:
                        out.append(cur)
                        cur = ""
                    out.append(ch)
                else:
                    cur += ch
            if cur:
                out.append(cur)
        return out


class WordPieceTokenizer:
    def __init__(self, vocab: Dict[str, int], unk_token: str = "[UNK]", max_chars_per_word: int = 100):
        self.vocab, self.unk, self.max_chars = vocab, unk_token, max_chars_per_word

    def tokenize(self, word: str) -> List[str]:
        if len(word) > self.max_chars:
            return [self.unk]
        pieces, start = [], 0
        while start < len(word):
            end, cur = len(word), None
            while start < end:
                sub = ("##" if start > 0 else "") + word[start:end]
                if sub in self.vocab:
                    cur = sub
                    break
                end -= 1
            if cur is None:
                return [self.unk]
            pieces.append(cur)
            start = end
        return pieces


class BertTokenizer:
    """`encode(text, text_pair)` -> ids with [CLS] / [SEP]; `encode_plus` also returns token type ids and the attention mask"""

    def __init__(self, vocab_file: str, do_lower_case: bool = True, unk_token="[UNK]", sep_token="[SEP]", pad_token="[PAD]", cls_token="[CLS]",
                 mask_token="[MASK]"):
        self.vocab: Dict[str, int] = {}
        with open(vocab_file, encoding="utf-8") as f:
            for i, line in enumerate(f):
                self.vocab[line.rstrip("\n")] = i
        self.inv = {i: t for t, i in self.vocab.items()}
        self.specials = (unk_token, sep_token, pad_token, cls_token, mask_token)
        self.basic = BasicTokenizer(do_lower_case, never_split=self.specials)
        self.wordpiece = WordPieceTokenizer(self.vocab, unk_token)
        self.unk_id, self.sep_id, self.pad_id = self.vocab[unk_token], self.vocab[sep_token], self.vocab[pad_token]
        self.cls_id, self.mask_id = self.vocab[cls_token], self.vocab[mask_token]
        self.bos_id, self.eos_id = self.cls_id, self.sep_id
        self.vocab_size = len(self.vocab)

    def tokenize(self, text: str) -> List[str]:
        out: List[str] = []
        for tok in self.basic.tokenize(text):
            out.extend([tok] if tok in self.specials else self.wordpiece.tokenize(tok))
        return out

    def convert_tokens_to_ids(self, tokens: Sequence[str]) -> List[int]:
        return [self.vocab.get(t, self.unk_id) for t in tokens]

    def convert_ids_to_tokens(self, ids: Sequence[int]) -> List[str]:
        return [self.inv.get(int(i), self.specials[0]) for i in ids]

    def encode(self, text: str, text_pair: Optional[str] = None, add_special_tokens: bool = True, max_length: Optional[int] = None) -> List[int]:
        return self.encode_plus(text, text_pair, add_special_tokens, max_length)[0]

    def encode_plus(self, text: str, text_pair: Optional[str] = None, add_special_tokens: bool = True, max_length: Optional[int] = None,
                    padding: bool = False) -> Tuple[List[int], List[int], List[int]]:
        """-> (input ids, token type ids, attention mask); over-long pairs are truncated longest-first"""
        a = self.convert_tokens_to_ids(self.tokenize(text))
        b = self.convert_tokens_to_ids(self.tokenize(text_pair)) if text_pair is not None else None
        extra = (3 if b is not None else 2) if add_special_tokens else 0
        if max_length is not None:
            while len(a) + (len(b) if b else 0) + extra > max_length:
                if b and len(b) >= len(a):
                    b.pop()
                else:
                    a.pop()
        if add_special_tokens:
            ids = [self.cls_id] + a + [self.sep_id] + ((b + [self.sep_id]) if b is not None else [])
            types = [0] * (len(a) + 2) + ([1] * (len(b) + 1) if b is not None else [])
        else:
            ids, types = a + (b or []), [0] * len(a) + [1] * len(b or [])
        mask = [1] * len(ids)
        if padding and max_length is not None and len(ids) < max_length:
            n = max_length - len(ids)
            ids, types, mask = ids + [self.pad_id] * n, types + [0] * n, mask + [0] * n
        return ids, types, mask

    def decode(self, ids: Sequence[int], skip_special_tokens: bool = True) -> str:
        toks = [t for t in self.convert_ids_to_tokens(ids) if not (skip_special_tokens and t in self.specials)]
        text = " ".join(toks).replace(" ##", "")
        return text

    @property
    def pad(self):
        return self.pad_id

    @property
    def eod(self):
        return self.sep_id
