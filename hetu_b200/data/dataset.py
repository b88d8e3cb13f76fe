"""Datasets (ref: python/hetu/data/dataset.py, examples/hydraulis indexed datasets)."""
from __future__ import annotations

import json
from typing import Optional, Sequence

import numpy as np


class JsonDataset:
    """one JSON object per line (or a JSON list); `key` selects the text field; tokenised lazily and cached"""

    def __init__(self, json_file: str, key: str = "text", tokenizer=None, max_seq_len: int = 1024, pad_id: int = 0, cache: bool = True):
        self.key, self.tokenizer, self.max_seq_len, self.pad_id = key, tokenizer, max_seq_len, pad_id
        with open(json_file) as f:
            head = f.read(1)
            f.seek(0)
            self.records = json.load(f) if head == "[" else [json.loads(l) for l in f if l.strip()]
        self._cache = {} if cache else None

    def __len__(self):
        return len(self.records)

    def tokens(self, i: int) -> np.ndarray:
        if self._cache is not None and i in self._cache:
            return self._cache[i]
        rec = self.records[i]
        text = rec[self.key] if isinstance(rec, dict) else rec
        ids = self.tokenizer.encode(text) if self.tokenizer is not None else [ord(c) % 256 for c in text]
        ids = np.asarray(ids[: self.max_seq_len + 1], dtype=np.int64)
        if self._cache is not None:
            self._cache[i] = ids
        return ids

    def __getitem__(self, i):
        return self.tokens(i)


class SyntheticDataset:
    """random token streams with a configurable length distribution (benchmarks / tests; no network needed)"""

    def __init__(self, num_samples: int, vocab_size: int, max_seq_len: int, min_seq_len: Optional[int] = None, seed: int = 0,
                 length_distribution: str = "fixed"):
        self.n, self.vocab, self.max_len = num_samples, vocab_size, max_seq_len
        rng = np.random.RandomState(seed)
        if length_distribution == "fixed":
            self.lens = np.full(num_samples, max_seq_len + 1)
        elif length_distribution == "uniform":
            self.lens = rng.randint(min_seq_len or 16, max_seq_len + 2, num_samples)
        else:  # long-tailed (log-normal), typical of SFT corpora
            self.lens = np.clip(rng.lognormal(np.log(max_seq_len / 8), 1.0, num_samples).astype(int), min_seq_len or 16, max_seq_len + 1)
        self.seed = seed

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        return np.random.RandomState(self.seed * 1000003 + i).randint(0, self.vocab, self.lens[i]).astype(np.int64)


class IndexedTokenDataset:
    """Megatron-style flat token file (.bin, int32/uint16) + index of document offsets (.idx.npy)"""

    def __init__(self, prefix: str, dtype=np.int32):
        self.data = np.memmap(prefix + ".bin", dtype=dtype, mode="r")
        self.offsets = np.load(prefix + ".idx.npy")

    @staticmethod
    def build(prefix: str, docs: Sequence[Sequence[int]], dtype=np.int32):
        offs = np.zeros(len(docs) + 1, dtype=np.int64)
        for i, d in enumerate(docs):
            offs[i + 1] = offs[i] + len(d)
        flat = np.concatenate([np.asarray(d, dtype=dtype) for d in docs]) if docs else np.zeros(0, dtype)
        flat.tofile(prefix + ".bin")
        np.save(prefix + ".idx.npy", offs)

    def __len__(self):
        return len(self.offsets) - 1

    def __getitem__(self, i):
        return np.asarray(self.data[self.offsets[i]:self.offsets[i + 1]], dtype=np.int64)
