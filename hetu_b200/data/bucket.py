"""Length bucketing, padding and sequence packing (varlen attention with cu_seqlens), including context-parallel aware
packing (ref: python/hetu/data/bucket.py:86-270, generate_cp_pack_data :193)."""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np


def get_sorted_batch_and_len(batch: Sequence[np.ndarray], pad_id: int):
    lens = np.array([len(s) for s in batch])
    order = np.argsort(-lens, kind="stable")
    return [batch[i] for i in order], lens[order]


def pad_sequences(seqs: Sequence[np.ndarray], max_len: int, pad_id: int, alignment: int = 1) -> np.ndarray:
    width = min(max_len, max(len(s) for s in seqs))
    width = (width + alignment - 1) // alignment * alignment
    out = np.full((len(seqs), width), pad_id, dtype=np.int64)
    for i, s in enumerate(seqs):
        n = min(len(s), width)
        out[i, :n] = s[:n]
    return out


def pack_sequences(seqs: Sequence[np.ndarray], max_len: int, pad_id: int, alignment: int = 128) -> List[Tuple[np.ndarray, np.ndarray]]:
    """First-fit-decreasing bin packing into rows of at most max_len tokens.
    Returns [(tokens [<= max_len], cu_seqlens [n+1])] -- every sequence is padded to `alignment` so that attention tiles
    never straddle two documents."""
    def al(n):
        return (n + alignment - 1) // alignment * alignment
    items = sorted(((min(len(s), max_len), i) for i, s in enumerate(seqs)), reverse=True)
    bins: List[List[int]] = []
    space: List[int] = []
    for n, i in items:
        need = min(al(n), max_len)
        for b in range(len(bins)):
            if space[b] >= need:
                bins[b].append(i)
                space[b] -= need
                break
        else:
            bins.append([i])
            space.append(max_len - need)
    out = []
    for b in bins:
        toks, cu = [], [0]
        for i in b:
            s = seqs[i][:max_len]
            pad = min(al(len(s)), max_len) - len(s)
            toks.append(np.concatenate([s, np.full(pad, pad_id, dtype=np.int64)]))
            cu.append(cu[-1] + len(s) + pad)
        out.append((np.concatenate(toks), np.asarray(cu, dtype=np.int32)))
    return out


def generate_cp_pack_data(tokens: np.ndarray, cu_seqlens: np.ndarray, cp: int, pattern: str = "SYM") -> List[Tuple[np.ndarray, np.ndarray]]:
    """Split one packed row over a context-parallel ring.  SYM: every document is cut into 2*cp chunks and rank i takes
    chunks i and 2cp-1-i (balanced causal work); NORMAL: contiguous cp chunks.  Returns per-rank (tokens, cu_seqlens)."""
    per_rank_tok: List[List[np.ndarray]] = [[] for _ in range(cp)]
    per_rank_cu: List[List[int]] = [[0] for _ in range(cp)]
    for a, b in zip(cu_seqlens[:-1], cu_seqlens[1:]):
        doc = tokens[a:b]
        parts = 2 * cp if pattern == "SYM" else cp
        assert len(doc) % parts == 0, "documents must be padded to a multiple of the CP chunk count"
        chunks = np.split(doc, parts)
        for r in range(cp):
            mine = [chunks[r], chunks[2 * cp - 1 - r]] if pattern == "SYM" else [chunks[r]]
            piece = np.concatenate(mine)
            per_rank_tok[r].append(piece)
            per_rank_cu[r].append(per_rank_cu[r][-1] + len(piece))
    return [(np.concatenate(t), np.asarray(c, dtype=np.int32)) for t, c in zip(per_rank_tok, per_rank_cu)]


class Bucket:
    """Collects a global batch, then emits padded or packed micro-batches (ref: data/bucket.py Bucket)."""

    def __init__(self, pad_id: int, max_seq_len: int, alignment: int = 128):
        self.pad_id, self.max_seq_len, self.alignment = pad_id, max_seq_len, alignment
        self._batch: List[np.ndarray] = []
        self._packed = None
        self._padded = None

    def add_data(self, seq: np.ndarray, valid_len: Optional[int] = None):
        self._batch.append(np.asarray(seq[: (valid_len or len(seq))], dtype=np.int64))

    def __len__(self):
        return len(self._batch)

    def pad_data(self):
        self._padded = pad_sequences(self._batch, self.max_seq_len, self.pad_id, self.alignment)
        return self._padded

    def pack_data(self):
        self._packed = pack_sequences(self._batch, self.max_seq_len, self.pad_id, self.alignment)
        return self._packed

    def packed_batch_size(self):
        return len(self._packed or [])

    def packed_cu_seqlens_list(self):
        return [cu for _, cu in (self._packed or [])]

    def packed_batch(self):
        return [t for t, _ in (self._packed or [])]

    def padding_stats(self) -> Dict[str, float]:
        real = sum(len(s) for s in self._batch)
        padded = self._padded.size if self._padded is not None else 0
        packed = sum(len(t) for t, _ in (self._packed or []))
        return {"real_tokens": real, "padded_tokens": padded, "packed_tokens": packed,
                "pad_efficiency": real / padded if padded else 0.0, "pack_efficiency": real / packed if packed else 0.0}

    @staticmethod
    def by_length(seqs: Sequence[np.ndarray], bucket_sizes: Sequence[int]) -> Dict[int, List[np.ndarray]]:
        """HotSPa-style sequence-length buckets: sizes sorted descending, 0 = catch-all"""
        sizes = sorted(bucket_sizes, reverse=True)
        out: Dict[int, List[np.ndarray]] = {s: [] for s in sizes}
        for s in seqs:
            for lo in sizes:
                if len(s) > lo or lo == sizes[-1]:
                    out[lo].append(s)
                    break
        return out
