"""Turns a list of token sequences into next-token (input, label) pairs (ref: python/hetu/data/data_collator.py)."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


class DataCollatorForLanguageModel:
    def __init__(self, tokenizer=None, ignore_index: int = -1):
        self.tokenizer, self.ignore_index = tokenizer, ignore_index

    def __call__(self, samples: Sequence) -> List[Tuple[np.ndarray, np.ndarray]]:
        """each sample is tokens [n+1] or (tokens, labels); returns [(inputs [n], labels [n])]"""
        out = []
        for s in samples:
            if isinstance(s, tuple):
                ids, lab = np.asarray(s[0]), np.asarray(s[1])
                out.append((ids[:-1], lab[1:]))
            else:
                ids = np.asarray(s)
                out.append((ids[:-1], ids[1:]))
        return out
