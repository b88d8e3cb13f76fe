"""Chat / prompt templates for SFT (ref: python/hetu/data/messages/*): turns messages into token ids + a loss mask
that trains only on assistant tokens."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Sequence, Tuple

import numpy as np


@dataclass
class PromptTemplate:
    template: str = "{instruction}\n{input}\n"

    def render(self, **fields) -> str:
        return self.template.format(**{k: fields.get(k, "") for k in ("instruction", "input", "output", "system")})


@dataclass
class ChatTemplate:
    system_prefix: str = "<|system|>\n"
    user_prefix: str = "<|user|>\n"
    assistant_prefix: str = "<|assistant|>\n"
    turn_suffix: str = "\n"

    def render(self, messages: Sequence[Dict[str, str]]) -> List[Tuple[str, bool]]:
        """-> [(text, is_assistant)]"""
        out = []
        for m in messages:
            role = m["role"]
            prefix = {"system": self.system_prefix, "user": self.user_prefix, "assistant": self.assistant_prefix}[role]
            out.append((prefix, False))
            out.append((m["content"] + self.turn_suffix, role == "assistant"))
        return out


def build_chat_sample(messages, tokenizer, template: ChatTemplate = None, max_len: int = 2048, ignore_index: int = -1):
    """-> (input_ids, labels) with labels = ignore_index outside assistant spans (next-token shifted by the trainer)"""
    template = template or ChatTemplate()
    ids, mask = [], []
    for text, is_asst in template.render(messages):
        t = tokenizer.encode(text, add_special_tokens=False)
        ids += t
        mask += [is_asst] * len(t)
    ids, mask = ids[:max_len], mask[:max_len]
    ids = np.asarray(ids, dtype=np.int64)
    labels = np.where(np.asarray(mask), ids, ignore_index)
    return ids, labels
