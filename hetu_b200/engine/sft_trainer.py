"""Supervised fine-tuning on chat / instruction data, optionally with LoRA adapters
(ref: python/hetu/engine/sft_trainer.py:13-114, sft_config.py)."""
from __future__ import annotations

import json

import numpy as np

from ..data.messages import ChatTemplate, PromptTemplate, build_chat_sample
from .trainer import Trainer
from .trainer_config import SFTConfig


class _SFTDataset:
    def __init__(self, records, tokenizer, cfg: SFTConfig):
        self.records, self.tok, self.cfg = records, tokenizer, cfg
        self.template = ChatTemplate()
        self.prompt = PromptTemplate(cfg.prompt_template) if cfg.prompt_template else PromptTemplate()

    def __len__(self):
        return len(self.records)

    def __getitem__(self, i):
        r = self.records[i]
        max_len = int(self.cfg.max_seq_length or 1024) + 1
        if "messages" in r:
            msgs = r["messages"]
        elif "instruction" in r or "output" in r:   # alpaca
            msgs = [{"role": "user", "content": self.prompt.render(**r)}, {"role": "assistant", "content": r.get("output", "")}]
        else:
            text = r[self.cfg.dataset_text_field or "text"]
            ids = np.asarray(self.tok.encode(text)[:max_len], np.int64)
            return ids, ids
        ids, labels = build_chat_sample(msgs, self.tok, self.template, max_len)
        if self.cfg.train_on_prompt:
            labels = ids.copy()
        return ids, labels


class SFTTrainer(Trainer):
    def __init__(self, sft_config: SFTConfig, model, tokenizer, optimizer=None, train_dataset=None, peft_config=None, **kwargs):
        super().__init__(sft_config, model, tokenizer, optimizer, train_dataset, **kwargs)
        self.peft_config = peft_config
        if self.train_dataset is None and sft_config.train_dataset_path:
            self.train_dataset = self._prepare_dataset(sft_config.train_dataset_path)
        elif isinstance(self.train_dataset, list):
            self.train_dataset = _SFTDataset(self.train_dataset, self.tokenizer, sft_config)

    def _prepare_dataset(self, path: str):
        with open(path) as f:
            head = f.read(1)
            f.seek(0)
            recs = json.load(f) if head == "[" else [json.loads(l) for l in f if l.strip()]
        return _SFTDataset(recs, self.tokenizer, self.pretrain_config)

    def create_define_graph(self):
        cfg: SFTConfig = self.pretrain_config
        if cfg.lora_rank > 0 or self.peft_config is not None:
            from ..peft import LoraConfig, wrap_model_factory
            pc = self.peft_config or LoraConfig(rank=cfg.lora_rank, lora_alpha=cfg.lora_alpha, lora_dropout=cfg.lora_dropout,
                                                target_modules=list(cfg.lora_target_modules))
            self.model_wrapper = wrap_model_factory(self.model_wrapper, pc)
        return super().create_define_graph()
