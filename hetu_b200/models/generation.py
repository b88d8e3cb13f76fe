"""Text generation for the causal LM heads (GPT / Llama): greedy, temperature, top-k and nucleus sampling, stop tokens, batched
prompts of different lengths.  The decoder runs one fixed-shape forward graph per step over the whole window (causal attention makes
the padded tail irrelevant), so the same code path works under any parallel strategy the model was built with; there is no KV cache.
(the reference has no generation utility: SFT / chat examples need one to show a tuned model's output)"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import numpy as np
import torch


def _filter_logits(logits: np.ndarray, top_k: int, top_p: float) -> np.ndarray:
    """[B, V] -> logits with everything outside the top-k / nucleus set at -inf"""
    out = logits.copy()
    if top_k and top_k > 0:
        k = min(int(top_k), out.shape[-1])
        kth = np.partition(out, -k, axis=-1)[:, -k][:, None]
        out[out < kth] = -np.inf
    if top_p and 0.0 < top_p < 1.0:
        order = np.argsort(-out, axis=-1)
        sorted_logits = np.take_along_axis(out, order, -1)
        probs = np.exp(sorted_logits - sorted_logits[:, :1])
        probs /= probs.sum(-1, keepdims=True)
        cum = np.cumsum(probs, -1)
        drop = cum - probs > top_p                               # keep the smallest prefix whose mass reaches top_p
        sorted_logits[drop] = -np.inf
        np.put_along_axis(out, order, sorted_logits, -1)
    return out


class Generator:
    """builds the forward graph once (`window` tokens per sequence, `batch` sequences) and decodes prompts with it

        gen = Generator(lambda: GPTLMHeadModel(cfg, ds_cfg), batch=2, window=64)      # inside no graph context: it opens its own
        gen.load_state_dict(trained_model.state_dict())
        out = gen.generate([[1, 5, 9], [7]], max_new_tokens=20, temperature=0.8, top_k=40)
    """

    def __init__(self, model_factory: Callable, batch: int, window: int, pad_id: int = 0, autocast: Optional[str] = None):
        import contextlib
        from .. import core
        self.batch, self.window, self.pad_id = int(batch), int(window), int(pad_id)
        ctx = core.autocast(autocast) if autocast else contextlib.nullcontext()
        self._gctx = core.graph("define_and_run", create_new=True, prefix="generate")
        self.graph = self._gctx.__enter__()
        try:
            with ctx:
                self.model = model_factory()
                n = self.batch * self.window
                self.ids = core.placeholder("int64", [n], name="gen_ids")
                self.pos = core.placeholder("int64", [n], name="gen_pos")
                self.logits = self.model(self.ids, self.pos, None, seq_len=self.window)
        finally:
            self._gctx.__exit__(None, None, None)
        if hasattr(self.model, "eval"):
            self.model.eval()
        self._positions = torch.arange(self.window).repeat(self.batch)

    def load_state_dict(self, state, strict: bool = False):
        return self.model.load_state_dict(state, strict=strict)

    def step_logits(self, tokens: np.ndarray) -> np.ndarray:
        """tokens [batch, window] -> logits [batch, window, vocab]"""
        out = self.graph.run(self.logits, [self.logits], {self.ids: torch.as_tensor(tokens.reshape(-1)), self.pos: self._positions})[0]
        return out.float().cpu().numpy().reshape(self.batch, self.window, -1)

    def generate(self, prompts: Sequence[Sequence[int]], max_new_tokens: int = 32, temperature: float = 0.0, top_k: int = 0, top_p: float = 1.0,
                 eos_id: Optional[int] = None, seed: int = 0) -> List[List[int]]:
        """-> prompt + continuation per sequence (stops at `eos_id`, at `max_new_tokens` or when the window is full)"""
        assert 0 < len(prompts) <= self.batch, f"{len(prompts)} prompts for a batch of {self.batch}"
        rng = np.random.RandomState(seed)
        tokens = np.full((self.batch, self.window), self.pad_id, np.int64)
        lens = np.ones(self.batch, np.int64)
        for i, p in enumerate(prompts):
            p = list(p)[-self.window + 1:] or [self.pad_id]
            tokens[i, :len(p)] = p
            lens[i] = len(p)
        done = np.array([i >= len(prompts) for i in range(self.batch)])
        start = lens.copy()
        for _ in range(int(max_new_tokens)):
            if done.all() or (lens[~done] >= self.window).all():
                break
            logits = self.step_logits(tokens)
            last = logits[np.arange(self.batch), lens - 1]                     # next-token distribution of every sequence
            if temperature and temperature > 0:
                z = _filter_logits(last / float(temperature), top_k, top_p)
                z = z - z.max(-1, keepdims=True)
                p = np.exp(z)
                p /= p.sum(-1, keepdims=True)
                nxt = np.array([rng.choice(p.shape[-1], p=p[i]) for i in range(self.batch)])
            else:
                nxt = last.argmax(-1)
            for i in range(self.batch):
                if done[i] or lens[i] >= self.window:
                    done[i] = True
                    continue
                tokens[i, lens[i]] = nxt[i]
                lens[i] += 1
                if eos_id is not None and int(nxt[i]) == int(eos_id):
                    done[i] = True
        return [tokens[i, :lens[i]].tolist() for i in range(len(prompts))]
