"""Encoder-decoder Transformer (Vaswani et al.) for sequence-to-sequence tasks: sinusoidal positions, post-LayerNorm encoder and
decoder stacks, causal decoder self-attention, encoder-decoder cross attention, label-smoothed cross entropy, greedy decoding.
`model(src [B, S], tgt_in [B, T], tgt_out [B, T])` -> (loss, logits [B * T, V]); padding id 0 is masked out of attention and loss.
(ref: hetu/v1/examples/nlp/hetu_transformer.py, hparams.py)"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

from .. import ops
from ..core import from_numpy
from ..nn import Dropout, Embedding, LayerNorm, Linear, Module, ModuleList


@dataclass
class TransformerConfig:
    src_vocab_size: int = 32000
    tgt_vocab_size: int = 32000
    d_model: int = 512
    num_heads: int = 8
    d_ff: int = 2048
    num_encoder_layers: int = 6
    num_decoder_layers: int = 6
    max_len: int = 512
    dropout: float = 0.1
    label_smoothing: float = 0.1
    pad_id: int = 0
    share_embeddings: bool = False          # one table for source, target and the output projection (joint vocabulary)


def _sinusoid(max_len: int, d: int) -> np.ndarray:
    pos = np.arange(max_len)[:, None]
    i = np.arange(d)[None, :]
    angle = pos / np.power(10000.0, (2 * (i // 2)) / d)
    return np.where(i % 2 == 0, np.sin(angle), np.cos(angle)).astype(np.float32)


class MultiHeadAttention(Module):
    """projections + scaled dot-product attention between a query sequence [B, Sq, d] and a memory [B, Sk, d] (both flattened to
    [B * S, d]); `key_mask` [B, Sk] marks real tokens, `causal` hides the future.  Unmasked equal-length attention runs the fused
    flash kernels, everything else the composed path."""

    def __init__(self, d_model, num_heads, dropout, name):
        super().__init__()
        self.h, self.d = num_heads, d_model // num_heads
        self.q, self.k, self.v, self.o = (Linear(d_model, d_model, name=f"{name}_{n}") for n in "qkvo")
        self.p = dropout

    def forward(self, x, mem, b, sq, sk, key_mask=None, causal=False):
        h, d = self.h, self.d
        q = ops.reshape(self.q(x), [b, sq, h, d])
        k = ops.reshape(self.k(mem), [b, sk, h, d])
        v = ops.reshape(self.v(mem), [b, sk, h, d])
        p = float(self.p) if self.training else 0.0
        if key_mask is None and sq == sk:
            a = ops.attn(q, k, v, p_dropout=p, is_causal=causal)
        else:
            qt = ops.reshape(ops.transpose(q, [0, 2, 1, 3]), [b * h, sq, d])
            kt = ops.reshape(ops.transpose(k, [0, 2, 1, 3]), [b * h, sk, d])
            vt = ops.reshape(ops.transpose(v, [0, 2, 1, 3]), [b * h, sk, d])
            scores = ops.reshape(ops.bmm(qt, ops.transpose(kt, [0, 2, 1])) * (1.0 / math.sqrt(d)), [b, h, sq, sk])
            if key_mask is not None:
                scores = scores + (ops.reshape(key_mask, [b, 1, 1, sk]) - 1.0) * 1e30
            if causal:
                future = from_numpy(np.triu(np.ones((sq, sk), np.float32), 1 + (sk - sq)).reshape(1, 1, sq, sk) * -1e30)
                scores = scores + future
            probs = ops.softmax(ops.reshape(scores, [b * h, sq, sk]), -1)
            if p > 0:
                probs = ops.dropout(probs, p)
            a = ops.transpose(ops.reshape(ops.bmm(probs, vt), [b, h, sq, d]), [0, 2, 1, 3])
        return self.o(ops.reshape(a, [b * sq, h * d]))


class EncoderLayer(Module):
    def __init__(self, c: TransformerConfig, i: int):
        super().__init__()
        self.attn = MultiHeadAttention(c.d_model, c.num_heads, c.dropout, f"tf_enc{i}_attn")
        self.ln1, self.ln2 = LayerNorm(c.d_model, name=f"tf_enc{i}_ln1"), LayerNorm(c.d_model, name=f"tf_enc{i}_ln2")
        self.f1, self.f2 = Linear(c.d_model, c.d_ff, name=f"tf_enc{i}_f1"), Linear(c.d_ff, c.d_model, name=f"tf_enc{i}_f2")
        self.drop = Dropout(c.dropout)

    def forward(self, x, b, s, mask):
        x = self.ln1(x + self.drop(self.attn(x, x, b, s, s, key_mask=mask)))
        return self.ln2(x + self.drop(self.f2(self.f1(x, act="relu"))))


class DecoderLayer(Module):
    def __init__(self, c: TransformerConfig, i: int):
        super().__init__()
        self.self_attn = MultiHeadAttention(c.d_model, c.num_heads, c.dropout, f"tf_dec{i}_self")
        self.cross_attn = MultiHeadAttention(c.d_model, c.num_heads, c.dropout, f"tf_dec{i}_cross")
        self.ln1, self.ln2, self.ln3 = (LayerNorm(c.d_model, name=f"tf_dec{i}_ln{k}") for k in (1, 2, 3))
        self.f1, self.f2 = Linear(c.d_model, c.d_ff, name=f"tf_dec{i}_f1"), Linear(c.d_ff, c.d_model, name=f"tf_dec{i}_f2")
        self.drop = Dropout(c.dropout)

    def forward(self, y, memory, b, t, s, tgt_mask, src_mask):
        y = self.ln1(y + self.drop(self.self_attn(y, y, b, t, t, key_mask=tgt_mask, causal=True)))
        y = self.ln2(y + self.drop(self.cross_attn(y, memory, b, t, s, key_mask=src_mask)))
        return self.ln3(y + self.drop(self.f2(self.f1(y, act="relu"))))


class Transformer(Module):
    def __init__(self, config: TransformerConfig):
        super().__init__()
        c = self.config = config
        self.src_embed = Embedding(c.src_vocab_size, c.d_model, name="tf_src_embed")
        self.tgt_embed = self.src_embed if c.share_embeddings else Embedding(c.tgt_vocab_size, c.d_model, name="tf_tgt_embed")
        self.encoder = ModuleList([EncoderLayer(c, i) for i in range(c.num_encoder_layers)])
        self.decoder = ModuleList([DecoderLayer(c, i) for i in range(c.num_decoder_layers)])
        self.out_proj = None if c.share_embeddings else Linear(c.d_model, c.tgt_vocab_size, bias=False, name="tf_out_proj")
        self.drop = Dropout(c.dropout)
        self._pe = _sinusoid(c.max_len, c.d_model)

    def _embed(self, table, ids, b, s):
        pe = from_numpy(np.tile(self._pe[:s], (b, 1)))
        return self.drop(table(ops.reshape(ids, [b * s])) * math.sqrt(self.config.d_model) + pe)

    def encode(self, src, src_mask=None):
        b, s = src.shape
        x = self._embed(self.src_embed, src, b, s)
        for layer in self.encoder:
            x = layer(x, b, s, src_mask)
        return x

    def decode(self, tgt_in, memory, src_len, src_mask=None, tgt_mask=None):
        b, t = tgt_in.shape
        y = self._embed(self.tgt_embed, tgt_in, b, t)
        for layer in self.decoder:
            y = layer(y, memory, b, t, src_len, tgt_mask, src_mask)
        if self.out_proj is None:
            return ops.linear(y, self.tgt_embed.weight, None, trans_b=True)
        return self.out_proj(y)

    def forward(self, src, tgt_in, tgt_out=None, src_mask=None, tgt_mask=None):
        logits = self.decode(tgt_in, self.encode(src, src_mask), src.shape[1], src_mask, tgt_mask)
        if tgt_out is None:
            return logits
        b, t = tgt_in.shape
        labels = ops.reshape(tgt_out, [b * t])
        eps, v = float(self.config.label_smoothing), self.config.tgt_vocab_size
        nll = ops.softmax_cross_entropy_sparse(logits, labels, ignored_index=self.config.pad_id, reduction="mean")
        if eps <= 0:
            return nll, logits
        # label smoothing: (1 - eps) * NLL + eps * mean over the vocabulary of -log p, on non-pad positions
        logp = ops.log_softmax(logits, -1)
        keep = ops.reshape(ops.not_equal(labels, self.config.pad_id), [b * t, 1])          # 1 on real tokens, 0 on padding
        uniform = ops.sum(ops.mean(logp, [1], True) * keep * -1.0) / ops.clamp(ops.sum(keep), 1.0, 1e30)
        return nll * (1.0 - eps) + uniform * eps, logits

    def greedy_decode(self, graph, src_np: np.ndarray, max_len: int, bos_id: int = 1, eos_id: int = 2):
        """host loop over target positions with a fixed-shape decoder input (padding beyond the current position is masked by
        causality) -> [B, <= max_len] token ids"""
        import torch
        from ..core import placeholder
        b, s = src_np.shape
        was_training = self.training
        self.eval()
        S = placeholder("int64", [b, s], name="tf_greedy_src")
        T = placeholder("int64", [b, max_len], name="tf_greedy_tgt")
        logits = self.forward(S, T)
        out = np.full((b, max_len), self.config.pad_id, np.int64)
        out[:, 0] = bos_id
        done = np.zeros(b, bool)
        for t in range(1, max_len):
            lg = graph.run(logits, [logits], {S: torch.as_tensor(src_np), T: torch.as_tensor(out)})[0].float().cpu().numpy().reshape(b, max_len, -1)
            nxt = lg[:, t - 1].argmax(-1)
            out[:, t] = np.where(done, self.config.pad_id, nxt)
            done |= nxt == eos_id
            if done.all():
                break
        self.train(was_training)
        return out
