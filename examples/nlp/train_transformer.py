"""Encoder-decoder Transformer on a synthetic translation task (target = reversed source), with padding masks, label smoothing and
greedy decoding at the end.

    python examples/nlp/train_transformer.py --steps 300

(ref: hetu/v1/examples/nlp/hetu_transformer.py, train_hetu_transformer.py)"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import hetu_b200 as ht
from hetu_b200.models import Transformer, TransformerConfig

ap = argparse.ArgumentParser()
ap.add_argument("--vocab", type=int, default=32); ap.add_argument("--seq", type=int, default=10); ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--d-model", type=int, default=64); ap.add_argument("--layers", type=int, default=2); ap.add_argument("--heads", type=int, default=4)
ap.add_argument("--steps", type=int, default=300); ap.add_argument("--lr", type=float, default=2e-3)
a = ap.parse_args()
V, S, B, BOS, EOS, PAD = a.vocab, a.seq, a.batch, 1, 2, 0
cfg = TransformerConfig(src_vocab_size=V, tgt_vocab_size=V, d_model=a.d_model, num_heads=a.heads, d_ff=4 * a.d_model, num_encoder_layers=a.layers,
                        num_decoder_layers=a.layers, max_len=S + 2, dropout=0.1, label_smoothing=0.1)
rng = np.random.RandomState(0)


def batch():
    lens = rng.randint(3, S + 1, B)
    src, tin, tout = np.zeros((B, S), np.int64), np.zeros((B, S + 1), np.int64), np.zeros((B, S + 1), np.int64)
    for i, n in enumerate(lens):
        seq = rng.randint(3, V, n)
        src[i, :n], tin[i, :n + 1], tout[i, :n + 1] = seq, np.concatenate([[BOS], seq[::-1]]), np.concatenate([seq[::-1], [EOS]])
    return src, tin, tout


with ht.graph("define_and_run", create_new=True) as g:
    model = Transformer(cfg)
    SRC, TIN, TOUT = (ht.placeholder("int64", shp, name=n) for n, shp in (("src", [B, S]), ("tgt_in", [B, S + 1]), ("tgt_out", [B, S + 1])))
    SM, TM = ht.placeholder("float32", [B, S], name="src_mask"), ht.placeholder("float32", [B, S + 1], name="tgt_mask")
    loss, logits = model(SRC, TIN, TOUT, src_mask=SM, tgt_mask=TM)
    train = ht.AdamOptimizer(lr=a.lr).minimize(loss)
    for step in range(a.steps):
        src, tin, tout = batch()
        out = g.run(loss, [loss, logits, train], {SRC: torch.as_tensor(src), TIN: torch.as_tensor(tin), TOUT: torch.as_tensor(tout),
                                                  SM: torch.as_tensor((src != PAD).astype(np.float32)), TM: torch.as_tensor((tin != PAD).astype(np.float32))})
        if step % 50 == 0 or step == a.steps - 1:
            pred = out[1].float().cpu().numpy().reshape(B, S + 1, V).argmax(-1)
            real = tout != PAD
            print(f"step {step} loss {float(out[0]):.4f} token-acc {float((pred[real] == tout[real]).mean()):.3f}", flush=True)
    src = rng.randint(3, V, (4, S))
    dec = model.greedy_decode(g, src, max_len=S + 2, bos_id=BOS, eos_id=EOS)
    for s_, d_ in zip(src, dec):
        print("source", s_.tolist(), "-> decoded", d_[1:S + 1].tolist())
