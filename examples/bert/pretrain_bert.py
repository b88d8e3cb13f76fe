"""BERT pre-training (masked-LM + next-sentence) on synthetic sentence pairs, data parallel over the ranks of the launch.

    python examples/bert/pretrain_bert.py --steps 30
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 examples/bert/pretrain_bert.py --size base --bf16

(ref: hetu/v1/examples/nlp/bert/train_hetu_bert.py, train_hetu_bert_dp.py, create_pretraining_data.py)"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import hetu_b200 as ht
from hetu_b200.models import BertConfig, BertForPreTraining

ap = argparse.ArgumentParser()
ap.add_argument("--size", default="tiny", choices=["tiny", "base", "large"])
ap.add_argument("--seq", type=int, default=32)
ap.add_argument("--batch", type=int, default=8, help="sequences per rank")
ap.add_argument("--steps", type=int, default=30)
ap.add_argument("--lr", type=float, default=1e-3)
ap.add_argument("--mask-prob", type=float, default=0.15)
ap.add_argument("--bf16", action="store_true")
a = ap.parse_args()
world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
ht.init_comm_group(world)
cfg = {"tiny": BertConfig(vocab_size=1000, hidden_size=64, num_hidden_layers=2, num_attention_heads=4, intermediate_size=128, max_position_embeddings=a.seq),
       "base": BertConfig.base(max_position_embeddings=max(a.seq, 128)), "large": BertConfig.large(max_position_embeddings=max(a.seq, 128))}[a.size]
B, S, V = a.batch, a.seq, cfg.vocab_size
CLS, SEP, MASK = 1, 2, 3


def make_batch(rng):
    """sentence pairs: the second half continues the first (label 0) or is random (label 1); 15 % of the tokens are masked
    (80 % [MASK], 10 % random, 10 % kept), their original ids are the MLM labels, everything else is ignored (-100)"""
    half = (S - 3) // 2
    ids = np.zeros((B, S), np.int64); tt = np.zeros((B, S), np.int64); nsp = rng.randint(0, 2, B)
    for b in range(B):
        first = rng.randint(10, V, half)
        second = (first + 1) % (V - 10) + 10 if nsp[b] == 0 else rng.randint(10, V, half)
        row = np.concatenate([[CLS], first, [SEP], second, [SEP]])
        ids[b, :len(row)] = row
        tt[b, half + 2:len(row)] = 1
    labels = np.full((B, S), -100, np.int64)
    pick = (rng.rand(B, S) < a.mask_prob) & (ids >= 10)
    labels[pick] = ids[pick]
    r = rng.rand(B, S)
    ids = np.where(pick & (r < 0.8), MASK, np.where(pick & (r >= 0.9), rng.randint(10, V, (B, S)), ids))
    return ids, tt, labels, nsp


import contextlib
with ht.graph("define_and_run", create_new=True) as g, (ht.autocast("bfloat16") if a.bf16 else contextlib.nullcontext()):
    model = BertForPreTraining(cfg)
    X, T = ht.placeholder("int64", [B, S], name="ids"), ht.placeholder("int64", [B, S], name="token_types")
    Y, N = ht.placeholder("int64", [B, S], name="mlm_labels"), ht.placeholder("int64", [B], name="nsp_labels")
    loss, mlm_logits, nsp_logits = model(X, T, masked_lm_labels=Y, next_sentence_label=N)
    train = ht.AdamOptimizer(lr=a.lr, weight_decay=0.01).minimize(loss)
rng = np.random.RandomState(rank)
for step in range(a.steps):
    ids, tt, labels, nsp = make_batch(rng)
    out = g.run(loss, [loss, nsp_logits, train], {X: torch.as_tensor(ids), T: torch.as_tensor(tt), Y: torch.as_tensor(labels), N: torch.as_tensor(nsp)},
                grad_scale=1.0 / world)
    if rank == 0 and (step % 10 == 0 or step == a.steps - 1):
        acc = float((out[1].float().argmax(1).cpu().numpy() == nsp).mean())
        print(f"step {step} loss {float(out[0]):.4f} nsp-acc {acc:.2f}", flush=True)
