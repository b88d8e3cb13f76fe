"""LobRA end to end on one device: two fine-tuning tasks share a frozen GPT; each step the batch scheduler builds padded
micro-batches whose rows are grouped by task, the task ranges become the [tokens, tasks] routing mask of the multi-LoRA layers.

    python examples/lobra/train_multi_lora.py

(ref: examples/lobra/scripts/llama_lora_multi_task.py, trainer/batch_scheduler.py)"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import hetu_b200 as ht
from hetu_b200.engine import lobra as L
from hetu_b200.models import GPTConfig, GPTLMHeadModel, generate_ds_parallel_config
from hetu_b200.peft import get_peft_model
from hetu_b200.peft.lora.config import LoraConfig

ht.init_comm_group(1)
ht.set_seed(0)
VOCAB, MAX_TOKENS, TASKS, ROWS = 96, 256, 2, 4
rng = np.random.RandomState(0)


def sample(task, n):
    """task 0: counting sequences (short); task 1: alternating pairs (longer)"""
    out = []
    for _ in range(n):
        ln = rng.randint(8, 17) if task == 0 else rng.randint(24, 49)
        start = rng.randint(0, VOCAB)
        seq = (start + np.arange(ln)) % VOCAB if task == 0 else np.where(np.arange(ln) % 2 == 0, start, (start + 7) % VOCAB)
        out.append(list(seq))
    return out


buckets = [16, 32, 64]
POOL = {t: sample(t, 6) for t in range(TASKS)}      # a small fixed fine-tuning set per tenant
cfg = GPTConfig(vocab_size=VOCAB, n_positions=64, n_embd=64, n_layer=2, n_head=4)
seq_sym = ht.IntSymbol(64)
with ht.graph("define_and_run", create_new=True) as g:
    model = get_peft_model(GPTLMHeadModel(cfg, [generate_ds_parallel_config(2, 1, 1, 1, 1, zero=False)]), LoraConfig(rank=8, num_tasks=TASKS, init_std=0.05))
    ids, pos, lab = (ht.placeholder("int64", [ROWS * 64], name=n) for n in ("ids", "pos", "lab"))
    mask = ht.placeholder("float32", [ROWS * 64, TASKS], name="task_mask")
    model.set_task_mask(mask)
    loss = model(ids, pos, lab, seq_len=seq_sym)
    train = ht.AdamOptimizer(lr=3e-2).minimize(loss)

for step in range(60):
    rows = [(t, s, L.bucket_of(len(s), buckets)) for t in range(TASKS) for s in POOL[t]]
    step_loss, n_mb = 0.0, 0
    for mb in L.greedy_local_batch_scheduler(rows, MAX_TOKENS, TASKS, pad_id=0):
        b, w = mb.batch_size, mb.seq_length
        x = mb.batch_data
        y = np.roll(x, -1, axis=1)
        y[x == 0] = -1                                           # padding positions carry no loss
        y[:, -1] = -1
        m = np.zeros((b, w, TASKS), np.float32)
        for t in mb.task_id():                                   # rows [offset, offset + size) of the micro-batch belong to task t
            m[mb.batch_offset_list[t]:mb.batch_offset_list[t] + mb.batch_size_list[t], :, t] = 1.0
        feed = {ids: torch.as_tensor(x.reshape(-1)), pos: torch.as_tensor(np.tile(np.arange(w), b)), lab: torch.as_tensor(y.reshape(-1)),
                mask: torch.as_tensor(m.reshape(b * w, TASKS))}
        step_loss += float(g.run(loss, [loss, train], feed, int_symbol_dict={seq_sym: w})[0])
        n_mb += 1
    if step % 10 == 0 or step == 59:
        print(f"step {step:2d}: {n_mb} micro-batches, mean loss {step_loss / n_mb:.4f}")
