"""Kernel-time accounting of the flagship training step (GPT-2 1.3B, 1 GPU) with the CUPTI activity tracer of torch.profiler:
per-kernel device time, their sum, and the wall time of the same steps -> how much of the step is GPU idle (launch gaps).
Not a bench number (tracing perturbs the run); it apportions the step."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hetu_b200 as ht
from hetu_b200.models import GPTConfig, GPTLMHeadModel, generate_ds_parallel_config

os.environ.setdefault("HETU_B200_STRICT", "1")
ht.init_comm_group(1)
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
cfg = GPTConfig.gpt2_1p3b()
S, B = 1024, 16
T = B * S
with ht.graph("define_and_run", create_new=True) as g, ht.autocast("bfloat16"):
    dsc = [generate_ds_parallel_config(cfg.n_layer, 1, 1, 1, 1, zero=True)]
    model = GPTLMHeadModel(cfg, dsc)
    ic = ht.nn.parallel.config2ds(dsc[0]["input"])
    ids = ht.parallel_placeholder("int64", [T], [ic[0]], device_group_hierarchy=[ic[1]], name="input_ids")
    pos = ht.parallel_placeholder("int64", [T], [ic[0]], device_group_hierarchy=[ic[1]], name="position_ids")
    lab = ht.parallel_placeholder("int64", [T], [ic[0]], device_group_hierarchy=[ic[1]], name="labels")
    loss = model(ids, pos, lab, seq_len=S)
    train_op = ht.AdamOptimizer(lr=1e-4, beta1=0.9, beta2=0.95, weight_decay=0.1).minimize(loss)
x = torch.randint(0, cfg.vocab_size, (T,), device=dev)
p = torch.arange(S, device=dev).repeat(B)
y = torch.roll(x, -1)
step = lambda: g.run(loss, [loss, train_op], {ids: x, pos: p, lab: y})[0]   # noqa: E731
for _ in range(4):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
N = 3
e0.record()
for _ in range(N):
    step()
e1.record()
torch.cuda.synchronize()
plain_ms = e0.elapsed_time(e1) / N
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    e0.record()
    for _ in range(N):
        step()
    e1.record()
    torch.cuda.synchronize()
traced_ms = e0.elapsed_time(e1) / N
rows = [(e.key, e.count, e.device_time_total) for e in prof.key_averages() if e.device_time_total > 0 and e.device_type == torch.autograd.DeviceType.CUDA]
rows.sort(key=lambda r: -r[2])
tot = sum(r[2] for r in rows)
print(f"step {plain_ms:.2f} ms untraced, {traced_ms:.2f} ms traced; kernel time sum {tot / N / 1e3:.2f} ms/step "
      f"-> idle {traced_ms - tot / N / 1e3:.2f} ms/step")
print(f"{'kernel':100s} {'n/step':>7s} {'ms/step':>9s} {'share':>6s}")
for k, c, t in rows[:45]:
    print(f"{k[:100]:100s} {c / N:7.1f} {t / N / 1e3:9.3f} {100 * t / tot:5.1f}%")
