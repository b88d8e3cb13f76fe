"""Per-op-type device time of one GPT-2 1.3B training step (executor profiler: synchronises after every op)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hetu_b200 as ht
from hetu_b200.models import GPTConfig, GPTLMHeadModel, generate_ds_parallel_config
torch.cuda.set_device(0)
cfg = GPTConfig.gpt2_1p3b()
B, S = int(os.environ.get("B", "16")), 1024
T = B * S
with ht.graph("define_and_run", create_new=True) as g, ht.autocast("bfloat16"):
    model = GPTLMHeadModel(cfg, [generate_ds_parallel_config(cfg.n_layer, 1, 1, 1, 1, zero=True)])
    ids, pos, lab = (ht.placeholder("int64", [T], name=n) for n in ("ids", "pos", "lab"))
    loss = model(ids, pos, lab, seq_len=S)
    train_op = ht.AdamOptimizer(lr=1e-4).minimize(loss)
x = torch.randint(0, cfg.vocab_size, (T,), device="cuda"); p = torch.arange(S, device="cuda").repeat(B)
feed = {ids: x, pos: p, lab: torch.roll(x, -1)}
for _ in range(3):
    g.run(loss, [loss, train_op], feed)
torch.cuda.synchronize()
import time
t0 = time.perf_counter(); 
for _ in range(3): g.run(loss, [loss, train_op], feed)
torch.cuda.synchronize(); print("step ms (async)", (time.perf_counter() - t0) / 3 * 1e3)
g.set_profile(True)
g.run(loss, [loss, train_op], feed)
g.set_profile(False)
agg = {}
for name, ms in g.op_times():
    agg.setdefault(name.split(":")[0], []).append(ms)
summ = {"by_optype": sorted(((k, sum(v), len(v)) for k, v in agg.items()), key=lambda r: -r[1]), "breakdown": g.step_breakdown()}
tot = sum(r[1] for r in summ["by_optype"])
print("sum of per-op (synchronised) ms", tot)
for name, ms, n in summ["by_optype"][:30]:
    print(f"{name:32s} {ms:9.3f} ms  x{n}")
print(json.dumps(summ["breakdown"]))
# the same, grouped by op name with the layer index masked (which GEMM of the block is slow?)
import re
byname = {}
for name, ms in g.op_times():
    byname.setdefault(re.sub(r"\d+", "#", name), []).append(ms)
print("---- by masked op name (mean ms per instance)")
for k, v in sorted(byname.items(), key=lambda kv: -sum(kv[1]))[:40]:
    print(f"{k:72s} {sum(v):9.3f} ms  x{len(v):3d}  mean {sum(v) / len(v) * 1e3:8.1f} us")
