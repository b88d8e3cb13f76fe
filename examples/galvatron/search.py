"""Galvatron: search the best layer-wise hybrid-parallel plan for a model under a memory budget."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from hetu_b200.planner import GalvatronSearchEngine, HardwareProfile, LayerProfile, galvatron_plan_to_ds_parallel_config

ap = argparse.ArgumentParser()
ap.add_argument("--gpus", type=int, default=8); ap.add_argument("--mem-gb", type=float, default=180)
ap.add_argument("--layers", type=int, default=32); ap.add_argument("--hidden", type=int, default=4096); ap.add_argument("--ffn", type=int, default=11008)
ap.add_argument("--heads", type=int, default=32); ap.add_argument("--seq", type=int, default=4096); ap.add_argument("--vocab", type=int, default=32000)
ap.add_argument("--out", default="galvatron_plan.json")
a = ap.parse_args()
layer = LayerProfile.transformer(a.hidden, a.ffn, a.seq, a.heads, swiglu=True)
eng = GalvatronSearchEngine(a.layers, a.gpus, layer, HardwareProfile(), vocab=a.vocab, hidden=a.hidden, seq=a.seq, memory_mb=a.mem_gb * 1024)
plan = eng.search(batch_sizes=(8, 16, 32, 64, 128))
if plan is None:
    raise SystemExit("no feasible plan under this memory budget")
print(json.dumps({k: v for k, v in plan.items() if k != "strategies"}, indent=1))
print("per-layer:", [(s.tp, s.dp, s.sdp, int(s.ckpt)) for s in plan["strategies"]][:8], "...")
GalvatronSearchEngine.save(plan, a.out)
json.dump(galvatron_plan_to_ds_parallel_config(plan, a.gpus), open(a.out.replace(".json", "_ds_parallel_config.json"), "w"))
print("wrote", a.out)
