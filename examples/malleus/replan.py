"""Malleus: a device slows down -> regroup tensor-parallel groups, move layers and micro-batches away from it."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from hetu_b200.engine import StrategyModel, TrainerCtxs, TrainerStrategyArgs

ctx = TrainerCtxs(normal_layers=8, normal_mbn=8)
old = TrainerStrategyArgs(dp=2, tp=2, pp=2, rank_to_device_mapping={i: i for i in range(8)})
slowdown = {i: 1.0 for i in range(8)}
slowdown[5] = 3.0                     # measured by engine.Straggler.run_profile()
model = StrategyModel(ctx, old, slowdown)
strategy, ds_parallel_config = model.make_plans()
for pl in model.plans:
    print("pipeline:", [(g.devices, f"x{g.layer_time:.2f}") for g in pl["groups"]], "layers", pl["layers"], "micro-batches", pl["micro_batches"])
print("estimated step time", round(model.estimate_time(model.plans), 2), "(healthy homogeneous: ", (8 + 2 - 1) * 8 * 1.0, ")")
