"""Hydraulis: per-step dispatch of a variable-length global batch over several parallel strategies.

A long-context strategy (large tp, few replicas) and a short-sequence data-parallel strategy share the devices through hot
switching; every step the planner decides which sequences run under which strategy and on which replica.
(ref: examples/hydraulis/strategy/dynamic_scip.py, llama_trainer.py)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from hetu_b200.engine import HydraulisPlanner, StrategyCost

# profiled (seq_len, ms) points of ONE replica per strategy on 8 GPUs
prof_dp8 = [(256, 9.0), (1024, 38.0), (4096, 190.0), (8192, 520.0)]                 # dp8 tp1: cheap, at most 8K tokens
prof_tp4 = [(1024, 14.0), (4096, 62.0), (16384, 330.0), (32768, 980.0)]             # dp2 tp4: 32K tokens
prof_tp8 = [(4096, 40.0), (32768, 560.0), (131072, 5200.0)]                         # dp1 tp8: 128K tokens
strategies = [StrategyCost.fit("dp8_tp1", 8, prof_dp8, max_seq=8192), StrategyCost.fit("dp2_tp4", 2, prof_tp4, max_seq=32768, switch_ms=25.0),
              StrategyCost.fit("dp1_tp8", 1, prof_tp8, max_seq=131072, switch_ms=25.0)]
planner = HydraulisPlanner(strategies)
rng = np.random.RandomState(0)
for step in range(3):
    lens = np.concatenate([rng.lognormal(6.5, 0.9, 120), rng.lognormal(9.5, 0.6, 6)]).astype(int).clip(32, 131072)
    plan = planner.plan([int(v) for v in lens])
    naive = sum(strategies[2].seq_ms(int(n)) for n in lens)            # everything under the only strategy that fits every sequence
    print(f"step {step}: {len(lens)} sequences, longest {lens.max()} tokens -> {plan['makespan_ms'] / 1e3:.2f} s "
          f"(single long-context strategy: {naive / 1e3:.2f} s)")
    for name, per in zip(plan["strategies"], plan["per_strategy"]):
        if per["indices"]:
            print(f"    {name}: {len(per['indices'])} sequences over {len(per['replicas'])} replicas, {per['ms'] / 1e3:.2f} s")
