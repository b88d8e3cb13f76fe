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


# ---------------------------------------------------------------------------------------------------------------------
# exact formulations (the reference solves these with pyscipopt / pulp; here scipy's HiGHS MILP interface):
#  (1) heterogeneous data-parallel pipelines that run CONCURRENTLY: which pipeline takes which sequence (min makespan incl.
#      the 1F1B bubble of a pp-stage pipeline);  (2) inside one pipeline: how many micro-batches and which sequences share one
from hetu_b200.engine.hydraulis import batching_strategy_ilp, dispatch_batch, dispatch_batch_ilp

pipes = [StrategyCost("tp8_pp1", 1, a=1.0e-3, b=2.0e-8, max_seq=32768), StrategyCost("tp4_pp2", 1, a=1.3e-3, b=3.5e-8, max_seq=16384),
         StrategyCost("tp2_pp2", 1, a=1.9e-3, b=7.0e-8, max_seq=8192), StrategyCost("tp2_pp2b", 1, a=1.9e-3, b=7.0e-8, max_seq=8192)]
stages = [1, 2, 2, 2]
lens = [int(v) for v in np.concatenate([rng.lognormal(7.2, 0.8, 40), [30000, 14000, 9000]]).clip(64, 32768)]
ilp = dispatch_batch_ilp(lens, pipes, stages)
greedy = dispatch_batch(lens, pipes, sequential=False)
print(f"\nconcurrent pipelines, {len(lens)} sequences: MILP makespan {ilp['makespan_ms']:.1f} ms ({ilp['status']}), "
      f"greedy {greedy['makespan_ms']:.1f} ms (without bubble term)")
for st, per in zip(pipes, ilp["per_strategy"]):
    toks = sum(lens[i] for i in per["indices"])
    print(f"    {st.name}: {len(per['indices'])} sequences, {toks} tokens, {per['ms']:.1f} ms")
mine = [lens[i] for i in ilp["per_strategy"][2]["indices"]]
pack = batching_strategy_ilp(mine, pipes[2], pp=2, max_tokens=8192, min_tokens=1024)
print(f"    micro-batching of pipeline tp2_pp2: {pack['num_micro_batches']} packed micro-batches, slowest {pack['max_micro_batch_ms']:.1f} ms, "
      f"pipeline end-to-end {pack['e2e_ms']:.1f} ms; tokens per micro-batch {[sum(mine[i] for i in mb) for mb in pack['micro_batches']]}")

# planner and trainers are separate processes in production: plans travel through the rendezvous server's KV store
from hetu_b200.rpc import DeviceControllerServer, KeyValueStoreClient
from hetu_b200.rpc.kv_store import ProducerConsumer
srv = DeviceControllerServer(1, port=24911).start()
try:
    chan = ProducerConsumer(KeyValueStoreClient("127.0.0.1:24911"), "hydraulis_plan")
    producer_side = HydraulisPlanner(strategies, producer=chan)
    producer_side.plan(lens[:16])
    got = chan.consume(0)
    print(f"    plan for step {got['step']} received through the KV store: makespan {got['makespan_ms'] / 1e3:.2f} s")
finally:
    srv.shutdown()
