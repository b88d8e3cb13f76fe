"""Ampelos: elastic / fault-tolerant training.  A 2-node x 8-GPU job (dp2 x tp4 x pp2) loses one GPU and sees a straggler:

  1. the elastic server surveys the nodes (`nvidia-smi` over ssh in production, a canned survey here),
  2. `ElasticStrategy.plan_from_nodes` asks the Ampelos planner (engine/strategy_ampelos.py) for the best plan on the GPUs
     that are really there -- narrower tensor-parallel groups, pipelines of different depth, uneven micro-batches,
  3. the worker command line is rewritten for the new plan, the rendezvous server hands out ranks according to the plan's
     rank -> physical-GPU mapping, and the data loader restarts from the consumed-sample count of the last checkpoint.

    python examples/ampelos/elastic_replan.py                # planning walk-through (no GPUs needed)
    python examples/ampelos/elastic_replan.py --run          # + a real restart loop with 4 local worker processes

(ref: examples/ampelos, python/hetu/rpc/{heturpc_elastic_server,elastic_arg_parser,pssh_start_elastic}.py,
 python/hetu/engine/strategy_ampelos.py)"""
import argparse
import os
import subprocess
import sys
import tempfile

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from hetu_b200.data import SyntheticDataset
from hetu_b200.data.dataloader import build_data_loader
from hetu_b200.engine import AmpelosStrategyModel, TrainerCtxs, TrainerStrategyArgs
from hetu_b200.rpc import ElasticServer, ElasticStrategy
from hetu_b200.rpc.elastic_server import available_gpus, detect_node_info

ap = argparse.ArgumentParser()
ap.add_argument("--run", action="store_true", help="also run the restart loop with local worker processes")
a = ap.parse_args()

# ---------------------------------------------------------------- 1. survey
ROW = "{i}, NVIDIA B200, 183359, {free}, 120, 0"
SURVEY = {"node0": "\n".join(ROW.format(i=i, free=181000) for i in range(8)),
          "node1": "\n".join(ROW.format(i=i, free=181000) for i in range(8) if i != 5)}     # GPU 5 of node1 fell off the bus


def runner(node, cmd):
    return SURVEY[node] if "nvidia-smi" in cmd else node + "\n"


info = detect_node_info(["node0", "node1"], runner)
gpus, smallest = available_gpus(info)
print("survey:", {v["hostname"]: [g["local_idx"] for g in v["gpus"]] for v in gpus.values()}, f"(smallest GPU {smallest} MiB)")

# ---------------------------------------------------------------- 2. re-plan
es = ElasticStrategy(tp=4, pp=2, dp=2, num_layers=32, global_micro_batches=32, memory_bound_layers=24,
                     straggler_ratios={2: 1.8})                                              # GPU 2 of node0 runs 1.8x slower
plan = es.plan_from_nodes(info)
print(f"\nnew plan: dp {plan['dp']}  tp {plan['tp']}  pp(max) {plan['pp']}  on {plan['num_devices']} GPUs, estimated step {plan['estimated_time']:.1f}")
for i, (ls, mb) in enumerate(zip(plan["hetero_layers"], plan["micro_batch_num_list"])):
    print(f"  pipeline {i}: layers per stage {ls}, {mb} micro-batches")
print("  candidates considered:")
for c in plan["candidates"]:
    print("   ", c)
print("  ranks per host:", plan["host_to_ranks"])

# the same planner, called directly (what MalleusTrainer / the elastic server do internally)
ctxs = TrainerCtxs(normal_layers=16, normal_mbn=16, memory_bound=24)
old = TrainerStrategyArgs(dp=2, tp=4, pp=2, hetero_layers=[[16, 16], [16, 16]], rank_to_device_mapping={r: r for r in range(16)})
m = AmpelosStrategyModel(ctxs, old, {d: (1.8 if d == 2 else 1.0) for d in range(16) if d != 13}, dead_devices=[13])
st, _ = m.make_plans()
assert st.hetero_layers == plan["hetero_layers"]

# ---------------------------------------------------------------- 3. restart plumbing
cmd = "python train_hetu.py --dp 2 --tp 4 --pp 2 --num_gpus 16 --hetero_stages [2,2] --steps 1000 --global_batch_size 512"
print("\nold command:", cmd)
print("new command:", es.renew_step(es.replace_cmd(cmd, plan), remaining_steps=1000 - 380))
ds = SyntheticDataset(2048, 259, 64)
loader = build_data_loader(ds, consumed_samples=380 * 4, global_batch_size=4)                # resume where step 380 left off
first = next(iter(loader))
print(f"data loader resumes at sample {380 * 4}: next batch has {len(first)} samples, loader.consumed_yielded = {loader.consumed_yielded}")

if a.run:
    # generation 0: rank 3 dies after a moment; the controller stops the others, re-plans for 3 survivors (tp shrinks to 1)
    w = tempfile.NamedTemporaryFile("w", suffix=".py", delete=False)
    w.write("import sys, time\ngen, rank, world = map(int, sys.argv[1:4])\nprint(f'  [gen {gen}] rank {rank}/{world} up', flush=True)\n"
            "if gen == 0 and rank == 3: time.sleep(0.3); sys.exit(1)\ntime.sleep(1.0)\n")
    w.close()

    def launch(gen, plan, addr):
        return [subprocess.Popen([sys.executable, w.name, str(gen), str(r), str(plan["num_devices"])]) for r in range(plan["num_devices"])]
    server = ElasticServer(launch, 4, ElasticStrategy(tp=2, pp=1), port=24810, max_restarts=2)
    rc = server.run()
    print("elastic run finished with", rc, "generations:", [(g["gen"], g["plan"]["dp"], g["plan"]["tp"]) for g in server.generations])
