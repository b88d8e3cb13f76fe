"""Galvatron profiling on real B200s -> cost model -> searched plan for Llama-2 7B (BASELINE config #3).

Launch on N GPUs (torchrun).  Measures through this framework:
  * hardware: NCCL all-reduce bus bandwidth per group size, p2p bandwidth between two ranks, compute / communication overlap
    coefficient (ref: tools/Galvatron/galvatron/core/profiler.py:405-533, profile_hardware/profile_overlap.py);
  * model: forward ms and activation MB per Llama-2 7B layer by layer-count differencing on one GPU (ref: profiler.py:243-403);
then runs the layer-wise dynamic-programming search (csrc/planner/dp_core.cc) with the MEASURED profile and writes
hardware profile, model profile and the plan (reference JSON schema + ds_parallel_config) under hetu_b200/planner/profiles/."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import hetu_b200 as ht
from hetu_b200.models import LlamaConfig, LlamaLMHeadModel, generate_ds_parallel_config
from hetu_b200.planner import GalvatronSearchEngine, galvatron_plan_to_ds_parallel_config
from hetu_b200.planner.profiler import HardwareProfiler, ModelProfiler, profile_overlap_coefficient

ht.init_comm_group()
rank, world = dist.get_rank(), dist.get_world_size()
dev = torch.device("cuda", torch.cuda.current_device())
out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "hetu_b200", "planner", "profiles")
os.makedirs(out_dir, exist_ok=True)

# ---------------------------------------------------------------- hardware
hp = HardwareProfiler(size_mb=256, iters=10)
hw = hp.profile(world, gpus_per_node=8)
# p2p: rank 0 -> rank 1, 256 MiB messages
x = torch.ones(256 * 2**20 // 2, device=dev, dtype=torch.bfloat16)
dist.barrier()
if world >= 2 and rank < 2:
    for _ in range(2):
        (dist.send(x, 1) if rank == 0 else dist.recv(x, 0))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        (dist.send(x, 1) if rank == 0 else dist.recv(x, 0))
    torch.cuda.synchronize()
    hw.p2p_bw = 0.25 / ((time.perf_counter() - t0) / 10)
t = torch.tensor([hw.p2p_bw if rank == 1 else 0.0], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
hw.p2p_bw = float(t) if float(t) > 0 else hw.p2p_bw
hw.overlap_coe = profile_overlap_coefficient(list(range(world)))
free, total = torch.cuda.mem_get_info()
hw.mem_mb = total / 2**20 * 0.92

# ---------------------------------------------------------------- model (rank 0 alone; the others wait)
SEQ = int(os.environ.get("PROFILE_SEQ", "2048"))
measured = None
if rank == 0:
    def build(num_layers):
        cfg = LlamaConfig.llama2_7b()
        cfg.num_hidden_layers = num_layers
        cfg.max_position_embeddings = SEQ
        with ht.graph("define_and_run", create_new=True) as g, ht.autocast("bfloat16"):
            dsc = [generate_ds_parallel_config(num_layers, 1, 1, 1, 1, zero=False, devices=[0])]
            model = LlamaLMHeadModel(cfg, dsc)
            ids = ht.placeholder("int64", [SEQ], name="ids")
            pos = ht.placeholder("int64", [SEQ], name="pos")
            lab = ht.placeholder("int64", [SEQ], name="lab")
            loss = model(ids, pos, lab, seq_len=SEQ)

        def feed(bsz):
            xx = torch.randint(0, cfg.vocab_size, (SEQ,), device=dev)
            return {ids: xx, pos: torch.arange(SEQ, device=dev), lab: torch.roll(xx, -1)}
        return g, loss, feed
    try:
        measured = ModelProfiler(build, SEQ, bsz=1, warmup=3, iters=10).profile((2, 4))
    except Exception as e:   # noqa: BLE001 -- the hardware profile is still useful
        measured = {"error": f"{type(e).__name__}: {str(e)[:300]}"}
dist.barrier()

if rank == 0:
    HardwareProfiler.save(hw, os.path.join(out_dir, f"b200_hardware_{world}gpu.json"))
    json.dump(measured, open(os.path.join(out_dir, f"llama2_7b_layer_seq{SEQ}_bf16.json"), "w"), indent=1)
    cfg = LlamaConfig.llama2_7b()
    mp = ModelProfiler(None, SEQ)
    layer = mp.to_layer_profile(cfg.hidden_size, cfg.intermediate_size, cfg.num_attention_heads,
                                measured if measured and "error" not in measured else None)
    eng = GalvatronSearchEngine(cfg.num_hidden_layers, 8, layer, hw, vocab=cfg.vocab_size, hidden=cfg.hidden_size, seq=SEQ)
    plan = eng.search(batch_sizes=(8, 16, 32, 64))
    # the BASELINE config fixes pp = 2: also report the best plan under that constraint
    fixed = max((r for gbs in (8, 16, 32, 64) for ch in (2, 4, 8) if (r := eng.evaluate(2, gbs, ch))), key=lambda r: r["throughput_samples_per_s"], default=None)
    for name, p in (("searched", plan), ("pp2", fixed)):
        if p is None:
            continue
        GalvatronSearchEngine.save(p, os.path.join(out_dir, f"galvatron_config_llama2-7b_8gpus_{name}.json"))
        json.dump(galvatron_plan_to_ds_parallel_config(p, 8), open(os.path.join(out_dir, f"llama2-7b_8gpus_{name}_ds_parallel_config.json"), "w"))
    print("GALVATRON " + json.dumps({"hardware": {"allreduce_bw": hw.allreduce_bw, "p2p_bw": hw.p2p_bw, "overlap_coe": hw.overlap_coe},
                                     "layer": measured, "searched": {k: v for k, v in (plan or {}).items() if k != "strategies"},
                                     "pp2": {k: v for k, v in (fixed or {}).items() if k != "strategies"}}))
dist.barrier()
dist.destroy_process_group()
