#!/usr/bin/env python
"""Flagship benchmark: GPT-2 1.3B bf16 training throughput (tokens/s), sharded data parallel (OSDP / ZeRO) over N B200s.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29500 \
        bench.py --gpus 8 --steps 10 --warmup 3

Prints ONE JSON line on rank 0 (see the contract in the task description).  `value` is device-timed (CUDA events,
max over ranks); `e2e` repeats the measurement through the public API with a pinned-host -> device copy of every
step's inputs and a device -> host read of the loss inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", type=str, default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch-per-gpu", type=int, default=int(os.environ.get("BENCH_BATCH_PER_GPU", "16")))
    ap.add_argument("--seq-len", type=int, default=1024)
    ap.add_argument("--model", type=str, default=os.environ.get("BENCH_MODEL", "gpt2-1.3b"))
    ap.add_argument("--tp", type=int, default=int(os.environ.get("BENCH_TP", "1")))
    return ap.parse_args()


class ClockSampler(threading.Thread):
    """samples SM clocks / throttle reasons of this rank's GPU with nvidia-smi while the timed region runs"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def reference_arm(args):
    """Runs the unmodified reference runtime installed under baseline/_ref (tools/Galvatron of PKU-DAIR/Hetu: PyTorch +
    cuBLAS + FlashAttention-2 + NCCL/FSDP) on the same metric and config -- see baseline/reference_gpt.py and DESIGN.md
    section 4.  Nothing of hetu_b200 is imported on this path."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "baseline"))
    import reference_gpt
    try:
        return reference_gpt.run(args, ClockSampler)
    except Exception as e:   # the contract: report why and exit 0
        import traceback
        traceback.print_exc()
        if int(os.environ.get("RANK", "0")) == 0:
            print(json.dumps({"impl": "reference", "unavailable": f"{type(e).__name__}: {str(e)[:300]}"}))
        return 0


def main():
    args = parse()
    if args.impl == "reference":
        return reference_arm(args)
    import torch
    import hetu_b200 as ht
    from hetu_b200.models import GPTConfig, GPTLMHeadModel, generate_ds_parallel_config

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus or world == 1, f"launched with WORLD_SIZE={world} but --gpus {args.gpus}"
    ht.init_comm_group(world)
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    os.environ.setdefault("HETU_B200_STRICT", "1")

    if args.model == "gpt2-1.3b":
        cfg = GPTConfig.gpt2_1p3b()
    elif args.model == "gpt2-small":
        cfg = GPTConfig.gpt2_small()
    else:
        raise SystemExit(f"unknown model {args.model}")
    if int(os.environ.get("BENCH_SP", "0")):
        cfg.sequence_parallel = True
    S, B = args.seq_len, args.batch_per_gpu
    tp = args.tp
    dp = world // tp
    T = B * S
    def build_graph():
        with ht.graph("define_and_run", create_new=True) as g, ht.autocast("bfloat16"):
            dsc = [generate_ds_parallel_config(cfg.n_layer, world, dp, tp, 1, zero=True)]
            model = GPTLMHeadModel(cfg, dsc)
            ids_cfg = ht.nn.parallel.config2ds(dsc[0]["input"])
            ids = ht.parallel_placeholder("int64", [T * dp], [ids_cfg[0]], device_group_hierarchy=[ids_cfg[1]], name="input_ids")
            pos = ht.parallel_placeholder("int64", [T * dp], [ids_cfg[0]], device_group_hierarchy=[ids_cfg[1]], name="position_ids")
            lab = ht.parallel_placeholder("int64", [T * dp], [ids_cfg[0]], device_group_hierarchy=[ids_cfg[1]], name="labels")
            loss = model(ids, pos, lab, seq_len=S)
            opt = ht.AdamOptimizer(lr=1e-4, beta1=0.9, beta2=0.95, weight_decay=0.1)
            train_op = opt.minimize(loss)
        return g, model, ids, pos, lab, loss, train_op

    g, model, ids, pos, lab, loss, train_op = build_graph()

    gen = torch.Generator().manual_seed(1234 + rank)
    n_host = 4
    host_ids = [torch.randint(0, cfg.vocab_size, (T,), generator=gen).pin_memory() for _ in range(n_host)]
    host_pos = torch.arange(S).repeat(B).pin_memory()
    host_lab = [torch.roll(h, -1).pin_memory() for h in host_ids]
    dev_pos = host_pos.to(dev)
    dev_ids = [h.to(dev) for h in host_ids]
    dev_lab = [h.to(dev) for h in host_lab]
    grad_scale = 1.0 / dp

    def step_device(i):
        return g.run(loss, [loss, train_op], {ids: dev_ids[i % n_host], pos: dev_pos, lab: dev_lab[i % n_host]}, grad_scale=grad_scale)[0]

    def step_e2e(i):
        a = host_ids[i % n_host].to(dev, non_blocking=True)
        b = host_lab[i % n_host].to(dev, non_blocking=True)
        p = host_pos.to(dev, non_blocking=True)
        out = g.run(loss, [loss, train_op], {ids: a, pos: p, lab: b}, grad_scale=grad_scale)[0]
        return float(out.float().cpu())     # device -> host read of the step's loss

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t.item())

    for i in range(max(args.warmup, 3)):
        last = step_device(i)
    barrier()
    launches0 = ht._C.kernel_launch_count()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        last = step_device(i)
    e1.record()
    barrier()
    dev_ms = max_over_ranks(e0.elapsed_time(e1))
    launches = ht._C.kernel_launch_count() - launches0
    # end-to-end loop through the public API with per-step H2D input copies and a D2H loss read
    for i in range(2):
        step_e2e(i)
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for i in range(args.steps):
        lv = step_e2e(i)
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3))
    sampler.stop_flag = True
    sampler.join(timeout=2)

    verify = None
    if world > 1 and os.environ.get("BENCH_VERIFY", "1") != "0":
        # multi-GPU correctness, outside every timed region: a fresh copy of the model trained 3 steps on the fused
        # (symmetric-memory / NVLS) path and another on the NCCL twin of the same graph, same weights (name-seeded
        # initialisation) and data -- losses and parameters must agree
        del g, model
        torch.cuda.empty_cache()

        def short_run(env):
            old = {k: os.environ.get(k) for k in env}
            os.environ.update(env)
            try:
                g2, m2, i2, p2, l2, ls2, tr2 = build_graph()
                losses = [float(g2.run(ls2, [ls2, tr2], {i2: dev_ids[k % n_host], p2: dev_pos, l2: dev_lab[k % n_host]},
                                       grad_scale=grad_scale)[0].float().cpu()) for k in range(3)]
                named = dict(m2.named_parameters())
                keep = {n: g2.get_param(named[n]).float().clone() for n in sorted(named)[:: max(len(named) // 12, 1)]}
            finally:
                for k, v in old.items():
                    if v is None:
                        os.environ.pop(k, None)
                    else:
                        os.environ[k] = v
            return losses, keep
        fused_l, fused_p = short_run({})
        twin_l, twin_p = short_run({"HETU_ZERO_FUSED": "0", "HETU_TP_FUSED": "0", "HETU_TP_FUSED_AG": "0"})
        # AdamW moves every element by about lr per step whatever the gradient's size, so an element whose gradient is ~0 can go
        # the other way on the two paths (different bf16 summation order): the meaningful numbers are how FEW elements differ by
        # an update-sized amount and the mean difference, next to the worst element
        lr_steps = 1e-4 * 3
        tot = sum(fused_p[n].numel() for n in fused_p)
        diff = torch.cat([(fused_p[n] - twin_p[n]).abs().reshape(-1) for n in fused_p])
        verify = {"steps": 3, "loss_fused": fused_l, "loss_nccl_twin": twin_l,
                  "max_loss_delta": max_over_ranks(max(abs(a - b) for a, b in zip(fused_l, twin_l))),
                  "max_param_diff": max_over_ranks(float(diff.max())), "mean_param_diff": float(diff.mean()),
                  "frac_elems_diff_over_half_lr": float((diff > 0.5e-4).float().mean()), "lr_times_steps": lr_steps,
                  "param_elems_compared": int(tot), "params_compared": len(fused_p)}
    tokens_per_step = T * dp
    value = tokens_per_step * args.steps / (dev_ms / 1e3)
    e2e_value = tokens_per_step * args.steps / (e2e_ms / 1e3)
    flops_per_token = 6 * cfg.num_parameters() + 6 * cfg.n_layer * cfg.n_embd * S  # dense + causal attention (half of the 12*L*h*S of full attention), fwd+bwd
    if rank == 0:
        print(json.dumps({
            "metric": "tokens/sec (device-timed, max over ranks) GPT-2 1.3B DP+TP at 1/2/4/8 B200",
            "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic (random token ids, random-init weights)", "impl": "ours",
            "config": {"model": args.model, "layers": cfg.n_layer, "hidden": cfg.n_embd, "heads": cfg.n_head, "vocab": cfg.vocab_size,
                       "global_batch": B * dp, "seq_len": S, "parallelism": f"dp{dp}" + (f"tp{tp}" if tp > 1 else "") + (" sp" if cfg.sequence_parallel else "") + " zero (OSDP)",
                       "l2": "working set per step >> 126 MB L2 (inputs larger than L2)", "optimizer": "AdamW fp32 master"},
            "e2e": {"value": e2e_value, "unit": "tokens/s", "h2d_bytes_per_step": int(3 * T * 8), "d2h_bytes_per_step": 4,
                    "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(launches), "final_loss": float(last.float().cpu()), "e2e_final_loss": lv,
            "model_tflops_per_gpu": value / world * flops_per_token / 1e12,
            "aten_fallbacks": int(ht._C.fallback_count()), "verify": verify,
            "clocks": sampler.summary(),
        }))
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
