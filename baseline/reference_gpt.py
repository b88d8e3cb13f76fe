"""Reference arm of bench.py: the UNMODIFIED reference runtime that is installable offline --
`tools/Galvatron` of PKU-DAIR/Hetu (package `hetu-galvatron`, pip-installed into baseline/_ref) -- training the same
GPT-2 1.3B config through its own public API and stock code path:

    initialize_galvatron(model_args, mode='train_dist')           galvatron/core/arguments.py
    config_from_meta / set_model_config                            galvatron/models/gpt/meta_configs/config_utils.py
    flash_attn.models.gpt.GPTLMHeadModel                           (the reference's model class, FlashAttention-2 kernels)
    construct_hybrid_parallel_model(...)                           galvatron/models/gpt/GPTModel_hybrid_parallel.py
    model.forward_backward(batch, iter, profiler); Adam.step()     galvatron/models/gpt/train_dist.py:58-77 (the loop body)

i.e. PyTorch + cuBLAS + FlashAttention-2 + NCCL (FSDP sharded data parallel) -- the reference's execution model.  None of
hetu_b200's models, kernels or engine is imported on this path (bench.py asserts that).  The loop below is
train_dist.py's loop with the tqdm / profiler calls replaced by the benchmark's timing protocol.

Two compatibility shims are needed for the image (both documented in DESIGN.md section 4, neither edits the reference):
`baseline/torch_compat.py` (private FSDP names of torch 2.0.1 that galvatron imports) and
`baseline/shims/fused_dense_lib.py` (flash-attention's optional cuBLASLt extension, absent from the wheel).
"""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def unavailable(why):
    if int(os.environ.get("RANK", "0")) == 0:
        print(json.dumps({"impl": "reference", "unavailable": why}))
    return 0


def run(args, ClockSampler):
    if not os.path.isdir(os.path.join(REF, "galvatron")):
        return unavailable("baseline/_ref/galvatron missing: run `python -m pip install --no-index --no-build-isolation --no-deps "
                           "--target baseline/_ref <copy of /root/reference/tools/Galvatron>` (see DESIGN.md section 4)")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("RANK", "0")
    os.environ.setdefault("WORLD_SIZE", "1")
    os.environ.setdefault("LOCAL_RANK", "0")
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(HERE, "shims"))
    sys.path.insert(0, HERE)
    import torch
    import torch_compat
    torch_compat.apply()
    assert torch.cuda.is_available(), "the reference arm needs a GPU"
    torch.cuda.set_device(local_rank)

    B, S = args.batch_per_gpu, args.seq_len
    if args.model != "gpt2-1.3b":
        return unavailable(f"reference arm only implements the headline config gpt2-1.3b, not {args.model}")
    hidden, layers, heads, vocab = 2048, 24, 16, 50304
    # sharded data parallel like the measured arm: FSDP SHARD_GRAD_OP (ZeRO-2: parameters all-gathered for the step,
    # gradients reduce-scattered, optimizer states sharded) -- same communication volume as hetu_b200's ZeRO path and
    # the faster of the reference's two sharded modes; REF_DP_TYPE=zero3 selects FULL_SHARD.  One GPU: plain ddp.
    dp_type = os.environ.get("REF_DP_TYPE", "zero2") if world > 1 else "ddp"
    sys.argv = [
        "train_dist.py",
        "--model_size", "gpt-1.5b", "--set_model_config_manually", "1", "--set_layernum_manually", "0",
        "--vocab_size", str(vocab), "--hidden_size", str(hidden), "--num_hidden_layers", str(layers),
        "--num_attention_heads", str(heads), "--seq_length", str(S),
        "--global_train_batch_size", str(B * world), "--epochs", "1", "--lr", "1e-4", "--adam_weight_decay", "0.1",
        "--dropout_prob", "0.0", "--check_loss", "0", "--profile", "0", "--save_profiled_memory", "0",
        "--pp_deg", "1", "--global_tp_deg", "1", "--global_tp_consec", "1", "--sdp", "1" if dp_type == "zero3" else "0",
        "--global_checkpoint", "0", "--chunks", "1", "--pipeline_type", "pipedream_flush",
        "--default_dp_type", dp_type, "--mixed_precision", "bf16", "--use-flash-attn",
        "--initialize_on_meta", "0", "--local-rank", str(local_rank),
    ]
    import galvatron  # noqa: F401  (puts its bundled megatron on sys.path)
    from galvatron.core import initialize_galvatron, GalvatronProfiler
    from galvatron.models.gpt.arguments import model_args
    from galvatron.models.gpt.GPTModel_hybrid_parallel import get_hybrid_parallel_configs, construct_hybrid_parallel_model
    from galvatron.models.gpt.meta_configs import config_from_meta, set_model_config
    from galvatron.utils import set_seed
    from flash_attn.models.gpt import GPTLMHeadModel
    from torch.optim import Adam

    gargs = initialize_galvatron(model_args, mode="train_dist")
    set_seed()
    dev = torch.device("cuda", local_rank)
    config = config_from_meta(gargs.model_size)
    config = set_model_config(config, gargs)
    config.head_dim = hidden // heads          # the meta json carries the 1.5b head_dim (50); only the softmax scale reads it
    hp_configs = get_hybrid_parallel_configs(model_config=config, training_args=gargs)
    gpt_model = GPTLMHeadModel(config, device="meta" if gargs.initialize_on_meta else "cpu")
    model = construct_hybrid_parallel_model(model=gpt_model, model_config=config, training_args=gargs,
                                            hybrid_parallel_configs=hp_configs)
    optimizer = Adam(model.parameters(), lr=gargs.lr, weight_decay=gargs.adam_weight_decay)
    profiler = GalvatronProfiler(gargs)          # profile=0: every hook is a no-op, as in the stock loop
    n_params = sum(p.numel() for p in model.parameters())

    loaded = [m for m in sys.modules if m == "hetu_b200" or m.startswith("hetu_b200.")]
    assert not loaded, f"hetu_b200 must not be on the reference path: {loaded}"

    gen = torch.Generator().manual_seed(1234 + rank)
    n_host = 4
    host_ids = [torch.randint(0, vocab, (B, S), generator=gen).pin_memory() for _ in range(n_host)]
    dev_ids = [h.to(dev) for h in host_ids]
    it = [0]

    def train_step(input_ids):
        loss = model.forward_backward([input_ids], it[0], profiler)   # stock: returns the loss as a python float (D2H)
        optimizer.step()
        optimizer.zero_grad()
        it[0] += 1
        return loss

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

    warm = max(args.warmup, 3)
    for i in range(warm):
        last = train_step(dev_ids[i % n_host])
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        last = train_step(dev_ids[i % n_host])
    e1.record()
    barrier()
    dev_ms = max_over_ranks(e0.elapsed_time(e1))
    # end to end: per-step pinned-host -> device input copy; the loss read back to the host is part of forward_backward
    for i in range(2):
        train_step(host_ids[i % n_host].to(dev, non_blocking=True))
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for i in range(args.steps):
        lv = train_step(host_ids[i % n_host].to(dev, non_blocking=True))
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3))
    sampler.stop_flag = True
    sampler.join(timeout=2)

    tokens_per_step = B * S * world
    value = tokens_per_step * args.steps / (dev_ms / 1e3)
    e2e_value = tokens_per_step * args.steps / (e2e_ms / 1e3)
    native = sorted({os.path.basename(getattr(m, "__file__", "") or "") for m in list(sys.modules.values())
                     if (getattr(m, "__file__", "") or "").endswith(".so") and
                     any(k in (getattr(m, "__file__", "") or "") for k in ("flash_attn", "galvatron", "hetu"))})
    if rank == 0:
        print(json.dumps({
            "metric": "tokens/sec (device-timed, max over ranks) GPT-2 1.3B DP+TP at 1/2/4/8 B200",
            "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic (random token ids, random-init weights)", "impl": "reference",
            "reference": "PKU-DAIR/Hetu tools/Galvatron (hetu-galvatron 1.0.0, unmodified, baseline/_ref): "
                         "flash_attn GPTLMHeadModel + FSDP(" + dp_type + ") + FlashAttention-2 + cuBLAS + NCCL, torch " + torch.__version__,
            "config": {"model": args.model, "layers": layers, "hidden": hidden, "heads": heads, "vocab": vocab,
                       "global_batch": B * world, "seq_len": S, "parallelism": f"dp{world} {dp_type} (FSDP)",
                       "params": int(n_params * (world if dp_type != "ddp" else 1)),
                       "l2": "working set per step >> 126 MB L2 (inputs larger than L2)", "optimizer": "Adam fp32 master (FSDP mixed precision bf16)"},
            "e2e": {"value": e2e_value, "unit": "tokens/s", "h2d_bytes_per_step": int(B * S * 8), "d2h_bytes_per_step": 4,
                    "ms_per_step": e2e_ms / args.steps},
            "final_loss": float(last), "e2e_final_loss": float(lv),
            "reference_native_so": native, "clocks": sampler.summary(),
        }))
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    return 0
