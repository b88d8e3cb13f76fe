#!/bin/bash
# 8 GPUs: the other BASELINE configurations (first with few layers to validate the path, then the full models)
mkdir -p gpurun_out; export PYTHONPATH=/root/repo
run() { local tag=$1; shift; timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) scripts/bench_configs.py "$@" > gpurun_out/cfg_$tag.log 2>&1; echo "$tag rc=$?"; grep -E "CONFIG|Error|error" gpurun_out/cfg_$tag.log | tail -3 | cut -c1-500; }
run llama7b_4l llama2-7b --layers 4 --steps 3
run llama7b llama2-7b --steps 4
run llama8b_fp8 llama3-8b --steps 4
run llama8b_bf16 llama3-8b --steps 4 --no-fp8
run moe gpt-moe --steps 6 --batch 8
