#!/bin/bash
mkdir -p gpurun_out; export PYTHONPATH=/root/repo
timeout 120 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29791 scripts/bench_configs.py llama2-7b --steps 3 --warmup 3 > gpurun_out/cfg_llama7b_4gpu.log 2>&1; echo "llama7b rc=$?"; grep -E "CONFIG|Error|error" gpurun_out/cfg_llama7b_4gpu.log | tail -4 | cut -c1-600
