#!/bin/bash
# 8-GPU validation: fused ZeRO numerics at 8 ranks, flagship bench at 8 and 4 GPUs (fused) + 8 (NCCL path), MoE EP=8
mkdir -p gpurun_out; export PYTHONPATH=/root/repo
run() { local n=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) "$@"; }
run 8 tests/workers/zero_fused_worker.py > gpurun_out/zero_fused_8.log 2>&1; echo "zero8 rc=$?"; grep -E "ZEROFUSED|Error|error" gpurun_out/zero_fused_8.log | tail -4 | cut -c1-600
run 8 bench.py --gpus 8 --steps 8 --warmup 3 > gpurun_out/bench_8.log 2>&1; echo "bench8 rc=$?"; tail -1 gpurun_out/bench_8.log | cut -c1-330
run 4 bench.py --gpus 4 --steps 8 --warmup 3 > gpurun_out/bench_4.log 2>&1; echo "bench4 rc=$?"; tail -1 gpurun_out/bench_4.log | cut -c1-330
HETU_ZERO_FUSED=0 run 8 bench.py --gpus 8 --steps 8 --warmup 3 > gpurun_out/bench_8_nccl.log 2>&1; echo "bench8 nccl rc=$?"; tail -1 gpurun_out/bench_8_nccl.log | cut -c1-330
run 8 tests/workers/moe_fused_worker.py > gpurun_out/moe_fused_8.log 2>&1; echo "moe8 rc=$?"; grep -E "MOEFUSED|Error|error" gpurun_out/moe_fused_8.log | tail -4 | cut -c1-600
run 8 tests/workers/symm_worker.py > gpurun_out/symm_8.log 2>&1; echo "symm8 rc=$?"; grep -E "SYMM|Error|error" gpurun_out/symm_8.log | tail -3 | cut -c1-900
