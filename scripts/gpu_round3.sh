#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=/root/repo
cd /root/repo
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 600 python bench.py --gpus 1 --steps 4 --warmup 3 --model gpt2-small --batch-per-gpu 8 > gpurun_out/bench_small.log 2>&1; echo "rc=$?" >> gpurun_out/bench_small.log
timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/bench_1p3b.log 2>&1; echo "rc=$?" >> gpurun_out/bench_1p3b.log
tail -30 gpurun_out/pytest_gpu.log; tail -5 gpurun_out/smoke.log; tail -5 gpurun_out/bench_small.log; tail -8 gpurun_out/bench_1p3b.log
