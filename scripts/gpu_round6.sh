#!/bin/bash
# round 6: GPU numerics tests, per-op profile and the 1-GPU flagship bench after the elementwise / fusion work
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
python scripts/profile_step.py > gpurun_out/profile_step.log 2>&1; echo "profile rc=$?"; tail -25 gpurun_out/profile_step.log
python bench.py --gpus 1 --steps 8 --warmup 3 > gpurun_out/bench_1.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_1.log
