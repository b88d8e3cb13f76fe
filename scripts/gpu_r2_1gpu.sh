#!/bin/bash
# round-2 single-GPU bundle: kernel numerics, step kernel accounting, bench (ours + reference)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q > gpurun_out/pytest_kernels.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_kernels.log
timeout 400 python scripts/profile_step_torch.py > gpurun_out/step_kernels.log 2>&1; echo "profile rc=$?"; head -40 gpurun_out/step_kernels.log | cut -c1-150
timeout 300 python bench.py --steps 8 --warmup 3 > gpurun_out/ours1.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/ours1.log | cut -c1-900
