#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "linear or mlp" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python scripts/profile_step.py > gpurun_out/profile_step.log 2>&1; echo "profile rc=$?"; tail -70 gpurun_out/profile_step.log
