#!/bin/bash
mkdir -p gpurun_out; cd /root/repo; export PYTHONPATH=/root/repo
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python scripts/profile_step.py > gpurun_out/profile_step.log 2>&1; echo "rc=$?" >> gpurun_out/profile_step.log
tail -15 gpurun_out/pytest_gpu.log; cat gpurun_out/profile_step.log | tail -40
