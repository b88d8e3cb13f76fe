#!/bin/bash
mkdir -p gpurun_out; cd /root/repo; export PYTHONPATH=/root/repo
timeout 600 python scripts/profile_step.py > gpurun_out/profile_step.log 2>&1; echo "rc=$?" >> gpurun_out/profile_step.log
cat gpurun_out/profile_step.log | tail -40
bash scripts/gpu_symm.sh 2
