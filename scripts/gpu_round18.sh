#!/bin/bash
mkdir -p gpurun_out; export PYTHONPATH=/root/repo
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 200 python bench.py --gpus 1 --steps 8 --warmup 3 > gpurun_out/bench_1.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_1.log | cut -c1-300
