#!/bin/bash
# end-of-round 1-GPU validation, third pass: whole GPU test suite + the bench contract
mkdir -p gpurun_out
export HETU_BACKTRACE=1
export PYTHONPATH=$PWD:$PYTHONPATH
timeout 400 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu_final3.log 2>&1; echo "pytest gpu rc=$?"; tail -8 gpurun_out/pytest_gpu_final3.log | cut -c1-600
timeout 200 python bench.py --steps 6 --warmup 3 > gpurun_out/bench_final3_1.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_final3_1.log | cut -c1-1200
