#!/bin/bash
# end-of-round 1-GPU validation, second pass: the whole GPU test suite without -x (multi-GPU tests skip on one device)
mkdir -p gpurun_out
export HETU_BACKTRACE=1
export PYTHONPATH=$PWD:$PYTHONPATH
timeout 400 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu_final2.log 2>&1; echo "pytest gpu rc=$?"; tail -12 gpurun_out/pytest_gpu_final2.log | cut -c1-600
