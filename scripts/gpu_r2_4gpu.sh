#!/bin/bash
# round-2 bundle on 4 GPUs of one box
N=${1:-4}
mkdir -p gpurun_out
export HETU_BACKTRACE=1
export PYTHONPATH=$PWD:$PYTHONPATH
timeout 600 python -m pytest tests/test_kernels_gpu.py -q > gpurun_out/pytest_kernels.log 2>&1; echo "pytest kernels rc=$?"; tail -3 gpurun_out/pytest_kernels.log | cut -c1-300
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  tests/workers/symm_worker.py > gpurun_out/symm_nvls_$N.log 2>&1; echo "symm rc=$?"
grep -E "SYMM|Error|error" gpurun_out/symm_nvls_$N.log | tail -4 | cut -c1-1800
timeout 1500 python -m pytest tests/test_features_multi_gpu.py -q -x --timeout 700 > gpurun_out/pytest_features_$N.log 2>&1; echo "pytest features rc=$?"; tail -6 gpurun_out/pytest_features_$N.log | cut -c1-400
HETU_TP_FUSED=0 HETU_TP_FUSED_AG=0 bash scripts/gpu_pp_debug.sh $N 2 2 4; mv gpurun_out/pp_${N}_tp2_pp2.log gpurun_out/pp_${N}_tp2_pp2_nofuse.log
bash scripts/gpu_pp_debug.sh $N 2 2 4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/bench_ours_$N.log 2>&1; echo "bench ours rc=$?"; tail -2 gpurun_out/bench_ours_$N.log | cut -c1-2200
