#!/bin/bash
# second 4-GPU bundle: new kernels (varlen single launch, padded head sizes), ring attention on NCCL, tp2 x pp2 with lazy
# NCCL init, bucketed NVLS ZeRO bench
N=4
mkdir -p gpurun_out
export HETU_BACKTRACE=1
export PYTHONPATH=$PWD:$PYTHONPATH
timeout 300 python -m pytest tests/test_kernels_gpu.py -q > gpurun_out/pytest_kernels.log 2>&1; echo "pytest kernels rc=$?"; tail -6 gpurun_out/pytest_kernels.log | cut -c1-300
timeout 400 python -m pytest tests/test_features_multi_gpu.py -q -k "ring_attention" --timeout 150 > gpurun_out/pytest_ring_$N.log 2>&1; echo "pytest ring rc=$?"; tail -5 gpurun_out/pytest_ring_$N.log | cut -c1-400
sed -i 's/timeout 300 python -m torch.distributed.run/timeout 150 python -m torch.distributed.run/' scripts/gpu_pp_debug.sh
bash scripts/gpu_pp_debug.sh $N 2 2 4
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/bench_ours_${N}_bucketed.log 2>&1; echo "bench ours rc=$?"; tail -1 gpurun_out/bench_ours_${N}_bucketed.log | cut -c1-700
