#!/bin/bash
# round-2 bundle on N GPUs of one box: 1-GPU kernel numerics first, then NVLS symmetric collectives, pipeline-parallel
# bring-up, and the N-GPU bench (fused ZeRO / NVLS path incl. the fused-vs-NCCL verify block) + reference arm
N=${1:-2}
mkdir -p gpurun_out
export HETU_BACKTRACE=1
export PYTHONPATH=$PWD:$PYTHONPATH
timeout 900 python -m pytest tests/test_kernels_gpu.py -q > gpurun_out/pytest_kernels.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_kernels.log | cut -c1-300
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  tests/workers/symm_worker.py > gpurun_out/symm_nvls_$N.log 2>&1; echo "symm rc=$?"
grep -E "SYMM|Error|error" gpurun_out/symm_nvls_$N.log | tail -6 | cut -c1-1500
bash scripts/gpu_pp_debug.sh $N ${2:-1} ${3:-2} ${4:-4}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/bench_ours_$N.log 2>&1; echo "bench ours rc=$?"; tail -2 gpurun_out/bench_ours_$N.log | cut -c1-1800
HETU_ZERO_NVLS=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
  bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/bench_ours_${N}_ipc.log 2>&1; echo "bench ours (peer-store path) rc=$?"; tail -1 gpurun_out/bench_ours_${N}_ipc.log | cut -c1-600
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 \
  bench.py --impl reference --gpus $N --steps 6 --warmup 3 > gpurun_out/bench_ref_$N.log 2>&1; echo "bench ref rc=$?"; tail -3 gpurun_out/bench_ref_$N.log | cut -c1-900
