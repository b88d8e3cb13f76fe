#!/bin/bash
# 2-GPU round: new GEMM store epilogue numerics, fused ZeRO path vs NCCL path, 1- and 2-GPU flagship bench
mkdir -p gpurun_out; export PYTHONPATH=/root/repo
python -m pytest tests/test_kernels_gpu.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tests/workers/zero_fused_worker.py > gpurun_out/zero_fused_$N.log 2>&1; echo "zero rc=$?"
grep -E "ZEROFUSED|Error|error|rc=" gpurun_out/zero_fused_$N.log | tail -12
python bench.py --gpus 1 --steps 8 --warmup 3 > gpurun_out/bench_1.log 2>&1; echo "bench1 rc=$?"; tail -1 gpurun_out/bench_1.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/bench_$N.log 2>&1; echo "bench$N rc=$?"
tail -2 gpurun_out/bench_$N.log
HETU_ZERO_FUSED=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/bench_${N}_nccl.log 2>&1; echo "bench$N nccl rc=$?"
tail -1 gpurun_out/bench_${N}_nccl.log
