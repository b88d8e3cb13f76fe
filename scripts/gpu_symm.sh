#!/bin/bash
# multi-GPU: symmetric memory collectives + fused GEMM->RS, then the 2-GPU OSDP bench
mkdir -p gpurun_out; cd /root/repo; export PYTHONPATH=/root/repo
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/workers/symm_worker.py > gpurun_out/symm_$N.log 2>&1; echo "rc=$?" >> gpurun_out/symm_$N.log
grep -E "SYMM|rc=|Error|error" gpurun_out/symm_$N.log | tail -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_$N.log 2>&1; echo "rc=$?" >> gpurun_out/bench_$N.log
tail -5 gpurun_out/bench_$N.log
