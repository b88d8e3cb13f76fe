#!/bin/bash
mkdir -p gpurun_out; export PYTHONPATH=/root/repo
N=${1:-2}
python scripts/fp8_debug.py > gpurun_out/fp8_debug.log 2>&1; echo "fp8dbg rc=$?"; cat gpurun_out/fp8_debug.log | tail -20
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tests/workers/moe_fused_worker.py > gpurun_out/moe_fused_$N.log 2>&1; echo "moe rc=$?"
grep -E "MOEFUSED|Error|error" gpurun_out/moe_fused_$N.log | tail -6
BENCH_TP=2 BENCH_SP=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/bench_${N}_tp2sp.log 2>&1; echo "bench tp2sp rc=$?"; tail -1 gpurun_out/bench_${N}_tp2sp.log | cut -c1-400
BENCH_TP=2 BENCH_SP=1 HETU_TP_FUSED=0 HETU_ZERO_FUSED=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/bench_${N}_tp2sp_nccl.log 2>&1; echo "bench tp2sp nccl rc=$?"; tail -1 gpurun_out/bench_${N}_tp2sp_nccl.log | cut -c1-400
