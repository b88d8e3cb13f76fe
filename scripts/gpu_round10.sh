#!/bin/bash
mkdir -p gpurun_out; export PYTHONPATH=/root/repo
N=${1:-2}
bash scripts/gpu_gemm_test.sh > /dev/null 2>&1; grep -E "8192|rc=" gpurun_out/gemm_test.log | grep -v "^rc=0" | tail -30
for M in gpt llama; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 tests/workers/tp_fused_worker.py $N $M > gpurun_out/tp_fused_${M}_$N.log 2>&1; echo "tp $M rc=$?"
grep -E "TPFUSED|Error|error" gpurun_out/tp_fused_${M}_$N.log | tail -6
done
BENCH_TP=2 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/bench_${N}_tp2.log 2>&1; echo "bench tp2 rc=$?"; tail -1 gpurun_out/bench_${N}_tp2.log
BENCH_TP=2 HETU_TP_FUSED=0 HETU_ZERO_FUSED=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/bench_${N}_tp2_nccl.log 2>&1; echo "bench tp2 nccl rc=$?"; tail -1 gpurun_out/bench_${N}_tp2_nccl.log
