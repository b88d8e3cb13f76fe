#!/bin/bash
mkdir -p gpurun_out; export PYTHONPATH=/root/repo
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python bench.py --gpus 1 --steps 8 --warmup 3 > gpurun_out/bench_1.log 2>&1; echo "bench1 rc=$?"; tail -1 gpurun_out/bench_1.log | cut -c1-400
N=2
BENCH_TP=2 BENCH_SP=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/bench_${N}_tp2sp_ag3.log 2>&1; echo "bench tp2sp (AG->GEMM v3) rc=$?"; tail -1 gpurun_out/bench_${N}_tp2sp_ag3.log | cut -c1-330
BENCH_TP=2 BENCH_SP=1 HETU_TP_FUSED_AG=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29573 bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/bench_${N}_tp2sp_noag3.log 2>&1; echo "bench tp2sp (no AG fusion) rc=$?"; tail -1 gpurun_out/bench_${N}_tp2sp_noag3.log | cut -c1-330
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29574 tests/workers/tp_fused_worker.py $N llama > gpurun_out/tp_fused_ag3.log 2>&1; grep -E "TPFUSED|Error" gpurun_out/tp_fused_ag3.log | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29575 bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/bench_2.log 2>&1; echo "bench dp2 rc=$?"; tail -1 gpurun_out/bench_2.log | cut -c1-330
