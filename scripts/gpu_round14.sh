#!/bin/bash
mkdir -p gpurun_out; export PYTHONPATH=/root/repo
N=2
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
for M in gpt llama; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 tests/workers/tp_fused_worker.py $N $M > gpurun_out/tp_fused_ag_${M}_$N.log 2>&1; echo "tp+ag $M rc=$?"
grep -E "TPFUSED|Error|error" gpurun_out/tp_fused_ag_${M}_$N.log | tail -4 | cut -c1-500
done
BENCH_TP=2 BENCH_SP=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/bench_${N}_tp2sp_ag.log 2>&1; echo "bench tp2sp (AG->GEMM + GEMM->RS) rc=$?"; tail -1 gpurun_out/bench_${N}_tp2sp_ag.log | cut -c1-330
BENCH_TP=2 BENCH_SP=1 HETU_TP_FUSED_AG=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29553 bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/bench_${N}_tp2sp_noag.log 2>&1; echo "bench tp2sp (GEMM->RS only) rc=$?"; tail -1 gpurun_out/bench_${N}_tp2sp_noag.log | cut -c1-330
