#!/bin/bash
mkdir -p gpurun_out; export PYTHONPATH=/root/repo
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
N=2
BENCH_TP=2 BENCH_SP=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/bench_${N}_tp2sp_ag2.log 2>&1; echo "bench tp2sp (AG->GEMM v2) rc=$?"; tail -1 gpurun_out/bench_${N}_tp2sp_ag2.log | cut -c1-330
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29563 tests/workers/tp_fused_worker.py $N gpt > gpurun_out/tp_fused_ag2.log 2>&1; grep -E "TPFUSED|Error" gpurun_out/tp_fused_ag2.log | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_sm100 --launch-skip 2 -c 1 -o gpurun_out/gemm_rs_full -f python scripts/ncu_targets.py gemm_rs > gpurun_out/ncu_gemm_rs.log 2>&1; echo "ncu gemm_rs rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_sm100 --launch-skip 2 -c 1 -o gpurun_out/gemm_fp8_full -f python scripts/ncu_targets.py fp8 > gpurun_out/ncu_fp8.log 2>&1; echo "ncu fp8 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd -c 1 -o gpurun_out/attn_fwd_full -f python scripts/ncu_targets.py attn > gpurun_out/ncu_attnfwd.log 2>&1; echo "ncu attn rc=$?"
cuobjdump -sass -fun 'gemm_bf16_sm100_kernel' hetu_b200/_C*.so 2>/dev/null | grep -E "Function|UTCHMMA|UTCQMMA|UTMALDG|UTCBAR|LDTM|STS|LDS|STG|LDG|ATOMG|RED|MEMBAR|SYNCS" | awk '{ if ($1=="Function") print; else c[$2]++ } END { for (k in c) print c[k], k }' | sort -rn | head -60 > gpurun_out/sass_gemm_summary.txt
ls -la gpurun_out/*.ncu-rep | tail -5
