#!/bin/bash
# end-of-round 1-GPU validation: the whole GPU test suite (multi-GPU tests skip on one device), the driver's bench contract,
# and ncu --set full captures of the round-2 kernels
mkdir -p gpurun_out
export HETU_BACKTRACE=1
export PYTHONPATH=$PWD:$PYTHONPATH
timeout 420 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest gpu rc=$?"; tail -6 gpurun_out/pytest_gpu_final.log | cut -c1-600
timeout 200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_final_1.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_final_1.log | cut -c1-1500
for t in norm varlen quant generic; do
  timeout 150 ncu --set full --clock-control none --import-source on -k regex:'dropout_add|norm_bwd|attn_fwd|attn_bwd_dkv|quantize|reduce_rows|reduce_cols|softmax_rows|strided_copy|unary_kernel' \
    -c 8 -f -o gpurun_out/r2_$t python scripts/ncu_targets_r2.py $t > gpurun_out/ncu_r2_$t.log 2>&1; echo "ncu $t rc=$?"; tail -2 gpurun_out/ncu_r2_$t.log | cut -c1-200
done
ls -la gpurun_out/*.ncu-rep 2>/dev/null | tail -5
