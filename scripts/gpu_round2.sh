#!/bin/bash
mkdir -p gpurun_out
bash scripts/gpu_attn_test.sh > /dev/null 2>&1
# re-time GEMM with the 8-warp epilogue
OUT=gpurun_out/gemm_test2.log
: > $OUT
run() { timeout 120 ./build/gemm_test "$@" >> $OUT 2>&1; echo "rc=$? for $*" >> $OUT; }
run 2 0 0 8192 6144 2048 0 0 20
run 2 0 0 8192 8192 2048 1 0 20
run 2 0 0 8192 2048 8192 2 0 20
run 2 0 1 8192 8192 2048 0 0 20
run 2 1 1 2048 8192 8192 3 1 20
run 2 0 0 8192 50304 2048 0 0 10
run 2 0 0 1000 520 264 1 0
run 1 0 0 1000 520 264 2 0
# one ncu capture of the 2-CTA GEMM (qkv shape)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 3 -c 1 -o gpurun_out/prof_gemm ./build/gemm_test 2 0 0 8192 6144 2048 0 0 1 > gpurun_out/ncu_gemm.log 2>&1
cat gpurun_out/attn_test.log | tail -70
cat $OUT
