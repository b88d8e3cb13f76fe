#!/bin/bash
# tcgen05 flash-attention fwd/bwd correctness grid + timing at the GPT-2 1.3B / Llama shapes.
mkdir -p gpurun_out
OUT=gpurun_out/attn_test.log
: > $OUT
run() { timeout 120 ./build/attn_test "$@" >> $OUT 2>&1; echo "rc=$? for $*" >> $OUT; }
run 1 1 1 128 128 64 0
run 1 1 1 128 128 64 1
run 1 2 2 256 256 64 1
run 1 2 2 256 256 128 1
run 2 4 4 512 512 128 1
run 2 4 2 384 384 128 1
run 1 2 2 200 200 128 1
run 1 2 2 200 328 64 0
run 1 2 1 256 512 128 1
# timing: GPT-2 1.3B (16 heads x 128, seq 1024, batch 8) and Llama-7B-like (32 heads x 128, seq 2048/4096)
run 8 16 16 1024 1024 128 1 20 1
run 4 32 32 2048 2048 128 1 20 1
run 2 32 32 4096 4096 128 1 10 1
run 8 16 16 1024 1024 64 1 20 1
tail -n 100 $OUT
