#!/bin/bash
# Runs the standalone tcgen05 GEMM harness over a grid of layouts/shapes.
# Each config is wrapped in its own timeout so a hung kernel cannot eat the box.
mkdir -p gpurun_out
OUT=gpurun_out/gemm_test.log
: > $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $OUT 2>&1
run() { timeout 120 ./build/gemm_test "$@" >> $OUT 2>&1; echo "rc=$? for $*" >> $OUT; }
for G in 1 2; do
  # NT / NN / TN / TT, small and ragged shapes
  run $G 0 0 256 256 128
  run $G 0 0 512 768 512
  run $G 0 1 512 768 512
  run $G 1 0 512 768 512
  run $G 1 1 512 768 512
  run $G 0 0 1000 520 264
  run $G 0 1 1000 520 264
  run $G 1 1 1000 520 264
  run $G 0 0 512 768 512 1 0
  run $G 0 0 512 768 512 2 0
  run $G 1 1 512 768 512 3 1
  run $G 0 0 512 768 512 3 0
done
# GPT-2 1.3B shapes (T = 8192 tokens, h = 2048), timed against cuBLAS
for G in 1 2; do
  run $G 0 0 8192 6144 2048 0 0 20    # qkv fwd
  run $G 0 0 8192 8192 2048 0 0 20    # fc1 fwd
  run $G 0 0 8192 2048 8192 0 0 20    # fc2 fwd
  run $G 0 1 8192 2048 8192 0 0 20    # fc1 dgrad
  run $G 1 1 8192 2048 8192 0 1 20    # fc1 wgrad (fp32 out)
  run $G 0 0 8192 8192 2048 1 0 20    # fc1 fwd + bias + gelu + pre-act store
done
tail -n 120 $OUT
