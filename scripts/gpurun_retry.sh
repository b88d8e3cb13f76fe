#!/bin/bash
# usage: gpurun_retry.sh OUTFILE TIMEOUT [--gpus N] -- command...   (retries while the pod reports "busy"/transient)
OUT=$1; shift; TMO=$1; shift
for i in $(seq 1 40); do
  gpurun --timeout $TMO "$@" > $OUT 2>&1
  if grep -q "status=transient" $OUT || grep -q "exit code 3" $OUT; then sleep 90; continue; fi
  break
done
