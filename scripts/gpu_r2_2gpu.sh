#!/bin/bash
# round-2 multi-GPU checks in one call: NVLS symmetric collectives, pipeline-parallel bring-up
N=${1:-2}
mkdir -p gpurun_out
export HETU_BACKTRACE=1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  tests/workers/symm_worker.py > gpurun_out/symm_nvls_$N.log 2>&1; echo "symm rc=$?"
grep -E "SYMM|Error|error" gpurun_out/symm_nvls_$N.log | tail -6
bash scripts/gpu_pp_debug.sh $N ${2:-1} ${3:-2} ${4:-4}
