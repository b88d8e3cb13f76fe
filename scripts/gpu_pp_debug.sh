#!/bin/bash
# pipeline-parallel bring-up on B200: N ranks, Llama-2 7B shapes with a reduced layer count
N=${1:-2}; TP=${2:-1}; PP=${3:-2}; LAYERS=${4:-4}
export HETU_BACKTRACE=1
export PYTHONPATH=$PWD:$PYTHONPATH
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  scripts/bench_configs.py llama2-7b --tp $TP --pp $PP --layers $LAYERS --seq 1024 --batch 4 --steps 3 --warmup 3 \
  > gpurun_out/pp_${N}_tp${TP}_pp${PP}.log 2>&1
echo "rc=$?"
grep -n "CONFIG\|HetuError\|backtrace\|_C.cpython\|Error" gpurun_out/pp_${N}_tp${TP}_pp${PP}.log | head -60
