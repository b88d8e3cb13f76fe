#!/bin/bash
# round-2 8-GPU validation: NVLS ZeRO bench at 8 ranks (with the fused-vs-NCCL verify block), BASELINE config #3
# (Llama-2 7B dp2 x tp2 x pp2), fused multi-GPU equivalence workers (2 ranks), new 1-GPU tests, Galvatron profiling + search
mkdir -p gpurun_out
export HETU_BACKTRACE=1
export PYTHONPATH=$PWD:$PYTHONPATH
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus 8 --steps 6 --warmup 3 > gpurun_out/bench_ours_8.log 2>&1; echo "bench ours rc=$?"; tail -1 gpurun_out/bench_ours_8.log | cut -c1-2400
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 \
  scripts/bench_configs.py llama2-7b --seq 2048 --batch 8 --steps 4 --warmup 3 > gpurun_out/cfg_llama7b_8.log 2>&1; echo "llama7b rc=$?"
grep -n "CONFIG\|HetuError" gpurun_out/cfg_llama7b_8.log | head -4 | cut -c1-900; grep -n "backtrace" -A12 gpurun_out/cfg_llama7b_8.log | head -16 | cut -c1-200
CUDA_VISIBLE_DEVICES=0,1 timeout 400 python -m pytest tests/test_fused_multi_gpu.py -q -x --timeout 300 > gpurun_out/pytest_fused_2.log 2>&1; echo "pytest fused(2 ranks) rc=$?"; tail -4 gpurun_out/pytest_fused_2.log | cut -c1-400
CUDA_VISIBLE_DEVICES=0 timeout 300 python -m pytest tests/test_kernels_gpu.py -q > gpurun_out/pytest_kernels.log 2>&1; echo "pytest kernels rc=$?"; tail -4 gpurun_out/pytest_kernels.log | cut -c1-300
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 \
  scripts/profile_galvatron.py > gpurun_out/galvatron_8.log 2>&1; echo "galvatron rc=$?"; grep "GALVATRON" gpurun_out/galvatron_8.log | cut -c1-2500
mkdir -p gpurun_out/planner_profiles; cp hetu_b200/planner/profiles/*.json gpurun_out/planner_profiles/ 2>/dev/null
