#!/bin/bash
# 2-GPU follow-up: fused multi-GPU equivalence workers (MoE fused all-to-all included) + Galvatron profiling / search
mkdir -p gpurun_out
export HETU_BACKTRACE=1
export PYTHONPATH=$PWD:$PYTHONPATH
CUDA_VISIBLE_DEVICES=0 timeout 150 python scripts/gemm_tile_sweep.py > gpurun_out/gemm_tile_sweep.log 2>&1; echo "tile sweep rc=$?"; grep GEMMTILE gpurun_out/gemm_tile_sweep.log | cut -c1-330
CUDA_VISIBLE_DEVICES=0 timeout 120 python -m pytest tests/test_kernels_gpu.py -q -k "narrow or linear" > gpurun_out/pytest_gemm_narrow.log 2>&1; echo "pytest narrow rc=$?"; tail -3 gpurun_out/pytest_gemm_narrow.log | cut -c1-300
timeout 300 python -m pytest tests/test_fused_multi_gpu.py -q --timeout 280 > gpurun_out/pytest_fused_2.log 2>&1; echo "pytest fused(2 ranks) rc=$?"; tail -4 gpurun_out/pytest_fused_2.log | cut -c1-400
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
  scripts/profile_galvatron.py > gpurun_out/galvatron_2.log 2>&1; echo "galvatron rc=$?"; grep "GALVATRON" gpurun_out/galvatron_2.log | cut -c1-2500
mkdir -p gpurun_out/planner_profiles; cp hetu_b200/planner/profiles/*.json gpurun_out/planner_profiles/ 2>/dev/null
