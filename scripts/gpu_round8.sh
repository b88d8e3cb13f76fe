#!/bin/bash
# launch list of one training step (kernel durations under ncu: NOT a bench number) + full captures of the top kernels
mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "norm or linear" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python bench.py --gpus 1 --steps 8 --warmup 3 > gpurun_out/bench_1.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_1.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 3600 -c 1300 --csv --log-file gpurun_out/launches.csv \
    python bench.py --gpus 1 --steps 1 --warmup 3 > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
python scripts/summarize_launches.py gpurun_out/launches.csv > gpurun_out/launches_summary.txt 2>&1; head -40 gpurun_out/launches_summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_sm100 --launch-skip 400 -c 3 -o gpurun_out/gemm_full -f \
    python bench.py --gpus 1 --steps 1 --warmup 3 > gpurun_out/ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_ --launch-skip 60 -c 4 -o gpurun_out/attn_full -f \
    python bench.py --gpus 1 --steps 1 --warmup 3 > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn rc=$?"
ls -la gpurun_out/*.ncu-rep
