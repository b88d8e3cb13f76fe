#!/bin/bash
mkdir -p gpurun_out; export PYTHONPATH=/root/repo
python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "fp8 or attention or adam or gpt_block" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
cuobjdump -sass hetu_b200/_C*.so 2>/dev/null | grep -E "UTCQMMA|UTCHMMA|UTCOMMA|UTCMMA|UTMALDG|UTMASTG|LDTM|STTM" | awk '{print $2}' | sort | uniq -c | sort -rn | head -20
python - <<'PY'
import torch, time, math, hetu_b200 as ht
# fp8 vs bf16 GEMM timing through the public op (M=16384 tokens)
def bench(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
for (M, N, K) in [(16384, 8192, 2048), (16384, 2048, 8192), (8192, 14336, 4096), (8192, 4096, 14336)]:
    x = (torch.randn(M, K, device="cuda") ).to(torch.bfloat16); w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).to(torch.bfloat16)
    X, W = ht.from_numpy(x), ht.from_numpy(w)
    t_bf = bench(lambda: ht.linear(X, W, None))
    t_f8 = bench(lambda: ht.linear_fp8(X, W, None))
    qx = ht._C  # noqa
    fl = 2.0 * M * N * K / 1e9
    print(f"FP8BENCH M={M} N={N} K={K}: bf16 {t_bf:.3f} ms ({fl / t_bf:.0f} TFLOP/s)   fp8 incl. quantisation {t_f8:.3f} ms ({fl / t_f8:.0f} TFLOP/s)")
PY
