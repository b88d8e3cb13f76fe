"""256- vs 128-column accumulator tiles of the tcgen05 GEMM on the GPT-2 1.3B / Llama shapes (device-timed, L2 flushed by
rotating over operand sets larger than L2); numerics of the narrow tile against an fp32 reference."""
import json
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hetu_b200 as ht
C = ht._C
dev = torch.device("cuda")
torch.manual_seed(0)


def time_it(fn, sets, iters=30):
    for i in range(5):
        fn(*sets[i % len(sets)])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(*sets[i % len(sets)])
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


rows = []
# (M, N, K, a_mn, b_mn): forward y = x W^T (K-major both), dgrad dx = dy W (B MN-major), wgrad dW = dy^T x (both MN-major)
shapes = [(8192, 2048, 2048, 0, 0), (8192, 2048, 8192, 0, 0), (8192, 6144, 2048, 0, 0), (8192, 8192, 2048, 0, 0),
          (8192, 2048, 6144, 0, 1), (8192, 2048, 8192, 0, 1), (8192, 8192, 2048, 0, 1),
          (6144, 2048, 8192, 1, 1), (2048, 2048, 8192, 1, 1), (8192, 2048, 8192, 1, 1), (2048, 8192, 8192, 1, 1),
          (4096, 4096, 4096, 0, 0), (16384, 4096, 4096, 0, 1), (16384, 4096, 11008, 0, 1)]
for (M, N, K, amn, bmn) in shapes:
    nset = max(2, int(200e6 // ((M * K + N * K) * 2)) + 1)
    sets = []
    for _ in range(nset):
        a = (torch.randn((K, M) if amn else (M, K), device=dev) * 0.05).bfloat16()
        b = (torch.randn((K, N) if bmn else (N, K), device=dev) * 0.05).bfloat16()
        sets.append((a, b))
    a, b = sets[0]
    ref = ((a.float().t() if amn else a.float()) @ (b.float() if bmn else b.float().t()))
    o128 = C.gemm(a, b, bool(amn), bool(bmn), block_n=128).float()
    o256 = C.gemm(a, b, bool(amn), bool(bmn), block_n=256).float()
    err128 = float((o128 - ref).abs().max() / ref.abs().max()); err256 = float((o256 - ref).abs().max() / ref.abs().max())
    t256 = time_it(lambda x, y: C.gemm(x, y, bool(amn), bool(bmn), block_n=256), sets)
    t128 = time_it(lambda x, y: C.gemm(x, y, bool(amn), bool(bmn), block_n=128), sets)
    tcub = time_it(lambda x, y: torch.matmul(x.t() if amn else x, y if bmn else y.t()), sets)
    fl = 2.0 * M * N * K
    rows.append({"M": M, "N": N, "K": K, "a_mn": amn, "b_mn": bmn, "ms_bn256": round(t256, 4), "ms_bn128": round(t128, 4), "ms_cublas": round(tcub, 4),
                 "tflops_bn256": round(fl / t256 / 1e9, 1), "tflops_bn128": round(fl / t128 / 1e9, 1), "tflops_cublas": round(fl / tcub / 1e9, 1),
                 "relerr_bn128": err128, "relerr_bn256": err256, "bit_equal": bool(torch.equal(o128, o256))})
    print("GEMMTILE " + json.dumps(rows[-1]), flush=True)
