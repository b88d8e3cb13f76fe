import torch, math, hetu_b200 as ht
from hetu_b200 import _C
torch.manual_seed(0)
for (M, N, K) in [(1024, 768, 512), (256, 256, 128), (256, 256, 256)]:
    x = torch.randn(M, K, device="cuda").to(torch.bfloat16); w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).to(torch.bfloat16)
    qx, sx = _C.quantize_rowwise_e4m3(x); qw, sw = _C.quantize_rowwise_e4m3(w)
    dx = qx.view(torch.float8_e4m3fn).float() * sx[:, None]; dw = qw.view(torch.float8_e4m3fn).float() * sw[:, None]
    def fq(t):
        tf = t.float(); sc = (tf.abs().amax(-1, keepdim=True) / 448.0).clamp_min(1e-30)
        return (tf / sc).to(torch.float8_e4m3fn).float() * sc
    print(f"M={M} N={N} K={K}: quant x diff {float((dx - fq(x)).abs().max()):.5f}  w diff {float((dw - fq(w)).abs().max()):.5f}  "
          f"scale diff {float((sx - x.float().abs().amax(-1) / 448).abs().max()):.2e}")
    for G in (1, 2):
        y = _C.gemm_fp8(qx, sx, qw, sw, True, G)
        ref = dx @ dw.t()
        e = (y - ref).abs()
        print(f"   G={G} gemm vs dequant matmul: max err {float(e.max()):.5f} mean {float(e.mean()):.6f} ref max {float(ref.abs().max()):.3f}; "
              f"worst at {divmod(int(e.argmax()), N)}")
    qt, st = _C.quantize_transpose_e4m3(w)
    dt = qt.view(torch.float8_e4m3fn).float() * st[:, None]
    print("   transpose quant diff", float((dt - fq(w.t().contiguous())).abs().max()))
