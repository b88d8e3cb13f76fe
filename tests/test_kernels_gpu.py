"""Numerics of the hand-written sm_100a kernels against plain PyTorch fp32 references of the same op.
Every test goes through the public op API on an eager graph with CUDA bf16 tensors; HETU_B200_STRICT=1 (set by
conftest for gpu tests) turns any silent ATen fallback into an error, so a pass means the native kernel ran."""
import math

import numpy as np
import pytest
import torch

import hetu_b200 as ht

pytestmark = pytest.mark.gpu


def dev(t):
    return t.cuda()


def bf(*shape, scale=1.0, seed=0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(torch.bfloat16).cuda()


def leaf(t, grad=True):
    return ht.from_numpy(t, requires_grad=grad)


def close(a, b, atol, rtol):
    a, b = a.float().cpu(), b.float().cpu()
    err = (a - b).abs()
    tol = atol + rtol * b.abs()
    assert bool((err <= tol).all()), f"max err {err.max().item():.5f} (tol {atol}+{rtol}*|ref|), ref max {b.abs().max().item():.4f}"


def launches():
    return ht._C.kernel_launch_count()


@pytest.mark.parametrize("M,N,K", [(256, 512, 384), (1000, 520, 264), (2048, 1024, 2048)])
@pytest.mark.parametrize("act", ["none", "gelu"])
def test_linear_fwd_bwd(M, N, K, act):
    x, w, b = bf(M, K, seed=1), bf(N, K, scale=1 / math.sqrt(K), seed=2), bf(N, seed=3)
    n0 = launches()
    X, W, B = leaf(x), leaf(w), leaf(b)
    y = ht.linear(X, W, B, act=act)
    loss = ht.sum(y * leaf(bf(M, N, seed=4), False))
    loss.backward()
    assert launches() - n0 >= 3   # fwd GEMM + dgrad + wgrad at least
    xr, wr, br = x.float().requires_grad_(), w.float().requires_grad_(), b.float().requires_grad_()
    yr = xr @ wr.t() + br
    if act == "gelu":
        yr = torch.nn.functional.gelu(yr)
    (yr * bf(M, N, seed=4).float()).sum().backward()
    close(torch.as_tensor(y.numpy()), yr.detach(), 0.03, 0.02)
    close(torch.as_tensor(X.grad.numpy()), xr.grad, 0.06, 0.03)
    # dW / db sum over M rows of bf16 products: error grows ~ sqrt(M) * eps_bf16 * |terms|
    close(torch.as_tensor(W.grad.numpy()), wr.grad, 0.02 * math.sqrt(M), 0.03)
    close(torch.as_tensor(B.grad.numpy()), br.grad, 0.03 * math.sqrt(M), 0.03)


@pytest.mark.parametrize("M,N,K,amn,bmn", [(512, 384, 256, False, False), (1000, 520, 264, False, True), (768, 2048, 512, True, True),
                                           (2048, 2048, 1024, True, False)])
def test_gemm_narrow_accumulator_tile_matches_wide_tile_and_fp32(M, N, K, amn, bmn):
    """the 128-column accumulator tile (wave-quantisation variant) against the 256-column one and an fp32 matmul, all four
    operand majors, ragged edges, with a fused bias + GELU epilogue"""
    C = ht._C
    torch.manual_seed(0)
    a = (torch.randn((K, M) if amn else (M, K), device="cuda") * 0.1).bfloat16()
    b = (torch.randn((K, N) if bmn else (N, K), device="cuda") * 0.1).bfloat16()
    bias = (torch.randn(N, device="cuda") * 0.1).bfloat16()
    ref = (a.float().t() if amn else a.float()) @ (b.float() if bmn else b.float().t())
    for kw in ({}, {"bias": bias, "act": "gelu"}):
        o128 = C.gemm(a, b, amn, bmn, block_n=128, **kw).float()
        o256 = C.gemm(a, b, amn, bmn, block_n=256, **kw).float()
        want = torch.nn.functional.gelu(ref + bias.float()) if kw else ref
        assert (o128 - want).abs().max() <= 2e-2 * want.abs().max() + 1e-3
        assert torch.equal(o128, o256)          # same K order per output element -> bit-identical
    o32 = C.gemm(a, b, amn, bmn, out_fp32=True, block_n=128)
    assert o32.dtype == torch.float32 and (o32 - ref).abs().max() <= 1e-2 * ref.abs().max() + 1e-3


def test_linear_residual_epilogue():
    x, w, r = bf(512, 256, seed=1), bf(384, 256, scale=0.06, seed=2), bf(512, 384, seed=3)
    y = ht.linear(leaf(x, False), leaf(w, False), None, residual=leaf(r, False))
    close(torch.as_tensor(y.numpy()), x.float() @ w.float().t() + r.float(), 0.03, 0.02)


@pytest.mark.parametrize("B,S,H,Hkv,D,causal", [(2, 256, 4, 4, 128, True), (1, 384, 4, 2, 64, True), (2, 200, 2, 2, 128, False),
                                                (2, 256, 4, 2, 32, True), (1, 300, 3, 3, 96, True)])   # 32 / 96: zero-padded to 64 / 128
def test_flash_attention_fwd_bwd(B, S, H, Hkv, D, causal):
    q, k, v, g = bf(B, S, H, D, seed=1), bf(B, S, Hkv, D, seed=2), bf(B, S, Hkv, D, seed=3), bf(B, S, H, D, seed=4)
    Q, K, V = leaf(q), leaf(k), leaf(v)
    n0 = ht._C.attn_launch_count()
    o = ht.attn(Q, K, V, is_causal=causal)
    ht.sum(o * leaf(g, False)).backward()
    assert ht._C.attn_launch_count() - n0 >= 4   # fwd + delta + dq + dkv
    qr, kr, vr = q.float().requires_grad_(), k.float().requires_grad_(), v.float().requires_grad_()
    kk = kr.repeat_interleave(H // Hkv, 2)
    vv = vr.repeat_interleave(H // Hkv, 2)
    orf = torch.nn.functional.scaled_dot_product_attention(qr.transpose(1, 2), kk.transpose(1, 2), vv.transpose(1, 2),
                                                           is_causal=causal).transpose(1, 2)
    (orf * g.float()).sum().backward()
    close(torch.as_tensor(o.numpy()), orf.detach(), 0.02, 0.02)
    close(torch.as_tensor(Q.grad.numpy()), qr.grad, 0.03, 0.03)
    close(torch.as_tensor(K.grad.numpy()), kr.grad, 0.03, 0.03)
    close(torch.as_tensor(V.grad.numpy()), vr.grad, 0.03, 0.03)


@pytest.mark.parametrize("rows,cols", [(512, 2048), (300, 768), (64, 8192), (1001, 4096), (37, 1024), (16384, 2048)])
def test_layernorm_and_rmsnorm(rows, cols):
    x, w, b, g = bf(rows, cols, seed=1), bf(cols, seed=2) * 0.1 + 1, bf(cols, seed=3) * 0.1, bf(rows, cols, seed=4)
    X, W, Bb = leaf(x), leaf(w), leaf(b)
    y = ht.layer_norm(X, W, Bb, eps=1e-5)
    ht.sum(y * leaf(g, False)).backward()
    xr, wr, br = x.float().requires_grad_(), w.float().requires_grad_(), b.float().requires_grad_()
    yr = torch.nn.functional.layer_norm(xr, (cols,), wr, br, 1e-5)
    (yr * g.float()).sum().backward()
    close(torch.as_tensor(y.numpy()), yr.detach(), 0.03, 0.02)
    close(torch.as_tensor(X.grad.numpy()), xr.grad, 0.05, 0.03)
    close(torch.as_tensor(W.grad.numpy()), wr.grad, 0.5, 0.03)
    close(torch.as_tensor(Bb.grad.numpy()), br.grad, 0.5, 0.03)
    X2, W2 = leaf(x), leaf(w)
    y2 = ht.rms_norm(X2, W2, eps=1e-6)
    ht.sum(y2 * leaf(g, False)).backward()
    xr2, wr2 = x.float().requires_grad_(), w.float().requires_grad_()
    yr2 = xr2 * torch.rsqrt(xr2.pow(2).mean(-1, keepdim=True) + 1e-6) * wr2
    (yr2 * g.float()).sum().backward()
    close(torch.as_tensor(y2.numpy()), yr2.detach(), 0.03, 0.02)
    close(torch.as_tensor(X2.grad.numpy()), xr2.grad, 0.05, 0.03)
    close(torch.as_tensor(W2.grad.numpy()), wr2.grad, 0.5, 0.03)


def test_softmax_cross_entropy_and_embedding():
    T, V, H = 512, 2048, 256
    logits = bf(T, V, seed=1)
    labels = torch.randint(0, V, (T,), generator=torch.Generator().manual_seed(5)).cuda()
    labels[::7] = -1
    Lg = leaf(logits)
    loss = ht.softmax_cross_entropy_sparse(Lg, ht.from_numpy(labels), ignored_index=-1, reduction="mean")
    loss.backward()
    lr = logits.float().requires_grad_()
    lref = torch.nn.functional.cross_entropy(lr, labels, ignore_index=-1)
    lref.backward()
    assert abs(float(loss.numpy()) - float(lref)) < 2e-2
    close(torch.as_tensor(Lg.grad.numpy()), lr.grad, 2e-4, 0.03)
    # non-unit seed (device-scalar scale kernel) and per-token losses
    Lg2 = leaf(logits)
    (ht.softmax_cross_entropy_sparse(Lg2, ht.from_numpy(labels), ignored_index=-1, reduction="mean") * 3.0).backward()
    close(torch.as_tensor(Lg2.grad.numpy()), 3.0 * lr.grad, 6e-4, 0.03)
    Lg3 = leaf(logits)
    wts = bf(T, seed=9).abs()
    ht.sum(ht.softmax_cross_entropy_sparse(Lg3, ht.from_numpy(labels), ignored_index=-1, reduction="none") * leaf(wts, False)).backward()
    lr3 = logits.float().requires_grad_()
    (torch.nn.functional.cross_entropy(lr3, labels, ignore_index=-1, reduction="none") * wts.float()).sum().backward()
    close(torch.as_tensor(Lg3.grad.numpy()), lr3.grad, 0.02, 0.03)
    table = bf(V, H, seed=2)
    ids = torch.randint(0, V, (T,), generator=torch.Generator().manual_seed(6)).cuda()
    Tb = leaf(table)
    e = ht.embedding_lookup(Tb, ht.from_numpy(ids))
    g = bf(T, H, seed=7)
    ht.sum(e * leaf(g, False)).backward()
    close(torch.as_tensor(e.numpy()), table.float()[ids], 1e-6, 0)
    ref = torch.zeros(V, H, device="cuda").index_add_(0, ids, g.float())
    close(torch.as_tensor(Tb.grad.numpy()), ref, 0.05, 0.02)


def test_activations_swiglu_rotary():
    x, g = bf(512, 1024, seed=1), bf(512, 1024, seed=2)
    for name, ref in [("gelu", torch.nn.functional.gelu), ("relu", torch.relu), ("silu", torch.nn.functional.silu)]:
        X = leaf(x)
        y = getattr(ht, name)(X)
        ht.sum(y * leaf(g, False)).backward()
        xr = x.float().requires_grad_()
        (ref(xr) * g.float()).sum().backward()
        close(torch.as_tensor(y.numpy()), ref(x.float()), 0.02, 0.02)
        close(torch.as_tensor(X.grad.numpy()), xr.grad, 0.03, 0.03)
    X = leaf(x)
    y = ht.swiglu(X)
    ht.sum(y * leaf(g[:, :512].contiguous(), False)).backward()
    xr = x.float().requires_grad_()
    yr = torch.nn.functional.silu(xr[:, :512]) * xr[:, 512:]
    (yr * g[:, :512].float()).sum().backward()
    close(torch.as_tensor(y.numpy()), yr.detach(), 0.03, 0.02)
    close(torch.as_tensor(X.grad.numpy()), xr.grad, 0.05, 0.03)
    # rotary (half-split convention)
    B, S, Hh, D = 2, 64, 4, 64
    q = bf(B, S, Hh, D, seed=3)
    y = ht.rotary(leaf(q, False))
    half = D // 2
    inv = 10000.0 ** (-torch.arange(half, dtype=torch.float32) * 2 / D)
    ang = torch.arange(S, dtype=torch.float32)[:, None] * inv[None]
    cs, sn = ang.cos()[None, :, None, :].cuda(), ang.sin()[None, :, None, :].cuda()
    qa, qb = q.float()[..., :half], q.float()[..., half:]
    ref = torch.cat([qa * cs - qb * sn, qb * cs + qa * sn], -1)
    close(torch.as_tensor(y.numpy()), ref, 0.03, 0.02)


def test_mlp_backward_fuses_activation_grad_into_dgrad():
    """linear(gelu) -> linear: the second linear's dgrad GEMM applies gelu'(pre) in its epilogue (no unary_act_bwd pass)"""
    M, H, F = 1024, 512, 2048
    x, w1, b1, w2 = bf(M, H, seed=1), bf(F, H, scale=1 / math.sqrt(H), seed=2), bf(F, seed=3), bf(H, F, scale=1 / math.sqrt(F), seed=4)
    g = bf(M, H, seed=5)
    for act, ref in [("gelu", torch.nn.functional.gelu), ("relu", torch.relu)]:
        X, W1, B1, W2 = leaf(x), leaf(w1), leaf(b1), leaf(w2)
        y = ht.linear(ht.linear(X, W1, B1, act=act), W2, None)
        ht.sum(y * leaf(g, False)).backward()
        xr, w1r, b1r, w2r = (t.float().requires_grad_() for t in (x, w1, b1, w2))
        yr = ref(xr @ w1r.t() + b1r) @ w2r.t()
        (yr * g.float()).sum().backward()
        close(torch.as_tensor(y.numpy()), yr.detach(), 0.05, 0.03)
        close(torch.as_tensor(X.grad.numpy()), xr.grad, 0.1, 0.04)
        close(torch.as_tensor(W1.grad.numpy()), w1r.grad, 0.03 * math.sqrt(M), 0.04)
        close(torch.as_tensor(B1.grad.numpy()), b1r.grad, 0.04 * math.sqrt(M), 0.04)
        close(torch.as_tensor(W2.grad.numpy()), w2r.grad, 0.03 * math.sqrt(M), 0.04)


@pytest.mark.parametrize("Hq,Hkv,D", [(8, 2, 64), (4, 4, 128), (8, 1, 128)])
def test_packed_grouped_qkv_rotary_attention(Hq, Hkv, D):
    """kv-head-major packed projection [g: q x rep, k, v]: rotary + flash attention read/write it through head slots"""
    from hetu_b200.ops_extra import attn_packed, rotary_packed
    B, S = 2, 256
    rep = Hq // Hkv
    T = B * S
    qkv = bf(T, (Hq + 2 * Hkv) * D, seed=1)
    g = bf(T, Hq * D, seed=2)
    X = leaf(qkv)
    r = rotary_packed(X, S, Hq, Hkv, D, layout="hqkv")
    o = attn_packed(r, S, Hq, Hkv, D, is_causal=True, layout="hqkv")
    ht.sum(o * leaf(g, False)).backward()
    xr = qkv.float().requires_grad_()
    x5 = xr.view(B, S, Hkv, rep + 2, D)
    half = D // 2
    inv = 10000.0 ** (-torch.arange(half, dtype=torch.float32, device="cuda") * 2 / D)
    ang = torch.arange(S, dtype=torch.float32, device="cuda")[:, None] * inv[None]
    cs, sn = ang.cos()[None, :, None, None, :], ang.sin()[None, :, None, None, :]

    def rot(t):
        a, b = t[..., :half], t[..., half:]
        return torch.cat([a * cs - b * sn, b * cs + a * sn], -1)
    q = rot(x5[:, :, :, :rep]).reshape(B, S, Hq, D)
    k = rot(x5[:, :, :, rep:rep + 1]).reshape(B, S, Hkv, D).repeat_interleave(rep, 2)
    v = x5[:, :, :, rep + 1].repeat_interleave(rep, 2)
    orf = torch.nn.functional.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), is_causal=True)
    orf = orf.transpose(1, 2).reshape(T, Hq * D)
    (orf * g.float()).sum().backward()
    close(torch.as_tensor(o.numpy()), orf.detach(), 0.03, 0.03)
    close(torch.as_tensor(X.grad.numpy()), xr.grad, 0.06, 0.04)


@pytest.mark.parametrize("T,H,D,bounds", [
    (640, 4, 64, [0, 128, 328, 640]),
    (1536, 2, 128, [0, 37, 38, 300, 1000, 1129, 1536]),      # 1-token document, documents crossing several 128-row tiles
])
def test_varlen_packed_attention_native(T, H, D, bounds):
    """documents of a packed row attend only to themselves: ONE fwd launch and one delta + dQ + dK/dV launch set for the whole
    packed buffer (block-diagonal causal mask from per-token document bounds inside the kernels)"""
    from hetu_b200.ops_extra import attn_packed
    qkv, g = bf(T, 3 * H * D, seed=1), bf(T, H * D, seed=2)
    cu = torch.tensor(bounds + [T, T], dtype=torch.int32).cuda()
    X = leaf(qkv)
    n0 = ht._C.attn_launch_count()
    o = attn_packed(X, T, H, H, D, is_causal=True, layout="hqkv", cu_seqlens=ht.from_numpy(cu))
    assert ht._C.attn_launch_count() - n0 == 1, "variable-length forward must be a single kernel launch"
    ht.sum(o * leaf(g, False)).backward()
    xr = qkv.float().requires_grad_()
    y = xr.view(T, H, 3, D)
    outs = []
    for a, b in zip(bounds[:-1], bounds[1:]):
        q, k, v = (y[a:b, :, i].transpose(0, 1).unsqueeze(0) for i in range(3))
        outs.append(torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=True)[0].transpose(0, 1).reshape(b - a, H * D))
    ref = torch.cat(outs)
    (ref * g.float()).sum().backward()
    close(torch.as_tensor(o.numpy()), ref.detach(), 0.02, 0.02)
    close(torch.as_tensor(X.grad.numpy()), xr.grad, 0.04, 0.04)


def test_swiglu_interleaved():
    x, g = bf(512, 1024, seed=1), bf(512, 512, seed=2)
    X = leaf(x)
    y = ht.swiglu(X, interleaved=True)
    ht.sum(y * leaf(g, False)).backward()
    xr = x.float().requires_grad_()
    yr = torch.nn.functional.silu(xr[:, 0::2]) * xr[:, 1::2]
    (yr * g.float()).sum().backward()
    close(torch.as_tensor(y.numpy()), yr.detach(), 0.03, 0.02)
    close(torch.as_tensor(X.grad.numpy()), xr.grad, 0.05, 0.03)


@pytest.mark.parametrize("M,N,K,act", [(1024, 768, 512, "none"), (2048, 1024, 2048, "gelu"), (384, 520, 272, "none")])
def test_linear_fp8_row_scaled(M, N, K, act):
    """e4m3 operands with one fp32 scale per 1 x K block, tcgen05 kind::f8f6f4, scales applied in the epilogue"""
    x, w, b, g = bf(M, K, seed=1), bf(N, K, scale=1 / math.sqrt(K), seed=2), bf(N, seed=3), bf(M, N, seed=4)
    X, W, B = leaf(x), leaf(w), leaf(b)
    n0 = ht._C.gemm_launch_count()
    y = ht.linear_fp8(X, W, B, act=act)
    ht.sum(y * leaf(g, False)).backward()
    assert ht._C.gemm_launch_count() - n0 >= 3

    def fq(t):      # the same quantisation in torch: per-row scale, round through e4m3
        tf = t.float()
        sc = (tf.abs().amax(-1, keepdim=True) / 448.0).clamp_min(1e-30)
        return (tf / sc).to(torch.float8_e4m3fn).float() * sc
    # exact check of the tensor-core product: dequantise what the quantisation kernel produced and multiply in fp32
    qx, sx = ht._C.quantize_rowwise_e4m3(x)
    qw, sw = ht._C.quantize_rowwise_e4m3(w)
    dx, dw = qx.view(torch.float8_e4m3fn).float() * sx[:, None], qw.view(torch.float8_e4m3fn).float() * sw[:, None]
    pre = dx @ dw.t() + b.float()
    yq = torch.nn.functional.gelu(pre) if act == "gelu" else pre
    close(torch.as_tensor(y.numpy()), yq, 0.03, 0.02)
    # and the quantiser itself against the torch formula (per-row amax / 448, round-to-nearest-even)
    # (torch divides by a python scalar through a reciprocal multiply, so its scale can differ in the last bit and flip the
    #  rounding of an element sitting on a tie: allow a vanishing fraction of one-ulp differences)
    assert float(((dx - fq(x)).abs() > 1e-6).float().mean()) < 1e-4 and float(((dw - fq(w)).abs() > 1e-6).float().mean()) < 1e-4
    assert torch.allclose(sx, x.float().abs().amax(-1) / 448.0, rtol=1e-6)
    xr, wr = x.float().requires_grad_(), w.float().requires_grad_()
    yr = xr @ wr.t() + b.float()
    yr = torch.nn.functional.gelu(yr) if act == "gelu" else yr
    (yr * g.float()).sum().backward()
    # against the unquantised reference: e4m3 carries ~2^-4 relative precision per element, errors average over K
    assert float((torch.as_tensor(y.numpy()).float().cpu() - yr.detach().cpu()).abs().mean()) < 0.03
    gerr = (torch.as_tensor(X.grad.numpy()).float().cpu() - xr.grad.cpu()).abs().mean() / xr.grad.abs().mean().cpu()
    assert float(gerr) < 0.06, float(gerr)
    # the weight gradient sees the fp8 forward only through act'(pre): compare in the mean (element-wise outliers follow the
    # e4m3 rounding of pre-activations near the GELU knee)
    werr = (torch.as_tensor(W.grad.numpy()).float().cpu() - wr.grad.cpu()).abs().mean() / wr.grad.abs().mean().cpu()
    assert float(werr) < 0.05, float(werr)


def test_scaled_masked_softmax_kernels():
    B, H, Sq, Sk = 2, 4, 64, 96
    x, g = bf(B, H, Sq, Sk, seed=1), bf(B, H, Sq, Sk, seed=2)
    X = leaf(x)
    y = ht.scaled_upper_triang_masked_softmax(X, 0.25)
    ht.sum(y * leaf(g, False)).backward()
    xr = x.float().requires_grad_()
    s = (xr * 0.25).masked_fill(torch.ones(Sq, Sk, dtype=torch.bool, device="cuda").tril(Sk - Sq).logical_not(), float("-inf"))
    yr = torch.softmax(s, -1)
    (yr * g.float()).sum().backward()
    close(torch.as_tensor(y.numpy()), yr.detach(), 0.01, 0.02)
    close(torch.as_tensor(X.grad.numpy()), xr.grad, 0.01, 0.03)
    m = (torch.rand(B, 1, Sq, Sk, device="cuda") > 0.7)
    y2 = ht.scaled_masked_softmax(leaf(x, False), ht.from_numpy(m), 2.0)
    close(torch.as_tensor(y2.numpy()), torch.softmax((x.float() * 2).masked_fill(m, float("-inf")), -1), 0.01, 0.02)


def test_fused_adam_matches_torch():
    n = 4096 * 33 + 5
    p = torch.randn(n, generator=torch.Generator().manual_seed(1)).cuda()
    with ht.graph("define_and_run", create_new=True) as g, ht.autocast("bfloat16"):
        w = ht.parallel_parameter(ht.provided_initializer(p), [n], None, requires_grad=True, name="w")
        x = ht.placeholder("bfloat16", [n], name="x")
        loss = ht.sum(w * x)
        opt = ht.AdamOptimizer(lr=1e-2, beta1=0.9, beta2=0.95, eps=1e-8, weight_decay=0.1)
        train = opt.minimize(loss)
        xs = torch.randn(n, generator=torch.Generator().manual_seed(2)).cuda().to(torch.bfloat16)
        pt = p.to(torch.bfloat16).float().clone().requires_grad_()
        topt = torch.optim.AdamW([pt], lr=1e-2, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.1)
        for _ in range(3):
            g.run(loss, [loss, train], {x: xs})
            topt.zero_grad()
            (pt * xs.float()).sum().backward()
            topt.step()
        master = g.get_param(opt.get_states(w)["master"])
    close(master, pt.detach(), 2e-4, 1e-3)


def test_moe_dispatch_combine_roundtrip():
    T, H, E, k = 1024, 256, 8, 2
    x = bf(T, H, seed=1)
    logits = bf(T, E, seed=2)
    cap = int(math.ceil(k * T / E * 1.25))
    gates, idx, loc, aux = ht.moe_gate(leaf(logits, False), k, cap)
    disp = ht.moe_dispatch(leaf(x, False), idx, loc, E, cap)
    y = ht.moe_combine(disp, idx, loc, gates)
    gt, it, lt = torch.as_tensor(gates.numpy()).cuda(), torch.as_tensor(idx.numpy()).cuda().long(), torch.as_tensor(loc.numpy()).cuda()
    # identity experts: y = sum_k gate_k * x for kept assignments
    ref = x.float() * (gt * (lt >= 0)).sum(-1, keepdim=True)
    close(torch.as_tensor(y.numpy()), ref, 0.03, 0.02)
    # capacity respected and slots unique per expert
    for e in range(E):
        slots = lt[(it == e) & (lt >= 0)]
        assert slots.numel() == slots.unique().numel() and slots.numel() <= cap


def test_moe_balanced_assignment_kernel_matches_host_algorithm():
    """BASE balanced assignment on the device (csrc/kernels/moe.cu, E rounds of choose/accept without host round trips) must
    produce the same assignment as the plain-loop host implementation of the same algorithm"""
    g = torch.Generator().manual_seed(5)
    for T, E in [(4096, 8), (1000, 16), (16384, 64)]:
        scores = torch.randn(T, E, generator=g)
        scores[:, 1] += 1.5
        cap = (T + E - 1) // E
        before = launches()
        _, idx, loc, _ = ht.make_op("moe_balance_assign", [ht.from_numpy(scores.cuda())], {"capacity": cap})
        assert launches() > before
        _, ridx, rloc, _ = ht.make_op("moe_balance_assign", [ht.from_numpy(scores)], {"capacity": cap})     # CPU tensors: host loops
        i_ = torch.as_tensor(idx.numpy()).cpu()[:, 0]
        l_ = torch.as_tensor(loc.numpy()).cpu()[:, 0]
        assert torch.equal(i_, torch.as_tensor(ridx.numpy()).cpu()[:, 0])
        assert torch.equal(l_, torch.as_tensor(rloc.numpy()).cpu()[:, 0])
        counts = torch.bincount(i_.long(), minlength=E)
        assert int(counts.max()) <= cap and int(counts.sum()) == T


@pytest.mark.parametrize("rms", [False, True])
@pytest.mark.parametrize("cols", [1024, 2048, 5120])
def test_fused_dropout_add_norm(rms, cols):
    """one-pass z = residual + dropout(x); y = norm(z) (csrc/kernels/norm.cu) against an fp32 reference; with p > 0 the mask
    of the fused kernel must be the one the standalone dropout kernel produces for the same op seed (the backward relies on it)"""
    rows = 777
    x, r = bf(rows, cols, seed=1), bf(rows, cols, seed=2)
    gamma, beta = (torch.rand(cols) + 0.5).to(torch.bfloat16).cuda(), bf(cols, seed=3)
    before = launches()
    X, R = leaf(x), leaf(r)
    y, z = ht.dropout_add_norm(X, leaf(gamma), None if rms else leaf(beta), residual=R, p=0.0, eps=1e-5, rms=rms)
    assert launches() > before
    zt = (x.float() + r.float()).to(torch.bfloat16).float()
    if rms:
        yt = zt * torch.rsqrt(zt.pow(2).mean(-1, keepdim=True) + 1e-5) * gamma.float()
    else:
        yt = torch.nn.functional.layer_norm(zt, (cols,), gamma.float(), beta.float(), 1e-5)
    close(torch.as_tensor(z.numpy()), zt, 1e-6, 1e-6)
    close(torch.as_tensor(y.numpy()), yt, 0.03, 0.02)
    w = bf(rows, cols, seed=4)
    ht.sum(y * leaf(w, False) + z * leaf(w, False)).backward()
    xr = x.float().clone().requires_grad_(True)
    zz = (xr + r.float())
    yy = zz * torch.rsqrt(zz.pow(2).mean(-1, keepdim=True) + 1e-5) * gamma.float() if rms else \
        torch.nn.functional.layer_norm(zz, (cols,), gamma.float(), beta.float(), 1e-5)
    ((yy + zz) * w.float()).sum().backward()
    close(torch.as_tensor(X.grad.numpy()), xr.grad, 0.06, 0.04)
    close(torch.as_tensor(R.grad.numpy()), xr.grad, 0.06, 0.04)
    # dropout: ones in, z = mask / (1 - p); the gradient through z must carry exactly the same mask
    X1 = leaf(torch.ones(rows, cols, dtype=torch.bfloat16).cuda())
    _, z1 = ht.dropout_add_norm(X1, leaf(gamma), None if rms else leaf(beta), residual=None, p=0.25, rms=rms)
    ht.sum(z1).backward()
    zn, gn = torch.as_tensor(z1.numpy()).float(), torch.as_tensor(X1.grad.numpy()).float()
    keep = (zn > 0).float().mean().item()
    assert abs(keep - 0.75) < 0.01 and torch.equal(zn > 0, gn > 0)


@pytest.mark.parametrize("kind", ["int8", "nf4", "fp4"])
def test_blockwise_quantisation_kernels_match_host_reference(kind):
    """native absmax quantise / dequantise (csrc/kernels/quant_block.cu) against the ATen formulation of the same scheme run on
    the CPU, and the 4-bit matmul (dequantise + tcgen05 GEMM) against an fp32 product of the dequantised weight"""
    g = torch.Generator().manual_seed(2)
    w = (torch.randn(384, 512, generator=g) * 0.3)
    before = launches()
    q, a = ht.quantization(ht.from_numpy(w.to(torch.bfloat16).cuda()), kind, 64)
    assert launches() > before
    qr, ar = ht.quantization(ht.from_numpy(w.to(torch.bfloat16).float()), kind, 64)       # CPU tensors: host path
    close(torch.as_tensor(a.numpy()), torch.as_tensor(ar.numpy()), 1e-6, 1e-6)
    qa, qb = torch.as_tensor(q.numpy()).cpu(), torch.as_tensor(qr.numpy()).cpu()
    assert (qa != qb).float().mean().item() < 1e-3           # identical codes up to exact .5 ties of the rounding mode
    d = ht.dequantization(q, a, "bfloat16", 64, shape=[384, 512], quant_type=kind)
    dr = ht.dequantization(qr, ar, "float32", 64, shape=[384, 512], quant_type=kind)
    close(torch.as_tensor(d.numpy()), torch.as_tensor(dr.numpy()), 0.01, 0.01)
    if kind != "int8":
        x = bf(256, 512, seed=9)
        y = ht.matmul4bit(leaf(x, False), q, a, 64, kind, weight_shape=[384, 512])
        ref = x.float() @ torch.as_tensor(d.numpy()).float().cuda().t()
        close(torch.as_tensor(y.numpy()), ref, 0.08, 0.02)


def test_native_pools_as_the_tensor_allocator():
    """HETU_NATIVE_ALLOCATOR=1: the framework's own pool -- caching, best-fit-with-coalescing or stream-ordered
    (HETU_MEMORY_POOL) -- replaces PyTorch's CUDA allocator (pluggable allocator); a small GPT trains on each with the same
    losses, and the pool's statistics show the traffic"""
    import os
    import subprocess
    import sys
    code = r"""
import os, sys, json, torch
import hetu_b200 as ht
from hetu_b200.models import GPTConfig, GPTLMHeadModel, generate_ds_parallel_config
cfg = GPTConfig(vocab_size=1024, n_positions=256, n_embd=256, n_layer=2, n_head=2)
with ht.graph("define_and_run", create_new=True) as g, ht.autocast("bfloat16"):
    model = GPTLMHeadModel(cfg, [generate_ds_parallel_config(2, 1, 1, 1, 1)])
    ids, pos, lab = (ht.placeholder("int64", [512], name=n) for n in ("ids", "pos", "lab"))
    loss = model(ids, pos, lab, seq_len=256)
    op = ht.AdamOptimizer(lr=1e-3).minimize(loss)
    x = torch.randint(0, 1024, (512,), generator=torch.Generator().manual_seed(0)).cuda()
    p = torch.arange(256).repeat(2).cuda()
    ls = [float(g.run(loss, [loss, op], {ids: x, pos: p, lab: torch.roll(x, -1)})[0]) for _ in range(6)]
torch.cuda.synchronize()
al = ht._C.tensor_allocator("cuda:0") if os.environ.get("HETU_NATIVE_ALLOCATOR") == "1" else None
st = al.stats() if al is not None else None
print("RESULT " + json.dumps({"losses": ls, "allocs": st["num_alloc"] if st else 0, "hits": st["cache_hits"] if st else 0,
                              "reserved": st["reserved"] if st else 0, "kind": al.kind if al is not None else "torch",
                              "summary": ht.memory_pool_summary("cuda:0") if al is not None else "",
                              "torch_reserved": torch.cuda.memory_reserved() if st is None else 0}))
"""
    import json
    outs = {}
    for flag, kind in (("0", "torch"), ("1", "caching"), ("1", "bfc"), ("1", "stream_ordered")):
        env = dict(os.environ, HETU_NATIVE_ALLOCATOR=flag, HETU_MEMORY_POOL=kind if flag == "1" else "caching",
                   PYTHONPATH=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
        line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
        assert r.returncode == 0 and line, kind + "\n" + r.stdout[-2000:] + r.stderr[-3000:]
        outs[kind] = json.loads(line[0][7:])
    a = outs["torch"]
    for kind in ("caching", "bfc", "stream_ordered"):
        b = outs[kind]
        assert b["kind"] == kind and b["allocs"] > 100 and b["reserved"] > 0, b
        if kind != "stream_ordered":
            assert b["hits"] > 0, b
        for x, y in zip(a["losses"], b["losses"]):
            assert abs(x - y) < 1e-3 * max(1.0, abs(x)), (kind, a["losses"], b["losses"])


def _close_to(a, b, dtype):
    rtol, atol = (2e-5, 2e-6) if dtype == torch.float32 else (1.6e-2, 1e-2)
    assert a.dtype == b.dtype and a.shape == b.shape, (a.dtype, b.dtype, a.shape, b.shape)
    af, bf = a.float(), b.float()
    both_nan = torch.isnan(af) & torch.isnan(bf)
    ok = both_nan | (af == bf) | ((af - bf).abs() <= atol + rtol * bf.abs())
    assert bool(ok.all()), float(((af - bf).abs() * (~ok)).max())


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
def test_generic_unary_family_matches_the_library(dtype):
    """every function of the generic unary kernel against the same PyTorch function computed in fp32, on a size with a ragged tail
    and on a non-contiguous view"""
    C = ht._C
    F = torch.nn.functional
    g = torch.Generator(device="cuda").manual_seed(0)
    x = (torch.randn(3, 1237, device="cuda", generator=g) * 2.0).to(dtype)
    pos = x.abs() + 0.1
    cases = {
        "neg": (x, (), torch.neg), "reciprocal": (pos, (), torch.reciprocal), "abs": (x, (), torch.abs), "ceil": (x, (), torch.ceil),
        "floor": (x, (), torch.floor), "round": (x, (), torch.round), "exp": (x, (), torch.exp), "log": (pos, (), torch.log),
        "sqrt": (pos, (), torch.sqrt), "rsqrt": (pos, (), torch.rsqrt), "sin": (x, (), torch.sin), "cos": (x, (), torch.cos),
        "clamp": (x, (-0.5, 0.7), lambda t: torch.clamp(t, -0.5, 0.7)), "sigmoid": (x, (), torch.sigmoid), "tanh": (x, (), torch.tanh),
        "leakyrelu": (x, (0.1,), lambda t: F.leaky_relu(t, 0.1)), "elu": (x, (1.3, 1.0), lambda t: F.elu(t, 1.3)),
        "hardshrink": (x, (0.4,), lambda t: F.hardshrink(t, 0.4)), "hardsigmoid": (x, (), F.hardsigmoid),
        "hardtanh": (x, (-0.8, 0.9), lambda t: F.hardtanh(t, -0.8, 0.9)), "hardswish": (x, (), F.hardswish),
        "logsigmoid": (x, (), F.logsigmoid), "softplus": (x, (2.0, 5.0), lambda t: F.softplus(t, 2.0, 5.0)), "mish": (x, (), F.mish),
        "softshrink": (x, (0.3,), lambda t: F.softshrink(t, 0.3)), "pow": (pos, (1.7,), lambda t: torch.pow(t, 1.7)),
        "add_scalar": (x, (0.37,), lambda t: t + 0.37), "mul_scalar": (x, (-1.9,), lambda t: t * -1.9), "rsub_scalar": (x, (2.5,), lambda t: 2.5 - t),
        "rdiv_scalar": (pos, (3.0,), lambda t: 3.0 / t), "div_scalar": (x, (7.0,), lambda t: t / 7.0),
    }
    assert set(cases) == set(C.GENERIC_UNARY)
    for name, (inp, params, fn) in cases.items():
        p = list(params) + [0.0] * (2 - len(params))
        out = C.g_unary(C.GENERIC_UNARY[name], inp, p[0], p[1])
        assert out is not None, name
        _close_to(out, fn(inp.float()).to(dtype), dtype)
    view = x.t()                                           # non-contiguous input goes through the strided copy first
    out = C.g_unary(C.GENERIC_UNARY["exp"], view)
    _close_to(out, torch.exp(view.float()).to(dtype), dtype)
    assert C.g_unary(C.GENERIC_UNARY["exp"], x.cpu()) is None and C.g_unary(0, torch.arange(4, device="cuda")) is None


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_generic_binary_broadcasting_reduce_softmax_concat_cast_fill(dtype):
    C = ht._C
    g = torch.Generator(device="cuda").manual_seed(1)
    rnd = lambda *s: torch.randn(*s, device="cuda", generator=g).to(dtype)
    a = rnd(4, 33, 65)
    fns = {"add": torch.add, "sub": torch.sub, "mul": torch.mul, "div": torch.div, "max": torch.maximum, "min": torch.minimum}
    for b in (rnd(4, 33, 65), rnd(65), rnd(33, 1), rnd(4, 1, 65), rnd(1), rnd(4, 33, 65).transpose(0, 1).contiguous().transpose(0, 1)):
        for name, fn in fns.items():
            bb = b.abs() + 0.5 if name == "div" else b
            out = C.g_binary(C.GENERIC_BINARY[name], a, bb)
            assert out is not None
            _close_to(out, fn(a.float(), bb.float()).to(dtype), dtype)
    base, expo = a.abs() + 0.1, rnd(65)
    _close_to(C.g_binary(C.GENERIC_BINARY["pow"], base, expo), torch.pow(base.float(), expo.float()).to(dtype), dtype)
    big = rnd(8192, 1031)
    _close_to(C.g_binary(C.GENERIC_BINARY["add"], big, big), (big.float() + big.float()).to(dtype), dtype)       # flat 16-byte path + tail
    assert C.g_binary(0, a, a.float() if dtype != torch.float32 else a.bfloat16()) is None                        # mixed dtypes: not ours

    # reductions: last dim, first dim, middle dim, several dims, non-adjacent dims, everything; long single row (two-pass)
    x = rnd(6, 50, 70)
    lib = {"sum": torch.sum, "mean": torch.mean, "max": torch.amax, "min": torch.amin}
    for axes in ([2], [0], [1], [1, 2], [0, 1], [0, 2], [0, 1, 2], [-1]):
        for keep in (False, True):
            for name, fn in lib.items():
                out = C.g_reduce(C.GENERIC_REDUCE[name], x, axes, keep)
                assert out is not None
                ref = fn(x.float(), dim=axes, keepdim=keep).to(dtype)
                if name in ("max", "min"):
                    assert torch.equal(out, ref)
                elif dtype == torch.float32:
                    torch.testing.assert_close(out, ref, rtol=1e-4, atol=1e-4)       # summation order differs from the library's
                else:
                    torch.testing.assert_close(out.float(), ref.float(), rtol=2e-2, atol=2e-2 * max(1.0, float(ref.float().abs().max())))
    small = rnd(5, 6) * 0.3 + 1.0
    torch.testing.assert_close(C.g_reduce(C.GENERIC_REDUCE["prod"], small, [1], False).float(), torch.prod(small.float(), 1), rtol=3e-2, atol=1e-3)
    long_row = rnd(3, 300_000)
    torch.testing.assert_close(C.g_reduce(C.GENERIC_REDUCE["sum"], long_row, [1], False).float(), long_row.float().sum(1), rtol=1e-2 if dtype != torch.float32 else 1e-4,
                               atol=4.0 if dtype != torch.float32 else 5e-2)
    tall = rnd(200_000, 40)
    torch.testing.assert_close(C.g_reduce(C.GENERIC_REDUCE["mean"], tall, [0], False).float(), tall.float().mean(0), rtol=1e-2, atol=1e-2 if dtype != torch.float32 else 1e-5)
    nanx = x.clone(); nanx[1, 2, 3] = float("nan")
    assert bool(torch.isnan(C.g_reduce(C.GENERIC_REDUCE["max"], nanx, [2], False)[1, 2]))

    # softmax / log-softmax along every dim, short and long rows
    for t, dims in ((rnd(7, 33, 19), (0, 1, 2, -1)), (rnd(5, 50304), (1,)), (rnd(300, 200), (0, 1))):
        for d in dims:
            _close_to(C.g_softmax(False, t, d), torch.softmax(t.float(), d).to(dtype), dtype)
            lo = C.g_softmax(True, t, d)
            torch.testing.assert_close(lo.float(), torch.log_softmax(t.float(), d), rtol=2e-2 if dtype != torch.float32 else 1e-5,
                                       atol=6e-2 if dtype != torch.float32 else 1e-5)

    # concat along every dim (one input is a non-contiguous view), strided copy, casts, fill
    parts = [rnd(3, 5, 8), rnd(3, 5, 8).transpose(0, 1).contiguous().transpose(0, 1), rnd(3, 5, 8)]
    for d in (0, 1, 2, -1):
        out = C.g_concat(parts, d)
        assert out is not None and torch.equal(out, torch.cat(parts, d))
    odd = [rnd(4, 3), rnd(4, 7), rnd(4, 1)]
    assert torch.equal(C.g_concat(odd, 1), torch.cat(odd, 1))
    v = rnd(6, 10, 14).permute(2, 0, 1)[1:, :, ::2]
    assert torch.equal(C.g_contiguous(v), v.contiguous())
    ints = torch.randint(-1000, 1000, (777,), device="cuda")
    assert torch.equal(C.g_cast(ints, "float32"), ints.float()) and torch.equal(C.g_cast(ints, "int32"), ints.int())
    f = rnd(1025) * 50
    for name, td in (("float32", torch.float32), ("bfloat16", torch.bfloat16), ("float16", torch.float16), ("int64", torch.int64), ("int32", torch.int32)):
        if td == dtype:
            continue
        assert torch.equal(C.g_cast(f, name), f.to(td)), name
    for name, td in (("float32", torch.float32), ("bfloat16", torch.bfloat16), ("int64", torch.int64)):
        out = C.g_full([3, 1001], name, 3.0)
        assert out.dtype == td and out.shape == (3, 1001) and bool((out == 3).all())


def test_gpt_block_training_matches_fp32_reference():
    """tiny GPT: native bf16 training vs the same graph on CPU fp32 -- loss curves must agree to bf16 accuracy"""
    from hetu_b200.models import GPTConfig, GPTLMHeadModel, generate_ds_parallel_config
    cfg = GPTConfig(vocab_size=512, n_positions=128, n_embd=256, n_layer=2, n_head=2)
    B, S = 2, 128
    rng = np.random.RandomState(0)
    X = rng.randint(0, 512, (B * S,))
    P = np.tile(np.arange(S), B)
    L = np.roll(X, -1)

    def train(native):
        ht.set_seed(3)
        import contextlib
        ctx = ht.autocast("bfloat16") if native else contextlib.nullcontext()
        import os
        if not native:
            os.environ["HETU_B200_FORCE_CPU"] = "1"
        try:
            with ht.graph("define_and_run", create_new=True) as g, ctx:
                model = GPTLMHeadModel(cfg, [generate_ds_parallel_config(cfg.n_layer, 1, 1, 1, 1)])
                ids, pos, lab = (ht.placeholder("int64", [B * S], name=n) for n in ("ids", "pos", "lab"))
                loss = model(ids, pos, lab, seq_len=S)
                train_op = ht.AdamOptimizer(lr=1e-3).minimize(loss)
                out = []
                for _ in range(5):
                    out.append(float(g.run(loss, [loss, train_op], {ids: X, pos: P, lab: L})[0]))
            return out
        finally:
            os.environ.pop("HETU_B200_FORCE_CPU", None)

    n0 = launches()
    native = train(True)
    assert launches() - n0 > 100
    ref = train(False)
    for a, b in zip(native, ref):
        assert abs(a - b) < 0.08, (native, ref)
    assert native[-1] < native[0]
