"""Op library on CPU: forward vs NumPy/PyTorch, backward (eager .backward() and define-and-run gradients) vs torch
autograd -- the reference's tests/test_cpu_ops.py / test_graphcpu_ops.py strategy."""
import numpy as np
import pytest
import torch

import hetu_b200 as ht

rng = np.random.RandomState(0)


def arr(*shape, pos=False):
    a = rng.randn(*shape).astype(np.float32)
    return np.abs(a) + 0.5 if pos else a


UNARY = [
    ("abs", torch.abs, False), ("exp", torch.exp, False), ("log", torch.log, True), ("sqrt", torch.sqrt, True),
    ("rsqrt", torch.rsqrt, True), ("sin", torch.sin, False), ("cos", torch.cos, False), ("sigmoid", torch.sigmoid, False),
    ("tanh", torch.tanh, False), ("relu", torch.relu, False), ("gelu", torch.nn.functional.gelu, False),
    ("silu", torch.nn.functional.silu, False), ("reciprocal", torch.reciprocal, True), ("neg", torch.neg, False),
    ("mish", torch.nn.functional.mish, False), ("hardswish", torch.nn.functional.hardswish, False),
    ("hardsigmoid", torch.nn.functional.hardsigmoid, False), ("logsigmoid", torch.nn.functional.logsigmoid, False),
    ("ceil", torch.ceil, False), ("floor", torch.floor, False), ("round", torch.round, False),
]


@pytest.mark.parametrize("name,ref,pos", UNARY)
def test_unary_forward_backward(name, ref, pos):
    x = arr(5, 7, pos=pos)
    X = ht.from_numpy(x, requires_grad=True)
    y = getattr(ht, name)(X)
    np.testing.assert_allclose(y.numpy(), ref(torch.tensor(x)).numpy(), rtol=1e-5, atol=1e-6)
    if name in ("ceil", "floor", "round"):
        return
    ht.sum(y).backward()
    xt = torch.tensor(x, requires_grad=True)
    ref(xt).sum().backward()
    np.testing.assert_allclose(X.grad.numpy(), xt.grad.numpy(), rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("shape_a,shape_b", [((4, 5), (4, 5)), ((4, 5), (5,)), ((3, 1, 5), (4, 5))])
@pytest.mark.parametrize("op", ["add", "sub", "mul", "div"])
def test_binary_broadcast(op, shape_a, shape_b):
    a, b = arr(*shape_a), arr(*shape_b, pos=True)
    A, B = ht.from_numpy(a, requires_grad=True), ht.from_numpy(b, requires_grad=True)
    y = getattr(ht, op)(A, B)
    ta, tb = torch.tensor(a, requires_grad=True), torch.tensor(b, requires_grad=True)
    ty = {"add": ta + tb, "sub": ta - tb, "mul": ta * tb, "div": ta / tb}[op]
    np.testing.assert_allclose(y.numpy(), ty.detach().numpy(), rtol=1e-5, atol=1e-6)
    ht.sum(y * y).backward()
    (ty * ty).sum().backward()
    np.testing.assert_allclose(A.grad.numpy(), ta.grad.numpy(), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(B.grad.numpy(), tb.grad.numpy(), rtol=1e-4, atol=1e-4)


def test_scalar_arithmetic_and_operators():
    x = arr(3, 4)
    X = ht.from_numpy(x)
    np.testing.assert_allclose((X * 2.0 + 1.0).numpy(), x * 2 + 1, rtol=1e-6)
    np.testing.assert_allclose((3.0 - X).numpy(), 3 - x, rtol=1e-6)
    np.testing.assert_allclose((1.0 / ht.from_numpy(np.abs(x) + 1)).numpy(), 1 / (np.abs(x) + 1), rtol=1e-6)
    np.testing.assert_allclose((-X).numpy(), -x)
    np.testing.assert_allclose(ht.pow(ht.from_numpy(np.abs(x) + 1), 1.5).numpy(), (np.abs(x) + 1) ** 1.5, rtol=1e-5)


@pytest.mark.parametrize("mode", ["sum", "mean", "max", "min", "prod"])
@pytest.mark.parametrize("axes,keep", [(None, False), ([1], False), ([0, 2], True)])
def test_reduce(mode, axes, keep):
    x = arr(3, 4, 5)
    y = ht.reduce(ht.from_numpy(x), mode, axes, keep)
    fn = {"sum": np.sum, "mean": np.mean, "max": np.max, "min": np.min, "prod": np.prod}[mode]
    ref = fn(x, axis=tuple(axes) if axes else None, keepdims=keep)
    np.testing.assert_allclose(y.numpy(), ref, rtol=1e-4, atol=1e-5)


def test_shape_ops():
    x = arr(2, 3, 4)
    X = ht.from_numpy(x, requires_grad=True)
    np.testing.assert_allclose(ht.reshape(X, [6, 4]).numpy(), x.reshape(6, 4))
    np.testing.assert_allclose(ht.transpose(X, [2, 0, 1]).numpy(), x.transpose(2, 0, 1))
    np.testing.assert_allclose(ht.slice(X, [0, 1, 1], [2, 2, 2]).numpy(), x[:, 1:3, 1:3])
    parts = ht.split(X, 2, dim=2)
    np.testing.assert_allclose(parts[1].numpy(), x[:, :, 2:])
    np.testing.assert_allclose(ht.concat([X, X], 1).numpy(), np.concatenate([x, x], 1))
    np.testing.assert_allclose(ht.repeat(X, [1, 2, 1]).numpy(), np.tile(x, (1, 2, 1)))
    np.testing.assert_allclose(ht.roll(X, [1], [2]).numpy(), np.roll(x, 1, 2))
    np.testing.assert_allclose(ht.pad(X, [1, 1]).numpy(), np.pad(x, ((0, 0), (0, 0), (1, 1))))
    np.testing.assert_allclose(ht.broadcast(ht.from_numpy(x[0]), [5, 3, 4], [0]).numpy(), np.broadcast_to(x[0], (5, 3, 4)))
    np.testing.assert_allclose(ht.triu(ht.from_numpy(x[0])).numpy(), np.triu(x[0]))
    # gradient through reshape / transpose / slice / split+concat
    y = ht.concat(ht.split(ht.transpose(ht.reshape(X, [6, 4]), [1, 0]), 2, dim=0), 1)
    ht.sum(y * y).backward()
    xt = torch.tensor(x, requires_grad=True)
    yt = torch.cat(torch.chunk(xt.reshape(6, 4).t(), 2, 0), 1)
    (yt * yt).sum().backward()
    np.testing.assert_allclose(X.grad.numpy(), xt.grad.numpy(), rtol=1e-5)


def test_linear_and_matmul_grads():
    x, w, b = arr(6, 5), arr(4, 5), arr(4)
    X, W, B = (ht.from_numpy(v, requires_grad=True) for v in (x, w, b))
    y = ht.linear(X, W, B, act="gelu")
    ht.sum(y * y).backward()
    xt, wt, bt = (torch.tensor(v, requires_grad=True) for v in (x, w, b))
    yt = torch.nn.functional.gelu(xt @ wt.t() + bt)
    (yt * yt).sum().backward()
    for a, t in ((X, xt), (W, wt), (B, bt)):
        np.testing.assert_allclose(a.grad.numpy(), t.grad.numpy(), rtol=1e-4, atol=1e-5)
    A, Bm = ht.from_numpy(arr(3, 4), requires_grad=True), ht.from_numpy(arr(5, 4), requires_grad=True)
    c = ht.matmul(A, Bm, trans_b=True)
    assert c.shape == [3, 5]
    ht.sum(c).backward()
    assert A.grad.shape == [3, 4] and Bm.grad.shape == [5, 4]


def test_norms_embedding_losses():
    x, w, b = arr(4, 8), arr(8), arr(8)
    X, W, B = (ht.from_numpy(v, requires_grad=True) for v in (x, w, b))
    y = ht.layer_norm(X, W, B, eps=1e-5)
    ht.sum(y * y).backward()
    xt, wt, bt = (torch.tensor(v, requires_grad=True) for v in (x, w, b))
    yt = torch.nn.functional.layer_norm(xt, (8,), wt, bt, 1e-5)
    (yt * yt).sum().backward()
    np.testing.assert_allclose(y.numpy(), yt.detach().numpy(), rtol=1e-4, atol=1e-5)
    for a, t in ((X, xt), (W, wt), (B, bt)):
        np.testing.assert_allclose(a.grad.numpy(), t.grad.numpy(), rtol=1e-3, atol=1e-4)
    X2, W2 = ht.from_numpy(x, requires_grad=True), ht.from_numpy(w, requires_grad=True)
    y2 = ht.rms_norm(X2, W2, eps=1e-6)
    ht.sum(y2 * y2).backward()
    xt2, wt2 = torch.tensor(x, requires_grad=True), torch.tensor(w, requires_grad=True)
    yt2 = xt2 * torch.rsqrt(xt2.pow(2).mean(-1, keepdim=True) + 1e-6) * wt2
    (yt2 * yt2).sum().backward()
    np.testing.assert_allclose(X2.grad.numpy(), xt2.grad.numpy(), rtol=1e-3, atol=1e-4)
    # sparse CE with ignore index
    logits, labels = arr(6, 10), np.array([1, 2, -1, 4, 9, 0])
    Lg = ht.from_numpy(logits, requires_grad=True)
    loss = ht.softmax_cross_entropy_sparse(Lg, ht.from_numpy(labels), ignored_index=-1)
    loss.backward()
    lt = torch.tensor(logits, requires_grad=True)
    ref = torch.nn.functional.cross_entropy(lt, torch.tensor(labels), ignore_index=-1)
    ref.backward()
    np.testing.assert_allclose(loss.numpy(), ref.item(), rtol=1e-5)
    np.testing.assert_allclose(Lg.grad.numpy(), lt.grad.numpy(), rtol=1e-4, atol=1e-6)
    # embedding
    table, ids = arr(10, 4), np.array([1, 3, 3, 7])
    Tb = ht.from_numpy(table, requires_grad=True)
    e = ht.embedding_lookup(Tb, ht.from_numpy(ids))
    ht.sum(e * e).backward()
    tt = torch.tensor(table, requires_grad=True)
    et = torch.nn.functional.embedding(torch.tensor(ids), tt)
    (et * et).sum().backward()
    np.testing.assert_allclose(Tb.grad.numpy(), tt.grad.numpy(), rtol=1e-5)
    # dense losses
    p, t = np.clip(np.abs(arr(5, 3)) / 3, 0.05, 0.95), (rng.rand(5, 3) > 0.5).astype(np.float32)
    np.testing.assert_allclose(ht.binary_cross_entropy(ht.from_numpy(p), ht.from_numpy(t)).numpy(),
                               torch.nn.functional.binary_cross_entropy(torch.tensor(p), torch.tensor(t)).item(), rtol=1e-5)
    np.testing.assert_allclose(ht.mse_loss(ht.from_numpy(p), ht.from_numpy(t)).numpy(), ((p - t) ** 2).mean(), rtol=1e-5)


def test_attention_rotary_swiglu_cpu():
    B, S, H, D = 2, 8, 2, 16
    q, k, v = arr(B, S, H, D), arr(B, S, H, D), arr(B, S, H, D)
    Q, K, V = (ht.from_numpy(t, requires_grad=True) for t in (q, k, v))
    o = ht.attn(Q, K, V, is_causal=True)
    ht.sum(o * o).backward()
    qt, kt, vt = (torch.tensor(t, requires_grad=True) for t in (q, k, v))
    ot = torch.nn.functional.scaled_dot_product_attention(qt.transpose(1, 2), kt.transpose(1, 2), vt.transpose(1, 2),
                                                          is_causal=True).transpose(1, 2)
    (ot * ot).sum().backward()
    np.testing.assert_allclose(o.numpy(), ot.detach().numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(Q.grad.numpy(), qt.grad.numpy(), rtol=1e-3, atol=1e-4)
    np.testing.assert_allclose(V.grad.numpy(), vt.grad.numpy(), rtol=1e-3, atol=1e-4)
    # rotary is orthogonal: backward = inverse rotation
    X = ht.from_numpy(q, requires_grad=True)
    y = ht.rotary(X)
    ht.sum(y * ht.from_numpy(k)).backward()
    np.testing.assert_allclose(np.linalg.norm(y.numpy()), np.linalg.norm(q), rtol=1e-5)
    yk = ht.rotary(ht.from_numpy(q.copy()))
    np.testing.assert_allclose((yk.numpy() * k).sum(), (q * X.grad.numpy()).sum(), rtol=1e-4)
    x = arr(4, 12)
    Xs = ht.from_numpy(x, requires_grad=True)
    ys = ht.swiglu(Xs)
    ht.sum(ys).backward()
    xt = torch.tensor(x, requires_grad=True)
    (torch.nn.functional.silu(xt[:, :6]) * xt[:, 6:]).sum().backward()
    np.testing.assert_allclose(Xs.grad.numpy(), xt.grad.numpy(), rtol=1e-4, atol=1e-5)


def test_cnn_ops():
    x, w, b = arr(2, 3, 8, 8), arr(4, 3, 3, 3), arr(4)
    X, W, B = (ht.from_numpy(v, requires_grad=True) for v in (x, w, b))
    y = ht.maxpool(ht.relu(ht.conv2d(X, W, B, padding=1, stride=1)), 2, 2, 0, 2)
    ht.sum(y).backward()
    xt, wt, bt = (torch.tensor(v, requires_grad=True) for v in (x, w, b))
    yt = torch.nn.functional.max_pool2d(torch.relu(torch.nn.functional.conv2d(xt, wt, bt, padding=1)), 2, 2)
    yt.sum().backward()
    np.testing.assert_allclose(y.numpy(), yt.detach().numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(W.grad.numpy(), wt.grad.numpy(), rtol=1e-3, atol=1e-4)
    np.testing.assert_allclose(ht.avgpool(ht.from_numpy(x), 2, 2, 0, 2).numpy(),
                               torch.nn.functional.avg_pool2d(torch.tensor(x), 2, 2).numpy(), rtol=1e-5)
    np.testing.assert_allclose(ht.einsum("ij,kj->ik", ht.from_numpy(arr(3, 4)), ht.from_numpy(arr(5, 4))).shape, [3, 5])


def test_moe_ops_cpu():
    T, H, E, k = 32, 8, 4, 2
    x, logits = arr(T, H), arr(T, E)
    gates, idx, loc, aux = ht.moe_gate(ht.from_numpy(logits), k, 12)
    X = ht.from_numpy(x, requires_grad=True)
    disp = ht.moe_dispatch(X, idx, loc, E, 12)
    y = ht.moe_combine(disp, idx, loc, gates)
    g_, i_, l_ = gates.numpy(), idx.numpy(), loc.numpy()
    ref = x * (g_ * (l_ >= 0)).sum(-1, keepdims=True)
    np.testing.assert_allclose(y.numpy(), ref, rtol=1e-5, atol=1e-6)
    ht.sum(y).backward()
    np.testing.assert_allclose(X.grad.numpy(), np.ones_like(x) * (g_ * (l_ >= 0)).sum(-1, keepdims=True), rtol=1e-5, atol=1e-6)
    probs = torch.softmax(torch.tensor(logits), -1)
    np.testing.assert_array_equal(np.sort(i_, 1), np.sort(torch.topk(probs, k, -1).indices.numpy(), 1))


def test_moe_balanced_assignment_gives_every_expert_its_share():
    """BASE layers: exactly T/E tokens per expert, slots 0..cap-1 used once, tokens prefer their best expert when it has room
    (ref: hetu/v1/python/hetu/layers/BalanceGate.py balance_assignment_op)"""
    rng = np.random.RandomState(3)
    T, E = 64, 4
    scores = rng.randn(T, E).astype(np.float32)
    scores[:, 0] += 2.0                                    # everybody's favourite: must be rationed
    _, idx, loc, _ = ht.make_op("moe_balance_assign", [ht.from_numpy(scores)], {"capacity": T // E})
    i_, l_ = idx.numpy()[:, 0], loc.numpy()[:, 0]
    assert sorted(np.bincount(i_, minlength=E).tolist()) == [T // E] * E
    for e in range(E):
        assert sorted(l_[i_ == e].tolist()) == list(range(T // E))
    # expert 0 keeps the T/E highest-scoring of its proposers (all tokens propose to it in round one)
    fav = np.argsort(-scores[:, 0], kind="stable")[:T // E]
    assert set(np.nonzero(i_ == 0)[0].tolist()) == set(fav.tolist())
    # uneven T: nobody is dropped, no expert exceeds ceil(T / E)
    _, idx2, loc2, _ = ht.make_op("moe_balance_assign", [ht.from_numpy(scores[:61])], {"capacity": 16})
    assert (idx2.numpy() >= 0).all() and np.bincount(idx2.numpy()[:, 0], minlength=E).max() <= 16 and (loc2.numpy() >= 0).all()


def test_moe_balance_loss_reaches_the_router_weights():
    """the load-balancing loss must be differentiable w.r.t. the gate weights (the aux output of the moe_gate op is not)"""
    from hetu_b200.models.moe.moe_model import TopKGate, BalanceGate
    with ht.graph("define_and_run", create_new=True) as g:
        x = ht.placeholder("float32", [16, 8], name="x")
        gate = TopKGate(8, 4, k=2, capacity_factor=2.0, name="g0")
        gates, idx, loc, aux, cap = gate(x)
        d_aux, = ht.gradients(aux, [gate.wg])
        assert d_aux is not None
        xv = torch.randn(16, 8, generator=torch.Generator().manual_seed(0))
        aux_v, dw, idx_v = g.run(None, [aux, d_aux, idx], {x: xv})
        w = g.get_param(gate.wg).clone().requires_grad_(True)
        probs = torch.softmax(xv @ w.t(), -1)
        ce = torch.nn.functional.one_hot(idx_v[:, 0].long(), 4).float().mean(0)
        ref = (probs.mean(0) * ce).sum() * 4
        ref.backward()
        np.testing.assert_allclose(float(aux_v), float(ref), rtol=1e-5)
        np.testing.assert_allclose(dw.numpy(), w.grad.numpy(), rtol=1e-4, atol=1e-6)
        assert float(dw.abs().sum()) > 0
        bg = BalanceGate(8, 4, name="g1")
        gates, idx, loc, aux, cap = bg(x)
        assert aux is None and cap == 4
        iv = g.run(None, [idx], {x: xv})[0]
        assert np.bincount(iv.numpy()[:, 0], minlength=4).tolist() == [4, 4, 4, 4]


def test_blockwise_quantisation_and_matmul4bit():
    torch.manual_seed(0)
    w = torch.randn(48, 64)
    W = ht.from_numpy(w)
    for kind, tol in [("int8", 0.02), ("nf4", 0.45), ("fp4", 0.75)]:
        q, a = ht.quantization(W, kind, 32)
        d = ht.dequantization(q, a, "float32", 32, shape=[48, 64], quant_type=kind)
        assert tuple(a.shape) == (96,)
        assert (torch.as_tensor(d.numpy()) - w).abs().max().item() < tol
    x = torch.randn(8, 64)
    X = ht.from_numpy(x, requires_grad=True)
    q, a = ht.quantization(W, "nf4", 32)
    y = ht.matmul4bit(X, q, a, 32, "nf4", weight_shape=[48, 64])
    ht.sum(y).backward()
    wd = torch.as_tensor(ht.dequantization(q, a, "float32", 32, shape=[48, 64], quant_type="nf4").numpy())
    assert (torch.as_tensor(y.numpy()) - x @ wd.t()).abs().max().item() < 1e-4
    assert (torch.as_tensor(X.grad.numpy()) - wd.sum(0)).abs().max().item() < 1e-4


def test_linear_fp8_emulation_matches_quantised_reference():
    import math
    torch.manual_seed(0)
    M, K, N = 64, 64, 48
    x, w, b = torch.randn(M, K), torch.randn(N, K) / math.sqrt(K), torch.randn(N)
    X, W, B = [ht.from_numpy(t, requires_grad=True) for t in (x, w, b)]
    y = ht.linear_fp8(X, W, B, act="gelu")
    ht.sum(y).backward()

    def fq(t):
        sc = (t.abs().amax(-1, keepdim=True) / 448.0).clamp_min(1e-30)
        return (t / sc).to(torch.float8_e4m3fn).float() * sc
    ref = torch.nn.functional.gelu(fq(x) @ fq(w).t() + b)
    assert (torch.as_tensor(y.numpy()) - ref).abs().max().item() < 1e-4
    xr = x.clone().requires_grad_()
    torch.nn.functional.gelu(xr @ w.t() + b).sum().backward()
    rel = (torch.as_tensor(X.grad.numpy()) - xr.grad).abs().mean() / xr.grad.abs().mean()
    assert rel < 0.08


def test_varlen_packed_attention_equals_per_document_attention():
    from hetu_b200.ops_extra import attn_packed
    torch.manual_seed(0)
    T, H, D = 48, 2, 8
    qkv, g = torch.randn(T, 3 * H * D), torch.randn(T, H * D)
    cu = torch.tensor([0, 16, 40, 48, 48], dtype=torch.int32)       # trailing repeat = padding entry
    X = ht.from_numpy(qkv, requires_grad=True)
    o = attn_packed(X, T, H, H, D, is_causal=True, layout="hqkv", cu_seqlens=ht.from_numpy(cu))
    ht.sum(o * ht.from_numpy(g)).backward()
    xr = qkv.clone().requires_grad_()
    y = xr.view(T, H, 3, D)
    outs = []
    for a, b in [(0, 16), (16, 40), (40, 48)]:
        q, k, v = (y[a:b, :, i].transpose(0, 1).unsqueeze(0) for i in range(3))
        outs.append(torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=True)[0].transpose(0, 1).reshape(b - a, H * D))
    ref = torch.cat(outs)
    (ref * g).sum().backward()
    assert (torch.as_tensor(o.numpy()) - ref).abs().max().item() < 1e-5
    assert (torch.as_tensor(X.grad.numpy()) - xr.grad).abs().max().item() < 1e-5


def test_attn_varlen_ops_and_tensor_method_surface():
    """hetu.attn_varlen / attn_varlen_qkvpacked (cu_seqlens packing) == per-document attention; every op of the reference's
    ops.yml is also reachable as a Tensor method"""
    torch.manual_seed(0)
    T, H, Hkv, D = 40, 4, 2, 8
    q, k, v = torch.randn(T, H, D), torch.randn(T, Hkv, D), torch.randn(T, Hkv, D)
    bounds = [0, 8, 24, 40]
    cu = ht.from_numpy(torch.tensor(bounds, dtype=torch.int32))
    Q, K, V = (ht.from_numpy(t, requires_grad=True) for t in (q, k, v))
    o = ht.attn_varlen(Q, K, V, cu, cu, 16, 16, is_causal=True)
    ht.sum(o * o).backward()
    qr, kr, vr = (t.clone().requires_grad_() for t in (q, k, v))
    outs = []
    for a, b in zip(bounds[:-1], bounds[1:]):
        qq = qr[a:b].transpose(0, 1).unsqueeze(0)
        kk = kr[a:b].transpose(0, 1).repeat_interleave(H // Hkv, 0).unsqueeze(0)
        vv = vr[a:b].transpose(0, 1).repeat_interleave(H // Hkv, 0).unsqueeze(0)
        outs.append(torch.nn.functional.scaled_dot_product_attention(qq, kk, vv, is_causal=True)[0].transpose(0, 1))
    ref = torch.cat(outs)
    (ref * ref).sum().backward()
    assert (torch.as_tensor(o.numpy()) - ref).abs().max().item() < 1e-5
    for got, want in ((Q, qr), (K, kr), (V, vr)):
        assert (torch.as_tensor(got.grad.numpy()) - want.grad).abs().max().item() < 1e-4
    # [T, 3, H, D] packed form
    qkv = torch.randn(T, 3, H, D)
    o2 = ht.attn_varlen_qkvpacked(ht.from_numpy(qkv), cu, 16)
    outs = []
    for a, b in zip(bounds[:-1], bounds[1:]):
        qq, kk, vv = (qkv[a:b, i].transpose(0, 1).unsqueeze(0) for i in range(3))
        outs.append(torch.nn.functional.scaled_dot_product_attention(qq, kk, vv, is_causal=True)[0].transpose(0, 1).reshape(b - a, H * D))
    assert (torch.as_tensor(o2.numpy()) - torch.cat(outs)).abs().max().item() < 1e-5
    names = """abs abs_ add add_ as_strided attn attn_qkvpacked attn_varlen attn_varlen_qkvpacked avgpool batch_norm binary_cross_entropy
        bmm broadcast ceil ceil_ checknumeric comm concat contiguous conv2d data_transfer dequantization diagonal div div_ dot dropout dropout_
        dropout2d dropout2d_ einsum elu embedding_lookup exp exp_ flash_attn floor floor_ fused_layernorm fused_rmsnorm gather hardshrink
        hardsigmoid hardswish hardtanh index_add_ instance_norm interpolate kl_div layer_norm leakyrelu leakyrelu_ linear log log_ logsigmoid
        masked_fill matmul matmul4bit maxpool mean mish mse_loss mul mul_ neg neg_ nll_loss norm onehot pad parallel_attn pow pow_ quantization
        range_mask reciprocal reciprocal_ reduce relu relu_ repeat reshape rms_norm roll rotary round round_ rsqrt rsqrt_ sigmoid sigmoid_ silu
        sin sin_ slice softmax softmax_cross_entropy softmax_cross_entropy_sparse softplus softshrink split sqrt sqrt_ sub sub_ sum swiglu tanh
        tanh_ transpose triu vocab_parallel_cross_entropy where where_""".split()
    assert [n for n in names if not hasattr(ht, n)] == [] and [n for n in names if not hasattr(ht.Tensor, n)] == []
    x = ht.from_numpy(torch.tensor([[1.0, -2.0], [3.0, 4.0]]))
    assert torch.allclose(torch.as_tensor(x.matmul(x).numpy()), torch.tensor([[-5.0, -10.0], [15.0, 10.0]]))
    assert torch.allclose(torch.as_tensor(x.add(x).softplus().numpy()), torch.nn.functional.softplus(torch.tensor([[2.0, -4.0], [6.0, 8.0]])))
    assert hasattr(ht, "Dataloader")


@pytest.mark.parametrize("rms", [False, True])
def test_dropout_add_norm_forward_and_gradients(rms):
    """fused z = residual + dropout(x); y = norm(z) (ref: RMSNorm.cu DropoutAddLn*): p = 0 must equal the unfused ops in value
    and in every gradient (x, residual, gamma, beta), including the gradient that arrives through the residual stream z"""
    rng = np.random.RandomState(1)
    x, r = rng.randn(6, 16).astype(np.float32), rng.randn(6, 16).astype(np.float32)
    g, b = rng.rand(16).astype(np.float32) + 0.5, rng.randn(16).astype(np.float32)
    X, R, G, B = (ht.from_numpy(a, requires_grad=True) for a in (x, r, g, b))
    y, z = ht.dropout_add_norm(X, G, None if rms else B, residual=R, p=0.0, eps=1e-5, rms=rms)
    wy, wz = rng.randn(6, 16).astype(np.float32), rng.randn(6, 16).astype(np.float32)
    ht.sum(y * ht.from_numpy(wy) + z * ht.from_numpy(wz)).backward()
    xt, rt, gt, bt = (torch.tensor(a, requires_grad=True) for a in (x, r, g, b))
    zt = xt + rt
    if rms:
        yt = zt * torch.rsqrt(zt.pow(2).mean(-1, keepdim=True) + 1e-5) * gt
    else:
        yt = torch.nn.functional.layer_norm(zt, (16,), gt, bt, 1e-5)
    (yt * torch.tensor(wy) + zt * torch.tensor(wz)).sum().backward()
    np.testing.assert_allclose(y.numpy(), yt.detach().numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(z.numpy(), zt.detach().numpy(), rtol=1e-6)
    np.testing.assert_allclose(X.grad.numpy(), xt.grad.numpy(), rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(R.grad.numpy(), rt.grad.numpy(), rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(G.grad.numpy(), gt.grad.numpy(), rtol=1e-3, atol=1e-5)
    if not rms:
        np.testing.assert_allclose(B.grad.numpy(), bt.grad.numpy(), rtol=1e-3, atol=1e-5)
    # with dropout: the gradient w.r.t. x is masked exactly like the forward (zero where the activation was dropped)
    X2 = ht.from_numpy(np.ones((6, 16), np.float32), requires_grad=True)
    y2, z2 = ht.dropout_add_norm(X2, G, None if rms else B, residual=None, p=0.5, rms=rms)
    ht.sum(z2).backward()
    zn, gn = z2.numpy(), X2.grad.numpy()
    assert set(np.unique(zn).round(4).tolist()) <= {0.0, 2.0} and 0.2 < (zn == 0).mean() < 0.8
    np.testing.assert_allclose(gn, zn, rtol=1e-6)               # d sum(z) / dx = mask / (1 - p) = z for x = 1


def test_attention_dropout_composed_path_matches_flash_semantics():
    """ref: FlashAttention.cu p_dropout -- dropout acts on the softmax probabilities (inverted scaling), the mask of the forward is
    replayed in backward; p_dropout == 0 keeps the fused op; GQA and causal masks as in the fused kernels"""
    torch.manual_seed(0)
    B, S, H, HKV, D = 2, 12, 4, 2, 8
    q, k, v = torch.randn(B, S, H, D), torch.randn(B, S, HKV, D), torch.randn(B, S, HKV, D)
    Q, K, V = (ht.from_numpy(t.clone(), requires_grad=True) for t in (q, k, v))
    from hetu_b200 import ops
    fused = ht.attn(Q, K, V, is_causal=True)
    composed, probs = ops._attn_with_dropout(Q, K, V, 0.0, -1.0, True)
    assert torch.allclose(torch.as_tensor(composed.numpy()), torch.as_tensor(fused.numpy()), atol=1e-5)
    # reference with an explicit mask: out = (dropmask * softmax / (1 - p)) @ v
    ht.set_seed(11)
    out, P = ht.attn(Q, K, V, p_dropout=0.3, is_causal=True, return_softmax=True)
    Pn = torch.as_tensor(P.numpy()).reshape(B, H, S, S)
    kk, vv = k.repeat_interleave(H // HKV, 2), v.repeat_interleave(H // HKV, 2)
    sc = torch.einsum("bshd,bthd->bhst", q, kk) / D ** 0.5
    sc = sc.masked_fill(torch.triu(torch.ones(S, S), 1).bool(), float("-inf"))
    sm = torch.softmax(sc, -1)
    keep = Pn != 0
    lower = torch.tril(torch.ones(S, S)).bool()
    assert torch.allclose(Pn[keep], (sm / 0.7)[keep], atol=1e-5) and not bool(Pn[:, :, ~lower].any())     # survivors are scaled by 1 / (1 - p)
    frac = float(keep[:, :, lower].float().mean())
    assert 0.6 < frac < 0.8                                                                                   # about 70 % kept
    want = torch.einsum("bhst,bthd->bshd", Pn, vv)
    assert torch.allclose(torch.as_tensor(out.numpy()), want, atol=1e-5)
    # backward replays the same mask: dV = P_dropped^T dO
    ht.sum(out).backward()
    dv = torch.einsum("bhst,bshd->bthd", Pn, torch.ones(B, S, H, D)).reshape(B, S, HKV, H // HKV, D).sum(3)
    assert torch.allclose(torch.as_tensor(V.grad.numpy()), dv, atol=1e-4)
    assert Q.grad is not None and float(torch.as_tensor(Q.grad.numpy()).abs().sum()) > 0
