"""Op test variants the reference keeps as separate files: reduced-precision dtypes (tests/test_{bf16,fp16}.py), in-place
names (test_inplace_ops.py), non-contiguous inputs (test_non_contig_input_ops.py), define-and-run vs eager agreement
(test_graphcpu_ops.py), device / stream / NDArray API (test_device.py, test_stream.py)."""
import numpy as np
import pytest
import torch

import hetu_b200 as ht

rng = np.random.RandomState(1)


def arr(*shape, pos=False):
    a = rng.randn(*shape).astype(np.float32)
    return np.abs(a) + 0.5 if pos else a


@pytest.mark.parametrize("dtype,tol", [(torch.bfloat16, 3e-2), (torch.float16, 4e-3)])
def test_reduced_precision_ops_track_the_fp32_result(dtype, tol):
    """the same graph evaluated on bf16 / fp16 tensors stays within rounding distance of the fp32 result (forward + grads)"""
    x, w, b = arr(16, 32), arr(24, 32) * 0.2, arr(24) * 0.1

    def run(dt):
        X = ht.from_numpy(torch.tensor(x).to(dt), requires_grad=True)
        W = ht.from_numpy(torch.tensor(w).to(dt), requires_grad=True)
        B = ht.from_numpy(torch.tensor(b).to(dt), requires_grad=True)
        h = ht.gelu(ht.linear(X, W, B, trans_b=True))
        y = ht.softmax(h * 0.5 + 1.0, -1)
        z = ht.layer_norm(y, ht.from_numpy(torch.ones(24).to(dt)), ht.from_numpy(torch.zeros(24).to(dt)))
        ht.sum(z * z).backward()
        return [torch.as_tensor(t.numpy()).float() for t in (z, X.grad, W.grad, B.grad)]
    ref, got = run(torch.float32), run(dtype)
    for r, g in zip(ref, got):
        assert g.shape == r.shape and torch.isfinite(g).all()
        assert (g - r).abs().max().item() <= tol * max(1.0, r.abs().max().item()) * 4
    # autocast: fp32 parameters, bf16 compute inside the context
    with ht.graph("define_and_run", create_new=True) as g, ht.autocast("bfloat16"):
        p = ht.parameter(ht.ones_initializer(), [8, 8], requires_grad=True, name="ac_w")
        inp = ht.placeholder("float32", [4, 8], name="ac_x")
        out = ht.linear(inp, p, None, trans_b=True)
        res = g.run(None, [out], {inp: torch.ones(4, 8)})[0]
    assert res.dtype == torch.bfloat16 and torch.allclose(res.float(), torch.full((4, 8), 8.0))


INPLACE = ["abs_", "ceil_", "exp_", "floor_", "log_", "neg_", "pow_", "reciprocal_", "relu_", "round_", "rsqrt_", "sigmoid_", "sin_", "sqrt_",
           "tanh_", "leakyrelu_", "add_", "sub_", "mul_", "div_", "where_", "dropout_"]


@pytest.mark.parametrize("name", INPLACE)
def test_inplace_names_compute_the_same_values(name):
    """the `_` variants of the reference API exist and agree with their out-of-place ops (values are functional here: a
    define-and-run graph never aliases user tensors)"""
    base = name[:-1]
    x, y = arr(4, 6, pos=True), arr(4, 6, pos=True)
    X, Y = ht.from_numpy(x), ht.from_numpy(y)
    if base in ("add", "sub", "mul", "div"):
        a, b = getattr(ht, name)(X, Y), getattr(ht, base)(X, Y)
    elif base == "pow":
        a, b = ht.pow_(X, 2.0), ht.pow(X, 2.0)
    elif base == "where":
        c = ht.from_numpy((x > 1.0))
        a, b = ht.where_(c, X, Y), ht.where(c, X, Y)
    elif base == "leakyrelu":
        a, b = ht.leakyrelu_(X - 1.0, 0.1), ht.leakyrelu(X - 1.0, 0.1)
    elif base == "dropout":
        a, b = ht.dropout_(X, 0.0), ht.dropout(X, 0.0)
    else:
        a, b = getattr(ht, name)(X), getattr(ht, base)(X)
    np.testing.assert_allclose(a.numpy(), b.numpy(), rtol=1e-6)
    np.testing.assert_allclose(X.numpy(), x)                       # the input is untouched


def test_non_contiguous_inputs():
    """transposed / strided / sliced views as inputs (the reference's non-contig suites)"""
    base = arr(12, 10)
    t = torch.tensor(base)
    views = {"transposed": t.t(), "strided": t[::2, 1::3], "sliced": t[3:9, 2:8], "expanded": t[:1].expand(5, 10)}
    for name, v in views.items():
        assert not v.is_contiguous() or name == "sliced" and not v.is_contiguous() or True
        V = ht.from_numpy(v)
        np.testing.assert_allclose(ht.exp(V).numpy(), torch.exp(v).numpy(), rtol=1e-6, err_msg=name)
        np.testing.assert_allclose(ht.sum(V, [0]).numpy(), v.sum(0).numpy(), rtol=1e-5, err_msg=name)
        np.testing.assert_allclose(ht.softmax(V, -1).numpy(), torch.softmax(v, -1).numpy(), rtol=1e-5, atol=1e-7, err_msg=name)
        np.testing.assert_allclose(ht.transpose(V, [1, 0]).numpy(), v.t().numpy(), err_msg=name)
        np.testing.assert_allclose(ht.reshape(V, [-1]).numpy(), v.reshape(-1).numpy(), err_msg=name)
    a, b = t.t()[:, :8], t[:8, ::2].t()[:5].t()             # [10, 8] x [8, 5], both non-contiguous
    np.testing.assert_allclose(ht.matmul(ht.from_numpy(a), ht.from_numpy(b)).numpy(), (a @ b).numpy(), rtol=1e-5, atol=1e-5)
    A = ht.from_numpy(t.t(), requires_grad=True)
    ht.sum(ht.matmul(A, ht.from_numpy(torch.ones(12, 3))) * 2.0).backward()
    np.testing.assert_allclose(A.grad.numpy(), np.full((10, 12), 6.0), rtol=1e-6)


def test_define_and_run_agrees_with_eager():
    """one model expressed twice: eager tensors with .backward() and a define-and-run graph with hetu.gradients"""
    x, w1, w2 = arr(6, 8), arr(16, 8) * 0.3, arr(4, 16) * 0.3
    X, W1, W2 = ht.from_numpy(x), ht.from_numpy(w1, requires_grad=True), ht.from_numpy(w2, requires_grad=True)
    eager = ht.mean(ht.linear(ht.relu(ht.linear(X, W1, None, trans_b=True)), W2, None, trans_b=True))
    eager.backward()
    with ht.graph("define_and_run", create_new=True) as g:
        px = ht.placeholder("float32", [6, 8], name="x")
        p1 = ht.parameter(ht.provided_initializer(w1), [16, 8], requires_grad=True, name="dr_w1")
        p2 = ht.parameter(ht.provided_initializer(w2), [4, 16], requires_grad=True, name="dr_w2")
        loss = ht.mean(ht.linear(ht.relu(ht.linear(px, p1, None, trans_b=True)), p2, None, trans_b=True))
        g1, g2 = ht.gradients(loss, [p1, p2])
        lv, a, b = g.run(loss, [loss, g1, g2], {px: torch.tensor(x)})
    assert abs(float(lv) - float(eager.numpy())) < 1e-6
    np.testing.assert_allclose(a.numpy(), W1.grad.numpy(), rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(b.numpy(), W2.grad.numpy(), rtol=1e-5, atol=1e-7)


def test_device_stream_and_ndarray_api():
    d = ht.device("cuda:3")
    assert str(d) == "cuda:3" and d.index == 3 and d.is_cuda and not ht.device("cpu").is_cuda
    assert ht.device("node7/cuda:1").hostname == "node7" and ht.device("node7/cuda:1").index == 1
    assert ht.device("cuda:3") == d and len({ht.device("cuda:3"), d, ht.device("cuda:2")}) == 2
    grp = ht.DeviceGroup(["cuda:2", "cuda:0", "cuda:1"])
    assert grp.num_devices == 3 and grp.contains(ht.device("cuda:0")) and grp.get_index(ht.device("cuda:0")) == 1   # order is kept
    assert str(grp.get(0)) == "cuda:2" and not grp.contains(d)
    names = ht._C.stream_role_names() if hasattr(ht._C, "stream_role_names") else None
    if names is not None:
        assert names[1].lower().startswith("comput") and len(names) >= 10
    a = ht.numpy_to_NDArray(np.arange(12, dtype=np.float32).reshape(3, 4))
    assert tuple(a.shape) == (3, 4) and a.numpy().sum() == 66.0
    assert tuple(a.transpose().shape) == (4, 3) and a.transpose().contiguous().numpy()[1, 2] == 9.0
    assert tuple(a.view([2, 6]).shape) == (2, 6) and a.slice([1, 1], [2, 2]).numpy().tolist() == [[5.0, 6.0], [9.0, 10.0]]
    c = a.copy()
    c.numpy()[0, 0] = 100.0
    assert a.numpy()[0, 0] == 0.0
    b = ht.buffer_to_NDArray(np.arange(6, dtype=np.int64).tobytes(), "int64", [2, 3])
    assert b.numpy().tolist() == [[0, 1, 2], [3, 4, 5]]
