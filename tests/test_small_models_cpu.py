"""Small end-to-end models against PyTorch (the reference's tests/test_simple_model.py / test_rnn.py / test_model.py): a CNN
classifier and an unrolled RNN built from graph ops train with SGD and must follow torch's loss curve step by step."""
import numpy as np
import torch

import hetu_b200 as ht

rng = np.random.RandomState(0)


def test_cnn_classifier_follows_the_torch_loss_curve():
    x = rng.randn(8, 3, 12, 12).astype(np.float32)
    y = rng.randint(0, 5, 8)
    w1, b1 = (rng.randn(6, 3, 3, 3) * 0.2).astype(np.float32), np.zeros(6, np.float32)
    w2, b2 = (rng.randn(8, 6, 3, 3) * 0.2).astype(np.float32), np.zeros(8, np.float32)
    wf, bf = (rng.randn(5, 8 * 3 * 3) * 0.1).astype(np.float32), np.zeros(5, np.float32)
    with ht.graph("define_and_run", create_new=True) as g:
        X = ht.placeholder("float32", [8, 3, 12, 12], name="x")
        Y = ht.placeholder("int64", [8], name="y")
        P = {n: ht.parameter(ht.provided_initializer(v), list(v.shape), requires_grad=True, name=f"cnn_{n}")
             for n, v in (("w1", w1), ("b1", b1), ("w2", w2), ("b2", b2), ("wf", wf), ("bf", bf))}
        h = ht.maxpool(ht.relu(ht.conv2d(X, P["w1"], P["b1"], padding=1)), 2, 2, 0, 2)          # [8, 6, 6, 6]
        h = ht.avgpool(ht.relu(ht.conv2d(h, P["w2"], P["b2"], padding=1)), 2, 2, 0, 2)          # [8, 8, 3, 3]
        logits = ht.linear(ht.reshape(h, [8, 72]), P["wf"], P["bf"], trans_b=True)
        loss = ht.softmax_cross_entropy_sparse(logits, Y, reduction="mean")
        train = ht.SGDOptimizer(lr=0.1).minimize(loss)
    tp = {n: torch.tensor(v, requires_grad=True) for n, v in (("w1", w1), ("b1", b1), ("w2", w2), ("b2", b2), ("wf", wf), ("bf", bf))}
    opt = torch.optim.SGD(tp.values(), lr=0.1)
    F = torch.nn.functional
    ours, ref = [], []
    for _ in range(6):
        ours.append(float(g.run(loss, [loss, train], {X: torch.tensor(x), Y: torch.tensor(y)})[0]))
        th = F.max_pool2d(F.relu(F.conv2d(torch.tensor(x), tp["w1"], tp["b1"], padding=1)), 2)
        th = F.avg_pool2d(F.relu(F.conv2d(th, tp["w2"], tp["b2"], padding=1)), 2)
        tl = F.cross_entropy(F.linear(th.reshape(8, 72), tp["wf"], tp["bf"]), torch.tensor(y))
        opt.zero_grad(); tl.backward(); opt.step()
        ref.append(float(tl))
    assert ours[-1] < ours[0]
    np.testing.assert_allclose(ours, ref, rtol=2e-4, atol=1e-5)
    np.testing.assert_allclose(g.get_param(P["w1"]).numpy(), tp["w1"].detach().numpy(), rtol=1e-3, atol=1e-5)


def test_unrolled_rnn_follows_the_torch_loss_curve():
    T, B, D, H = 5, 4, 6, 8
    xs = rng.randn(T, B, D).astype(np.float32)
    tgt = rng.randn(B, 3).astype(np.float32)
    wx, wh, bh = (rng.randn(D, H) * 0.3).astype(np.float32), (rng.randn(H, H) * 0.3).astype(np.float32), np.zeros(H, np.float32)
    wo = (rng.randn(H, 3) * 0.3).astype(np.float32)
    with ht.graph("define_and_run", create_new=True) as g:
        X = [ht.placeholder("float32", [B, D], name=f"x{t}") for t in range(T)]
        Y = ht.placeholder("float32", [B, 3], name="y")
        P = {n: ht.parameter(ht.provided_initializer(v), list(v.shape), requires_grad=True, name=f"rnn_{n}")
             for n, v in (("wx", wx), ("wh", wh), ("bh", bh), ("wo", wo))}
        h = ht.tanh(ht.matmul(X[0], P["wx"]) + P["bh"])
        for t in range(1, T):                                   # the same weights are used at every time step
            h = ht.tanh(ht.matmul(X[t], P["wx"]) + ht.matmul(h, P["wh"]) + P["bh"])
        loss = ht.mean(ht.mse_loss(ht.matmul(h, P["wo"]), Y, reduction="none"))
        train = ht.SGDOptimizer(lr=0.2).minimize(loss)
    tp = {n: torch.tensor(v, requires_grad=True) for n, v in (("wx", wx), ("wh", wh), ("bh", bh), ("wo", wo))}
    opt = torch.optim.SGD(tp.values(), lr=0.2)
    feed = {X[t]: torch.tensor(xs[t]) for t in range(T)}
    feed[Y] = torch.tensor(tgt)
    ours, ref = [], []
    for _ in range(6):
        ours.append(float(g.run(loss, [loss, train], feed)[0]))
        th = torch.tanh(torch.tensor(xs[0]) @ tp["wx"] + tp["bh"])
        for t in range(1, T):
            th = torch.tanh(torch.tensor(xs[t]) @ tp["wx"] + th @ tp["wh"] + tp["bh"])
        tl = ((th @ tp["wo"] - torch.tensor(tgt)) ** 2).mean()
        opt.zero_grad(); tl.backward(); opt.step()
        ref.append(float(tl))
    assert ours[-1] < ours[0]
    np.testing.assert_allclose(ours, ref, rtol=2e-4, atol=1e-6)
    np.testing.assert_allclose(g.get_param(P["wh"]).numpy(), tp["wh"].detach().numpy(), rtol=1e-3, atol=1e-5)


def test_eager_mode_mlp_matches_define_and_run():
    """the same MLP step in the eager graph (ops run immediately, .backward()) and in a define-and-run graph"""
    x, w, b = rng.randn(5, 7).astype(np.float32), (rng.randn(3, 7) * 0.4).astype(np.float32), rng.randn(3).astype(np.float32)
    W, Bi = ht.from_numpy(w, requires_grad=True), ht.from_numpy(b, requires_grad=True)
    out = ht.sum(ht.sigmoid(ht.linear(ht.from_numpy(x), W, Bi, trans_b=True)))
    out.backward()
    with ht.graph("define_and_run", create_new=True) as g:
        px = ht.placeholder("float32", [5, 7], name="x")
        pw = ht.parameter(ht.provided_initializer(w), [3, 7], requires_grad=True, name="em_w")
        pb = ht.parameter(ht.provided_initializer(b), [3], requires_grad=True, name="em_b")
        lo = ht.sum(ht.sigmoid(ht.linear(px, pw, pb, trans_b=True)))
        gw, gb = ht.gradients(lo, [pw, pb])
        vals = g.run(lo, [lo, gw, gb], {px: torch.tensor(x)})
    assert abs(float(vals[0]) - float(out.numpy())) < 1e-5
    np.testing.assert_allclose(vals[1].numpy(), W.grad.numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(vals[2].numpy(), Bi.grad.numpy(), rtol=1e-5, atol=1e-6)
