"""v1 legacy surface: static executor API, parameter server (C++), HET cache table, search strategies."""
import threading

import numpy as np
import pytest
import torch

import hetu_b200 as ht
from hetu_b200 import v1


def test_v1_executor_trains_logistic_regression():
    from hetu_b200.v1 import executor as ex
    ex.reset_graph()
    rng = np.random.RandomState(0)
    X = rng.randn(256, 8).astype(np.float32)
    wtrue = rng.randn(8, 1).astype(np.float32)
    Y = (X @ wtrue > 0).astype(np.float32)
    x = v1.placeholder_op("x", shape=[256, 8])
    y = v1.placeholder_op("y", shape=[256, 1])
    W = v1.Variable("W", value=np.zeros((8, 1), np.float32))
    b = v1.Variable("b", value=np.zeros((1,), np.float32))
    p = v1.sigmoid_op(v1.matmul_op(x, W) + b)
    loss = v1.reduce_mean_op(v1.binarycrossentropy_op(p, y), [0, 1])
    train = v1.AdamOptimizer(learning_rate=0.1).minimize(loss)
    exe = v1.Executor([loss, train], ctx=v1.cpu(0))
    losses = [float(exe.run(feed_dict={x: X, y: Y})[0].asnumpy()) for _ in range(40)]
    assert losses[-1] < 0.35 * losses[0]


def test_parameter_server_dense_sparse_ssp_preduce():
    ps0, ps1 = v1.PSContext(2, 0, "t1"), v1.PSContext(2, 1, "t1")
    assert ps0.server is ps1.server
    ps0.init_dense("w", np.ones(4), opt="sgd", lr=0.5)
    ps1.init_dense("w", np.zeros(4), opt="sgd", lr=0.5)        # first initialiser wins
    ps0.push("w", np.array([1, 2, 3, 4.0]))
    np.testing.assert_allclose(ps1.pull("w"), [0.5, 0.0, -0.5, -1.0])
    ps0.init_sparse("emb", np.zeros((10, 2)), opt="sgd", lr=1.0)
    ps0.sparse_push("emb", [3, 3, 7], np.array([[1, 1], [1, 1], [2, 2.0]]))
    np.testing.assert_allclose(ps1.sparse_pull("emb", [3, 7, 0], 2), [[-2, -2], [-2, -2], [0, 0]])
    assert ps0.server.row_versions(ps0.key("emb"), [3, 7, 0]) == [2, 1, 0]
    # BSP barrier + SSP with staleness 1: the fast worker may run at most 1 clock ahead
    ps0.ssp_init(1)
    order = []

    def fast():
        for c in range(1, 4):
            ps0.ssp_sync(c)
            order.append(("fast", c))

    def slow():
        import time
        for c in range(1, 4):
            time.sleep(0.05)
            ps1.ssp_sync(c)
            order.append(("slow", c))
    ts = [threading.Thread(target=fast), threading.Thread(target=slow)]
    [t.start() for t in ts]
    [t.join(10) for t in ts]
    for i, (who, c) in enumerate(order):
        if who == "fast":
            done_slow = max([cc for w, cc in order[:i] if w == "slow"], default=0)
            assert c - done_slow <= 2          # entered clock c only once the slow worker reached c - 1
    # partial reduce: both workers arrive inside the window -> averaged together
    res = {}

    def pr(ctx, val):
        res[ctx.worker_id] = ctx.preduce("g", np.array(val, np.float32), min_workers=2, wait_ms=200)
    ts = [threading.Thread(target=pr, args=(ps0, [2.0, 4.0])), threading.Thread(target=pr, args=(ps1, [4.0, 8.0]))]
    [t.start() for t in ts]
    [t.join(10) for t in ts]
    np.testing.assert_allclose(res[0][0], [3.0, 6.0])
    assert sorted(res[0][1]) == [0, 1] and sorted(res[1][1]) == [0, 1]


def test_cache_sparse_table_bounded_staleness():
    ps = v1.PSContext(1, 0, "t2")
    table = np.arange(40, dtype=np.float32).reshape(20, 2)
    ps.init_sparse("emb", table, opt="sgd", lr=1.0)
    cst = v1.CacheSparseTable(ps, "emb", 20, 2, limit=8, policy="LRU", bound=2, lr=1.0)
    e = cst.embedding_lookup(np.array([[1, 2], [2, 5]]))
    assert tuple(e.shape) == (2, 2, 2) and torch.allclose(e[0, 1], torch.tensor([4.0, 5.0]))
    again = cst.embedding_lookup(np.array([1, 2]))
    st = cst.stats()
    assert st["hits"] >= 2 and torch.allclose(again[0], torch.tensor([2.0, 3.0]))
    for _ in range(4):           # updates beyond the push bound reach the server
        cst.embedding_update(np.array([1]), np.ones((1, 2), np.float32))
    assert ps.sparse_pull("emb", [1], 2)[0, 0] < 2.0


def test_search_strategies_are_feasible_and_no_worse_than_data_parallel():
    layers = v1.strategies.transformer_layers(4, 1024, 4096, 1024, 8)
    n = 8
    dp = v1.DataParallel(n)
    base = dp.total_time(layers, dp.assign(layers))
    for S in (v1.OptCNNSearching(n), v1.FlexFlowSearching(n, budget=300)):
        pl = S.assign(layers)
        assert len(pl) == len(layers) and S.feasible(layers, pl)
        assert S.total_time(layers, pl) <= base * 1.0001
    meg = v1.MegatronLM(n, tp=2)
    pl = meg.assign(layers)
    assert pl[1].split == {"batch": 4, "out": 2} and pl[3].split == {"batch": 4, "in": 2} and pl[4].split == {"batch": 4, "out": 2}
    g = v1.GPipeSearching(n, num_stages=4, micro_batches=8)
    pl = g.assign(layers)
    stages = sorted({tuple(p.devices) for p in pl})
    assert len(stages) == 4 and g.estimate(layers) > 0
    pd = v1.PipeDreamSearching(n).assign(layers)
    po = v1.PipeOptSearching(n).assign(layers)
    assert len(pd) == len(po) == len(layers)


def test_v1_layers_dataloaders_schedulers_and_metrics():
    """the legacy training loop: dataloader_op nodes fed per executor pass ('train' / 'validate'), layers building the graph,
    an lr scheduler object as the optimizer's learning rate, numpy metrics on the outputs"""
    import hetu_b200.v1 as v1
    v1.reset_graph()
    rng = np.random.RandomState(0)
    X = rng.randn(256, 16).astype(np.float32)
    wtrue = rng.randn(16, 3).astype(np.float32)
    Y = np.eye(3, dtype=np.float32)[(X @ wtrue).argmax(1)]
    x = v1.dataloader_op([v1.Dataloader(X[:192], 32, "train"), v1.Dataloader(X[192:], 32, "validate")])
    y = v1.dataloader_op([v1.Dataloader(Y[:192], 32, "train"), v1.Dataloader(Y[192:], 32, "validate")])
    net = v1.layers.Sequence(v1.layers.Linear(16, 32, activation="relu", name="fc1"), v1.layers.LayerNorm(32, name="ln"),
                              v1.layers.Linear(32, 3, name="fc2"))
    logits = net(x)
    loss = v1.layers.SoftmaxCrossEntropyLoss()(logits, y)
    sched = v1.lr_scheduler.StepScheduler(0.05, step_size=30, gamma=0.5)
    train = v1.AdamOptimizer(learning_rate=sched).minimize(loss)
    ex = v1.Executor({"train": [loss, train], "validate": [loss, logits, y]})
    assert ex.get_batch_num("train") == 6 and ex.get_batch_num("validate") == 2
    first = last = None
    for epoch in range(12):
        for _ in range(ex.get_batch_num("train")):
            lv = float(ex.run("train", convert_to_numpy_ret_vals=True)[0])
            first = lv if first is None else first
            last = lv
    assert last < 0.5 * first and sched.cnt == 72 and sched.get() == pytest.approx(0.05 * 0.25)
    accs = []
    for _ in range(ex.get_batch_num("validate")):
        _, lg, yy = ex.run("validate", convert_to_numpy_ret_vals=True)
        accs.append(v1.metrics.accuracy(yy, v1.metrics.softmax_func(lg)))
        assert v1.metrics.f_score_one_hot(yy, lg, average="macro") > 0.5
    assert np.mean(accs) > 0.8
    # metrics against closed forms
    lab = np.array([0, 0, 1, 1])
    assert v1.metrics.auc(lab, np.array([0.1, 0.4, 0.35, 0.8])) == pytest.approx(0.75, abs=0.02)
    assert v1.metrics.auc(lab, np.array([0.1, 0.2, 0.8, 0.9])) == pytest.approx(1.0, abs=0.02)
    oh = np.eye(3)[[0, 1, 2, 2]]
    pr = np.eye(3)[[0, 2, 2, 2]]
    assert v1.metrics.precision_score_one_hot(oh, pr).tolist() == pytest.approx([1.0, 0.0, 2 / 3])
    assert v1.metrics.recall_score_one_hot(oh, pr, average="micro") == pytest.approx(0.75)
    # initializers: factories and direct constructors
    w = v1.init.he_normal((64, 32), name="w_he")
    assert tuple(w.shape) == (64, 32) and abs(float(np.std(w.numpy())) - np.sqrt(2.0 / 64)) < 0.03
    c = v1.init.GenConstant(2.5)((3,), name="c25")
    assert np.allclose(c.numpy(), 2.5)
    v1.reset_graph()
