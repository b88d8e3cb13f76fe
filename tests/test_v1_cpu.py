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


def test_onnx_export_import_round_trip(tmp_path):
    """hetu2onnx writes a real ONNX protobuf (wire format checked field by field), onnx2hetu rebuilds an equivalent graph:
    a CNN + MLP head + embedding branch gives identical outputs after the round trip"""
    import hetu_b200.v1 as v1
    from hetu_b200.v1.onnx import proto as P
    v1.reset_graph()
    rng = np.random.RandomState(0)
    x = v1.placeholder_op("image", [2, 3, 8, 8])
    ids = v1.placeholder_op("ids", [2], dtype="int64")
    h = v1.layers.Conv2d(3, 4, 3, padding=1, activation="relu", name="c1")(x)
    h = v1.layers.BatchNorm(4, name="bn")(h)
    h = v1.layers.MaxPool2d(2)(h)
    h = v1.pad_op(h, [[0, 0], [0, 0], [1, 0], [0, 1]])
    h = v1.layers.AvgPool2d(5)(h)                         # [2, 4, 1, 1]
    f = v1.array_reshape_op(h, [2, 4])
    e = v1.layers.Embedding(10, 4, name="emb")(ids)
    z = v1.concat_op(f, e, axis=1)                        # [2, 8]
    z = v1.layers.Linear(8, 16, activation="gelu", name="fc1")(z)
    z = v1.layers.LayerNorm(16, name="ln")(z)
    z = v1.layers.Linear(16, 6, name="fc2", weight_transpose=True)(z)
    z = v1.slice_op(z, [0, 1], [2, 4]) * 0.5 + 1.0
    z = v1.div_op(v1.leaky_relu_op(z, 0.2), v1.sqrt_op(v1.exp_op(z)))
    probs = v1.softmax_op(z)
    score = v1.reduce_sum_op(v1.matmul_op(probs, probs, trans_B=True), [1])
    feed = {x: rng.randn(2, 3, 8, 8).astype(np.float32), ids: np.array([3, 7])}
    ex = v1.Executor([probs, score])
    want = ex.run(feed_dict=feed, convert_to_numpy_ret_vals=True)
    path = v1.onnx.export([probs, score], str(tmp_path / "model.onnx"))
    # the file is a well-formed ModelProto
    m = P.dec_model(open(path, "rb").read())
    assert m["ir_version"] == 8 and m["opset"][""] == 20 and m["producer"] == "hetu_b200"
    types = [n["op_type"] for n in m["graph"]["nodes"]]
    for t in ("Conv", "BatchNormalization", "MaxPool", "Pad", "AveragePool", "Reshape", "Gather", "Concat", "Gemm", "Gelu", "LayerNormalization",
              "Slice", "LeakyRelu", "Softmax", "MatMul", "ReduceSum"):
        assert t in types, t
    assert sorted(i["name"].split("_")[0] for i in m["graph"]["inputs"]) == ["ids", "image"]
    assert [i["shape"] for i in m["graph"]["inputs"] if i["name"].startswith("image")] == [[2, 3, 8, 8]]
    assert [o["shape"] for o in m["graph"]["outputs"]] == [[2, 4], [2]]
    conv = next(n for n in m["graph"]["nodes"] if n["op_type"] == "Conv")
    assert conv["attrs"]["pads"] == [1, 1, 1, 1] and conv["attrs"]["strides"] == [1, 1]
    assert any(a.shape == (4, 3, 3, 3) for a in m["graph"]["initializers"].values())
    # import into a fresh graph and compare
    v1.reset_graph()
    inputs, outs = v1.onnx.load(path)
    by = {k.split("_")[0]: v for k, v in inputs.items()}
    ex2 = v1.Executor(outs)
    got = ex2.run(feed_dict={by["image"]: feed[x], by["ids"]: feed[ids]}, convert_to_numpy_ret_vals=True)
    for a, b in zip(got, want):
        assert a.shape == b.shape and np.allclose(a, b, atol=1e-5), np.abs(a - b).max()
    v1.reset_graph()
    # codec details: negative ints, packed floats, scalar tensors
    node = P.dec_node(P.enc_node("X", ["a"], ["b"], "n", {"i": -3, "f": 0.25, "s": "str", "ints": [1, -2, 3], "floats": [0.5, 1.5], "t": np.arange(6, dtype=np.int64).reshape(2, 3)}))
    assert node["attrs"]["i"] == -3 and node["attrs"]["f"] == 0.25 and node["attrs"]["s"] == "str" and node["attrs"]["ints"] == [1, -2, 3]
    assert node["attrs"]["floats"] == [0.5, 1.5] and node["attrs"]["t"].tolist() == [[0, 1, 2], [3, 4, 5]]
    name, arr = P.dec_tensor(P.enc_tensor("s", np.float32(2.5)))
    assert name == "s" and arr.shape == () and float(arr) == 2.5


def test_parameter_server_over_the_network_with_heturun_launcher(tmp_path):
    """multi-process PS job: the launcher hosts the native parameter server, 3 worker processes connect over TCP and run
    BSP regression + HET-cached sparse updates + partial reduce + SSP clocks through it"""
    import os
    import sys
    from hetu_b200 import _C
    from hetu_b200.v1.launcher import launch
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    # in-process smoke of the transport first: errors travel back as exceptions, arrays round-trip exactly
    ps = _C.ParameterServer(1)
    srv = _C.PsNetServer(ps, 0, "127.0.0.1")
    cl = _C.PsNetClient("127.0.0.1", srv.port)
    cl.init_dense(7, [1.0, 2.0, 3.0], _C.PsOptimizer.SGD, 0.5)
    cl.push_dense(7, [2.0, 2.0, 2.0])
    assert cl.pull_dense(7) == [0.0, 1.0, 2.0] and ps.pull_dense(7) == [0.0, 1.0, 2.0] and cl.num_workers() == 1
    with pytest.raises(RuntimeError):
        cl.pull_dense(12345)
    assert cl.pull_dense(7) == [0.0, 1.0, 2.0]            # the connection survives a failed request
    assert srv.requests >= 6
    del cl
    srv.stop()
    logs = tmp_path / "logs"
    logs.mkdir()
    env_keep = dict(os.environ)
    os.environ.update({"PYTHONPATH": root, "HETU_B200_FORCE_CPU": "1", "CUDA_VISIBLE_DEVICES": "", "OMP_NUM_THREADS": "1"})
    try:
        codes = launch([sys.executable, os.path.join(root, "tests", "workers", "ps_net_worker.py")],
                       {"shared": {"DMLC_PS_ROOT_URI": "127.0.0.1"}, "launch": {"worker": 3, "server": 1, "scheduler": 1}}, log_dir=str(logs), timeout=240)
    finally:
        os.environ.clear()
        os.environ.update(env_keep)
    text = "\n".join((logs / f"worker{w}.log").read_text() for w in range(3))
    assert codes == [0, 0, 0], text
    lines = sorted(l for l in text.splitlines() if l.startswith("PSNET"))
    assert len(lines) == 3
    for w, l in enumerate(lines):
        assert f"worker={w} " in l and "preduce=[1.0, 1.0, 1.0, 1.0]" in l and "partners=[0, 1, 2]" in l
        assert float(l.split("err=")[1].split()[0]) < 0.05


def test_sharded_parameter_servers_behind_the_scheduler_with_heturun(tmp_path):
    """`server: 2` in the launch file: the launcher runs the native scheduler and two server roles, 3 worker processes register,
    shard the dense weight over the servers' key ranges and an embedding table over row % 2, synchronise through scheduler
    barriers and converge"""
    import os
    import sys
    from hetu_b200.v1.launcher import launch
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    logs = tmp_path / "logs"
    logs.mkdir()
    env_keep = dict(os.environ)
    os.environ.update({"PYTHONPATH": root, "HETU_B200_FORCE_CPU": "1", "CUDA_VISIBLE_DEVICES": "", "OMP_NUM_THREADS": "1"})
    try:
        codes = launch([sys.executable, os.path.join(root, "tests", "workers", "ps_sharded_worker.py")],
                       {"shared": {"DMLC_PS_ROOT_URI": "127.0.0.1"}, "launch": {"worker": 3, "server": 2, "scheduler": 1}}, log_dir=str(logs), timeout=240)
    finally:
        os.environ.clear()
        os.environ.update(env_keep)
    text = "\n".join((logs / f"worker{w}.log").read_text() for w in range(3))
    assert codes == [0, 0, 0], text
    lines = sorted(l for l in text.splitlines() if l.startswith("PSSHARD"))
    assert len(lines) == 3, text
    for l in lines:
        assert float(l.split("err=")[1].split()[0]) < 0.05 and "emb7=3.0" in l and "own=1.0" in l and "dead=0" in l, l


def test_embedding_compression_trainer_schedules_and_rate_sizing():
    """ref: tools/EmbeddingMemoryCompression run_compressed.py + methods/scheduler -- the harness sizes a method for a target rate,
    trains a CTR model on the planted-signal stream, runs the method's schedule and reports AUC + achieved compression"""
    from hetu_b200.tools.emb_compress.trainer import CompressionTrainer, SyntheticCTR, plan_for_rate
    # sizing: the planned constructor arguments land near the requested parameter budget
    from hetu_b200.tools.emb_compress import build_compressed_embedding
    import hetu_b200 as ht
    with ht.graph("define_and_run", create_new=True):
        for m in ("hash", "robe", "adapt", "dpq"):
            e = build_compressed_embedding(m, 8000, 16, **plan_for_rate(m, 8000, 16, 0.1))
            assert 5.0 <= e.compression_ratio() <= 16.0, (m, e.compression_ratio())
    data = SyntheticCTR(2000, 4, 3)
    f = data.frequency(20000)
    assert f.sum() > 0 and f[:500].sum() > 0                     # every field contributes ids; the head of each field dominates
    common = dict(num_embeddings=2000, dim=8, num_fields=4, num_dense=3, batch_size=128, compress_rate=0.25, lr=0.02)
    full = CompressionTrainer(None, "wdl", **common).run(steps=50, eval_batches=3)
    assert full["auc"] > 0.62 and full["stage1_loss"][1] < full["stage1_loss"][0] and full["ratio"] == 1.0
    hashed = CompressionTrainer("hash", "wdl", **common).run(steps=50, eval_batches=3)
    assert 3.5 <= hashed["ratio"] <= 4.5 and 0.5 < hashed["auc"] <= full["auc"] + 0.05
    # DeepLight: the mask refresh prunes towards the target while training continues
    dl = CompressionTrainer("deeplight", "deepfm", **common)
    r = dl.run(steps=40, eval_batches=2)
    assert 0.0 < r["schedule"]["sparsity"] <= 0.75 and r["effective_ratio"] > 1.0
    w, mask = dl.value_of(dl.embedding.weight), dl.value_of(dl.embedding.mask)
    assert abs(float((mask == 0).mean()) - r["schedule"]["sparsity"]) < 1e-6
    # AutoDim: search stage picks a width, stage 2 retrains a table of that width
    ad = CompressionTrainer("autodim", "deepfm", **common)
    r = ad.run(steps=25, eval_batches=2)
    assert r["schedule"]["selected_dim"] in (2, 4, 8) and "stage1" in r and "stage2_loss" in r
    assert ad.embedding.dim == r["schedule"]["selected_dim"]
    # OptEmbed: supernet with sampled widths, then the width search
    oe = CompressionTrainer("optembed", "deepfm", **common).run(steps=25, eval_batches=2)
    assert oe["schedule"]["searched_dim"] in (2, 4, 6, 8) and len(oe["schedule"]["candidate_auc"]) >= 3


def test_v1_executor_parameter_server_and_hybrid_comm_modes(tmp_path):
    """`Executor(..., comm_mode='PS')`: variables live on the server, the step fetches gradients, the server applies the optimizer
    (3 worker processes under heturun end with identical weights and a falling loss); `comm_mode='Hybrid'`: embedding tables
    (is_embed) go through sparse push / pull, dense variables keep the local optimizer"""
    import os
    import sys
    from hetu_b200.v1.launcher import launch
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    logs = tmp_path / "logs"
    logs.mkdir()
    env_keep = dict(os.environ)
    os.environ.update({"PYTHONPATH": root, "HETU_B200_FORCE_CPU": "1", "CUDA_VISIBLE_DEVICES": "", "OMP_NUM_THREADS": "1"})
    try:
        codes = launch([sys.executable, os.path.join(root, "tests", "workers", "v1_ps_executor_worker.py")],
                       {"shared": {"DMLC_PS_ROOT_URI": "127.0.0.1"}, "launch": {"worker": 3, "server": 1}}, log_dir=str(logs), timeout=240)
    finally:
        os.environ.clear()
        os.environ.update(env_keep)
    text = "\n".join((logs / f"worker{w}.log").read_text() for w in range(3))
    assert codes == [0, 0, 0], text
    lines = sorted(l for l in text.splitlines() if l.startswith("V1PS"))
    assert len(lines) == 3, text
    sums = {l.split("wsum=")[1].split()[0] for l in lines}
    assert len(sums) == 1, lines                                  # BSP through the server: every worker holds the same weights
    for l in lines:
        first, last = float(l.split("first=")[1].split()[0]), float(l.split("last=")[1].split()[0])
        assert last < 0.6 * first, l

    # Hybrid in one process: the embedding table is served by the (in-process) parameter server, the dense layer trains locally
    import numpy as np
    import hetu_b200.v1 as v1
    from hetu_b200.v1 import executor as v1ex
    v1ex.reset_graph()
    rng = np.random.RandomState(0)
    ids = v1.placeholder_op("ids", [16], dtype="int64")
    tgt = v1.placeholder_op("tgt", [16, 1])
    table = v1.Variable("emb_table", value=rng.randn(50, 4).astype(np.float32) * 0.1, is_embed=True)
    wout = v1.Variable("w_out", value=rng.randn(4, 1).astype(np.float32) * 0.1)
    pred = v1.matmul_op(v1.embedding_lookup_op(table, ids), wout)
    loss = v1.reduce_mean_op(v1.mse_op(pred, tgt), [0, 1])
    train = v1.SGDOptimizer(learning_rate=0.2).minimize(loss)
    ps = v1.PSContext(1, 0, "hybrid_test")
    ex = v1.Executor([loss, train], comm_mode="Hybrid", ps=ps)
    t0 = ex.graph.get_param(table).clone()
    per_id = rng.randn(50).astype(np.float32)
    losses = []
    for step in range(80):
        i = rng.randint(0, 20, 16)                               # only ids < 20 ever appear
        out = ex.run(feed_dict={ids: i, tgt: per_id[i].reshape(16, 1)})
        losses.append(float(out[0].asnumpy()))
    t1 = ex.graph.get_param(table)
    assert losses[-1] < 0.3 * losses[0]
    assert not np.allclose(t1[:20].numpy(), t0[:20].numpy()) and np.array_equal(t1[20:].numpy(), t0[20:].numpy())
    served = ps.sparse_pull("emb_table", list(range(50)), 4)
    np.testing.assert_allclose(served, t1.numpy(), rtol=1e-5, atol=1e-6)     # the server's table is the worker's table
    v1ex.reset_graph()


def test_v1_strategies_emit_executable_ds_parallel_configs():
    """ref: hetu/v1 distributed_strategies -- the fixed and searching strategies run on a GPT layer graph and their placement
    becomes the ds_parallel_config of the DistributedStates executor (same dict the hand-written generator produces)"""
    from hetu_b200.models import generate_ds_parallel_config
    from hetu_b200.v1 import strategies as S
    L, H, F, SEQ, B = 8, 1024, 4096, 1024, 16
    layers = S.transformer_layers(L, H, F, SEQ, B)
    summ = S.summarize_placements(layers, S.MegatronLM(8, tp=4).assign(layers))
    assert (summ["pp"], summ["tp"], summ["dp"]) == (1, 4, 2)
    cfg = S.strategy_to_ds_parallel_config(S.MegatronLM(8, tp=4), L, H, F, SEQ, B)
    ref = generate_ds_parallel_config(L, 8, 2, 4, 1)
    assert cfg["blocks"] == ref["blocks"] and cfg["wte"] == ref["wte"] and cfg["searched_by"] == "MegatronLM" and cfg["estimated_step_s"] > 0
    cfg = S.strategy_to_ds_parallel_config(S.DataParallel(4), L, H, F, SEQ, B)
    assert (cfg["dp"], cfg["tp"], cfg["pp"]) == (4, 1, 1)
    g = S.GPipeSearching(8, num_stages=4, micro_batches=8)
    cfg = S.strategy_to_ds_parallel_config(g, L, H, F, SEQ, B)
    assert (cfg["dp"], cfg["tp"], cfg["pp"]) == (2, 1, 4)
    ranges = sorted(b["range"] for b in cfg["blocks"].values())
    assert ranges[0][0] == 0 and ranges[-1][1] == L - 1 and all(a[1] + 1 == b[0] for a, b in zip(ranges, ranges[1:]))
    # the searching strategies return something the generator accepts, and never worse than plain data parallelism under their model
    for strat in (S.FlexFlowSearching(8, budget=300), S.OptCNNSearching(8), S.PipeDreamSearching(8), S.PipeOptSearching(8)):
        cfg = S.strategy_to_ds_parallel_config(strat, L, H, F, SEQ, B)
        assert cfg["dp"] * cfg["tp"] * cfg["pp"] == 8 and len(cfg["devices"]) == 8 and cfg["estimated_step_s"] > 0


def test_v1_long_tail_op_constructors_match_numpy():
    """ref: hetu/v1/python/hetu/gpu_ops -- the `*_op` constructors beyond the common set, checked against numpy / torch"""
    import numpy as np
    import torch
    import hetu_b200.v1 as v1
    from hetu_b200.v1 import executor as v1ex
    v1ex.reset_graph()
    rng = np.random.RandomState(0)
    a, b, c = (rng.randn(4, 5).astype(np.float32) for _ in range(3))
    m1, m2 = rng.randn(4, 3).astype(np.float32), rng.randn(3, 5).astype(np.float32)
    A, B, C = (v1.Variable(n, value=v, trainable=False) for n, v in (("lt_a", a), ("lt_b", b), ("lt_c", c)))
    M1, M2 = v1.Variable("lt_m1", value=m1, trainable=False), v1.Variable("lt_m2", value=m2, trainable=False)
    pos = v1.Variable("lt_pos", value=np.abs(a) + 0.5, trainable=False)
    ids = v1.Variable("lt_ids", value=np.array([7, 12, 5, 30], np.int64), dtype="int64", trainable=False)
    nodes = {
        "addmm": (v1.addmm_op(C, M1, M2, alpha=0.5, beta=2.0), 0.5 * m1 @ m2 + 2.0 * c),
        "minus_byconst": (v1.minus_byconst_op(A, 3.0), 3.0 - a), "div_const": (v1.div_const_op(2.0, pos), 2.0 / (np.abs(a) + 0.5)),
        "const_pow": (v1.const_pow_op(A, 2.0), 2.0 ** a), "power": (v1.power_op(pos, 1.5), (np.abs(a) + 0.5) ** 1.5),
        "clamp": (v1.clamp_op(A, mmin=-0.3, mmax=0.4), np.clip(a, -0.3, 0.4)), "sign": (v1.sign_op(A), np.sign(a)), "bool": (v1.bool_op(A), (a != 0).astype(np.float32)),
        "max": (v1.max_op(A, B), np.maximum(a, b)), "min": (v1.min_op(A, B), np.minimum(a, b)),
        "reduce_mul": (v1.reduce_mul_op(pos, [1]), np.prod(np.abs(a) + 0.5, 1)), "reduce_norm1": (v1.reduce_norm1_op(A, [1]), np.abs(a).sum(1)),
        "reduce_norm2": (v1.reduce_norm2_op(A, [0]), np.sqrt((a * a).sum(0))), "reducesumaxiszero": (v1.reducesumaxiszero_op(A), a.sum(0)),
        "argmax": (v1.argmax_op(A, 1), a.argmax(1)), "argsort": (v1.argsort_op(A, 1, descending=True), np.argsort(-a, 1, kind="stable")),
        "topk_val": (v1.topk_val_op(A, 2, 1), -np.sort(-a, 1)[:, :2]), "topk_idx": (v1.topk_idx_op(A, 2, 1), np.argsort(-a, 1, kind="stable")[:, :2]),
        "cumsum": (v1.cumsum_with_bias_op(A, 1.0, dim=1), np.cumsum(a, 1) + 1.0), "tile": (v1.tile_op(A, [2, 1]), np.tile(a, (2, 1))),
        "roll": (v1.roll_op(A, 2, 1), np.roll(a, 2, 1)), "where_const": (v1.where_const_op(v1.bool_op(v1.relu_op(A)), A, -1.0), np.where(a > 0, a, -1.0)),
        "slice_assign": (v1.slice_assign_op(A, [1, 1], [2, 3], 9.0), (lambda t: (t.__setitem__((slice(1, 3), slice(1, 4)), 9.0), t)[1])(a.copy())),
        "log_softmax": (v1.log_softmax_op(A), torch.log_softmax(torch.tensor(a), -1).numpy()),
        "crossentropy": (v1.crossentropy_op(v1.softmax_op(A), v1.softmax_op(B)), -(torch.softmax(torch.tensor(b), -1) * torch.log_softmax(torch.tensor(a), -1)).sum(-1).numpy()),
        "bce_logits": (v1.binarycrossentropywithlogits_op(A, v1.bool_op(v1.relu_op(B))),
                       torch.nn.functional.binary_cross_entropy_with_logits(torch.tensor(a), torch.tensor((b > 0).astype(np.float32)), reduction="none").numpy()),
        "mod_hash": (v1.mod_hash_op(ids, 5), np.array([2, 2, 0, 0])), "div_hash": (v1.div_hash_op(ids, 5), np.array([1, 2, 1, 6])),
        "conv2d_reducesum": (v1.conv2d_reducesum_op(v1.array_reshape_op(A, [2, 2, 5, 1])), a.reshape(2, 2, 5, 1).sum((0, 2, 3))),
        "tril_lookup": (v1.tril_lookup_op(v1.array_reshape_op(v1.slice_op(A, [0, 0], [4, 4]), [1, 4, 4])), a[:, :4][np.tril_indices(4)].reshape(1, -1)),
        "min_dist": (v1.min_dist_op(A, B), ((a[:, None, :] - b[None]) ** 2).sum(-1).argmin(1)),
        "prune": (v1.prune_low_magnitude_op(A, 0.5), a * (np.abs(a) > np.sort(np.abs(a).reshape(-1))[9])),
    }
    names = list(nodes)
    ex = v1.Executor([nodes[n][0] for n in names])
    outs = ex.run(feed_dict={}, convert_to_numpy_ret_vals=True)
    for n, got in zip(names, outs):
        np.testing.assert_allclose(np.asarray(got, dtype=np.float64), np.asarray(nodes[n][1], dtype=np.float64), rtol=2e-5, atol=2e-5, err_msg=n)
    s1, s2 = v1.normal_sample_op([1000], 1.0, 2.0), v1.uniform_sample_op([1000], -1.0, 1.0)
    r = v1.Executor([s1, s2, v1.randint_sample_op([50], 3, 9), v1.gumbel_sample_op([10])]).run(feed_dict={}, convert_to_numpy_ret_vals=True)
    assert abs(r[0].mean() - 1.0) < 0.2 and abs(r[0].std() - 2.0) < 0.2 and -1.0 <= r[1].min() and r[1].max() <= 1.0 and 3 <= r[2].min() and r[2].max() < 9
    v1ex.reset_graph()


def test_v1_explicit_gradient_node_constructors_match_torch_autograd():
    """ref: hetu/v1/python/hetu/gpu_ops/{Relu,Conv2d,MaxPool,BatchNorm,Pad,Slice,Concat,...}.py -- `*_gradient_op(...)` built by hand
    give the same values as torch autograd of the forward"""
    import numpy as np
    import torch
    import torch.nn.functional as F
    import hetu_b200.v1 as v1
    from hetu_b200.v1 import executor as v1ex
    v1ex.reset_graph()
    rng = np.random.RandomState(1)

    def var(name, val, dtype=None):
        return v1.Variable("gn_" + name, value=val, trainable=False, **({"dtype": dtype} if dtype else {}))

    def tgrad(fn, x, dy):
        t = torch.tensor(x, requires_grad=True)
        y = fn(t)
        return torch.autograd.grad(y, t, torch.tensor(dy).reshape(y.shape))[0].numpy()

    x, dy = rng.randn(4, 6).astype(np.float32), rng.randn(4, 6).astype(np.float32)
    X, DY = var("x", x), var("dy", dy)
    pos = np.abs(x) + 0.5
    POS = var("pos", pos)
    img, w = rng.randn(2, 3, 8, 8).astype(np.float32), rng.randn(4, 3, 3, 3).astype(np.float32)
    dconv = rng.randn(2, 4, 8, 8).astype(np.float32)
    IMG, W, DCONV = var("img", img), var("w", w), var("dconv", dconv)
    dpool = rng.randn(2, 3, 4, 4).astype(np.float32)
    DPOOL = var("dpool", dpool)
    scale = rng.rand(3).astype(np.float32) + 0.5
    SCALE = var("scale", scale)
    dimg = rng.randn(2, 3, 8, 8).astype(np.float32)
    DIMG = var("dimg", dimg)
    idx = rng.randint(0, 6, (4, 3)).astype(np.int64)
    IDX = var("idx", idx, "int64")
    dg = rng.randn(4, 3).astype(np.float32)
    DG = var("dg", dg)
    tgt = rng.randint(0, 6, 4).astype(np.int64)
    TGT = var("tgt", tgt, "int64")
    bn = v1.batch_normalization_gradient_op(DIMG, IMG, SCALE, None, 1e-5)

    def bn_ref(which):
        t, s, b = torch.tensor(img, requires_grad=True), torch.tensor(scale, requires_grad=True), torch.zeros(3, requires_grad=True)
        y = F.batch_norm(t, None, None, s, b, True, 0.1, 1e-5)
        return torch.autograd.grad(y, [t, s, b], torch.tensor(dimg))[which].numpy()

    fwd_tanh, fwd_sm, fwd_lsm = np.tanh(x), torch.softmax(torch.tensor(x), -1).numpy(), torch.log_softmax(torch.tensor(x), -1).numpy()
    nodes = {
        "relu": (v1.relu_gradient_op(X, DY), dy * (x > 0)),
        "leaky_relu": (v1.leaky_relu_gradient_op(X, DY, 0.1), dy * np.where(x > 0, 1.0, 0.1)),
        "gelu": (v1.gelu_gradient_op(X, DY), tgrad(F.gelu, x, dy)),
        "tanh": (v1.tanh_gradient_op(var("tanh", fwd_tanh), DY), dy * (1 - fwd_tanh ** 2)),
        "abs": (v1.abs_gradient_op(DY, X), dy * np.sign(x)),
        "log": (v1.log_grad_op(DY, POS, 0.0), dy / pos),
        "pow": (v1.pow_gradient_op(POS, DY, 2.5), tgrad(lambda t: t ** 2.5, pos, dy)),
        "const_pow": (v1.const_pow_gradient_op(X, DY, 3.0), tgrad(lambda t: 3.0 ** t, x, dy)),
        "softmax": (v1.softmax_gradient_op(var("sm", fwd_sm), DY), tgrad(lambda t: torch.softmax(t, -1), x, dy)),
        "log_softmax": (v1.log_softmax_gradient_op(var("lsm", fwd_lsm), DY), tgrad(lambda t: torch.log_softmax(t, -1), x, dy)),
        "bce_logits": (v1.binarycrossentropywithlogits_gradient_op(X, var("lab", (pos > 1).astype(np.float32)), DY),
                       tgrad(lambda t: F.binary_cross_entropy_with_logits(t, torch.tensor((pos > 1).astype(np.float32)), reduction="none"), x, dy)),
        "nll": (v1.nll_loss_grad_op(var("one", np.ones(1, np.float32)), TGT, 6), tgrad(lambda t: F.nll_loss(t, torch.tensor(tgt)), x, np.ones((), np.float32))),
        "norm": (v1.norm_gradient_op(X, None, var("dn", dy[:, 0].copy()), 1, 2), tgrad(lambda t: t.norm(2, 1), x, dy[:, 0].copy())),
        "addmm_bias": (v1.addmm_gradient_op(var("bias", np.zeros((1, 6), np.float32)), DY, 2.0), 2.0 * dy.sum(0, keepdims=True)),
        "reshape": (v1.array_reshape_gradient_op(X, var("dflat", dy.reshape(-1))), dy),
        "repeat": (v1.repeat_gradient_op(X, var("drep", np.tile(dy, (3, 2)))), 6.0 * dy),
        "pad": (v1.pad_gradient_op(DY, [[1, 1], [2, 0]]), dy[1:3, 2:]),
        "slice": (v1.slice_gradient_op(var("dsl", dy[:2, :3].copy()), [1, 2], [4, 6]), np.pad(dy[:2, :3], [[1, 1], [2, 1]])),
        "concat0": (v1.concat_gradient_op(DY, var("c0", x[:, :2].copy()), 1, 0), dy[:, :2]),
        "concat1": (v1.concat_gradient_op(DY, var("c1", x[:, 2:].copy()), 1, 1), dy[:, 2:]),
        "concatenate": (v1.concatenate_gradient_op(DY, var("c2", x[:, 1:4].copy()), 1, offset=1), dy[:, 1:4]),
        "split": (v1.split_gradient_op(var("dsp", dy[:, 2:4].copy()), [1], [1], [3]), np.pad(dy[:, 2:4], [[0, 0], [2, 2]])),
        "gather": (v1.gather_gradient_op(X, DG, 1, IDX), tgrad(lambda t: t.gather(1, torch.tensor(idx)), x, dg)),
        "scatter1d": (v1.scatter1d_grad_op(var("d1", dy[:, 0].copy().reshape(4, 1)), var("i1", np.array([2, 0, 3], np.int64), "int64")), dy[[2, 0, 3], 0].reshape(3, 1)),
        "tril": (v1.tril_lookup_gradient_op(var("dtri", np.arange(1, 7, dtype=np.float32).reshape(1, 6))),
                 (lambda m: (m.__setitem__(np.tril_indices(3), np.arange(1, 7)), m)[1])(np.zeros((3, 3), np.float32)).reshape(1, 3, 3)),
        "interpolate": (v1.interpolate_grad_op(var("dint", np.ones((2, 3, 16, 16), np.float32)), IMG, "bilinear", False),
                        tgrad(lambda t: F.interpolate(t, size=(16, 16), mode="bilinear", align_corners=False), img, np.ones((2, 3, 16, 16), np.float32))),
        "conv_data": (v1.conv2d_gradient_of_data_op(W, DCONV, IMG, padding=1, stride=1), tgrad(lambda t: F.conv2d(t, torch.tensor(w), padding=1), img, dconv)),
        "conv_filter": (v1.conv2d_gradient_of_filter_op(IMG, DCONV, W, padding=1, stride=1), tgrad(lambda t: F.conv2d(torch.tensor(img), t, padding=1), w, dconv)),
        "max_pool": (v1.max_pool2d_gradient_op(None, DPOOL, IMG, 2, 2, 0, 2), tgrad(lambda t: F.max_pool2d(t, 2, 2), img, dpool)),
        "avg_pool": (v1.avg_pool2d_gradient_op(None, DPOOL, IMG, 2, 2, 0, 2), tgrad(lambda t: F.avg_pool2d(t, 2, 2), img, dpool)),
        "bn_data": (v1.batch_normalization_gradient_of_data_op(bn, IMG), bn_ref(0)),
        "bn_scale": (v1.batch_normalization_gradient_of_scale_op(bn, SCALE), bn_ref(1)),
        "bn_bias": (v1.batch_normalization_gradient_of_bias_op(bn, None), bn_ref(2)),
    }
    names = list(nodes)
    outs = v1.Executor([nodes[n][0] for n in names]).run(feed_dict={}, convert_to_numpy_ret_vals=True)
    for n, got in zip(names, outs):
        np.testing.assert_allclose(np.asarray(got, dtype=np.float64), np.asarray(nodes[n][1], dtype=np.float64), rtol=2e-4, atol=2e-5, err_msg=n)
    v1ex.reset_graph()


def test_v1_quantised_table_sparse_row_and_moe_gradient_constructors():
    """ref: hetu/v1/src/ops/{QuantizeEmbedding,SignedQuantize,SparseSet,AssignWithIndexedSlices,UniqueIndices}.cu and
    gpu_ops/{LayoutTransform,ReverseLayoutTransform}.py -- quantised lookups, ALPT rounding with its LSQ step gradient, row
    assignment, de-duplication and the hand-built MoE dispatch / combine gradients"""
    import numpy as np
    import torch
    import hetu_b200.v1 as v1
    from hetu_b200 import ops
    from hetu_b200.v1 import executor as v1ex
    v1ex.reset_graph()
    rng = np.random.RandomState(2)

    def var(name, val, dtype=None, trainable=False):
        return v1.Variable("qt_" + name, value=val, trainable=trainable, **({"dtype": dtype} if dtype else {}))

    q8 = rng.randint(0, 256, (10, 4)).astype(np.float32)
    qp = np.stack([rng.rand(10) * 0.1 + 0.01, rng.randn(10)], 1).astype(np.float32)
    ids = np.array([3, 7, 3, 9], np.int64)
    Q8, QP, IDS = var("q8", q8), var("qp", qp), var("ids", ids, "int64")
    step = (rng.rand(10, 1) * 0.05 + 0.01).astype(np.float32)
    STEP = var("step", step)
    s8 = rng.randint(-128, 128, (10, 4)).astype(np.float32)
    S8 = var("s8", s8)
    look = (rng.randn(4, 4) * 90).astype(np.float32)        # some values beyond the int8 range
    LOOK, ST4 = var("look", look), var("st4", step[ids])
    new_rows = rng.randn(4, 4).astype(np.float32)
    NEW = var("new", new_rows)
    uniq = np.array([3, 7, 9, -1], np.int64)
    UNIQ = var("uniq", uniq, "int64")
    table = rng.randn(10, 4).astype(np.float32)
    TABLE = var("table", table)
    inverse = np.array([0, 1, 0, 2], np.int64)
    INV = var("inv", inverse, "int64")
    rows = rng.randn(4, 4).astype(np.float32)
    ROWS = var("rows", rows)

    assigned = table.copy()
    assigned[[3, 7, 9]] = new_rows[:3]
    lsq = np.where(look >= 127, 127.0, np.where(look <= -128, -128.0, np.floor(look + 0.5) - look))
    rounded = np.clip(np.floor(np.clip(look, -128, 127) + 0.5), -128, 127) * step[ids] + 0.25
    dedup = np.zeros((4, 4), np.float32)
    for j in range(3):
        dedup[j] = rows[inverse == j].mean(0)
    dgrad = np.zeros((4, 4), np.float32)
    np.add.at(dgrad, inverse, rows)
    unified_ids = np.array([2, -1, 11, 5], np.int64)
    uni = np.where(((unified_ids >= 0) & (unified_ids < 10))[:, None], q8[np.clip(unified_ids, 0, 9)] * 0.02 - 1.5, 0.0)
    nodes = {
        "quantized_lookup": (v1.quantized_embedding_lookup_op(Q8, IDS, QP, 8), q8[ids] * qp[ids, :1] + qp[ids, 1:]),
        "unified_lookup": (v1.unified_quantized_embedding_lookup_op(Q8, var("uids", unified_ids, "int64"), 0.02, -1.5, 8), uni),
        "alpt_lookup": (v1.alpt_embedding_lookup_op(S8, IDS, STEP, 0.25, 8), s8[ids] * step[ids] + 0.25),
        "alpt_rounding": (v1.alpt_rounding_op(LOOK, ST4, 0.25, 8), rounded),
        "alpt_scale_grad": (v1.alpt_scale_gradient_op(LOOK, 8), lsq),
        "assign_rows": (v1.assign_with_indexedslices_op(TABLE, UNIQ, NEW), assigned),
        "sparse_set": (v1.sparse_set_op(TABLE, UNIQ, NEW), assigned),
        "assign_unified": (v1.assign_quantized_embedding_op(Q8, var("u3", np.array([1, 4], np.int64), "int64"), var("n2", np.array([[0.0, 0.5, 1.0, 5.2]] * 2, np.float32)), 8,
                                                            scale=0.02, minele=0.0),
                           (lambda t: (t.__setitem__([1, 4], np.array([0, 25, 50, 255], np.float32)), t)[1])(q8.copy())),
        "dedup_lookup": (v1.deduplicate_lookup_op(ROWS, (INV, None)), dedup),
        "dedup_grad": (v1.deduplicate_grad_op(ROWS, (INV, None)), dgrad),
        "sum_sparse": (v1.sum_sparse_gradient_op([10, 4], (IDS, ROWS), TABLE), (lambda t: (np.add.at(t, ids, rows), t)[1])(table.copy())),
        "slice_by_matrix": (v1.slice_by_matrix_op(var("cube", np.arange(24, dtype=np.float32).reshape(2, 3, 4)), var("i1", np.array([1, 0], np.int64), "int64"),
                                                  var("i2", np.array([2, 1], np.int64), "int64")), np.arange(24, dtype=np.float32).reshape(2, 3, 4)[[1, 0], [2, 1]]),
        "slice_assign_matrix": (v1.slice_assign_matrix_op(TABLE, ROWS, [2, 1], [2, 2], [1, 0], [2, 2]),
                                (lambda t: (t.__setitem__((slice(2, 4), slice(1, 3)), rows[1:3, 0:2]), t)[1])(table.copy())),
    }
    names = list(nodes)
    outs = v1.Executor([nodes[n][0] for n in names]).run(feed_dict={}, convert_to_numpy_ret_vals=True)
    for n, got in zip(names, outs):
        np.testing.assert_allclose(np.asarray(got, dtype=np.float64), np.asarray(nodes[n][1], dtype=np.float64), rtol=2e-5, atol=2e-5, err_msg=n)

    # ALPT: the looked-up row passes the gradient straight through, the step gets dy * lsq
    v1ex.reset_graph()
    L2, S2 = var("look2", look, trainable=True), var("st2", step[ids], trainable=True)
    y = v1.alpt_rounding_op(L2, S2, 0.25, 8)
    gl, gs = v1.gradients(v1.reduce_sum_op(y * var("dy2", rows), [0, 1]), [L2, S2])
    gl, gs = v1.Executor([gl, gs]).run(feed_dict={}, convert_to_numpy_ret_vals=True)
    np.testing.assert_allclose(gl, rows * step[ids], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(gs, (rows * lsq).sum(1, keepdims=True), rtol=1e-4, atol=1e-4)

    # MoE layout transforms: the hand-built gradient nodes equal autodiff of dispatch / combine
    v1ex.reset_graph()
    T, D, E, CAP = 6, 4, 3, 3
    x = rng.randn(T, D).astype(np.float32)
    expert = np.array([0, 2, 1, 0, 2, 0], np.int64)
    loc = np.array([0, 0, 0, 1, 1, 2], np.int64)
    gate = rng.rand(T).astype(np.float32)
    X, IDX, LOC, GATE = var("mx", x, trainable=True), var("me", expert, "int64"), var("ml", loc, "int64"), var("mg", gate, trainable=True)
    dslots = rng.randn(E * CAP, D).astype(np.float32)
    DS = var("mds", dslots)
    disp = v1.layout_transform_op(X, IDX, LOC, CAP, E)
    auto_dx = v1.gradients(v1.reduce_sum_op(disp * DS, [0, 1]), [X])[0]
    hand_dx = v1.layout_transform_gradient_op(DS, IDX, LOC, CAP)
    Y = var("my", dslots, trainable=True)
    dtok = rng.randn(T, D).astype(np.float32)
    DT = var("mdt", dtok)
    comb = v1.reverse_layout_transform_op(Y, IDX, LOC, GATE, CAP, E)
    auto_dy, auto_dg = v1.gradients(v1.reduce_sum_op(comb * DT, [0, 1]), [Y, GATE])
    hand_dy = v1.reverse_layout_transform_gradient_data_op(DT, IDX, LOC, GATE, CAP, E)
    hand_dg = v1.reverse_layout_transform_gradient_gate_op(DT, Y, IDX, LOC, CAP)
    hand_ng = v1.reverse_layout_transform_no_gate_gradient_op(DT, IDX, LOC, CAP, E)
    r = v1.Executor([auto_dx, hand_dx, auto_dy, hand_dy, auto_dg, hand_dg, hand_ng]).run(feed_dict={}, convert_to_numpy_ret_vals=True)
    np.testing.assert_allclose(r[1], r[0], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(r[3].reshape(r[2].shape), r[2], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(r[5].reshape(-1), r[4].reshape(-1), rtol=1e-5, atol=1e-6)
    slots = expert * CAP + loc
    want = np.zeros((E * CAP, D), np.float32)
    want[slots] = dtok
    np.testing.assert_allclose(r[6].reshape(E * CAP, D), want, rtol=1e-6, atol=1e-6)
    v1ex.reset_graph()


def test_v1_pipeline_send_and_receive_nodes_between_two_processes():
    """ref: hetu/v1/python/hetu/gpu_ops/{PipelineSend,PipelineReceive}.py -- a hand-placed two-stage graph: the activation travels
    through `pipeline_send_op` / `pipeline_receive_op` (graph ops over the runtime's P2P channel)"""
    import json
    import os
    from dist_utils import run_workers
    here = os.path.dirname(os.path.abspath(__file__))
    ok, outs = run_workers(os.path.join(here, "workers", "v1_pipeline_p2p_worker.py"), 2, timeout=120)
    assert ok, "\n-----\n".join(o[-3000:] for o in outs)
    line = next(l for o in outs for l in o.splitlines() if l.startswith("P2P "))
    r = json.loads(line[4:])
    assert r["err"] < 1e-5 and r["raw"] == [0.0, 1.0, 2.0, 3.0, 4.0, 5.0]
    c = json.loads(next(l for o in outs for l in o.splitlines() if l.startswith("COMM "))[5:])
    assert c == {"sum": [3.0] * 3, "gather": [[0.0, 0.0], [1.0, 1.0]], "out": [4.0] * 3, "bc": [8.0]}


def test_v1_role_functions_run_a_scheduler_two_servers_and_two_workers(tmp_path):
    """ref: hetu/v1/python/hetu/gpu_ops/executor.py:100-137 -- scheduler_init / server_init / worker_init driven by the DMLC_*
    environment (five processes, one program); the workers see both servers, BSP pushes land on every shard"""
    import json
    import os
    import subprocess
    import sys
    from dist_utils import free_port
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = os.path.join(root, "tests", "workers", "v1_roles_worker.py")
    base = dict(os.environ, PYTHONPATH=root, HETU_B200_FORCE_CPU="1", CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="1",
                DMLC_PS_ROOT_URI="127.0.0.1", DMLC_PS_ROOT_PORT=str(free_port()), DMLC_NUM_SERVER="2", DMLC_NUM_WORKER="2")
    for k in ("HETU_PS_SCHEDULER", "HETU_PS_ADDRESS"):
        base.pop(k, None)
    procs = []
    for i, role in enumerate(["scheduler", "server", "server", "worker", "worker"]):
        out = open(tmp_path / f"{role}{i}.log", "w")
        procs.append((role, i, subprocess.Popen([sys.executable, script], env=dict(base, DMLC_ROLE=role), stdout=out, stderr=subprocess.STDOUT)))
        if role == "scheduler":
            import time
            time.sleep(1.0)                     # the scheduler's port must be listening before the others dial it
    codes = []
    for role, i, p in procs:
        try:
            codes.append(p.wait(150))
        except subprocess.TimeoutExpired:
            p.kill()
            codes.append(-9)
    logs = {f"{role}{i}": (tmp_path / f"{role}{i}.log").read_text() for role, i, _ in procs}
    assert codes == [0] * 5, json.dumps(logs)[-4000:]
    res = [json.loads(l[6:]) for t in logs.values() for l in t.splitlines() if l.startswith("ROLES ")]
    assert len(res) == 2 and all(r["servers"] == 2 and r["same"] for r in res)
    # three steps, both workers' gradients (1 and 2) applied with lr 0.5 -> -4.5; rows 0 and 1 pushed once, row 5 by both workers
    for r in res:
        assert abs(r["w"] - (-4.5)) < 1e-6, r
        assert r["emb"][0] == [-1.0, -1.0] and r["emb"][1] == [-1.0, -1.0] and r["emb"][2][0] <= -1.0


def test_v1_remaining_layer_classes_train_and_reduce():
    """ref: hetu/v1/python/hetu/layers/{moe_layer,hash_layer,ktop1_layer,sam_layer,loss,concatenate,batch_split_layer}.py -- the MoE
    layer family trains under a v1 optimizer, the loss layers reduce as declared, the gates resolve from ht.layers"""
    import numpy as np
    import hetu_b200.v1 as v1
    from hetu_b200.v1 import executor as v1ex
    L = v1.layers
    rng = np.random.RandomState(0)
    xs, ys = rng.randn(16, 8).astype(np.float32), rng.randn(16, 8).astype(np.float32)
    for cls, top in ((L.MoELayer, 2), (L.KTop1Layer, 1), (L.HashLayer, 1), (L.SAMLayer, 2)):
        v1ex.reset_graph()
        x, y = v1.placeholder_op("x", [16, 8]), v1.placeholder_op("y", [16, 8])
        block = L.BatchSplitOnlyLayer(L.Sequence(cls(8, 16, 4, top=top, capacity_factor=2.0), L.Linear(8, 8, name=f"{cls.__name__}_head")))
        out = block(x)
        assert v1ex.annotation(out, "layer_constraint") == "BatchSplitOnlyLayer" and block.split_dims == (0,) and L.ReserveSplitLayer.split_dims == (0, 1)
        loss = v1.reduce_mean_op(L.MSELoss()(out, y) if False else L.MAELoss("mean")(out, y), [0])
        train = v1.optim.AdamOptimizer(0.02).minimize(loss)
        ex = v1.Executor([loss, train])
        hist = [float(ex.run(feed_dict={x: xs, y: ys}, convert_to_numpy_ret_vals=True)[0]) for _ in range(25)]
        assert hist[-1] < 0.85 * hist[0], (cls.__name__, hist[0], hist[-1])
    v1ex.reset_graph()
    a, b = v1.Variable("la", value=xs, trainable=False), v1.Variable("lb", value=(ys > 0).astype(np.float32), trainable=False)
    nodes = [L.MAELoss("sum")(a, b), L.MAELoss("none")(a, b), L.BCEWithLogitsLoss("mean")(a, b), L.Concatenate(1)(a, b), L.Concatenate(0)(a)]
    r = v1.Executor(nodes).run(feed_dict={}, convert_to_numpy_ret_vals=True)
    t = (ys > 0).astype(np.float32)
    np.testing.assert_allclose(r[0], np.abs(xs - t).sum(0), rtol=1e-5)
    np.testing.assert_allclose(r[1], np.abs(xs - t), rtol=1e-5)
    np.testing.assert_allclose(r[2], (np.maximum(xs, 0) - xs * t + np.log1p(np.exp(-np.abs(xs)))).mean(0), rtol=1e-4, atol=1e-5)
    assert r[3].shape == (16, 16) and r[4].shape == (16, 8)
    from hetu_b200.models import moe
    assert L.TopKGate is moe.TopKGate and L.BalanceAssignmentGate is moe.BalanceGate and L.SAMGate is moe.SAMGate and L.HashGate is moe.HashGate
    v1ex.reset_graph()


def test_v1_adagrad_amsgrad_adamw_and_lamb_follow_the_reference_update_rules():
    """ref: hetu/v1/python/hetu/optimizer.py:335-760 -- four steps of every rule on a quadratic match torch.optim (AdaGrad, AMSGrad,
    AdamW) and the written-out LAMB recurrence"""
    import numpy as np
    import torch
    import hetu_b200.v1 as v1
    from hetu_b200.v1 import executor as v1ex
    rng = np.random.RandomState(3)
    w0, a = rng.randn(5, 4).astype(np.float32), rng.randn(5, 4).astype(np.float32)

    def ours(make):
        v1ex.reset_graph()
        w = v1.Variable("ow", value=w0.copy())
        loss = v1.reduce_sum_op(v1.mul_op(v1.mul_op(w - v1.Variable("oa", value=a, trainable=False), w), w), [0, 1])     # sum (w - a) w^2
        train = make().minimize(loss)
        ex = v1.Executor([loss, train])
        for _ in range(4):
            ex.run(feed_dict={})
        return ex.graph.get_param(w).float().numpy().copy()

    def theirs(make):
        w = torch.tensor(w0.copy(), requires_grad=True)
        opt = make([w])
        for _ in range(4):
            opt.zero_grad()
            ((w - torch.tensor(a)) * w * w).sum().backward()
            opt.step()
        return w.detach().numpy()

    np.testing.assert_allclose(ours(lambda: v1.AdaGradOptimizer(0.1, eps=1e-10)), theirs(lambda p: torch.optim.Adagrad(p, lr=0.1, eps=1e-10)), rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(ours(lambda: v1.AdaGradOptimizer(0.1, initial_accumulator_value=0.5, eps=1e-10)),
                               theirs(lambda p: torch.optim.Adagrad(p, lr=0.1, initial_accumulator_value=0.5, eps=1e-10)), rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(ours(lambda: v1.AdamWOptimizer(0.05, epsilon=1e-8, weight_decay=0.1)),
                               theirs(lambda p: torch.optim.AdamW(p, lr=0.05, eps=1e-8, weight_decay=0.1)), rtol=2e-4, atol=2e-5)
    # AMSGrad: running maximum of the bias-corrected second moment
    def amsgrad_ref():
        w = w0.astype(np.float64).copy()
        m, v, vh = np.zeros_like(w), np.zeros_like(w), np.zeros_like(w)
        for t in range(1, 5):
            g = 3 * w * w - 2 * a * w
            m, v = 0.9 * m + 0.1 * g, 0.999 * v + 0.001 * g * g
            vh = np.maximum(vh, v / (1 - 0.999 ** t))
            w = w - 0.05 * (m / (1 - 0.9 ** t)) / (np.sqrt(vh) + 1e-7)
        return w
    np.testing.assert_allclose(ours(lambda: v1.AMSGradOptimizer(0.05)), amsgrad_ref(), rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(ours(lambda: v1.AdamOptimizer(0.05, amsgrad=True)), amsgrad_ref(), rtol=2e-4, atol=2e-5)

    def lamb_ref():
        w = w0.astype(np.float64).copy()
        m, v = np.zeros_like(w), np.zeros_like(w)
        for t in range(1, 5):
            g = 3 * w * w - 2 * a * w
            m, v = 0.9 * m + 0.1 * g, 0.999 * v + 0.001 * g * g
            upd = (m / (1 - 0.9 ** t)) / (np.sqrt(v / (1 - 0.999 ** t)) + 1e-7) + 0.01 * w
            w = w - 0.05 * (np.linalg.norm(w) / np.linalg.norm(upd)) * upd
        return w
    np.testing.assert_allclose(ours(lambda: v1.LambOptimizer(0.05, weight_decay=0.01)), lamb_ref(), rtol=2e-4, atol=2e-5)
    v1ex.reset_graph()


def test_v1_profiler_and_simulator_measure_cache_and_model_transfers(tmp_path):
    """ref: hetu/v1/python/hetu/profiler.py -- HetuProfiler times nodes (whole and per op), HetuSimulator caches measured node times on
    disk and prices transfers / collectives / re-sharding with its link model"""
    import numpy as np
    import hetu_b200.v1 as v1
    from hetu_b200.v1 import executor as v1ex
    v1ex.reset_graph()
    x = v1.placeholder_op("px", [64, 128])
    w = v1.Variable("pw", value=np.random.randn(128, 256).astype(np.float32) * 0.05)
    y = v1.relu_op(v1.matmul_op(x, w))
    prof = v1.HetuProfiler([y], {x: [64, 128]})
    assert prof.profile(5) > 0
    per_op = prof.profile_all(3)
    assert any(k.startswith("matmul") for k in per_op) and any(k.startswith("unary_act") for k in per_op) and all(v >= 0 for v in per_op.values())
    log = prof.profile_n_log(str(tmp_path / "ops.log"), num_iterations=2)
    assert (tmp_path / "ops.log").read_text().count("ms") == len(log)
    ids = prof.get_lookup_sampler(100, ignore_rate=0.5)([1000])
    assert ids.max() < 100 and 0.3 < (ids == -1).mean() < 0.7
    v1ex.reset_graph()

    cache = str(tmp_path / "exetime.json")
    sim = v1.HetuSimulator(cache_path=cache)
    t1 = sim.get_node_time("matmul", [[64, 128], [128, 256]])
    assert t1 > 0 and sim.get_node_time("matmul", [[64, 128], [128, 256]]) == t1           # second call: cached
    assert sim.get_node_time("some_unknown_op", [[1024, 1024]], [1024, 1024]) > 0                  # memory-bound estimate
    sim.write_cache()
    again = v1.HetuSimulator(cache_path=cache)
    assert again.get_node_time("matmul", [[64, 128], [128, 256]]) == t1                      # survives through the file
    g0, g1 = v1.gpu(0), v1.gpu(1)
    far = v1.rgpu("other-host", 0)
    shape = [1024, 1024]
    assert sim.get_dev_distance(g0, g0) == 0 and sim.get_dev_distance(g0, g1) == 1 and sim.get_dev_distance(g0, far) == 2
    near_t, far_t = sim.get_comm_time(g0, g1, shape), sim.get_comm_time(g0, far, shape)
    assert sim.get_comm_time(g0, g0, shape) == 0 and 0 < near_t < far_t
    np.testing.assert_allclose(near_t, 8e-3 + 4 * 2 ** 20 / 900e9 * 1e3, rtol=1e-6)
    ar2, ar8 = sim.get_allreduce_time(shape, [v1.gpu(i) for i in range(2)]), sim.get_allreduce_time(shape, [v1.gpu(i) for i in range(8)])
    ag8 = sim.get_allreduce_time(shape, [v1.gpu(i) for i in range(8)], v1.NCCLOP.AllGather)
    assert 0 < ar2 < ar8 and ag8 < ar8 and sim.get_allreduce_time(shape, [g0]) == 0
    assert sim.get_split_shape({0: 2, 1: 4}, shape) == [512, 256] and sim.get_concatenate_shape([[2, 3], [5, 3]], 0) == [7, 3]
    assert sim.get_split_time(shape, [0], [0], [2]) < sim.get_concatenate_time([shape, shape], 0)
    assert sim.get_update_time(shape) > sim.get_update_time(shape, sparse_shape=[16, 1024])
    # re-sharding: rows split over 2 devices -> columns split over the same 2: each target pulls the quarter it lacks from the peer
    moved = sim.get_general_comm_time({0: 2}, {1: 2}, [g0, g1], [g0, g1], shape)
    np.testing.assert_allclose(moved, sim.get_comm_time(g0, g1, [512, 512]), rtol=1e-9)
    assert sim.get_general_comm_time({0: 2}, {0: 2}, [g0, g1], [g0, g1], shape) == 0
    assert sim.get_group_comm_time([(g0, g1, shape), (g0, g1, shape)]) > sim.get_group_comm_time([(g0, g1, shape)])
    prof2 = v1.NCCLProfiler()                                                                       # single process: nothing to move
    assert prof2.profile_allreduce(1024, [0]) == 0.0 and prof2.profile_sendrecv(1024, [0, 0]) == 0.0


def test_v1_arrays_indexed_slices_device_groups_and_node_status(tmp_path):
    """ref: hetu/v1/python/hetu/ndarray.py (array / sparse_array / IndexedSlices), context.py (DeviceGroup, NodeStatus), data.py"""
    import gzip
    import pickle
    import numpy as np
    import hetu_b200.v1 as v1
    a = v1.array(np.arange(6.0).reshape(2, 3), v1.gpu(0))
    assert a.shape == (2, 3) and a.asnumpy().dtype == np.float32 and v1.is_gpu_ctx(a.ctx) and not v1.is_gpu_ctx(v1.cpu(0))
    b = v1.empty((2, 3), v1.cpu(0))
    a.copyto(b)
    b[0] = np.zeros(3)
    np.testing.assert_array_equal(b.asnumpy(), [[0, 0, 0], [3, 4, 5]])
    assert v1.empty_like(a).shape == (2, 3) and v1.array(np.array([1, 2], np.int64), v1.cpu(0)).asnumpy().dtype == np.int64
    for form in ("csr", "coo"):
        sp = v1.sparse_array([1.0, 2.0, 3.0], ([1, 0, 1], [0, 1, 2]), (2, 3), form=form)
        np.testing.assert_array_equal(sp.to_dense().asnumpy(), [[0, 2, 0], [1, 0, 3]])
        idx, val = sp.coo()
        assert idx.shape == (2, 3) and val.numel() == 3
    sl = v1.IndexedSlices(np.array([1, 3, 1, -1]), np.arange(8, dtype=np.float32).reshape(4, 2), (5, 2))
    d = sl.deduplicate()
    assert d.indices.tolist() == [1, 3] and d.values.tolist() == [[4.0, 6.0], [2.0, 3.0]] and sl.get_sparse_shape() == (4, 2)
    np.testing.assert_array_equal(sl.asnumpy(), [[0, 0], [4, 6], [0, 0], [2, 3], [0, 0]])

    g = v1.DeviceGroup([("gpu:0", "gpu:1"), ("node2:gpu:0", "node2:gpu:1"), "cpu:0"])
    assert g.is_mp and g.mp_dev_num == 2 and g.worker_num == 2 and g.server_num == 1 and len(g.flat_workers()) == 4
    assert g.index(v1.rgpu("node2", 1)) == 1 and g.workers[1][0].hostname == "node2" and g == v1.DeviceGroup(g)
    g.check_mp_num(2)
    assert v1.DeviceGroup("gpu:3").get_only() == v1.gpu(3) and v1.DeviceGroup([v1.gpu(1), v1.gpu(0)]).get_sorted()[0] == v1.gpu(0)

    # row-split partial sums over 8 devices: 2 (split) x 2 (partial) x 2 (duplicate)
    n = v1.NodeStatus({0: 2}, dev_num=8, partial=2)
    assert (n.duplicate, n.partial, n.get_default_order(), n.get_loop_sizes()) == (2, 2, (-2, -1, 0), (4, 2, 1))
    assert n.map_dev_to_index(5, True) == {-2: 1, -1: 0, 0: 1} and n.map_dev_to_index(5) == {0: 1} and n.get_devices_by_dim(0, 1) == [1, 3, 5, 7]
    summed = n.remove_partial()
    assert summed.duplicate == 4 and summed.partial == 1 and n.check_allreduce(summed) and not n.check_allgather(summed)
    gathered = v1.NodeStatus({}, dev_num=8, duplicate=8)
    assert summed.check_allgather(gathered) and v1.NodeStatus({0: 4}, dev_num=4).check_allgather(v1.NodeStatus({}, dev_num=4, duplicate=4))
    rs_target = v1.NodeStatus({0: 4}, dev_num=8, duplicate=2, order=(-1, 0))
    assert n.combine_state((-2, 0)) == ({0: 4}, 2, 1) and n.check_reducescatter(rs_target)
    assert n.exchange_state(0, 1) == ({1: 2}, 2, 2) and n.exchange_order(0, 1) == (-2, -1, 1)
    ds = n.to_distributed_states()
    assert ds.device_num == 8 and dict(ds.states)[-2] == 2 and v1.NodeStatus.from_distributed_states(ds) == n
    assert n.valid(True) and n.valid_state() and {n: 1}[v1.NodeStatus({0: 2}, dev_num=8, partial=2)] == 1

    # data helpers: one-hot, augmentation, the MNIST archive format
    assert v1.data.convert_to_one_hot([0, 2], 3).tolist() == [[1, 0, 0], [0, 0, 1]]
    imgs = np.random.RandomState(0).rand(4, 3, 32, 32).astype(np.float32)
    aug = v1.data.data_augmentation(imgs, "train", flip=True, crop=True, crop_shape=(24, 24), whiten=True, rng=np.random.RandomState(1))
    assert aug.shape == (4, 3, 24, 24) and abs(aug.reshape(4, -1).mean(1)).max() < 1e-4
    assert v1.data.data_augmentation(imgs, "test", crop=True).shape == (4, 3, 24, 24)
    arch = tmp_path / "mnist.pkl.gz"
    with gzip.open(arch, "wb") as f:
        pickle.dump([(np.zeros((5, 784)), np.arange(5))] * 3, f)
    (tx, ty), _, _ = v1.data.mnist(str(arch))
    assert tx.shape == (5, 784) and ty.shape == (5, 10)
    assert v1.lr is v1.lr_scheduler and v1.BertTokenizer is not None


def test_v1_partial_reduce_groups_the_workers_that_are_ready(tmp_path):
    """ref: hetu/v1/python/hetu/preduce.py -- a late worker is left out of the round (the punctual two average among themselves),
    a round with everybody on time averages over all three"""
    import json
    import os
    import subprocess
    import sys
    import time
    from dist_utils import free_port
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = os.path.join(root, "tests", "workers", "v1_preduce_worker.py")
    base = dict(os.environ, PYTHONPATH=root, HETU_B200_FORCE_CPU="1", CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="1",
                DMLC_PS_ROOT_URI="127.0.0.1", DMLC_PS_ROOT_PORT=str(free_port()), DMLC_NUM_SERVER="1", DMLC_NUM_WORKER="3",
                MASTER_ADDR="127.0.0.1", MASTER_PORT=str(free_port()), WORLD_SIZE="3")
    for k in ("HETU_PS_SCHEDULER", "HETU_PS_ADDRESS", "RANK", "LOCAL_RANK"):
        base.pop(k, None)
    procs = []
    for i, role in enumerate(["scheduler", "server", "worker", "worker", "worker"]):
        env = dict(base, DMLC_ROLE=role)
        if role == "worker":
            env.update(RANK=str(i - 2), LOCAL_RANK=str(i - 2))
        out = open(tmp_path / f"{role}{i}.log", "w")
        procs.append((role, i, subprocess.Popen([sys.executable, script], env=env, stdout=out, stderr=subprocess.STDOUT)))
        if role == "scheduler":
            time.sleep(1.0)
    codes = []
    for role, i, p in procs:
        try:
            codes.append(p.wait(150))
        except subprocess.TimeoutExpired:
            p.kill()
            codes.append(-9)
    logs = {f"{role}{i}": (tmp_path / f"{role}{i}.log").read_text() for role, i, _ in procs}
    assert codes == [0] * 5, json.dumps(logs)[-4000:]
    res = {r["rank"]: r for r in (json.loads(l[8:]) for t in logs.values() for l in t.splitlines() if l.startswith("PREDUCE "))}
    assert res[0]["p1"] == res[1]["p1"] == [0, 1] and res[2]["p1"] == [2]
    assert res[0]["first"] == res[1]["first"] == [1.5] * 4 and res[2]["first"] == [3.0] * 4
    assert all(res[r]["p2"] == [0, 1, 2] and res[r]["second"] == [2.0] * 4 for r in range(3))


def test_v1_logger_memory_reuse_plan_and_stream_handles(tmp_path):
    """ref: hetu/v1/python/hetu/{logger,memory_pool,stream}.py"""
    import json
    import numpy as np
    import hetu_b200.v1 as v1
    from hetu_b200.v1 import executor as v1ex
    from hetu_b200.v1.memory_pool import nodes_of_graph
    log = v1.HetuLogger(file=str(tmp_path / "m.jsonl"), echo=False)
    log.log("loss", np.array([0.5]))
    log.wrapped_log("acc", [[0.25]])
    with pytest.raises(AssertionError):
        log.log("loss", 1.0)
    log.step()
    log.log("loss", 0.4)
    log.step()
    recs = [json.loads(l) for l in open(tmp_path / "m.jsonl")]
    assert [r["loss"] for r in recs] == [0.5, 0.4] and recs[0]["acc"] == 0.25 and recs[1]["_step"] == 1 and len(log.history) == 2
    quiet = v1.HetuLogger(rank=1, nrank=2, echo=False)                  # not rank 0: buffers, never emits
    quiet.log("x", 1.0)
    quiet.step()
    assert quiet.history == [] and not quiet.need_log
    w = v1.WandbLogger("proj", "run", file=str(tmp_path / "w.jsonl"))
    w.set_config({"lr": 0.1})
    w.log("loss", 1.0)
    w.step()
    assert json.loads(open(tmp_path / "w.jsonl").read())["loss"] == 1.0 and w.config == {"lr": 0.1} and w.name == "run"

    # a chain of same-shaped activations needs two buffers, not one per node; fetched nodes and parameters are never recycled
    v1ex.reset_graph()
    x = v1.placeholder_op("mx", [32, 64])
    wgt = v1.Variable("mw", value=np.zeros((64, 64), np.float32))
    h = x
    for _ in range(6):
        h = v1.relu_op(v1.matmul_op(h, wgt))
    ex = v1.Executor([h])
    pool = v1.HetuMemoryPool()
    plan = pool.plan_graph(ex.graph, [h])
    nodes = nodes_of_graph(ex.graph)
    acts = [n for n in nodes if n.type in ("matmul", "unary_act")]
    assert len(acts) == 12 and plan["buffers"] < len(nodes) and plan["allocated_bytes"] < plan["naive_bytes"]
    owners = {n.id for n in acts if n.id not in plan["reuse"]}
    assert len(owners) == 3 and h.id not in plan["reuse"].values() and all(v in {n.id for n in acts} for v in plan["reuse"].values())
    # replaying the plan never hands a buffer to a node while an unfinished reader still needs it
    owner_of = {n.id: plan["reuse"].get(n.id, n.id) for n in acts}
    for i, n in enumerate(acts):
        for later in acts[i + 1:]:
            if owner_of[later.id] == owner_of[n.id]:
                readers = [m for m in acts if n.id in m.inputs]
                assert all(acts.index(m) < acts.index(later) or m is later for m in readers)
                break
    assert pool.test_memory(["gpu0"], {"gpu0": nodes}) and not pool.test_memory(["gpu0"], {"gpu0": nodes}, capacity_bytes=1024)
    v1ex.reset_graph()

    a, b = v1.stream.create_event_handle(v1.cpu(0)), v1.stream.create_event_handle(v1.cpu(0))
    a.record()
    b.record(v1.stream.create_stream_handle(v1.cpu(0)))
    b.sync()
    assert b.time_since(a) >= 0
    ev = v1.stream.CSEvent(None, 3)
    ev.update_ts(7)
    assert ev.updated and ev.ts == 7
    ev.sync()
    assert not ev.updated


def test_v1_reversed_truncated_normal_raw_data_batch_indices_and_gnn_feeder(tmp_path):
    """ref: hetu/v1/python/hetu/initializers.py:231, dataloader.py:10,34,253"""
    import numpy as np
    import hetu_b200.v1 as v1
    from hetu_b200.v1 import executor as v1ex
    v1ex.reset_graph()
    v1.random.set_random_seed(5)
    w = v1.init.reversed_truncated_normal([200, 50], mean=1.0, stddev=0.5, name="rtn")
    gen = v1.init.GenReversedTruncatedNormal(0.0, 1.0)([64], name="rtn2")
    vals = v1.Executor([w, gen]).run(feed_dict={}, convert_to_numpy_ret_vals=True)
    z = (vals[0] - 1.0) / 0.5
    assert np.abs(z).min() >= 2.0 - 1e-5 and abs((z > 0).mean() - 0.5) < 0.05 and np.abs(vals[1]).min() >= 2.0 - 1e-5

    mm = np.memmap(tmp_path / "chunk.bin", dtype=np.float32, mode="w+", shape=(6, 3))
    mm[:] = np.arange(18).reshape(6, 3) + 100
    raw = v1.RawData([np.arange(12).reshape(4, 3), mm], dtype=np.float32)
    assert len(raw) == 10 and raw.shape == (10, 3) and raw[5].tolist() == [103, 104, 105]
    np.testing.assert_array_equal(raw[[0, 9, 4, 3]], [[0, 1, 2], [115, 116, 117], [100, 101, 102], [9, 10, 11]])
    np.testing.assert_array_equal(raw[3:6], [[9, 10, 11], [100, 101, 102], [103, 104, 105]])
    bi = v1.BatchIndices(5, need_shuffle=True, seed=1)
    first = [bi[k] for k in range(5)]
    second = [bi[k] for k in range(5)]
    assert sorted(first) == sorted(second) == list(range(5)) and first != second
    with pytest.raises(AssertionError):
        bi[2], bi[0]                                           # a new pass before the previous one finished

    # graph feeder: two nodes (features, normalised adjacency) read the CURRENT graph; step() rotates current <- next
    v1ex.reset_graph()
    v1.GNNDataLoaderOp.graph = v1.GNNDataLoaderOp.nxt_graph = None
    feats = v1.GNNDataLoaderOp(lambda g: g["x"], shape=[4, 2], name="gnn_x")
    adj = v1.GNNDataLoaderOp(lambda g: g["a"], shape=[4, 4], name="gnn_a")
    out = v1.matmul_op(adj.node, feats.node)
    ex = v1.Executor([out])
    g1 = {"x": np.ones((4, 2), np.float32), "a": np.eye(4, dtype=np.float32) * 2}
    g2 = {"x": np.full((4, 2), 3.0, np.float32), "a": np.eye(4, dtype=np.float32)}
    v1.GNNDataLoaderOp.step(g1)
    v1.GNNDataLoaderOp.step(g2)                               # current = g1, next = g2
    np.testing.assert_array_equal(ex.run(feed_dict={}, convert_to_numpy_ret_vals=True)[0], np.full((4, 2), 2.0))
    assert feats.get_next_arr()[0, 0] == 3.0
    v1.GNNDataLoaderOp.step(None)
    np.testing.assert_array_equal(ex.run(feed_dict={}, convert_to_numpy_ret_vals=True)[0], np.full((4, 2), 3.0))
    v1ex.reset_graph()


def test_v1_function_style_launcher_and_saved_search_plans(tmp_path):
    """ref: hetu/v1/python/hetu/launcher.py (launch(target, args): scheduler + server + workers as processes of one program) and
    distributed_strategies/base.py (BaseSearchingStrategy save / load)"""
    import argparse
    import json
    import os
    import sys
    import yaml
    from dist_utils import free_port
    from hetu_b200.v1 import launcher
    from hetu_b200.v1 import strategies as S
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "tests", "workers"))     # stays until the roles have started: spawn hands sys.path to the children
    import v1_launch_roles_target as target_mod
    cfg = tmp_path / "local.yml"
    cfg.write_text(yaml.safe_dump({"shared": {"DMLC_PS_ROOT_URI": "127.0.0.1", "DMLC_PS_ROOT_PORT": free_port(), "DMLC_NUM_WORKER": 2, "DMLC_NUM_SERVER": 1,
                                              "HETU_B200_FORCE_CPU": 1, "PYTHONPATH": root + os.pathsep + os.path.join(root, "tests", "workers")},
                                   "launch": {"worker": 2, "server": 1, "scheduler": 1}}))
    keep = dict(os.environ)
    try:
        for k in ("HETU_PS_SCHEDULER", "HETU_PS_ADDRESS"):
            os.environ.pop(k, None)
        codes = launcher.launch(target_mod.train, argparse.Namespace(config=str(cfg), out=str(tmp_path)), timeout=60)
    finally:
        os.environ.clear()
        os.environ.update(keep)
        sys.path.remove(os.path.join(root, "tests", "workers"))
    assert codes == [0, 0, 0, 0], codes
    for w in range(2):
        r = json.load(open(tmp_path / f"worker{w}.json"))
        assert r == {"w": [-2.0] * 4, "role": "worker"}

    layers = [S.LayerSpec(f"l{i}", flops=4e12, params=2e8, act=6e7) for i in range(6)] if hasattr(S, "LayerSpec") else None
    if layers is not None:
        path = str(tmp_path / "plan.json")
        first = S.BaseSearchingStrategy(S.FlexFlowSearching(8, budget=100), save_path=path)
        plan = first.assign(layers)
        saved = json.load(open(path))
        assert saved["strategy"] == "FlexFlowSearching" and len(saved["placements"]) == 6 and saved["estimated_step_s"] > 0
        again = S.BaseSearchingStrategy(S.FlexFlowSearching(8, budget=100, seed=9), load_path=path)
        assert [p.key() for p in again.assign(layers)] == [p.key() for p in plan] and again.loaded and not first.loaded


def test_embedding_compression_post_training_and_autosrh_schedules():
    """ref: tools/EmbeddingMemoryCompression/methods/scheduler/{compressor,switchinference,deduplication,quantize,autosrh}.py -- train
    the full table, compress it to the requested rate (prune to CSR / merge near-identical blocks / quantise), keep training; AutoSrh
    searches its gates then retrains under the frozen mask"""
    from hetu_b200.tools.emb_compress.trainer import CompressionTrainer
    common = dict(num_embeddings=1200, dim=8, num_fields=4, num_dense=3, batch_size=128, lr=0.02)
    r = CompressionTrainer("sparse", "wdl", compress_rate=0.3, **common).run(steps=30, eval_batches=2)
    assert r["stage1"]["ratio"] == 1.0 and 2.5 < r["ratio"] < 4.5 and r["schedule"]["sparsity"] > 0.8 and 0.0 <= r["auc"] <= 1.0
    r = CompressionTrainer("dedup", "wdl", compress_rate=0.5, **common).run(steps=30, eval_batches=2)
    assert 1.6 < r["ratio"] < 2.6 and r["schedule"]["tolerance"] > 0 and np.isfinite(r["stage2_loss"][1])
    tr = CompressionTrainer("quantize", "wdl", compress_rate=0.25, **common)
    r = tr.run(steps=30, eval_batches=2)
    assert 3.5 < r["ratio"] <= 4.0 and r["schedule"]["digit"] == 8 and np.isfinite(r["stage2_loss"][1])
    # stage 2 started from the trained table (not a fresh one): after its 30 steps the two are still strongly correlated
    assert np.corrcoef(tr.trained_table.ravel(), tr.value_of(tr.embedding.weight).ravel())[0, 1] > 0.8
    r = CompressionTrainer("autosrh", "wdl", compress_rate=0.25, method_kwargs={"nsplit": 4}, **common).run(steps=30, eval_batches=2)
    assert r["stage1"]["ratio"] <= 1.0 and 3.0 < r["ratio"] < 5.0 and "alpha_abs_mean" in r["schedule"]
