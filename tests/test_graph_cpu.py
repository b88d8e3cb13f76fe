"""Define-and-run graphs on CPU (BASELINE config #1: 2-layer MLP, world_size 1) plus executor features:
micro-batch accumulation, run levels, recompute, module API, symbolic shapes, multi-strategy shape plans."""
import math
import os
import numpy as np
import pytest
import torch

import hetu_b200 as ht


def build_mlp(g, din=16, dh=32, dout=4):
    x = ht.placeholder("float32", [8, din], name="x")
    y = ht.placeholder("int64", [8], name="y")
    fc1, fc2 = ht.nn.Linear(din, dh, name="fc1"), ht.nn.Linear(dh, dout, name="fc2")
    loss = ht.softmax_cross_entropy_sparse(fc2(fc1(x, act="relu")), y)
    return x, y, fc1, fc2, loss


def test_two_layer_mlp_matches_torch():
    ht.set_seed(1)
    X = np.random.RandomState(0).randn(8, 16).astype(np.float32)
    Y = np.random.RandomState(1).randint(0, 4, (8,))
    with ht.graph("define_and_run", create_new=True) as g:
        x, y, fc1, fc2, loss = build_mlp(g)
        opt = ht.SGDOptimizer(lr=0.1)
        train = opt.minimize(loss)
        w1, b1, w2, b2 = (g.get_param(p).clone() for p in (fc1.weight, fc1.bias, fc2.weight, fc2.bias))
        losses = [float(g.run(loss, [loss, train], {x: X, y: Y})[0]) for _ in range(5)]
    tw1, tb1, tw2, tb2 = (t.clone().requires_grad_() for t in (w1, b1, w2, b2))
    ref = []
    for _ in range(5):
        l = torch.nn.functional.cross_entropy(torch.relu(torch.tensor(X) @ tw1.t() + tb1) @ tw2.t() + tb2, torch.tensor(Y))
        ref.append(float(l))
        gs = torch.autograd.grad(l, [tw1, tb1, tw2, tb2])
        with torch.no_grad():
            for p, gr in zip((tw1, tb1, tw2, tb2), gs):
                p -= 0.1 * gr
    np.testing.assert_allclose(losses, ref, rtol=1e-5)


def test_micro_batches_equal_full_batch():
    X = np.random.RandomState(0).randn(8, 16).astype(np.float32)
    Y = np.random.RandomState(1).randint(0, 4, (8,))

    def run(mb):
        ht.set_seed(5)
        with ht.graph("define_and_run", create_new=True) as g:
            x = ht.placeholder("float32", [8 // mb, 16], name="x")
            y = ht.placeholder("int64", [8 // mb], name="y")
            fc = ht.nn.Linear(16, 4, name="fc")
            loss = ht.softmax_cross_entropy_sparse(fc(x), y)
            train = ht.AdamOptimizer(lr=1e-2).minimize(loss)
            for _ in range(3):
                out = g.run(loss, [loss, train], {x: X, y: Y}, num_micro_batches=mb)
            return g.get_param(fc.weight).clone()

    np.testing.assert_allclose(run(1).numpy(), run(4).numpy(), rtol=1e-4, atol=1e-6)


def test_run_levels_grad_then_update():
    X = np.random.RandomState(0).randn(8, 16).astype(np.float32)
    Y = np.random.RandomState(1).randint(0, 4, (8,))
    ht.set_seed(2)
    with ht.graph("define_and_run", create_new=True) as g:
        x, y, fc1, fc2, loss = build_mlp(g)
        train = ht.SGDOptimizer(lr=0.5).minimize(loss)
        w0 = g.get_param(fc2.weight).clone()
        g.run(loss, [loss, train], {x: X, y: Y}, run_level="grad")          # accumulate only
        assert torch.equal(w0, g.get_param(fc2.weight))
        assert g.accumulated_grad(fc2.weight) is not None
        g.run(loss, [loss, train], {x: X, y: Y}, run_level="update")        # second gradient + update
        assert not torch.equal(w0, g.get_param(fc2.weight))
        g.run(loss, [loss], {x: X, y: Y}, run_level="compute_only")
        with ht.run_level("topo"):
            assert g.run(loss, [loss, train], {x: X, y: Y}) == []


def test_recompute_gives_identical_gradients():
    X = np.random.RandomState(0).randn(8, 16).astype(np.float32)

    def run(rc):
        ht.set_seed(9)
        with ht.graph("define_and_run", create_new=True) as g:
            x = ht.placeholder("float32", [8, 16], name="x")
            a, b, c = ht.nn.Linear(16, 32, name="a"), ht.nn.Linear(32, 32, name="b"), ht.nn.Linear(32, 1, name="c")
            h = a(x, act="gelu")
            if rc:
                with ht.recompute([True]):
                    h = ht.layer_norm(b(h, act="gelu"), ht.ones([32], requires_grad=False), ht.zeros([32], requires_grad=False))
            else:
                h = ht.layer_norm(b(h, act="gelu"), ht.ones([32], requires_grad=False), ht.zeros([32], requires_grad=False))
            loss = ht.sum(c(h))
            train = ht.SGDOptimizer(lr=0.1).minimize(loss)
            g.run(loss, [loss, train], {x: X})
            bd = g.step_breakdown()
            return g.get_param(a.weight).clone(), bd.get("recomputed_ops", 0)

    w_plain, n0 = run(False)
    w_rc, n1 = run(True)
    assert n0 == 0 and n1 > 0
    np.testing.assert_allclose(w_plain.numpy(), w_rc.numpy(), rtol=1e-6)


def test_module_api_and_state_dict():
    with ht.graph("define_and_run", create_new=True) as g:
        m = ht.nn.Sequential(ht.nn.Linear(4, 8, name="l0"), ht.nn.ReLU(), ht.nn.Linear(8, 2, name="l1"))
        names = [n for n, _ in m.named_parameters()]
        assert names == ["0.weight", "0.bias", "2.weight", "2.bias"]
        sd = m.state_dict()
        assert sd["0.weight"].shape == (8, 4)
        sd["2.bias"] = torch.ones(2)
        m.load_state_dict(sd)
        x = ht.placeholder("float32", [3, 4], name="x")
        out = g.run(None, [m(x)], {x: np.zeros((3, 4), np.float32)})[0]
        l0b = sd["0.bias"]
        ref = torch.relu(l0b) @ sd["2.weight"].t() + 1.0
        np.testing.assert_allclose(out[0].numpy(), ref.numpy(), rtol=1e-5)
        subs = g.subgraphs()
        assert any(k.endswith("0") for k in subs) and all("fwd_ops" in v for v in subs.values())


def test_symbolic_reshape_follows_int_symbol():
    with ht.graph("define_and_run", create_new=True) as g:
        seq = ht.IntSymbol(4)
        x = ht.placeholder("float32", [2, 8], name="x")
        y = ht.reshape(x, [ht.IntSymbol(2) * (ht.IntSymbol(8) // seq), seq])
        out = g.run(None, [y], {x: np.arange(16, dtype=np.float32).reshape(2, 8)})[0]
        assert list(out.shape) == [4, 4]
        out = g.run(None, [y], {x: np.arange(16, dtype=np.float32).reshape(2, 8)}, int_symbol_dict={seq: 2})[0]
        assert list(out.shape) == [8, 2]


def test_gradients_api_and_eager_mix():
    with ht.graph("define_and_run", create_new=True) as g:
        x = ht.placeholder("float32", [3, 3], name="x")
        w = ht.parameter(ht.ones_initializer(), [3, 3], name="w")
        y = ht.sum(ht.matmul(x, w) * 2.0)
        gw, = ht.gradients(y, [w])
        out = g.run(None, [gw], {x: np.eye(3, dtype=np.float32)})[0]
        np.testing.assert_allclose(out.numpy(), 2 * np.ones((3, 3)))
    # eager graph keeps working afterwards
    a = ht.from_numpy(np.ones((2, 2), np.float32), requires_grad=True)
    ht.sum(a * 3.0).backward()
    np.testing.assert_allclose(a.grad.numpy(), 3 * np.ones((2, 2)))


def test_lr_schedule_and_grad_scaler():
    from hetu_b200.optim import OptimizerParamScheduler
    s = OptimizerParamScheduler(0.0, 1.0, 0.1, lr_warmup_steps=10, lr_decay_steps=110, lr_decay_style="cosine")
    assert abs(s.get_lr(5) - 0.5) < 1e-9 and abs(s.get_lr(10) - 1.0) < 1e-9
    assert abs(s.get_lr(60) - (0.1 + 0.9 * 0.5)) < 1e-9 and s.get_lr(500) == 0.1
    lin = OptimizerParamScheduler(0.0, 1.0, 0.0, 0, 100, "linear")
    assert abs(lin.get_lr(25) - 0.75) < 1e-9
    sc = ht.GradScaler(init_scale=8.0, growth_interval=2)
    sc.update(True)
    assert sc.get_scale() == 4.0
    sc.update(False); sc.update(False)
    assert sc.get_scale() == 8.0


def test_tensor_surface_timecost_splits_device_graph():
    """reference Tensor API: .graph / .device / .timecost(µb) under the profiler / .reset_data_from_splits"""
    with ht.graph("define_and_run", create_new=True) as g:
        x = ht.placeholder("float32", [4, 8], name="x")
        w = ht.parameter(ht.ones_initializer(), [8, 8], requires_grad=True, name="w_tc")
        y = ht.matmul(x, w, name="mm_tc")
        loss = ht.sum(y)
    assert y.graph is g and w.graph is g and w.device is not None
    with ht.profiler(enabled=True, graph=g):
        out = g.run(loss, [loss], {x: torch.ones(4, 8)})
    assert float(out[0]) == 4 * 8 * 8
    assert y.timecost(0) >= 0.0 and loss.timecost() >= 0.0
    w.reset_data_from_splits([np.full((3, 8), 2.0, dtype=np.float32), np.full((5, 8), 0.5, dtype=np.float32)])
    out = g.run(loss, [loss], {x: torch.ones(4, 8)})
    assert float(out[0]) == 4 * 8 * (3 * 2.0 + 5 * 0.5)


def test_lr_and_weight_decay_schedules_reach_the_update_ops():
    """AdamOptimizer(lr_warmup_steps, lr_decay_steps, lr_decay_style, start_wd/end_wd...) + step_lr(): with a constant gradient
    Adam moves every weight by exactly lr_t per step, so the trajectory exposes the schedule"""
    with ht.graph("define_and_run", create_new=True) as g:
        w = ht.parameter(ht.zeros_initializer(), [4], requires_grad=True, name="w_sched")
        loss = ht.sum(w)
        opt = ht.AdamOptimizer(lr=0.1, lr_warmup_steps=4, lr_decay_steps=8, lr_decay_style="linear", min_lr=0.02, eps=1e-12)
        train = opt.minimize(loss)
    want = [opt.scheduler.get_lr(s) for s in range(1, 11)]
    assert want[0] == pytest.approx(0.025) and want[3] == pytest.approx(0.1) and want[5] == pytest.approx(0.06) and want[9] == pytest.approx(0.02)
    pos, seen = 0.0, []
    for s in range(10):
        g.run(loss, [loss, train], {})
        cur = float(torch.as_tensor(w.numpy())[0])
        seen.append(pos - cur)
        pos = cur
        assert opt.step_lr() == pytest.approx(want[s + 1] if s + 1 < 10 else opt.scheduler.get_lr(11))
    assert seen == pytest.approx(want, rel=1e-4)
    # external control (v1 lr_scheduler classes): SGD step = lr * grad
    with ht.graph("define_and_run", create_new=True) as g2:
        v = ht.parameter(ht.zeros_initializer(), [2], requires_grad=True, name="v_sched")
        l2 = ht.sum(v)
        sgd = ht.SGDOptimizer(lr=1.0)
        t2 = sgd.minimize(l2)
    from hetu_b200.v1.lr_scheduler import ExponentialScheduler, MultiStepScheduler, ReduceOnPlateauScheduler, StepScheduler
    sch = StepScheduler(1.0, step_size=2, gamma=0.5)
    deltas, pos = [], 0.0
    for s in range(5):
        sgd.set_learning_rate(sch.get())
        g2.run(l2, [l2, t2], {})
        cur = float(torch.as_tensor(v.numpy())[0])
        deltas.append(pos - cur)
        pos = cur
        sch.step()
    assert deltas == pytest.approx([1.0, 1.0, 0.5, 0.5, 0.25])
    ms = MultiStepScheduler(1.0, [2, 3], 0.1)
    assert [round(ms.step(), 6) for _ in range(4)] == [1.0, 0.1, 0.01, 0.01]
    ex = ExponentialScheduler(2.0, 0.5)
    assert [ex.step() for _ in range(3)] == [1.0, 0.5, 0.25]
    rp = ReduceOnPlateauScheduler(1.0, factor=0.5, patience=1)
    assert [rp.step(x) for x in (1.0, 0.9, 0.9, 0.9, 0.9, 0.9)] == [1.0, 1.0, 1.0, 0.5, 0.5, 0.25]


def test_extra_nn_modules_pad_newgelu_single_config_parallel():
    from hetu_b200.models.parallel_config import generate_ds_parallel_config
    x = torch.randn(2, 3, 4, 5)
    y = ht.nn.ConstantPad2d((1, 2, 0, 3), 1.5)(ht.from_numpy(x))
    assert torch.equal(torch.as_tensor(y.numpy()), torch.nn.functional.pad(x, (1, 2, 0, 3), value=1.5))
    assert torch.equal(torch.as_tensor(ht.nn.ZeroPad2d(2)(ht.from_numpy(x)).numpy()), torch.nn.functional.pad(x, (2, 2, 2, 2)))
    g = ht.nn.NewGeLU()(ht.from_numpy(x))
    assert torch.allclose(torch.as_tensor(g.numpy()), torch.nn.functional.gelu(x, approximate="tanh"), atol=1e-6)
    cfg = generate_ds_parallel_config(2, 1, 1, 1, 1, zero=False)
    blk = cfg["blocks"]["blocks0-1"]
    with ht.graph("define_and_run", create_new=True) as gph:
        col = ht.nn.HtColumnParallelLinear(8, 16, blk["mlp"]["dense_h_to_4h"], gather_output=False, name="c1")
        row = ht.nn.HtRowParallelLinear(16, 8, ds_parallel_config=blk["mlp"]["dense_4h_to_h"], name="r1")
        ln = ht.nn.HtParallelLayerNorm(8, blk["layernorm1"], name="l1")
        inp = ht.placeholder("float32", [4, 8], name="inp")
        out = row(col(ln(inp)))
        res = gph.run(None, [out], {inp: torch.randn(4, 8)})[0]
    assert tuple(res.shape) == (4, 8) and type(col).__name__ == "HtColumnParallelLinear" and isinstance(col, ht.nn.HtMultiColumnParallelLinear)


def test_micro_batches_of_different_sequence_lengths():
    """graph.run(int_symbol_dict={seq_len: [16, 8]}, feeds of different widths per micro-batch): the executor sets the symbol
    and re-infers static shapes before every micro-batch; the step equals accumulating the two micro-batches in separate runs"""
    from hetu_b200.models import GPTConfig, GPTLMHeadModel, generate_ds_parallel_config
    rng = np.random.RandomState(0)
    cfg = GPTConfig(vocab_size=64, n_positions=16, n_embd=16, n_layer=2, n_head=2)
    widths = [16, 8]
    data = [(rng.randint(0, 64, (2, w)), np.tile(np.arange(w), (2, 1))) for w in widths]

    def build():
        ht.set_seed(11)
        with ht.graph("define_and_run", create_new=True) as g:
            sym = ht.IntSymbol(16)
            model = GPTLMHeadModel(cfg, [generate_ds_parallel_config(2, 1, 1, 1, 1, zero=False)])
            ids, pos, lab = (ht.placeholder("int64", [32], name=n) for n in ("ids", "pos", "lab"))
            loss = model(ids, pos, lab, seq_len=sym)
            train = ht.SGDOptimizer(lr=0.5).minimize(loss)
        return g, sym, ids, pos, lab, loss, train, model

    def feeds(ids, pos, lab, which):
        f = {ids: [], pos: [], lab: []}
        for k in which:
            x, p = data[k]
            f[ids].append(torch.as_tensor(x.reshape(-1)))
            f[pos].append(torch.as_tensor(p.reshape(-1)))
            f[lab].append(torch.as_tensor(np.roll(x, -1, axis=1).reshape(-1)))
        return f

    g, sym, ids, pos, lab, loss, train, model = build()
    out = g.run(loss, [loss, train], feeds(ids, pos, lab, [0, 1]), int_symbol_dict={sym: widths}, num_micro_batches=2)
    per_mb = [float(v) for v in out[0].reshape(-1)]
    after_a = {n: g.get_param(p).clone() for n, p in model.named_parameters()}

    g2, sym2, ids2, pos2, lab2, loss2, train2, model2 = build()
    with ht.run_level("grad"):
        l0 = g2.run(loss2, [loss2, train2], feeds(ids2, pos2, lab2, [0]), int_symbol_dict={sym2: 16}, grad_scale=0.5)
    l1 = g2.run(loss2, [loss2, train2], feeds(ids2, pos2, lab2, [1]), int_symbol_dict={sym2: 8}, grad_scale=0.5)
    assert per_mb[1] == pytest.approx(float(l1[0]), rel=1e-5) and len(per_mb) == 2
    after_b = {n: g2.get_param(p) for n, p in model2.named_parameters()}
    assert set(after_a) == set(after_b) and len(after_a) > 10
    for n in after_a:
        assert torch.allclose(after_a[n], after_b[n], atol=2e-6, rtol=1e-5), n
    moved = sum(float((after_a[n] - g.get_param(p)).abs().sum()) for n, p in model.named_parameters())
    assert moved == 0.0 and any(float(v.abs().sum()) > 0 for v in after_a.values())


def test_module_tree_helpers_buffers_children_apply():
    with ht.graph("define_and_run", create_new=True):
        seq = ht.nn.Sequential(ht.nn.Linear(4, 8, name="a"), ht.nn.ReLU(), ht.nn.Linear(8, 2, name="b"))
        bn = ht.nn.BatchNorm(3)
    assert [n for n, _ in seq.named_children()] == ["0", "1", "2"] and len(list(seq.children())) == 3
    seen = []
    seq.apply(lambda m: seen.append(type(m).__name__))
    assert seen == ["Linear", "ReLU", "Linear", "Sequential"]
    assert len(list(bn.buffers())) == len(list(bn.named_buffers())) >= 2
    assert seq.eval().training is False and seq.train().training is True and seq.to(None) is seq


def test_grad_scaler_minimize_unscales_checks_and_skips():
    """scaled-loss training must follow the unscaled trajectory; a non-finite gradient skips the step and backs the scale off
    (ref: hetu/graph/autocast/gradscaler.h, optimizer_update.cc SGDUpdateWithGradScaler)"""
    def train(scaler, feeds):
        with ht.graph("define_and_run", create_new=True) as g:
            x = ht.placeholder("float32", [4, 3], name="x")
            w = ht.parameter(ht.ones_initializer(), [3, 1], requires_grad=True, name="w_gs")
            loss = ht.sum(ht.matmul(x, w)) * 0.25
            opt = ht.SGDOptimizer(lr=0.01)
            train_op = scaler.minimize(opt, loss) if scaler is not None else opt.minimize(loss)
            ws = []
            for f in feeds:
                g.run(loss, [loss, train_op], {x: f})
                ws.append(g.get_param(w).clone())
        return ws

    ones = torch.ones(4, 3)
    plain = train(None, [ones, ones, ones])
    sc = ht.GradScaler(init_scale=1024.0, growth_interval=2)
    scaled = train(sc, [ones, ones, ones])
    for a, b in zip(plain, scaled):
        np.testing.assert_allclose(a.numpy(), b.numpy(), rtol=1e-6)
    assert abs(plain[0][0, 0].item() - 0.99) < 1e-6           # one SGD step of size lr * dL/dw = 0.01 * 1.0
    assert sc.get_scale() == 2048.0 and sc.skipped_steps() == 0   # grew once after two clean steps
    # an inf in the batch: the step is dropped, the scale halves, the weights stay
    sc2 = ht.GradScaler(init_scale=64.0, growth_interval=100)
    bad = ones.clone(); bad[0, 0] = float("inf")
    ws = train(sc2, [ones, bad, ones])
    np.testing.assert_allclose(ws[1].numpy(), ws[0].numpy())
    assert sc2.found_inf() is False and sc2.skipped_steps() == 1 and sc2.get_scale() == 32.0
    np.testing.assert_allclose(ws[2].numpy(), plain[1].numpy(), rtol=1e-6)


def test_define_by_run_graph_reuses_ops_evaluates_lazily_and_prunes():
    """ref: hetu/graph/define_by_run_graph.{h,cc}: FindReusableOp (identical pure ops are not duplicated), lazy evaluation of
    exactly the needed ancestry with cached results, pruning of ops nobody references any more"""
    with ht.graph("define_by_run", create_new=True) as g:
        a = ht.from_numpy(np.arange(6, dtype=np.float32).reshape(2, 3))
        b = ht.from_numpy(np.ones((2, 3), np.float32))
        n0 = g.num_ops
        s1 = ht.exp(a + b)
        s2 = ht.exp(a + b)                       # same expression: both ops are reused
        assert g.num_ops == n0 + 2 and g.reuse_hits() == 2 and s1.id == s2.id
        d1, d2 = ht.dropout(a, 0.5), ht.dropout(a, 0.5)
        assert d1.id != d2.id                    # random ops are never merged
        big = ht.matmul(s1, ht.from_numpy(np.ones((3, 4), np.float32)))
        side = ht.sum(ht.relu(a - 3.0))          # an unrelated branch
        assert s1.eager_data() is None           # nothing has run yet
        np.testing.assert_allclose(big.numpy(), np.exp(np.arange(6).reshape(2, 3) + 1.0) @ np.ones((3, 4)), rtol=1e-5)
        assert s1.eager_data() is not None and side.eager_data() is None      # only big's ancestry was evaluated
        np.testing.assert_allclose(float(side.numpy()), 1.0 + 2.0)
        live = g.num_live_ops()
        del big, side, d1, d2, s2
        removed = g.prune()
        # matmul + its const, the relu/sub/sum chain and both dropouts go; exp(a + b) stays (s1 is still referenced)
        assert removed >= 7 and g.num_live_ops() == live - removed
        np.testing.assert_allclose(s1.numpy(), np.exp(np.arange(6).reshape(2, 3) + 1.0), rtol=1e-5)
        again = ht.exp(a + b)
        assert again.id == s1.id                 # still reusable after pruning its consumers


def test_graphboard_exports_dict_dot_and_html(tmp_path):
    """ref: hetu/v1/python/graphboard -- the graph as data, as Graphviz source and as a standalone HTML page"""
    import json as _json
    from hetu_b200.utils import graphboard
    with ht.graph("define_and_run", create_new=True) as g:
        with ht.subgraph("encoder"):
            lin = ht.nn.Linear(8, 4, name="gb_lin")
            x = ht.placeholder("float32", [2, 8], name="gb_x")
            h = ht.relu(lin(x))
        loss = ht.sum(h)
        train = ht.SGDOptimizer(lr=0.1).minimize(loss)
    d = graphboard.graph_to_dict(g)
    types = [o["type"] for o in d["ops"]]
    assert "linear" in types and "unary_act" in types and any(t.endswith("_update") for t in types) and any(o["is_bwd"] for o in d["ops"])
    relu = next(o for o in d["ops"] if o["type"] == "unary_act" and not o["is_bwd"])
    assert d["tensors"][relu["outputs"][0]]["shape"] == [2, 4] and d["tensors"][relu["inputs"][0]]["producer"] is not None
    dot = graphboard.to_dot(g)
    assert dot.startswith("digraph") and "->" in dot and "cluster_" in dot and "linear" in dot and dot.count("[label=") == len(d["ops"])
    path = graphboard.to_html(g, str(tmp_path / "graph.html"), title="toy")
    page = open(path).read()
    assert "<table" in page and "gb_lin" in page and 'id="graph"' in page
    embedded = _json.loads(page.split('id="graph">')[1].split("</script>")[0])
    assert len(embedded["ops"]) == len(d["ops"])


def test_nn_init_functions_parameter_and_pretrained_round_trip(tmp_path):
    """ref: python/hetu/nn/{init,parameter}.py, models/utils/{config_utils,model_utils,common_utils}.py -- in-place initialisers with the
    right statistics, `nn.Parameter`, HuggingFace-style save_pretrained / from_pretrained with sharded safetensors"""
    from hetu_b200.models import GPTConfig, GPTLMHeadModel, generate_ds_parallel_config
    from hetu_b200.models.utils.pretrained import PreTrainedModel, parse_size, split_state_dict_into_shards
    from hetu_b200.nn import init
    ht.set_seed(7)
    with ht.graph("define_and_run", create_new=True) as g:
        lin = ht.nn.Linear(400, 300, name="init_lin")
        conv = ht.nn.Conv2d(8, 16, 3, name="init_conv")
        w = lin.weight
        init.xavier_uniform_(w)
        v = g.get_param(w)
        lim = math.sqrt(6.0 / (400 + 300))
        assert float(v.abs().max()) <= lim + 1e-6 and abs(float(v.std()) - lim / math.sqrt(3)) < 0.05 * lim
        init.kaiming_normal_(conv.weight, mode="fan_out", nonlinearity="relu")
        assert abs(float(g.get_param(conv.weight).std()) - math.sqrt(2.0 / (16 * 9))) < 0.02
        init.lecun_normal_(w)
        assert abs(float(g.get_param(w).std()) - math.sqrt(1.0 / 400)) < 0.005
        init.trunc_normal_(w, std=0.02, a=-0.04, b=0.04)
        assert float(g.get_param(w).abs().max()) <= 0.04 and abs(float(g.get_param(w).std()) - 0.0176) < 0.002
        init.constant_(lin.bias, 0.5); assert bool((g.get_param(lin.bias) == 0.5).all())
        init.zeros_(lin.bias); init.ones_(lin.bias); assert bool((g.get_param(lin.bias) == 1).all())
        init.uniform_(lin.bias, -2, 2); init.normal_(lin.bias, 1.0, 0.1)
        assert init.calculate_gain("relu") == math.sqrt(2.0) and abs(init.calculate_gain("leaky_relu", 0.2) - math.sqrt(2 / 1.04)) < 1e-9
        assert init._calculate_fan_in_and_fan_out(conv.weight) == (72, 144)
        p = ht.nn.Parameter(np.arange(6, dtype=np.float32).reshape(2, 3), name="my_param")
        assert p.requires_grad and list(p.shape) == [2, 3] and torch.equal(g.get_param(p), torch.arange(6.0).reshape(2, 3))
    # pretrained IO: config.json + sharded safetensors + index; reload under a fresh graph reproduces the logits
    files, index = split_state_dict_into_shards({"a": torch.zeros(1000), "b": torch.zeros(1000), "c": torch.zeros(10)}, max_shard_size=5000)
    assert list(files) == ["model-00001-of-00002.safetensors", "model-00002-of-00002.safetensors"] and index["weight_map"]["c"].startswith("model-00002")
    assert parse_size("5GB") == 5 * 10 ** 9 and parse_size("200MiB") == 200 * 2 ** 20
    cfg = GPTConfig(vocab_size=64, n_positions=16, n_embd=32, n_layer=2, n_head=2)
    ids = torch.randint(0, 64, (16,))
    with ht.graph("define_and_run", create_new=True) as g:
        m = GPTLMHeadModel(cfg, [generate_ds_parallel_config(2, 1, 1, 1, 1, zero=False)])
        x, pos = ht.placeholder("int64", [16], name="ids"), ht.placeholder("int64", [16], name="pos")
        logits = m(x, pos, None, seq_len=16)
        want = g.run(logits, [logits], {x: ids, pos: torch.arange(16)})[0].clone()
        written = m.save_pretrained(str(tmp_path / "gpt"), max_shard_size="20KB")
    assert len(written) > 1 and os.path.exists(tmp_path / "gpt" / "config.json") and os.path.exists(tmp_path / "gpt" / "model.safetensors.index.json")
    assert GPTConfig.from_pretrained(str(tmp_path / "gpt")) == cfg and isinstance(m, PreTrainedModel)
    with ht.graph("define_and_run", create_new=True) as g:
        m2 = GPTLMHeadModel.from_pretrained(str(tmp_path / "gpt"))
        x, pos = ht.placeholder("int64", [16], name="ids"), ht.placeholder("int64", [16], name="pos")
        logits = m2(x, pos, None, seq_len=16)
        got = g.run(logits, [logits], {x: ids, pos: torch.arange(16)})[0]
    assert torch.equal(got, want)
    with pytest.raises(ValueError, match="remote"):
        GPTLMHeadModel.from_pretrained("https://huggingface.co/gpt2")


def test_tokenizer_base_classes():
    from hetu_b200.data.tokenizers import BaseTokenizer, ByteTokenizer, PreTrainedTokenizer, SpecialToken
    tok = PreTrainedTokenizer(ByteTokenizer())
    assert isinstance(tok, BaseTokenizer) and tok.vocab_size == 259 and tok.pad == 256 and tok.eod == 258
    rows = tok.batch_encode(["hi", "hello"], padding=True)
    assert len(rows[0]) == len(rows[1]) and rows[0][-1] == tok.pad_id and tok.batch_decode(rows) == ["hi", "hello"]
    assert tok.batch_encode(["abcdef"], max_length=4)[0] == tok.encode("abcdef")[:4]
    assert "pad" in repr(SpecialToken("<pad>", 0, SpecialToken.PAD))


def test_profiler_views_by_optype_shape_and_module():
    """ref: hetu.profiler summary views -- op type, op type + shape, op instance, module (subgraph) with children rolled into parents"""
    with ht.graph("define_and_run", create_new=True) as g:
        with ht.subgraph("encoder"):
            with ht.subgraph("block0"):
                a = ht.nn.Linear(16, 32, name="pv_a")
            with ht.subgraph("block1"):
                b = ht.nn.Linear(32, 8, name="pv_b")
            x = ht.placeholder("float32", [4, 16], name="pv_x")
            with ht.subgraph("block0"):
                h = a(x, act="relu")
            with ht.subgraph("block1"):
                y = b(h)
        loss = ht.sum(y)
        with ht.profiler(graph=g) as prof:
            g.run(loss, [loss], {x: torch.randn(4, 16)})
        by_type = dict((k, (t, n)) for k, t, n in prof.summary("optype")["by_optype"])
        assert by_type["linear"][1] == 2
        shapes = [k for k, _, _ in prof.summary("optype_shape")["by_optype_shape"]]
        assert "linear [4, 32]" in shapes and "linear [4, 8]" in shapes
        mods = dict((k, n) for k, _, n in prof.summary("subgraph")["by_subgraph"])
        assert mods.get("encoder", 0) >= mods.get("encoder.block0", 0) + mods.get("encoder.block1", 0) >= 2
        assert any(k.startswith("linear:") for k, _, _ in prof.summary("op")["by_op"])
