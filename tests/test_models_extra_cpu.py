"""CTR models, GNN (spmm, 1.5-D partition), embedding compression methods."""
import numpy as np
import pytest
import torch

import hetu_b200 as ht
from hetu_b200.models import DCN, GCN, WDL, DeepFM, dist_gcn_15d_forward, normalise_adjacency, partition_15d
from hetu_b200.tools.emb_compress import METHODS, build_compressed_embedding


@pytest.mark.parametrize("cls", [WDL, DeepFM, DCN])
def test_ctr_models_learn_a_synthetic_rule(cls):
    rng = np.random.RandomState(0)
    B, F, D, N = 128, 4, 8, 50
    dense = rng.randn(B, 3).astype(np.float32)
    sparse = rng.randint(0, N, (B, F)) + np.arange(F) * 0
    label = ((dense[:, :1] + (sparse[:, :1] % 2) * 2.0 - 1.0) > 0).astype(np.float32)
    with ht.graph("define_and_run", create_new=True) as g:
        model = cls(N, D, num_fields=F, num_dense=3, hidden=(32, 32))
        d = ht.placeholder("float32", [B, 3], name="dense")
        s = ht.placeholder("int64", [B, F], name="sparse")
        y = ht.placeholder("float32", [B, 1], name="label")
        loss, logit = model(d, s, y)
        train = ht.AdamOptimizer(lr=0.02).minimize(loss)
    feed = {d: torch.as_tensor(dense), s: torch.as_tensor(sparse), y: torch.as_tensor(label)}
    losses = [float(g.run(loss, [loss, train], feed)[0]) for _ in range(60)]
    assert losses[-1] < 0.6 * losses[0], (cls.__name__, losses[0], losses[-1])


def test_gcn_trains_and_15d_partition_reproduces_the_dense_product():
    rng = np.random.RandomState(1)
    n, f, c = 40, 8, 3
    edges = rng.randint(0, n, (2, 120))
    idx, val = normalise_adjacency(edges, n)
    x = rng.randn(n, f).astype(np.float32)
    labels = (np.arange(n) % c).astype(np.int64)
    with ht.graph("define_and_run", create_new=True) as g:
        model = GCN(f, 16, c)
        I, V = ht.from_numpy(torch.as_tensor(idx)), ht.from_numpy(torch.as_tensor(val))
        X = ht.placeholder("float32", [n, f], name="x")
        Y = ht.placeholder("int64", [n], name="y")
        loss, _ = model(I, V, X, n, Y)
        train = ht.AdamOptimizer(lr=0.05).minimize(loss)
    feed = {X: torch.as_tensor(x), Y: torch.as_tensor(labels)}
    losses = [float(g.run(loss, [loss, train], feed)[0]) for _ in range(40)]
    assert losses[-1] < 0.7 * losses[0]
    # 1.5-D execution == dense A (X W)
    w = rng.randn(f, 5).astype(np.float32)
    A = np.zeros((n, n), np.float32)
    np.add.at(A, (idx[0], idx[1]), val)
    ref = A @ (x @ w)
    for p, rep in [(4, 2), (8, 4), (6, 1)]:
        np.testing.assert_allclose(dist_gcn_15d_forward(idx, val, x, w, p, rep), ref, rtol=1e-4, atol=1e-4)
        parts = partition_15d(n, p, rep)
        assert len(parts) == p and all(len(d["replica_group"]) == rep for d in parts)


@pytest.mark.parametrize("method,kw", [("hash", {"buckets": 64}), ("qr", {}), ("tt", {"rank": 4}), ("dhe", {"num_hashes": 16, "hidden": (32,)}),
                                       ("robe", {"array_size": 1024, "chunk": 4}), ("dpq", {"subspaces": 2, "codes": 16}),
                                       ("mgqe", {"subspaces": 2, "codes": 16, "rare_codes": 4}), ("mde", {}), ("autodim", {"candidates": (2, 4)}),
                                       ("pep", {}), ("deeplight", {}), ("optembed", {}), ("alpt", {}), ("adapt", {"hot": 50, "buckets": 64}),
                                       ("cafe", {"hot": 50, "buckets": 64})])
def test_embedding_compression_methods_forward_backward(method, kw):
    N, D = 1000, 8
    ids = torch.as_tensor(np.random.RandomState(0).randint(0, N, (6, 3)))
    with ht.graph("define_and_run", create_new=True) as g:
        emb = build_compressed_embedding(method, N, D, **kw)
        I = ht.placeholder("int64", [6, 3], name="ids")
        e = emb(I)
        loss = ht.mean(e * e, [0, 1, 2])
        train = ht.SGDOptimizer(lr=0.1).minimize(loss)
    out = g.run(loss, [e, loss, train], {I: ids})
    assert tuple(out[0].shape) == (6, 3, D) and np.isfinite(float(out[1]))
    l2 = float(g.run(loss, [loss, train], {I: ids})[0])
    assert l2 <= float(out[1]) + 1e-6          # a gradient step on ||e||^2 never increases it
    assert emb.compression_ratio() > 0
    if method in ("hash", "qr", "tt", "dhe", "robe", "adapt"):
        assert emb.compression_ratio() > 1.5


def test_autosrh_dedup_quantized_and_sparse_embeddings():
    """ref: tools/EmbeddingMemoryCompression/methods/layers/{autosrh,deduplication,quantize,sparse}.py"""
    rng = np.random.RandomState(0)
    N, D = 64, 8
    ids_np = rng.randint(0, N, (5, 3))
    ids = torch.as_tensor(ids_np)
    # AutoSrh: gates train with the table; after retrain() only the kept gates pass and the size accounts for the mask
    with ht.graph("define_and_run", create_new=True) as g:
        emb = build_compressed_embedding("autosrh", N, D, nsplit=4)
        I = ht.placeholder("int64", [5, 3], name="ids")
        e = emb(I)
        loss = ht.mean(e * e, [0, 1, 2]) + emb.l1_penalty() * 1e-3
        train = ht.SGDOptimizer(lr=0.1).minimize(loss)
        a0 = g.get_param(emb.alpha).clone()
        g.run(loss, [loss, train], {I: ids})
        assert not torch.equal(g.get_param(emb.alpha), a0)
        full = emb.num_parameters()
        alpha = rng.rand(4, D).astype(np.float32)
        kept = emb.retrain(alpha, keep_rate=0.25)
        assert abs(kept - 0.25) < 0.05 and emb.num_parameters() < 0.4 * full
        masked = g.run(None, [emb(I)], {I: ids})[0]
        w = g.get_param(emb.weight)
        thr = np.sort(alpha.reshape(-1))[-8]
        want = w[ids_np.reshape(-1)].numpy() * (alpha >= thr)[emb.group_np[ids_np.reshape(-1)]]
        np.testing.assert_allclose(masked.reshape(-1, D).numpy(), want, rtol=1e-6, atol=1e-7)
    # Dedup: identical blocks are stored once and every id still reads its original row
    table = rng.randn(N, D).astype(np.float32)
    table[8:12] = table[0:4]
    table[40:44] = table[0:4]
    with ht.graph("define_and_run", create_new=True) as g:
        emb = build_compressed_embedding("dedup", N, D, table=table, nemb_per_block=4)
        I = ht.placeholder("int64", [5, 3], name="ids")
        probe = torch.as_tensor(np.array([[0, 9, 42], [3, 11, 43], [5, 20, 63], [8, 40, 1], [10, 41, 2]]))
        out = g.run(None, [emb(I)], {I: probe})[0]
        np.testing.assert_allclose(out.numpy(), table[probe.numpy()], rtol=1e-6)
        assert emb.weight.shape[0] == N - 8 and emb.compression_ratio() > 1.1 and emb.remap_np[2] == emb.remap_np[10] == 0
    # Quantized: outputs sit on the grid, gradients pass straight through to the table
    for kw in ({"digit": 8, "scale": 0.01, "middle": 0.0}, {"digit": 16, "scale": 1e-4}, {"digit": 8, "use_qparam": True}):
        with ht.graph("define_and_run", create_new=True) as g:
            emb = build_compressed_embedding("quantize", N, D, **kw)
            I = ht.placeholder("int64", [5, 3], name="ids")
            e = emb(I)
            loss = ht.sum(e, [0, 1, 2])
            gw = ht.gradients(loss, [emb.weight])[0]
            out, grad = g.run(loss, [e, gw], {I: ids})
            w = g.get_param(emb.weight).numpy()[ids_np.reshape(-1)]
            got = out.reshape(-1, D).numpy()
            if kw.get("use_qparam"):
                step = (w.max(1, keepdims=True) - w.min(1, keepdims=True)) / 255.0
                assert np.abs(got - w).max() <= step.max() / 2 + 1e-6
                k = (got - w.min(1, keepdims=True)) / step
            else:
                assert np.abs(got - w).max() <= kw["scale"] / 2 + 1e-6
                k = got / kw["scale"]
            assert np.abs(k - np.round(k)).max() < 1e-2
            counts = np.bincount(ids_np.reshape(-1), minlength=N).astype(np.float32)
            np.testing.assert_allclose(grad.numpy(), np.repeat(counts[:, None], D, 1), rtol=1e-6)
            assert emb.compression_ratio() > (3.5 if kw["digit"] == 8 and not kw.get("use_qparam") else 1.9)
    # Sparse: CSR storage of a pruned table, the graph lookup and the host gather agree
    pruned = table * (rng.rand(N, D) < 0.2)
    with ht.graph("define_and_run", create_new=True) as g:
        emb = build_compressed_embedding("sparse", N, D, table=pruned)
        I = ht.placeholder("int64", [5, 3], name="ids")
        out = g.run(None, [emb(I)], {I: ids})[0]
        np.testing.assert_allclose(out.numpy(), pruned[ids_np], rtol=1e-6)
        np.testing.assert_allclose(emb.rows(ids_np), pruned[ids_np], rtol=1e-6)
        assert emb.nnz == int((pruned != 0).sum()) and emb.compression_ratio() > 1.5


def test_cafe_hot_sketch_promotes_frequent_ids():
    with ht.graph("define_and_run", create_new=True):
        emb = METHODS["cafe"](1000, 8, hot=4, buckets=32)
    for _ in range(5):
        emb.observe(np.array([7, 7, 7, 9, 9, 500]), np.ones(6))
    r = emb.remap(np.array([7, 9, 123]))
    assert r[0] < 4 and r[1] < 4 and r[2] >= 4
