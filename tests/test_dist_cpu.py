"""Multi-process (gloo, CPU) strategy-equivalence tests: every parallel strategy must reproduce the single-device
loss curve of the same model / data / seed."""
import json
import os

import numpy as np

import pytest

from dist_utils import run_workers

HERE = os.path.dirname(os.path.abspath(__file__))

WORKER = os.path.join(os.path.dirname(__file__), "workers", "gpt_parallel_worker.py")


def _losses(outs):
    for o in outs:
        for line in o.splitlines():
            if line.startswith("LOSSES "):
                return json.loads(line[len("LOSSES "):])
    raise AssertionError("no losses reported:\n" + "\n-----\n".join(outs))


_ref_cache = {}


def _reference(model="gpt"):
    if model not in _ref_cache:
        ok, outs = run_workers(WORKER, 1, [1, 1, 1, 0, 0, 1, model])
        assert ok, outs
        _ref_cache[model] = _losses(outs)
    return _ref_cache[model]


@pytest.mark.dist
@pytest.mark.parametrize("dp,tp,pp,zero,sp,mb", [
    (2, 1, 1, 0, 0, 1),     # data parallel (all-reduce)
    (2, 1, 1, 1, 0, 1),     # ZeRO / OSDP: reduce-scatter grads, sharded Adam, all-gather params
    (1, 2, 1, 0, 0, 1),     # Megatron tensor parallel
    (1, 2, 1, 0, 1, 1),     # + sequence parallel
    (1, 1, 2, 0, 0, 2),     # pipeline (1F1B, 2 micro-batches)
    (2, 2, 1, 1, 1, 1),     # dp2 x tp2 + zero + sp
    (1, 1, 1, 0, 0, 2),     # micro-batch accumulation on one device
    (2, 1, 2, 0, 0, 2),     # dp x pp: tied embedding gradient crosses stages un-reduced, reduced once by the owner
    (2, 2, 2, 1, 1, 2),     # 3-D: dp2 x tp2 x pp2 + zero + sp + micro-batches (8 ranks)
])
def test_strategy_matches_single_device(dp, tp, pp, zero, sp, mb):
    ref = _reference()
    ok, outs = run_workers(WORKER, dp * tp * pp, [dp, tp, pp, zero, sp, mb])
    assert ok, "\n-----\n".join(outs)
    got = _losses(outs)
    assert len(got) == len(ref)
    for a, b in zip(got, ref):
        assert abs(a - b) < 2e-3 * max(1.0, abs(b)), (got, ref)


@pytest.mark.dist
def test_llama_tp2_sp_matches_single_device():
    ref = _reference("llama")
    ok, outs = run_workers(WORKER, 2, [1, 2, 1, 0, 1, 1, "llama"])
    assert ok, "\n-----\n".join(outs)
    got = _losses(outs)
    for a, b in zip(got, ref):
        assert abs(a - b) < 2e-3 * max(1.0, abs(b)), (got, ref)


@pytest.mark.dist
def test_llama_3d_parallel_matches_single_device():
    """Llama (GQA, rotary, SwiGLU, untied head) under dp2 x tp2 x pp2 + ZeRO + sequence parallel + 2 micro-batches"""
    ref = _reference("llama")
    ok, outs = run_workers(WORKER, 8, [2, 2, 2, 1, 1, 2, "llama"])
    assert ok, "\n-----\n".join(outs)
    got = _losses(outs)
    for a, b in zip(got, ref):
        assert abs(a - b) < 2e-3 * max(1.0, abs(b)), (got, ref)


@pytest.mark.dist
def test_uneven_pipeline_layer_split_matches_single_device():
    """Malleus-style layer re-balancing: stage 0 holds 1 layer, stage 1 holds 3 (dp2 x pp2, 2 micro-batches)"""
    ref = _reference("gpt")
    ok, outs = run_workers(WORKER, 4, [2, 1, 2, 0, 0, 2, "gpt", "1,3"])
    assert ok, "\n-----\n".join(outs)
    got = _losses(outs)
    for a, b in zip(got, ref):
        assert abs(a - b) < 2e-3 * max(1.0, abs(b)), (got, ref)


MOE_WORKER = os.path.join(os.path.dirname(__file__), "workers", "moe_worker.py")


@pytest.mark.dist
@pytest.mark.parametrize("gate", ["topk"])
def test_moe_expert_parallel_matches_single_device(gate):
    """GPT-MoE with the experts sharded over the data-parallel ranks (dispatch / combine all-to-all inside the layout
    transform ops) reproduces the single-device loss curve when no token is dropped"""
    ok, outs = run_workers(MOE_WORKER, 1, [1, gate])
    assert ok, "\n".join(outs)
    ref = _losses(outs)
    ok, outs = run_workers(MOE_WORKER, 2, [2, gate])
    assert ok, "\n-----\n".join(outs)
    got = _losses(outs)
    for a, b in zip(got, ref):
        assert abs(a - b) < 2e-3 * max(1.0, abs(b)), (got, ref)


HOT_WORKER = os.path.join(os.path.dirname(__file__), "workers", "hot_switch_worker.py")


@pytest.mark.dist
def test_hot_switching_between_dp_and_tp_keeps_the_loss_curve():
    """HotSPa: parameters and optimizer states are re-sharded in place when the strategy changes between steps"""
    ok, outs = run_workers(HOT_WORKER, 2, ["single"])
    assert ok, "\n-----\n".join(outs)
    ref = _losses(outs)
    log = os.path.join(os.environ.get("TMPDIR", "/tmp"), f"hb_switch_{os.getpid()}.jsonl")
    for r in range(2):
        if os.path.exists(f"{log}.rank{r}"):
            os.remove(f"{log}.rank{r}")
    ok, outs = run_workers(HOT_WORKER, 2, ["switch"], env_extra={"HETU_SWITCH_PROFILE": "TIME", "HETU_SWITCH_LOG_FILE": log})
    assert ok, "\n-----\n".join(outs)
    got = _losses(outs)
    for r in range(2):          # HETU_SWITCH_PROFILE: one record per hot switch and rank
        rows = [json.loads(l) for l in open(f"{log}.rank{r}")]
        assert rows and all(x["rank"] == r and x["from"] != x["to"] and x["switch_ms"] >= 0 for x in rows)
        assert any(x["elems_sent"] > 0 for x in rows)
        os.remove(f"{log}.rank{r}")
    for a, b in zip(got, ref):
        assert abs(a - b) < 2e-3 * max(1.0, abs(b)), (got, ref)


CP_WORKER = os.path.join(os.path.dirname(__file__), "workers", "cp_worker.py")


@pytest.mark.dist
@pytest.mark.parametrize("pattern", ["SYM", "NORMAL"])
def test_context_parallel_ring_attention_matches_full_attention(pattern):
    """ring attention over 2 ranks (zig-zag SYM or contiguous NORMAL split): outputs and dq / dk / dv equal full causal
    attention on the gathered sequence (the worker asserts < 1e-4)"""
    ok, outs = run_workers(CP_WORKER, 2, [2, pattern])
    assert ok, "\n-----\n".join(outs)
    assert sum("CPERR" in o for o in outs) == 2


@pytest.mark.dist
@pytest.mark.parametrize("cp,pattern", [(2, "NORMAL"), (4, "SYM"), (1, "SYM")])
def test_context_parallel_varlen_attention_over_packed_rows(cp, pattern):
    """parallel_attn(cu_seqlens=...): documents of a packed row (crossing chunk and rank borders, one of length 1) attend only to
    themselves; outputs and dq / dk / dv equal per-document causal attention on the gathered row"""
    ok, outs = run_workers(CP_WORKER, cp, [cp, pattern, "varlen"])
    assert ok, "\n-----\n".join(outs)
    assert sum("CPERR" in o for o in outs) == cp


@pytest.mark.dist
def test_ring_attention_analysis_log(tmp_path):
    """HETU_PARALLEL_ATTN=ANALYSIS: one JSON line per ring-attention op and rank with per-round attention time, computed /
    mask-skipped blocks and rotated KV bytes; under the SYM split both ranks compute the same number of blocks"""
    log = str(tmp_path / "ring.jsonl")
    ok, outs = run_workers(CP_WORKER, 2, [2, "SYM"], env_extra={"HETU_PARALLEL_ATTN": "ANALYSIS", "HETU_PARALLEL_ATTN_LOG_FILE": log})
    assert ok, "\n-----\n".join(outs)
    per_rank = []
    for r in range(2):
        rows = [json.loads(l) for l in open(f"{log}.rank{r}")]
        assert rows and all(x["cp"] == 2 and x["pattern"] == "SYM" and len(x["rounds"]) == 2 for x in rows)
        first = rows[0]["rounds"]
        assert first[0]["kv_from"] == r and first[1]["kv_from"] == 1 - r
        assert first[0]["kv_bytes_sent"] > 0 and first[1]["kv_bytes_sent"] == 0
        per_rank.append(sum(x["blocks"] for x in first))
        assert sum(x["blocks"] + x["skipped"] for x in first) == 8          # 2 x 2 chunk pairs per round
    assert per_rank[0] == per_rank[1]                                        # zig-zag split balances the causal work


HETERO_WORKER = os.path.join(os.path.dirname(__file__), "workers", "hetero_worker.py")


@pytest.mark.dist
@pytest.mark.parametrize("layout,world,mb,kind", [("tp2_tp1", 3, 1, "gpt"), ("tp2pp2_tp1", 5, 2, "gpt"), ("tp2to1_tp1", 4, 2, "gpt"),
                                                  ("tp2to1_tp1", 4, 2, "llama")])
def test_heterogeneous_pipelines_reproduce_the_single_device_loss(layout, world, mb, kind):
    """Malleus / Ampelos unions: pipelines with different tp degrees (and stage counts) on unequal batch shares; parameter
    gradients are synchronised per finest common shard (grouped_all_reduce).  `tp2to1_tp1`: the stages of ONE pipeline have
    different tensor-parallel degrees -- activations / gradients cross the stage border through a re-sharding exchange and the
    tied embedding table is re-sharded on its way to the last stage; with 2 micro-batches the exchange runs inside 1F1B."""
    ref = _reference(kind)
    # the (tp2 x pp2) + tp1 case also runs Megatron sequence parallelism inside the tp2 pipeline: norm-weight gradients are
    # first reduced over the pipeline's tp group, then synchronised across pipelines (two chained deferred syncs)
    ok, outs = run_workers(HETERO_WORKER, world, [layout, mb, kind], env_extra={"HETERO_SP": "1"} if layout == "tp2pp2_tp1" else None)
    assert ok, "\n-----\n".join(outs)
    got = _losses(outs)
    for a, b in zip(got, ref):
        assert abs(a - b) < 2e-3 * max(1.0, abs(b)), (got, ref)


@pytest.mark.dist
def test_straggler_report_writes_per_rank_step_breakdown(tmp_path):
    """HETU_STRAGGLER=1: every run appends this rank's step breakdown (attention fwd/bwd, tensor-parallel collectives and
    their traffic, data-parallel gradient reduction, optimizer, other compute) to HETU_STRAGGLER_LOG_FILE.rank<r>"""
    log = str(tmp_path / "straggler.jsonl")
    ok, outs = run_workers(WORKER, 4, [2, 2, 1, 0, 0, 1, "gpt"], env_extra={"HETU_STRAGGLER": "1", "HETU_STRAGGLER_LOG_FILE": log})
    assert ok, "\n-----\n".join(outs)
    for r in range(4):
        rows = [json.loads(l) for l in open(f"{log}.rank{r}")]
        assert len(rows) == 4 and [x["step"] for x in rows] == [1, 2, 3, 4] and all(x["rank"] == r for x in rows)
        last = rows[-1]
        for k in ("attn_fwd_ms", "attn_bwd_ms", "tp_collective_ms", "tp_collective_bytes", "dp_grad_reduce_ms", "optimizer_ms",
                  "other_compute_ms", "compute_ms", "update_ms", "total_ms"):
            assert k in last and last[k] >= 0.0, (k, last)
        assert last["tp_collective_bytes"] > 0 and last["dp_grad_reduce_ms"] > 0 and last["attn_fwd_ms"] > 0
        parts = sum(last[k] for k in ("attn_fwd_ms", "attn_bwd_ms", "tp_collective_ms", "other_compute_ms", "update_ms") if k in last)
        assert parts <= last["total_ms"] * 1.05


GALVATRON_WORKER = os.path.join(os.path.dirname(__file__), "workers", "galvatron_worker.py")


@pytest.mark.dist
@pytest.mark.parametrize("tps,world", [("1,1,2,2", 2), ("1,4,2,1", 4)])
def test_galvatron_layerwise_strategies_reproduce_the_single_device_loss(tps, world):
    """layers of one model under different (tp, dp) degrees on the same devices (the plan format of the Galvatron search):
    activations are relocated at every layout change (dp -> tp: all-gather of the batch shards, tp -> dp: local slice)"""
    ref = _reference()
    ok, outs = run_workers(GALVATRON_WORKER, world, [tps, world])
    assert ok, "\n-----\n".join(outs)
    got = _losses(outs)
    for a, b in zip(got, ref):
        assert abs(a - b) < 2e-3 * max(1.0, abs(b)), (got, ref)


CP_MODEL_WORKER = os.path.join(os.path.dirname(__file__), "workers", "cp_model_worker.py")


@pytest.mark.dist
@pytest.mark.parametrize("cp,tp", [(2, 1), (2, 2)])
def test_whole_model_context_parallel_training_matches_single_device(cp, tp):
    """Llama with every sequence split over a CP ring (SYM chunks): ring attention in every layer, rotary at the original
    positions, parameter gradients reduced over the ring (and TP inside each ring member for cp x tp)"""
    ref = _reference("llama")
    ok, outs = run_workers(CP_MODEL_WORKER, cp * tp, [cp, tp])
    assert ok, "\n-----\n".join(outs)
    got = _losses(outs)
    for a, b in zip(got, ref):
        assert abs(a - b) < 2e-3 * max(1.0, abs(b)), (got, ref)


TRAINER_CP_WORKER = os.path.join(os.path.dirname(__file__), "workers", "trainer_cp_worker.py")


@pytest.mark.dist
def test_trainer_with_context_parallel_strategies(tmp_path):
    """`Trainer` on ds_parallel_configs with cp > 1: rows are cut into each ring member's SYM chunks, the model's ring is
    derived from the device layout [pp][dp][cp][tp], sequence length is symbolic; loss histories equal the cp = 1 run"""
    env = {"TRAINER_OUT": str(tmp_path)}
    ok, outs = run_workers(TRAINER_CP_WORKER, 1, [1, 1, 1], env_extra=env)
    assert ok, "\n-----\n".join(outs)
    ref = _losses(outs)
    for dp, cp, tp in ((2, 2, 1), (1, 2, 2)):
        ok, outs = run_workers(TRAINER_CP_WORKER, dp * cp * tp, [dp, cp, tp], env_extra=env)
        assert ok, "\n-----\n".join(outs)
        got = _losses(outs)
        for a, b in zip(got, ref):
            assert abs(a - b) < 1e-3 * max(1.0, abs(b)), ((dp, cp, tp), got, ref)
    # packing x context parallelism: packed rows (documents crossing chunk borders) through variable-length ring attention
    ok, outs = run_workers(TRAINER_CP_WORKER, 1, [1, 1, 1, "pack"], env_extra=env)
    assert ok, "\n-----\n".join(outs)
    ref = _losses(outs)
    ok, outs = run_workers(TRAINER_CP_WORKER, 2, [1, 2, 1, "pack"], env_extra=env)
    assert ok, "\n-----\n".join(outs)
    for a, b in zip(_losses(outs), ref):
        assert abs(a - b) < 2e-4 * max(1.0, abs(b)), (_losses(outs), ref)


TRAINER_HETERO_WORKER = os.path.join(os.path.dirname(__file__), "workers", "trainer_hetero_worker.py")


@pytest.mark.dist
def test_trainer_on_a_heterogeneous_strategy(tmp_path):
    """`Trainer(ds_parallel_configs=[hetero config], hetero_shares=[3, 1])`: (tp2 x pp2) + tp1 pipelines, every pipeline takes
    its share of each global batch; the loss history equals the single-device Trainer"""
    env = {"TRAINER_OUT": str(tmp_path)}
    ok, outs = run_workers(TRAINER_HETERO_WORKER, 1, ["single"], env_extra=env)
    assert ok, "\n-----\n".join(outs)
    ref = _losses(outs)
    ok, outs = run_workers(TRAINER_HETERO_WORKER, 5, ["hetero"], env_extra=env)
    assert ok, "\n-----\n".join(outs)
    got = _losses(outs)
    assert sum("INFO" in o and "True" in o for o in outs) == 5
    for a, b in zip(got, ref):
        assert abs(a - b) < 1e-3 * max(1.0, abs(b)), (got, ref)


@pytest.mark.dist
@pytest.mark.parametrize("worker,mode,world", [("trainer_rebuild_worker.py", "rebuild", 3), ("malleus_apply_worker.py", "malleus", 4),
                                               ("malleus_apply_worker.py", "shrink", 4)])
def test_replanning_a_running_job_keeps_the_loss_curve(tmp_path, worker, mode, world):
    """Trainer.rebuild (explicit) and MalleusTrainer(auto_apply=True) (straggler report -> plan -> rebuild inside the training
    loop): the job moves from a homogeneous strategy to a heterogeneous one through a split checkpoint; parameters, Adam states,
    step counters and the data position carry over, so the 4-step loss curve equals the single-device one.  `shrink`: a device
    becomes 50x slower, the planner dissolves its tensor-parallel group -- one tp2 pipeline over the healthy devices, the other
    two ranks idle (they own no graph but keep taking part in group creation and barriers)"""
    path = os.path.join(os.path.dirname(__file__), "workers", worker)
    ok, outs = run_workers(path, 1, ["single"], env_extra={"TRAINER_OUT": str(tmp_path / "single")})
    assert ok, "\n-----\n".join(outs)
    ref = _losses(outs)
    ok, outs = run_workers(path, world, [mode], env_extra={"TRAINER_OUT": str(tmp_path / mode)})
    assert ok, "\n-----\n".join(outs)
    got = _losses(outs)
    for a, b in zip(got, ref):
        assert abs(a - b) < 1e-3 * max(1.0, abs(b)), (got, ref)


@pytest.mark.dist
def test_grouped_slice_collectives_over_heterogeneous_holders():
    """grouped_all_reduce / grouped_reduce_scatter / grouped_all_gather (the Split* collectives of heterogeneous DP): a parameter
    held as 2 shards by a tp2 pipeline and whole by a tp1 pipeline is reduced / scattered / gathered slice by slice"""
    ok, outs = run_workers(os.path.join(os.path.dirname(__file__), "workers", "grouped_comm_worker.py"), 3, [])
    assert ok, "\n-----\n".join(outs)
    assert sum("GROUPED" in o and "True" in o for o in outs) == 3


@pytest.mark.dist
def test_single_strategy_parallel_modules_are_tp_invariant():
    """hetu.nn.{VocabParallelEmbedding, ColumnParallelLinear, RowParallelLinear, ParallelLayerNorm}(device_group, dp): outputs and
    the effect of one SGD step are the same under tp = 1 and tp = 2 (the reference's tests/test_parallel.py scenario)"""
    worker = os.path.join(os.path.dirname(__file__), "workers", "parallel_modules_worker.py")
    res = []
    for tp in (1, 2):
        ok, outs = run_workers(worker, tp, [tp])
        assert ok, "\n-----\n".join(outs)
        line = next(l for o in outs for l in o.splitlines() if l.startswith("PM "))
        res.append(json.loads(line[3:]))
    a, b = res
    assert abs(a["l0"] - b["l0"]) < 1e-5 and abs(a["l1"] - b["l1"]) < 1e-5 and a["l1"] < a["l0"]
    for k in ("y0", "y1"):
        assert max(abs(x - y) for x, y in zip(a[k], b[k])) < 1e-4


@pytest.mark.dist
@pytest.mark.parametrize("world,local", [(4, 2), (6, 3), (8, 2)])
def test_hierarchical_all_to_all_equals_flat(world, local):
    """two-level all-to-all (node group, layout transform, rail group) == flat all-to-all, values and gradient"""
    ok, outs = run_workers(os.path.join(os.path.dirname(__file__), "workers", "hall_to_all_worker.py"), world, [local])
    assert ok, "\n-----\n".join(outs)
    line = [l for o in outs for l in o.splitlines() if l.startswith("HA2A ")][0]
    assert json.loads(line[5:])["ok"]


@pytest.mark.dist
@pytest.mark.parametrize("zero", [0, 1])
def test_overlapped_gradient_reduce_matches_the_serial_path(zero):
    """HETU_OVERLAP_GRAD_REDUCE: the data-parallel all-reduce / ZeRO reduce-scatter of every parameter is launched from inside
    backward as soon as its gradient is final and only awaited by the update phase -- same loss curve as the serial path"""
    ref = _reference()
    ok, outs = run_workers(WORKER, 2, [2, 1, 1, zero, 0, 1], env_extra={"HETU_OVERLAP_GRAD_REDUCE": "ON"})
    assert ok, "\n-----\n".join(outs)
    got = _losses(outs)
    for a, b in zip(got, ref):
        assert abs(a - b) < 2e-3 * max(1.0, abs(b)), (got, ref)


@pytest.mark.dist
def test_comm_group_api_coalesce_reduce_gather_scatter():
    ok, outs = run_workers(os.path.join(os.path.dirname(__file__), "workers", "comm_api_worker.py"), 3)
    assert ok, "\n-----\n".join(outs)
    line = [l for o in outs for l in o.splitlines() if l.startswith("COMMAPI ")][0]
    assert json.loads(line[8:])["ok"]


@pytest.mark.parametrize("p,c,comm", [(4, 1, "groups"), (4, 2, "groups"), (6, 3, "groups"), (4, 2, "world"), (8, 2, "groups")])
def test_distgcn_15d_training_matches_single_process(p, c, comm):
    """ref: hetu/v1 DistGCN_15d -- p ranks, replication c: every step's loss and the final weights equal plain single-process training
    of the same GCN.  The p ranks run as threads of this process over `ThreadCollectives` (same all-gather / all-reduce interface as
    the process-group transport), so the partitioned forward + backward is checked without a network transport;
    tests/workers/distgcn_worker.py is the same run over real process groups (torchrun)."""
    import threading
    import torch
    from hetu_b200.models.gnn import DistGCN15D, ThreadCollectives, normalise_adjacency
    rng = np.random.RandomState(0)
    n, f, hdim, k, steps = 96, 12, 16, 4, 6
    edges = rng.randint(0, n, (2, 400))
    idx, val = normalise_adjacency(edges, n)
    x = rng.randn(n, f).astype(np.float32)
    y = rng.randint(0, k, n)
    mask = rng.rand(n) < 0.7
    ends = ThreadCollectives.world(p)
    models = [DistGCN15D(idx, val, x, y, [f, hdim, k], r, p, c, lr=0.5, seed=3, train_mask=mask, comm=comm, collectives=ends[r])
              for r in range(p)]
    losses, errors = [None] * p, []

    def run(r):
        try:
            losses[r] = [models[r].step() for _ in range(steps)]
        except Exception as e:                                   # noqa: BLE001 - reported below
            errors.append((r, repr(e)))
    threads = [threading.Thread(target=run, args=(r,), daemon=True) for r in range(p)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(120)
    assert not errors and all(l is not None for l in losses), errors
    A = torch.sparse_coo_tensor(torch.as_tensor(idx), torch.as_tensor(val), (n, n)).to_dense()
    g = torch.Generator().manual_seed(3)
    ws = [(torch.randn(a, b, generator=g) * (1.0 / np.sqrt(a))).requires_grad_() for a, b in ((f, hdim), (hdim, k))]
    ref = []
    X, Y, M = torch.as_tensor(x), torch.as_tensor(y), torch.as_tensor(mask)
    for _ in range(steps):
        h = torch.relu(A @ (X @ ws[0]))
        z = A @ (h @ ws[1])
        l = torch.nn.functional.cross_entropy(z[M], Y[M])
        ref.append(float(l))
        gs = torch.autograd.grad(l, ws)
        with torch.no_grad():
            for w, gg in zip(ws, gs):
                w -= 0.5 * gg
    for r in range(p):
        np.testing.assert_allclose(losses[r], ref, rtol=2e-4, atol=2e-5)
        assert max(float((a - b.detach()).abs().max()) for a, b in zip(models[r].weights, ws)) < 2e-4
    assert ref[-1] < ref[0]
    # replication trades memory for traffic: per step a rank gathers n/c feature rows and reduces n*c/p -- below the 1-D n rows
    # once 1/c + c/p < 1
    if comm == "groups" and c > 1 and c * c <= p:
        base = [DistGCN15D(idx, val, x, y, [f, hdim, k], r, p, 1, lr=0.5, seed=3, train_mask=mask, collectives=e)
                for r, e in enumerate(ThreadCollectives.world(p))]
        ts = [threading.Thread(target=base[r].step, daemon=True) for r in range(p)]
        [t.start() for t in ts]
        [t.join(60) for t in ts]
        ratio = (models[0].bytes_moved / steps) / base[0].bytes_moved
        assert abs(ratio - (1.0 / c + c / p)) < 1e-6
        assert (ratio < 1.0) == (1.0 / c + c / p < 1.0 - 1e-9)
