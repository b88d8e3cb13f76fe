"""Multi-process (gloo, CPU) strategy-equivalence tests: every parallel strategy must reproduce the single-device
loss curve of the same model / data / seed."""
import json
import os

import pytest

from dist_utils import run_workers

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
    ok, outs = run_workers(HOT_WORKER, 2, ["switch"])
    assert ok, "\n-----\n".join(outs)
    got = _losses(outs)
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


HETERO_WORKER = os.path.join(os.path.dirname(__file__), "workers", "hetero_worker.py")


@pytest.mark.dist
@pytest.mark.parametrize("layout,world", [("tp2_tp1", 3), ("tp2pp2_tp1", 5)])
def test_heterogeneous_pipelines_reproduce_the_single_device_loss(layout, world):
    """Malleus / Ampelos unions: pipelines with different tp degrees (and stage counts) on unequal batch shares; parameter
    gradients are synchronised per finest common shard (grouped_all_reduce)"""
    ref = _reference()
    ok, outs = run_workers(HETERO_WORKER, world, [layout])
    assert ok, "\n-----\n".join(outs)
    got = _losses(outs)
    for a, b in zip(got, ref):
        assert abs(a - b) < 2e-3 * max(1.0, abs(b)), (got, ref)
