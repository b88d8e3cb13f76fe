"""The executor features that round 1 only validated on CPU / gloo, run on B200s over NCCL (needs >= 2 GPUs; the 4-rank cases
need 4): pipeline parallel 1F1B, dp x tp x pp with ZeRO + sequence parallel, hot strategy switching, ring (context-parallel)
attention incl. packed variable-length rows, heterogeneous pipelines.  Same workers and tolerances as tests/test_dist_cpu.py."""
import json
import os

import pytest
import torch

from dist_utils import run_workers

pytestmark = pytest.mark.gpu
W = os.path.join(os.path.dirname(__file__), "workers")


def _gpus():
    return torch.cuda.device_count()


def _losses(outs):
    for o in outs:
        for line in o.splitlines():
            if line.startswith("LOSSES "):
                return json.loads(line[len("LOSSES "):])
    raise AssertionError("no losses reported:\n" + "\n-----\n".join(outs))


def _need(n):
    if _gpus() < n:
        pytest.skip(f"needs {n} GPUs")


_ref = {}


def _reference(model="gpt"):
    if model not in _ref:
        ok, outs = run_workers(os.path.join(W, "gpt_parallel_worker.py"), 1, [1, 1, 1, 0, 0, 1, model], force_cpu=False, timeout=600)
        assert ok, outs
        _ref[model] = _losses(outs)
    return _ref[model]


@pytest.mark.parametrize("dp,tp,pp,zero,sp,mb,model", [
    (1, 1, 2, 0, 0, 2, "gpt"),        # pipeline parallel, 1F1B with 2 micro-batches, NCCL p2p
    (2, 1, 1, 1, 0, 1, "gpt"),        # ZeRO
    (1, 2, 2, 1, 1, 2, "llama"),      # tp2 x pp2 + sequence parallel (the BASELINE config #3 shape on 4 GPUs)
    (2, 1, 2, 1, 0, 2, "gpt"),        # dp2 x pp2 + ZeRO inside a pipeline stage group
])
def test_strategies_on_gpus_match_the_single_gpu_loss(dp, tp, pp, zero, sp, mb, model):
    _need(dp * tp * pp)
    ref = _reference(model)
    ok, outs = run_workers(os.path.join(W, "gpt_parallel_worker.py"), dp * tp * pp, [dp, tp, pp, zero, sp, mb, model], force_cpu=False,
                           timeout=600)
    assert ok, "\n-----\n".join(outs)
    got = _losses(outs)
    for a, b in zip(got, ref):
        assert abs(a - b) < 5e-3 * max(1.0, abs(b)), (got, ref)


def test_hot_switching_on_gpus_keeps_the_loss_curve(tmp_path):
    _need(2)
    ok, outs = run_workers(os.path.join(W, "hot_switch_worker.py"), 2, ["single"], force_cpu=False, timeout=600)
    assert ok, "\n-----\n".join(outs)
    ref = _losses(outs)
    log = str(tmp_path / "switch.jsonl")
    ok, outs = run_workers(os.path.join(W, "hot_switch_worker.py"), 2, ["switch"], force_cpu=False, timeout=600,
                           env_extra={"HETU_SWITCH_PROFILE": "TIME", "HETU_SWITCH_LOG_FILE": log})
    assert ok, "\n-----\n".join(outs)
    for a, b in zip(_losses(outs), ref):
        assert abs(a - b) < 5e-3 * max(1.0, abs(b))
    rows = [json.loads(l) for r in range(2) for l in open(f"{log}.rank{r}")]
    assert rows and any(x["elems_sent"] > 0 for x in rows)
    print("HOTSWITCH " + json.dumps({"switches": len(rows), "max_switch_ms": max(x["switch_ms"] for x in rows)}))


@pytest.mark.parametrize("cp,pattern,varlen", [(2, "SYM", False), (2, "NORMAL", True), (4, "SYM", True)])
def test_ring_attention_on_gpus(cp, pattern, varlen):
    _need(cp)
    args = [cp, pattern] + (["varlen"] if varlen else [])
    ok, outs = run_workers(os.path.join(W, "cp_worker.py"), cp, args, force_cpu=False, timeout=600)
    assert ok, "\n-----\n".join(outs)
    assert sum("CPERR" in o for o in outs) == cp
