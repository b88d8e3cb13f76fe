"""Fused compute + collective paths on real GPUs (>= 2): every worker trains the same graph twice -- with the fused
symmetric-memory kernels and with the NCCL twin -- and the loss curves / parameters must agree."""
import json
import os

import pytest
import torch

from dist_utils import run_workers

pytestmark = pytest.mark.gpu
W = os.path.join(os.path.dirname(__file__), "workers")


def _world():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    return 2 if n < 4 else (4 if n < 8 else 8)


def _line(outs, tag):
    for o in outs:
        for l in o.splitlines():
            if l.startswith(tag + " "):
                return json.loads(l[len(tag) + 1:])
    raise AssertionError(f"no {tag} line:\n" + "\n-----\n".join(outs))


def test_fused_zero_matches_nccl_path():
    world = _world()
    ok, outs = run_workers(os.path.join(W, "zero_fused_worker.py"), world, force_cpu=False, timeout=900)
    assert ok, "\n-----\n".join(outs)
    r = _line(outs, "ZEROFUSED")
    assert r["ranks_identical"]
    assert r["max_param_diff"] < 0.02 and r["max_param_diff_accum"] < 0.02
    for a, b in zip(r["fused_losses"], r["ref_losses"]):
        assert abs(a - b) < 5e-3 * max(1.0, abs(b))


@pytest.mark.parametrize("model", ["gpt", "llama"])
def test_fused_tensor_parallel_gemm_collectives_match_nccl_path(model):
    world = 2 if _world() >= 2 else 0
    ok, outs = run_workers(os.path.join(W, "tp_fused_worker.py"), world, [world, model], force_cpu=False, timeout=900)
    assert ok, "\n-----\n".join(outs)
    r = _line(outs, "TPFUSED")
    assert r["symm_kernel_launches_fused_run"] > 0
    assert r["max_param_diff"] < 0.02
    for a, b in zip(r["fused_losses"], r["ref_losses"]):
        assert abs(a - b) < 5e-3 * max(1.0, abs(b))


def test_fused_expert_parallel_all_to_all_matches_nccl_path():
    world = _world()
    ok, outs = run_workers(os.path.join(W, "moe_fused_worker.py"), world, force_cpu=False, timeout=900)
    assert ok, "\n-----\n".join(outs)
    r = _line(outs, "MOEFUSED")
    assert r["symm_launches"] > 0
    for a, b in zip(r["fused_losses"], r["ref_losses"]):
        assert abs(a - b) < 5e-3 * max(1.0, abs(b))
