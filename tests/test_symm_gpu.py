"""Symmetric-memory collectives and the fused GEMM->reduce-scatter kernel vs NCCL (needs >= 2 GPUs)."""
import json
import os

import pytest
import torch

from dist_utils import run_workers

pytestmark = pytest.mark.gpu
WORKER = os.path.join(os.path.dirname(__file__), "workers", "symm_worker.py")


def test_symmetric_collectives_match_nccl():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else (4 if n < 8 else 8)
    ok, outs = run_workers(WORKER, world, force_cpu=False, timeout=900)
    assert ok, "\n-----\n".join(outs)
    line = [l for o in outs for l in o.splitlines() if l.startswith("SYMM ")][0]
    r = json.loads(line[5:])
    assert r["all_gather_ok"] and r["all_to_all_ok"]
    assert r["reduce_scatter_err"] < 0.05 and r["all_reduce_err"] < 0.05
    assert r["gemm_rs_err"] < 0.02 * max(1.0, r["gemm_rs_ref_max"])
    if r.get("multicast_supported"):
        # NVLS path (cuMulticast + multimem.*): must map on NVSwitch systems and agree with NCCL
        assert r.get("multicast_mapped"), r.get("multicast_error")
        assert r["mc_all_gather_ok"] and r["mc_reduce_scatter_err"] < 0.05 and r["mc_all_reduce_err"] < 0.05
    print(json.dumps(r, indent=1))
