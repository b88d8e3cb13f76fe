import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu)")
    config.addinivalue_line("markers", "dist: multi-process test (gloo on CPU, nccl on GPU)")


def _ensure_built():
    import glob
    if not glob.glob(os.path.join(ROOT, "hetu_b200", "_C*.so")):
        import subprocess
        subprocess.check_call([sys.executable, os.path.join(ROOT, "build.py")], cwd=ROOT)


_ensure_built()


@pytest.fixture(autouse=True)
def _gpu_strict(request):
    """GPU tests must run on the hand-written kernels: a silent ATen fallback on a CUDA bf16 hot op is an error."""
    if request.node.get_closest_marker("gpu") is not None:
        import torch
        if not torch.cuda.is_available():
            pytest.skip("no GPU")
        os.environ["HETU_B200_STRICT"] = "1"
    yield
