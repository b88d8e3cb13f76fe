"""Import-time compatibility for running the UNMODIFIED reference runtime (tools/Galvatron, written against
torch 2.0.1) on the image's torch 2.11.  Galvatron's pipeline module does `from torch.distributed.fsdp... import`
of private FSDP names that were renamed / moved since 2.0.1; the import fails before any user code runs.
This file only adds the old names back to torch's namespace (aliases to the renamed objects).  It does not
touch the reference sources and nothing here is used by hetu_b200.

Only the names needed to *import* galvatron.core.pipeline are provided; the functions that use them
(`fsdp_reduce_gradients` for pp>1 asynchronous gradient reduction) are not on the pp=1 path the benchmark times.
"""
import sys
import types


def apply():
    import torch  # noqa: F401
    import torch.distributed.fsdp._common_utils as cu
    import torch.distributed.fsdp._flat_param as flat_param
    import torch.distributed.fsdp._runtime_utils as ru
    import torch.distributed.utils as du

    # torch 2.0.1: torch.distributed.fsdp.flat_param  ->  2.1+: torch.distributed.fsdp._flat_param
    sys.modules.setdefault("torch.distributed.fsdp.flat_param", flat_param)
    # torch 2.0.1: torch.distributed.fsdp._utils {p_assert, _no_dispatch_record_stream}
    if "torch.distributed.fsdp._utils" not in sys.modules:
        m = types.ModuleType("torch.distributed.fsdp._utils")
        m.p_assert = du._p_assert
        m._no_dispatch_record_stream = cu._no_dispatch_record_stream
        sys.modules["torch.distributed.fsdp._utils"] = m
    # removed in 2.1 (the check moved into the hook registration)
    if not hasattr(ru, "_check_comm_hook"):
        def _check_comm_hook(comm_hook, comm_hook_state):
            assert comm_hook is not None, "Communication hook should not be None"
        ru._check_comm_hook = _check_comm_hook
