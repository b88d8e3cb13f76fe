from .module import Module, ModuleList, Sequential, ModuleDict  # noqa: F401
from .layers import *  # noqa: F401,F403
from .parallel import *  # noqa: F401,F403
from . import init  # noqa: F401,E402


def Parameter(data, requires_grad: bool = True, name: str = "param", dtype=None):  # noqa: N802
    """a trainable graph variable initialised from an array / tensor (ref: python/hetu/nn/parameter.py Parameter)"""
    import numpy as np
    import torch
    from ..core import parameter, provided_initializer
    arr = data.detach().cpu().numpy() if isinstance(data, torch.Tensor) else np.asarray(data)
    return parameter(provided_initializer(arr), list(arr.shape), dtype=dtype or str(arr.dtype), requires_grad=requires_grad, name=name)
