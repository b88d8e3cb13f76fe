"""In-place parameter initialisers (PyTorch-style, trailing underscore): fill an existing parameter of the current graph with
values drawn under the global seed.  `fan_in` / `fan_out` follow the [out, in, *kernel] weight convention.
(ref: python/hetu/nn/init.py, nn/functional.py -- uniform_ / normal_ / trunc_normal_ / constant_ / xavier_* / kaiming_* / lecun_*)"""
from __future__ import annotations

import math

import numpy as np
import torch

_counter = [0]


def _rng():
    from .. import core
    _counter[0] += 1
    seed = int(core._global_seed[0])
    return np.random.RandomState((seed * 1000003 + _counter[0]) % (2 ** 31))


def _assign(param, values: np.ndarray):
    from ..core import _graphs_by_id
    t = torch.as_tensor(np.asarray(values, dtype=np.float32)).reshape(list(param.shape))
    g = _graphs_by_id.get(param.graph_id)
    if g is not None:
        g.set_param(param, t)
    if param.eager_data() is not None:
        param.set_eager_data(g.get_param(param) if g is not None else t)
    return param


def _calculate_fan_in_and_fan_out(param):
    shape = list(param.shape)
    if len(shape) < 2:
        raise ValueError("fan in / fan out need at least a 2-D parameter")
    receptive = int(np.prod(shape[2:])) if len(shape) > 2 else 1
    return shape[1] * receptive, shape[0] * receptive


def _calculate_correct_fan(param, mode: str):
    fan_in, fan_out = _calculate_fan_in_and_fan_out(param)
    if mode not in ("fan_in", "fan_out", "avg"):
        raise ValueError(f"mode {mode} not in fan_in / fan_out / avg")
    return fan_in if mode == "fan_in" else fan_out if mode == "fan_out" else (fan_in + fan_out) / 2.0


def calculate_gain(nonlinearity: str, param=None) -> float:
    if nonlinearity in ("linear", "conv1d", "conv2d", "conv3d", "sigmoid"):
        return 1.0
    if nonlinearity == "tanh":
        return 5.0 / 3
    if nonlinearity == "relu":
        return math.sqrt(2.0)
    if nonlinearity == "leaky_relu":
        slope = 0.01 if param is None else float(param)
        return math.sqrt(2.0 / (1 + slope ** 2))
    if nonlinearity == "selu":
        return 3.0 / 4
    raise ValueError(f"unsupported nonlinearity {nonlinearity}")


def uniform_(param, a: float = 0.0, b: float = 1.0):
    return _assign(param, _rng().uniform(a, b, list(param.shape)))


def normal_(param, mean: float = 0.0, std: float = 1.0):
    return _assign(param, _rng().normal(mean, std, list(param.shape)))


def trunc_normal_(param, mean: float = 0.0, std: float = 1.0, a: float = -2.0, b: float = 2.0):
    r = _rng()
    out = r.normal(mean, std, list(param.shape))
    bad = (out < a) | (out > b)
    while bad.any():                                   # resample the tails
        out[bad] = r.normal(mean, std, int(bad.sum()))
        bad = (out < a) | (out > b)
    return _assign(param, out)


def constant_(param, val: float):
    return _assign(param, np.full(list(param.shape), val, np.float32))


def ones_(param):
    return constant_(param, 1.0)


def zeros_(param):
    return constant_(param, 0.0)


def generalized_xavier_(param, dist: str, mode: str, gain: float):
    """variance = gain / fan(mode): the common form behind the Xavier / Kaiming / LeCun families"""
    fan = _calculate_correct_fan(param, mode)
    var = gain / max(fan, 1.0)
    if dist == "uniform":
        lim = math.sqrt(3.0 * var)
        return uniform_(param, -lim, lim)
    if dist == "normal":
        return normal_(param, 0.0, math.sqrt(var))
    raise ValueError(f"dist {dist} not in uniform / normal")


def xavier_uniform_(param, gain: float = 1.0):
    return generalized_xavier_(param, "uniform", "avg", gain * gain)


def xavier_normal_(param, gain: float = 1.0):
    return generalized_xavier_(param, "normal", "avg", gain * gain)


def kaiming_uniform_(param, a: float = 0.0, mode: str = "fan_in", nonlinearity: str = "leaky_relu"):
    return generalized_xavier_(param, "uniform", mode, calculate_gain(nonlinearity, a) ** 2)


def kaiming_normal_(param, a: float = 0.0, mode: str = "fan_in", nonlinearity: str = "leaky_relu"):
    return generalized_xavier_(param, "normal", mode, calculate_gain(nonlinearity, a) ** 2)


def lecun_uniform_(param):
    return generalized_xavier_(param, "uniform", "fan_in", 1.0)


def lecun_normal_(param):
    return generalized_xavier_(param, "normal", "fan_in", 1.0)
