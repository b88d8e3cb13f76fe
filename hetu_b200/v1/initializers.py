"""v1 initializers: `init.xavier_normal(shape, name=...)` creates the Variable, `init.GenXavierNormal()` returns a factory a
layer calls later with the shape (ref: hetu/v1/python/hetu/initializers.py)."""
from __future__ import annotations

import math

import numpy as np

from .. import core


class BaseInit:
    def __init__(self, shape=None):
        self.shape = tuple(shape) if shape is not None else None

    def initializer(self):
        raise NotImplementedError

    def __call__(self, shape=None, name=None, trainable=True, dtype="float32", ctx=None):
        from .executor import Variable
        return Variable(name or "var", initializer=self.initializer(), shape=list(shape or self.shape), trainable=trainable, dtype=dtype)


class EmptyInit(BaseInit):
    def initializer(self): return core.zeros_initializer()                       # noqa: E704


class ConstantInit(BaseInit):
    def __init__(self, constant=0.0, shape=None):
        super().__init__(shape)
        self.constant = float(constant)

    def initializer(self): return core.constant_initializer(self.constant)       # noqa: E704


class ZerosInit(ConstantInit):
    def __init__(self, shape=None): super().__init__(0.0, shape)                  # noqa: E704


class OnesInit(ConstantInit):
    def __init__(self, shape=None): super().__init__(1.0, shape)                  # noqa: E704


class UniformInit(BaseInit):
    def __init__(self, low=-1.0, high=1.0, shape=None):
        super().__init__(shape)
        self.low, self.high = float(low), float(high)

    def initializer(self): return core.uniform_initializer(self.low, self.high)   # noqa: E704


class NormalInit(BaseInit):
    def __init__(self, mean=0.0, stddev=1.0, shape=None):
        super().__init__(shape)
        self.mean, self.stddev = float(mean), float(stddev)

    def initializer(self): return core.normal_initializer(self.mean, self.stddev)  # noqa: E704


class TruncatedNormalInit(NormalInit):
    def initializer(self): return core.truncated_normal_initializer(self.mean, self.stddev)   # noqa: E704


def _fans(shape):
    """dense weights are [in, out]; convolution filters [out_c, in_c, kh, kw]"""
    shape = list(shape)
    if len(shape) < 2:
        return shape[0], shape[0]
    if len(shape) == 2:
        return shape[0], shape[1]
    rf = int(np.prod(shape[2:]))
    return shape[1] * rf, shape[0] * rf


class _Xavier(BaseInit):
    normal = True

    def __init__(self, gain=1.0, mode="avg", shape=None):
        super().__init__(shape)
        assert mode in ("fan_in", "fan_out", "avg")
        self.gain, self.mode = float(gain), mode
        self._for = None

    def __call__(self, shape=None, name=None, trainable=True, dtype="float32", ctx=None):
        self._for = list(shape or self.shape)
        return super().__call__(shape, name, trainable, dtype, ctx)

    def initializer(self):
        fi, fo = _fans(self._for or self.shape)
        fan = {"fan_in": fi, "fan_out": fo, "avg": (fi + fo) / 2.0}[self.mode]
        if self.normal:
            return core.normal_initializer(0.0, math.sqrt(self.gain / fan))
        lim = math.sqrt(3.0 * self.gain / fan)
        return core.uniform_initializer(-lim, lim)


class GeneralXavierNormalInit(_Xavier):
    normal = True


class GeneralXavierUniformInit(_Xavier):
    normal = False


def _mk(cls, gain, mode):
    class _C(cls):
        def __init__(self, shape=None): super().__init__(gain, mode, shape)       # noqa: E704
    return _C


XavierNormalInit, XavierUniformInit = _mk(GeneralXavierNormalInit, 1.0, "avg"), _mk(GeneralXavierUniformInit, 1.0, "avg")
HeNormalInit, HeUniformInit = _mk(GeneralXavierNormalInit, 2.0, "fan_in"), _mk(GeneralXavierUniformInit, 2.0, "fan_in")
LecunNormalInit, LecunUniformInit = _mk(GeneralXavierNormalInit, 1.0, "fan_in"), _mk(GeneralXavierUniformInit, 1.0, "fan_in")


def nulls(shape, name=None, trainable=True, dtype="float32", ctx=None): return EmptyInit()(shape, name, trainable, dtype)            # noqa: E704
def zeros(shape, name=None, trainable=True, dtype="float32", ctx=None): return ZerosInit()(shape, name, trainable, dtype)            # noqa: E704
def ones(shape, name=None, trainable=True, dtype="float32", ctx=None): return OnesInit()(shape, name, trainable, dtype)              # noqa: E704
def constant(shape, fill_value=0.0, name=None, trainable=True, dtype="float32", ctx=None): return ConstantInit(fill_value)(shape, name, trainable, dtype)   # noqa: E704,E501
def truncated_normal(shape, mean=0.0, stddev=1.0, name=None, trainable=True, dtype="float32", ctx=None): return TruncatedNormalInit(mean, stddev)(shape, name, trainable, dtype)   # noqa: E704,E501
def random_normal(shape, mean=0.0, stddev=1.0, name=None, trainable=True, dtype="float32", ctx=None): return NormalInit(mean, stddev)(shape, name, trainable, dtype)   # noqa: E704,E501
def random_uniform(shape, minval=-1.0, maxval=1.0, name=None, trainable=True, dtype="float32", ctx=None): return UniformInit(minval, maxval)(shape, name, trainable, dtype)   # noqa: E704,E501
def general_xavier_normal(shape, gain, mode, name=None, trainable=True, dtype="float32", ctx=None): return GeneralXavierNormalInit(gain, mode)(shape, name, trainable, dtype)   # noqa: E704,E501
def general_xavier_uniform(shape, gain, mode, name=None, trainable=True, dtype="float32", ctx=None): return GeneralXavierUniformInit(gain, mode)(shape, name, trainable, dtype)   # noqa: E704,E501
def xavier_normal(shape, name=None, trainable=True, dtype="float32", ctx=None): return XavierNormalInit()(shape, name, trainable, dtype)      # noqa: E704
def xavier_uniform(shape, name=None, trainable=True, dtype="float32", ctx=None): return XavierUniformInit()(shape, name, trainable, dtype)    # noqa: E704
def he_normal(shape, name=None, trainable=True, dtype="float32", ctx=None): return HeNormalInit()(shape, name, trainable, dtype)              # noqa: E704
def he_uniform(shape, name=None, trainable=True, dtype="float32", ctx=None): return HeUniformInit()(shape, name, trainable, dtype)            # noqa: E704
def lecun_normal(shape, name=None, trainable=True, dtype="float32", ctx=None): return LecunNormalInit()(shape, name, trainable, dtype)        # noqa: E704
def lecun_uniform(shape, name=None, trainable=True, dtype="float32", ctx=None): return LecunUniformInit()(shape, name, trainable, dtype)      # noqa: E704


def GenEmpty(): return EmptyInit()                                        # noqa: E704
def GenZeros(): return ZerosInit()                                        # noqa: E704
def GenOnes(): return OnesInit()                                          # noqa: E704
def GenConstant(fill_value=0.0): return ConstantInit(fill_value)          # noqa: E704
def GenTruncatedNormal(mean=0.0, stddev=1.0): return TruncatedNormalInit(mean, stddev)   # noqa: E704
def GenNormal(mean=0.0, stddev=1.0): return NormalInit(mean, stddev)      # noqa: E704
def GenUniform(minval=-1.0, maxval=1.0): return UniformInit(minval, maxval)   # noqa: E704
def GenGeneralXavierNormal(gain, mode): return GeneralXavierNormalInit(gain, mode)     # noqa: E704
def GenGeneralXavierUniform(gain, mode): return GeneralXavierUniformInit(gain, mode)   # noqa: E704
def GenXavierNormal(): return XavierNormalInit()                          # noqa: E704
def GenXavierUniform(): return XavierUniformInit()                        # noqa: E704
def GenHeNormal(): return HeNormalInit()                                  # noqa: E704
def GenHeUniform(): return HeUniformInit()                                # noqa: E704
def GenLecunNormal(): return LecunNormalInit()                            # noqa: E704
def GenLecunUniform(): return LecunUniformInit()                          # noqa: E704


class ReversedTruncatedNormalInit(NormalInit):
    """normal samples from OUTSIDE two standard deviations (|z| >= 2): the tails that TruncatedNormalInit rejects
    (ref: hetu/v1/src/ops/Initializers.cu reversed_truncated_normal_kernel)"""

    def __call__(self, shape=None, name=None, trainable=True, dtype="float32", ctx=None):
        from .executor import Variable
        from .runtime_api import random as _rnd
        shape = list(shape or self.shape)
        n = int(np.prod(shape))
        rng = _rnd.get_np_rand(1)
        out = np.empty(0, np.float64)
        while out.size < n:                                      # ~4.6 % of draws land in the tails
            z = rng.standard_normal(max(4096, 32 * (n - out.size)))
            out = np.concatenate([out, z[np.abs(z) >= 2.0]])
        data = (out[:n] * self.stddev + self.mean).astype(np.float32).reshape(shape)
        return Variable(name or "reversed_truncated_normal_initializer", value=data, trainable=trainable, dtype=dtype)


def reversed_truncated_normal(shape, mean=0.0, stddev=1.0, name=None, trainable=True, dtype="float32", ctx=None): return ReversedTruncatedNormalInit(mean, stddev)(shape, name, trainable, dtype)   # noqa: E704,E501
def GenReversedTruncatedNormal(mean=0.0, stddev=1.0): return ReversedTruncatedNormalInit(mean, stddev)   # noqa: E704
