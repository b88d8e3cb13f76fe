"""v1 array helpers: `array` / `empty` / `sparse_array` / `IndexedSlices` over torch tensors.
(ref: hetu/v1/python/hetu/ndarray.py -- NDArray :150, array :485, empty :505, ND_Sparse_Array :590, sparse_array :641,
IndexedSlices :680)"""
from __future__ import annotations

import numpy as np
import torch

from .executor import _Ctx, cpu


def is_gpu_ctx(ctx) -> bool:
    return bool(ctx is not None and getattr(ctx, "kind", None) == "gpu")


def _device(ctx):
    if is_gpu_ctx(ctx) and torch.cuda.is_available():
        return torch.device("cuda", int(ctx.index))
    return torch.device("cpu")


class NDArray:
    """a tensor with the device context it was asked for (on a GPU-less machine GPU contexts fall back to host memory)"""

    def __init__(self, t: torch.Tensor, ctx: _Ctx = None):
        self.t, self.ctx = t, ctx or cpu(0)

    shape = property(lambda self: tuple(self.t.shape))
    dtype = property(lambda self: np.dtype(str(self.t.dtype).replace("torch.", "")) if self.t.dtype != torch.bfloat16 else np.dtype("float32"))

    def asnumpy(self):
        t = self.t.detach().cpu()
        return (t.float() if t.dtype == torch.bfloat16 else t).numpy()

    def copyto(self, target):
        """into another NDArray, or onto a device context (a new array)"""
        if isinstance(target, NDArray):
            target.t.copy_(self.t)
            return target
        return NDArray(self.t.to(_device(target)), target)

    def __setitem__(self, idx, value):
        v = value.t if isinstance(value, NDArray) else torch.as_tensor(np.asarray(value))
        self.t[idx] = v.to(self.t.device, self.t.dtype)

    def __getitem__(self, idx):
        return NDArray(self.t[idx], self.ctx)

    def reshape(self, shape):
        return NDArray(self.t.reshape(list(shape)), self.ctx)

    def __repr__(self):
        return f"NDArray(shape={self.shape}, ctx={self.ctx})"


def array(arr, ctx=None, dtype=np.float32, force32=True) -> NDArray:
    a = np.asarray(arr.asnumpy() if isinstance(arr, NDArray) else arr)
    if not force32:
        a = a.astype(dtype)
    elif a.dtype.kind == "f":
        a = a.astype(np.float32)          # integer ids keep their width (lookups take int64)
    return NDArray(torch.from_numpy(np.ascontiguousarray(a)).to(_device(ctx)), ctx)


def empty(shape, ctx=None, dtype=np.float32, force32=True) -> NDArray:
    td = getattr(torch, np.dtype(dtype).name)
    return NDArray(torch.empty(list(shape), dtype=td, device=_device(ctx)), ctx)


def empty_like(arr: NDArray) -> NDArray:
    return NDArray(torch.empty_like(arr.t), arr.ctx)


class ND_Sparse_Array:              # noqa: N801  (the reference's name)
    """CSR / COO matrix: `values`, (`row`, `col`) indices, `shape`; `.t` is the torch sparse tensor the graph's spmm consumes"""

    def __init__(self, values, indices, shape, form="csr", ctx=None):
        self.form, self.ctx, self.shape = form, ctx or cpu(0), tuple(int(s) for s in shape)
        self.nrow, self.ncol = self.shape
        v = torch.as_tensor(np.asarray(values, np.float32))
        r, c = (torch.as_tensor(np.asarray(i, np.int64)) for i in indices)
        dev = _device(ctx)
        if form == "csr":
            order = torch.argsort(r * self.ncol + c)        # the inputs are per-entry (COO) coordinates: sort and compress the rows
            r, c, v = r[order], c[order], v[order]
            r = torch.cat([torch.zeros(1, dtype=torch.int64), torch.cumsum(torch.bincount(r, minlength=self.nrow), 0)])
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                self.t = torch.sparse_csr_tensor(r, c, v, self.shape).to(dev)
        else:
            self.t = torch.sparse_coo_tensor(torch.stack([r, c]), v, self.shape).coalesce().to(dev)
        self.values, self.row, self.col = v, r, c

    def to_dense(self):
        return NDArray(self.t.to_dense(), self.ctx)

    def coo(self):
        """(indices [2, nnz], values [nnz]) -- what `ops.spmm` / `csrmm_op` take"""
        t = self.t.to_sparse_coo().coalesce()
        return t.indices(), t.values()


def sparse_array(values, indices, shape, form="csr", ctx=None) -> ND_Sparse_Array:
    return ND_Sparse_Array(values, indices, shape, form, ctx)


class IndexedSlices:
    """a sparse gradient: rows `values[i]` belong at `indices[i]` of a [dense_shape] tensor (ref: ndarray.py:680)"""

    def __init__(self, indices=None, values=None, dense_shape=None):
        self.indices, self.values, self.dense_shape = indices, values, tuple(dense_shape) if dense_shape is not None else None

    @staticmethod
    def _t(x):
        return x.t if hasattr(x, "t") and not isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x)) if not isinstance(x, torch.Tensor) else x

    def get_dense_shape(self):
        assert self.dense_shape is not None
        return self.dense_shape

    def get_sparse_shape(self):
        return tuple(self._t(self.values).shape)

    def deduplicate(self):
        """merge repeated ids (their rows add up): -> IndexedSlices with unique, sorted indices"""
        idx, val = self._t(self.indices).reshape(-1).long(), self._t(self.values)
        val = val.reshape(idx.numel(), -1)
        uniq, inverse = torch.unique(idx, return_inverse=True)
        summed = torch.zeros(uniq.numel(), val.shape[1], dtype=val.dtype).index_add_(0, inverse, val)
        keep = uniq >= 0
        return IndexedSlices(uniq[keep], summed[keep], self.dense_shape)

    def to_dense(self, stream=None):
        idx, val = self._t(self.indices).reshape(-1).long(), self._t(self.values)
        out = torch.zeros(self.get_dense_shape(), dtype=val.dtype)
        ok = idx >= 0
        out.index_add_(0, idx[ok], val.reshape(idx.numel(), -1)[ok].reshape(-1, *out.shape[1:]))
        return NDArray(out)

    def asnumpy(self):
        return self.to_dense().asnumpy()
