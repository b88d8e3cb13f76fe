"""Blockwise absmax quantisation: int8, NF4 and FP4 codes, and the 4-bit frozen-weight matmul used by QLoRA-style
fine-tuning (ref: hetu.quantization / hetu.dequantization / hetu.matmul4bit over bitsandbytes,
hetu/graph/ops/Quantization.cc).  8-bit floating point for training GEMMs lives in ops.linear_fp8."""
from __future__ import annotations

from typing import Sequence, Tuple

from ...core import make_op

_KIND = {"int8": "int8", "hetu.int8": "int8", "nf4": "nf4", "nfloat4": "nf4", "hetu.nfloat4": "nf4", "fp4": "fp4", "float4": "fp4",
         "hetu.float4": "fp4"}


def quantize_blockwise_op(x, dtype="int8", blocksize: int = 64, **kw):
    """-> (codes, absmax[ceil(n / blocksize)]); 4-bit codes are packed two per byte"""
    kind = _KIND[str(dtype).lower()]
    return tuple(make_op("quantize_blockwise", [x], {"kind": kind, "blocksize": int(blocksize)}, **kw))


def dequantize_blockwise_op(q, absmax, dtype="float32", blocksize: int = 64, shape: Sequence[int] = None, quant_type: str = None, **kw):
    kind = _KIND[str(quant_type).lower()] if quant_type else ("int8" if str(q.dtype).endswith("int8") else "nf4")
    shape = list(shape) if shape is not None else list(q.shape)
    return make_op("dequantize_blockwise", [q, absmax], {"kind": kind, "blocksize": int(blocksize), "shape": [int(s) for s in shape],
                                                          "dtype": str(dtype).replace("hetu.", "")}, **kw)[0]


def matmul4bit_op(x, w_q, absmax, blocksize: int = 64, quant_type: str = "nf4", weight_shape: Sequence[int] = None, **kw):
    """y = x @ dequant(w_q)^T for a frozen [out, in] weight stored as packed 4-bit codes"""
    assert weight_shape is not None, "matmul4bit needs weight_shape=[out_features, in_features]"
    return make_op("matmul4bit", [x, w_q, absmax], {"kind": _KIND[quant_type.lower()], "blocksize": int(blocksize),
                                                    "weight_shape": [int(s) for s in weight_shape]}, **kw)[0]


def quantize_weight_nf4(w, blocksize: int = 64):
    """helper for loading a frozen base model: returns (codes, absmax, shape)"""
    q, a = quantize_blockwise_op(w, "nf4", blocksize)
    return q, a, list(w.shape)
