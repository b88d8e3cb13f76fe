"""Import an ONNX model file as a v1 graph (ref: hetu/v1/python/hetu/onnx/onnx2hetu.py + X2hetu handlers)."""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np

from ... import ops
from . import proto as P

_UNARY = {"Relu": ops.relu, "Sigmoid": ops.sigmoid, "Tanh": ops.tanh, "Exp": ops.exp, "Log": ops.log, "Sqrt": ops.sqrt, "Abs": ops.abs, "Neg": ops.neg,
          "Gelu": ops.gelu, "Mish": ops.mish, "Softplus": ops.softplus, "Ceil": ops.ceil, "Floor": ops.floor, "Round": ops.round, "Sin": ops.sin,
          "Cos": ops.cos, "Reciprocal": ops.reciprocal, "Identity": lambda x: x, "Dropout": lambda x: x}
_BINARY = {"Add": ops.add, "Sub": ops.sub, "Mul": ops.mul, "Div": ops.div}
_NP_DT = {P.FLOAT: "float32", P.INT64: "int64", P.INT32: "int32", P.BOOL: "bool", P.FLOAT16: "float16", P.DOUBLE: "float64"}


def load(path: str, trainable: bool = True) -> Tuple[Dict[str, object], List[object]]:
    """-> ({input name: placeholder node}, [output nodes]); initializers become Variables (scalars / shape operands of
    Reshape, Slice, Pad, Reduce* are folded into attributes)"""
    from ..executor import Variable, placeholder_op
    with open(path, "rb") as f:
        m = P.dec_model(f.read())
    g = m["graph"]
    consts: Dict[str, np.ndarray] = dict(g["initializers"])
    val: Dict[str, object] = {}
    inputs: Dict[str, object] = {}
    for vi in g["inputs"]:
        if vi["name"] in consts:
            continue
        inputs[vi["name"]] = val[vi["name"]] = placeholder_op(vi["name"], [int(d) for d in vi["shape"]], dtype=_NP_DT.get(vi["elem_type"], "float32"))

    def tensor(name):
        if name not in val:
            arr = consts[name]
            val[name] = Variable(name, value=arr, trainable=trainable and arr.dtype.kind == "f", dtype=str(arr.dtype))
        return val[name]

    def const(name):
        assert name in consts, f"operand {name} must be a constant initializer"
        return consts[name]

    for n in g["nodes"]:
        ty, a, i, o = n["op_type"], n["attrs"], n["input"], n["output"]
        if ty in _UNARY:
            y = _UNARY[ty](tensor(i[0]))
        elif ty in _BINARY:
            sc = [k for k in range(2) if i[k] in consts and consts[i[k]].size == 1 and i[k] not in val]
            if sc:                                   # tensor (op) python scalar
                k = sc[0]
                other = tensor(i[1 - k])
                c = float(consts[i[k]].reshape(-1)[0])
                y = {"Add": lambda: other + c, "Mul": lambda: other * c, "Sub": lambda: (other - c) if k == 1 else (c - other),
                     "Div": lambda: (other / c) if k == 1 else (c / other)}[ty]()
            else:
                y = _BINARY[ty](tensor(i[0]), tensor(i[1]))
        elif ty == "LeakyRelu":
            y = ops.leakyrelu(tensor(i[0]), float(a.get("alpha", 0.01)))
        elif ty == "Pow":
            y = ops.pow(tensor(i[0]), float(np.asarray(const(i[1])).reshape(-1)[0]))
        elif ty == "MatMul":
            x, w = tensor(i[0]), tensor(i[1])
            y = ops.bmm(x, w) if len(x.shape) == 3 and len(w.shape) == 3 else ops.matmul(x, w)
        elif ty == "Gemm":
            x = tensor(i[0])
            if a.get("transA"):
                x = ops.transpose(x, [1, 0])
            assert float(a.get("alpha", 1.0)) == 1.0 and float(a.get("beta", 1.0)) == 1.0, "scaled Gemm is not supported"
            y = ops.linear(x, tensor(i[1]), tensor(i[2]) if len(i) > 2 else None, trans_b=bool(a.get("transB", 0)))
        elif ty == "Softmax":
            y = ops.softmax(tensor(i[0]), int(a.get("axis", -1)))
        elif ty == "Reshape":
            y = ops.reshape(tensor(i[0]), [int(v) for v in const(i[1])])
        elif ty == "Flatten":
            x = tensor(i[0])
            ax = int(a.get("axis", 1))
            y = ops.reshape(x, [int(np.prod(x.shape[:ax])), int(np.prod(x.shape[ax:]))])
        elif ty == "Transpose":
            x = tensor(i[0])
            y = ops.transpose(x, [int(p) for p in a.get("perm", list(range(len(x.shape)))[::-1])])
        elif ty == "Concat":
            y = ops.concat([tensor(k) for k in i], int(a.get("axis", 0)))
        elif ty == "Slice":
            x = tensor(i[0])
            starts, ends = [int(v) for v in const(i[1])], [int(v) for v in const(i[2])]
            axes = [int(v) for v in const(i[3])] if len(i) > 3 and i[3] else list(range(len(starts)))
            begin, size = [0] * len(x.shape), list(x.shape)
            for s, e, ax in zip(starts, ends, axes):
                e = min(e, x.shape[ax])
                begin[ax], size[ax] = s, e - s
            y = ops.slice(x, begin, size)
        elif ty in ("ReduceSum", "ReduceMean", "ReduceMax", "ReduceMin", "ReduceProd"):
            x = tensor(i[0])
            axes = [int(v) for v in const(i[1])] if len(i) > 1 and i[1] else [int(v) for v in (a.get("axes") or range(len(x.shape)))]
            y = ops.reduce(x, ty[6:].lower(), axes, bool(a.get("keepdims", 1)))
        elif ty == "LayerNormalization":
            y = ops.layer_norm(tensor(i[0]), tensor(i[1]), tensor(i[2]), eps=float(a.get("epsilon", 1e-5)))
        elif ty == "Conv":
            pads, st = a.get("pads", [0, 0, 0, 0]), a.get("strides", [1, 1])
            assert len(set(pads)) == 1 and len(set(st)) == 1, "asymmetric padding / stride is not supported"
            y = ops.conv2d(tensor(i[0]), tensor(i[1]), tensor(i[2]) if len(i) > 2 else None, padding=int(pads[0]), stride=int(st[0]))
        elif ty in ("MaxPool", "AveragePool"):
            k, pads, st = a["kernel_shape"], a.get("pads", [0, 0, 0, 0]), a.get("strides", [1, 1])
            fn = ops.maxpool if ty == "MaxPool" else ops.avgpool
            y = fn(tensor(i[0]), int(k[0]), int(k[1]), padding=int(pads[0]), stride=int(st[0]))
        elif ty == "BatchNormalization":
            y = ops.batch_norm(tensor(i[0]), tensor(i[1]), tensor(i[2]), tensor(i[3]), tensor(i[4]), momentum=1.0 - float(a.get("momentum", 0.9)),
                               eps=float(a.get("epsilon", 1e-5)))
        elif ty == "Gather":
            assert int(a.get("axis", 0)) == 0, "only row gathers (embedding lookups) are supported"
            y = ops.embedding_lookup(tensor(i[0]), tensor(i[1]))
        elif ty == "Pad":
            x = tensor(i[0])
            pads = [int(v) for v in const(i[1])]
            nd = len(x.shape)
            flat = []
            for d in range(nd - 1, -1, -1):
                flat += [pads[d], pads[nd + d]]
            y = ops.pad(x, flat, "constant", float(np.asarray(const(i[2])).reshape(-1)[0]) if len(i) > 2 and i[2] else 0.0)
        else:
            raise NotImplementedError(f"no importer for ONNX op '{ty}'")
        val[o[0]] = y
    return inputs, [val[vi["name"]] for vi in g["outputs"]]
