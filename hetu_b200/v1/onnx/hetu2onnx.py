"""Export a v1 (or any define-and-run) graph to an ONNX model file
(ref: hetu/v1/python/hetu/onnx/hetu2onnx.py + onnx_opset/*: per-op handlers that emit ONNX nodes)."""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np

from ... import core
from . import proto as P

_UNARY = {"relu": "Relu", "sigmoid": "Sigmoid", "tanh": "Tanh", "exp": "Exp", "log": "Log", "sqrt": "Sqrt", "abs": "Abs", "neg": "Neg",
          "gelu": "Gelu", "mish": "Mish", "softplus": "Softplus", "ceil": "Ceil", "floor": "Floor", "round": "Round", "sin": "Sin", "cos": "Cos",
          "reciprocal": "Reciprocal", "contiguous": "Identity", "stop_gradient": "Identity"}
_BINARY = {"add": "Add", "sub": "Sub", "mul": "Mul", "div": "Div"}
_DT = {"float32": P.FLOAT, "int64": P.INT64, "int32": P.INT32, "bool": P.BOOL, "float16": P.FLOAT16, "float64": P.DOUBLE, "bfloat16": P.FLOAT}


class _Exporter:
    def __init__(self, graph):
        self.g = graph
        self.nodes: List[bytes] = []
        self.inits: List[bytes] = []
        self.inputs: List[bytes] = []
        self.names: Dict[int, str] = {}
        self.done = set()
        self.uid = 0

    def fresh(self, base):
        self.uid += 1
        return f"{base}_{self.uid}"

    def name_of(self, t):
        if t.id not in self.names:
            self.names[t.id] = f"{t.name or 't'}_{t.id}"
        return self.names[t.id]

    def const(self, arr, base="const"):
        n = self.fresh(base)
        self.inits.append(P.enc_tensor(n, np.asarray(arr)))
        return n

    def emit(self, op_type, ins, outs, attrs=None):
        self.nodes.append(P.enc_node(op_type, ins, outs, self.fresh(op_type), attrs))

    def visit(self, t):
        if t.producer_id in self.done:
            return
        self.done.add(t.producer_id)
        info = self.g.op_info(t.producer_id)
        ty, a, ins, outs = info["type"], info["attrs"], info["inputs"], info["outputs"]
        for i in ins:
            self.visit(i)
        I = [self.name_of(i) for i in ins]          # noqa: E741
        O = [self.name_of(o) for o in outs]
        if ty == "placeholder":
            self.inputs.append(P.enc_value_info(O[0], _DT.get(a.get("dtype", "float32"), P.FLOAT), list(outs[0].shape)))
        elif ty == "variable":
            w = self.g.get_param(outs[0]).detach().float().cpu().numpy() if a.get("dtype", "float32") in ("float32", "bfloat16", "float16") \
                else self.g.get_param(outs[0]).detach().cpu().numpy()
            self.inits.append(P.enc_tensor(O[0], w))
        elif ty in _UNARY:
            self.emit(_UNARY[ty], I[:1], O[:1])
        elif ty == "unary_act":
            k = a["kind"]
            if k == "silu":                        # x * sigmoid(x)
                s = self.fresh("sig")
                self.emit("Sigmoid", I[:1], [s])
                self.emit("Mul", [I[0], s], O[:1])
            else:
                self.emit(_UNARY[k], I[:1], O[:1])
        elif ty == "rsqrt":
            s = self.fresh("sqrt")
            self.emit("Sqrt", I[:1], [s])
            self.emit("Reciprocal", [s], O[:1])
        elif ty == "leakyrelu":
            self.emit("LeakyRelu", I[:1], O[:1], {"alpha": float(a.get("alpha", 0.01))})
        elif ty == "pow":
            self.emit("Pow", [I[0], self.const(np.float32(a["exponent"]))], O[:1])
        elif ty in _BINARY:
            if len(I) == 2:
                self.emit(_BINARY[ty], I, O[:1])
            else:                                   # tensor (op) scalar; `from_const` = scalar (op) tensor
                c = self.const(np.float32(a["value"]))
                self.emit(_BINARY[ty], [c, I[0]] if a.get("from_const") else [I[0], c], O[:1])
        elif ty == "matmul":
            x, y = I
            if a.get("trans_a"):
                x2 = self.fresh("ta"); self.emit("Transpose", [x], [x2], {"perm": [1, 0]}); x = x2          # noqa: E702
            if a.get("trans_b"):
                y2 = self.fresh("tb"); self.emit("Transpose", [y], [y2], {"perm": [1, 0]}); y = y2          # noqa: E702
            self.emit("MatMul", [x, y], O[:1])
        elif ty == "bmm":
            self.emit("MatMul", I, O[:1])
        elif ty == "linear":
            assert not a.get("has_residual") and a.get("act", "none") in ("none", "relu", "gelu"), "fused residual export is not supported"
            gemm_out = O[0] if a.get("act", "none") == "none" else self.fresh("gemm")
            if len(ins[0].shape) == 2:
                self.emit("Gemm", I[:3] if a.get("has_bias") else I[:2], [gemm_out], {"transB": int(bool(a.get("trans_b")))})
            else:
                w = I[1]
                if a.get("trans_b"):
                    w = self.fresh("wt"); self.emit("Transpose", [I[1]], [w], {"perm": [1, 0]})           # noqa: E702
                mm = self.fresh("mm") if a.get("has_bias") else gemm_out
                self.emit("MatMul", [I[0], w], [mm])
                if a.get("has_bias"):
                    self.emit("Add", [mm, I[2]], [gemm_out])
            if a.get("act", "none") != "none":
                self.emit(_UNARY[a["act"]], [gemm_out], O[:1])
        elif ty == "softmax":
            self.emit("Softmax", I[:1], O[:1], {"axis": int(a.get("dim", -1))})
        elif ty == "reshape":
            self.emit("Reshape", [I[0], self.const(np.array(a["shape"], dtype=np.int64), "shape")], O[:1])
        elif ty == "transpose":
            self.emit("Transpose", I[:1], O[:1], {"perm": [int(p) for p in a["perm"]]})
        elif ty == "concat":
            self.emit("Concat", I, O[:1], {"axis": int(a.get("dim", 0))})
        elif ty == "slice":
            b, s = a["begin"], a["size"]
            self.emit("Slice", [I[0], self.const(np.array(b, dtype=np.int64)), self.const(np.array([x + y for x, y in zip(b, s)], dtype=np.int64)),
                                self.const(np.arange(len(b), dtype=np.int64))], O[:1])
        elif ty == "reduce":
            op = {"sum": "ReduceSum", "mean": "ReduceMean", "max": "ReduceMax", "min": "ReduceMin", "prod": "ReduceProd"}[a.get("mode", "sum")]
            axes = a.get("axes") or list(range(len(ins[0].shape)))
            self.emit(op, [I[0], self.const(np.array(axes, dtype=np.int64), "axes")], O[:1], {"keepdims": int(bool(a.get("keepdims")))})
        elif ty == "fused_norm":
            assert not a.get("rms"), "RMSNorm has no ONNX operator in opset 20"
            self.emit("LayerNormalization", I[:3], O[:1], {"axis": -1, "epsilon": float(a.get("eps", 1e-5))})
        elif ty == "conv2d":
            p, st = int(a.get("padding", 0)), int(a.get("stride", 1))
            self.emit("Conv", I, O[:1], {"pads": [p, p, p, p], "strides": [st, st]})
        elif ty in ("maxpool", "avgpool"):
            p, st = int(a.get("padding", 0)), int(a.get("stride", 1))
            attrs = {"kernel_shape": [int(a["kernel_H"]), int(a["kernel_W"])], "pads": [p, p, p, p], "strides": [st, st]}
            if ty == "avgpool":
                attrs["count_include_pad"] = 1
            self.emit("MaxPool" if ty == "maxpool" else "AveragePool", I[:1], O[:1], attrs)
        elif ty == "batch_norm":
            self.emit("BatchNormalization", I[:5], O[:1], {"epsilon": float(a.get("eps", 1e-5)), "momentum": 1.0 - float(a.get("momentum", 0.1))})
        elif ty == "dropout":
            self.emit("Identity", I[:1], O[:1])          # inference graph
        elif ty == "embedding_lookup":
            self.emit("Gather", [I[0], I[1]], O[:1], {"axis": 0})
        elif ty == "pad":
            flat = [int(v) for v in a["paddings"]]
            nd = len(ins[0].shape)
            begins, ends = [0] * nd, [0] * nd
            for k in range(len(flat) // 2):           # flat list is last-dimension-first
                begins[nd - 1 - k], ends[nd - 1 - k] = flat[2 * k], flat[2 * k + 1]
            self.emit("Pad", [I[0], self.const(np.array(begins + ends, dtype=np.int64), "pads"), self.const(np.float32(a.get("value", 0.0)))], O[:1])
        else:
            raise NotImplementedError(f"no ONNX handler for op '{ty}'")


def export(outputs: Sequence, path: str, graph_name: str = "hetu_graph", opset: int = 20) -> str:
    """write the sub-graph that computes `outputs` (tensors of one graph) to `path`; parameters become initializers with
    their current values, placeholders become graph inputs"""
    outputs = list(outputs)
    g = core._graphs_by_id[outputs[0].graph_id]
    ex = _Exporter(g)
    for o in outputs:
        ex.visit(o)
    outs = [P.enc_value_info(ex.name_of(o), _DT.get(o.dtype, P.FLOAT), list(o.shape)) for o in outputs]
    with open(path, "wb") as f:
        f.write(P.enc_model(graph_name, ex.nodes, ex.inits, ex.inputs, outs, opset))
    return path
