"""Self-contained ONNX protobuf codec (no `onnx` package needed): the handful of messages a model file consists of
(ModelProto, GraphProto, NodeProto, AttributeProto, TensorProto, ValueInfoProto, TypeProto) as plain dicts <-> wire bytes.
Field numbers follow onnx/onnx.proto (IR version 8)."""
from __future__ import annotations

import struct
from typing import Dict, List, Tuple

import numpy as np

# TensorProto.DataType
FLOAT, UINT8, INT8, INT32, INT64, BOOL, FLOAT16, DOUBLE, BFLOAT16 = 1, 2, 3, 6, 7, 9, 10, 11, 16
NP_TO_ONNX = {np.dtype("float32"): FLOAT, np.dtype("uint8"): UINT8, np.dtype("int8"): INT8, np.dtype("int32"): INT32, np.dtype("int64"): INT64,
              np.dtype("bool"): BOOL, np.dtype("float16"): FLOAT16, np.dtype("float64"): DOUBLE}
ONNX_TO_NP = {v: k for k, v in NP_TO_ONNX.items()}
# AttributeProto.AttributeType
A_FLOAT, A_INT, A_STRING, A_TENSOR, A_FLOATS, A_INTS, A_STRINGS = 1, 2, 3, 4, 6, 7, 8


# ------------------------------------------------------------------------------------------------------------- wire format
def _varint(v: int) -> bytes:
    if v < 0:
        v += 1 << 64
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        out.append(b | (0x80 if v else 0))
        if not v:
            return bytes(out)


def _key(field: int, wt: int) -> bytes:
    return _varint((field << 3) | wt)


def f_varint(field, v): return _key(field, 0) + _varint(int(v))                                   # noqa: E704
def f_bytes(field, b): return _key(field, 2) + _varint(len(b)) + bytes(b)                          # noqa: E704
def f_str(field, s): return f_bytes(field, s.encode("utf-8"))                                     # noqa: E704
def f_float(field, v): return _key(field, 5) + struct.pack("<f", float(v))                        # noqa: E704
def f_packed_ints(field, vs): return f_bytes(field, b"".join(_varint(int(v)) for v in vs))        # noqa: E704
def f_packed_floats(field, vs): return f_bytes(field, struct.pack(f"<{len(vs)}f", *[float(v) for v in vs]))   # noqa: E704


def parse(buf: bytes) -> List[Tuple[int, int, object]]:
    """-> [(field, wire type, value)]: varint -> int, 64-bit -> 8 bytes, length-delimited -> bytes, 32-bit -> 4 bytes"""
    out, i, n = [], 0, len(buf)
    while i < n:
        key, i = _read_varint(buf, i)
        field, wt = key >> 3, key & 7
        if wt == 0:
            v, i = _read_varint(buf, i)
        elif wt == 1:
            v, i = buf[i:i + 8], i + 8
        elif wt == 2:
            ln, i = _read_varint(buf, i)
            v, i = buf[i:i + ln], i + ln
        elif wt == 5:
            v, i = buf[i:i + 4], i + 4
        else:
            raise ValueError(f"unsupported wire type {wt}")
        out.append((field, wt, v))
    return out


def _read_varint(buf, i):
    shift = v = 0
    while True:
        b = buf[i]
        i += 1
        v |= (b & 0x7F) << shift
        if not b & 0x80:
            return v, i
        shift += 7


def _signed(v):
    return v - (1 << 64) if v >= (1 << 63) else v


def _ints(wt, v):
    if wt == 0:
        return [_signed(v)]
    out, i = [], 0
    while i < len(v):
        x, i = _read_varint(v, i)
        out.append(_signed(x))
    return out


def _floats(wt, v):
    return list(struct.unpack(f"<{len(v) // 4}f", v))


# ---------------------------------------------------------------------------------------------------------------- messages
def enc_tensor(name: str, arr: np.ndarray) -> bytes:
    arr = np.asarray(arr)
    shape = arr.shape                       # (ascontiguousarray would promote a scalar to 1-d)
    dt = NP_TO_ONNX[arr.dtype]
    return f_packed_ints(1, shape) + f_varint(2, dt) + f_str(8, name) + f_bytes(9, np.ascontiguousarray(arr).tobytes())


def dec_tensor(buf: bytes) -> Tuple[str, np.ndarray]:
    dims, dt, name, raw, fdata, i32, i64 = [], FLOAT, "", None, [], [], []
    for f, wt, v in parse(buf):
        if f == 1: dims += _ints(wt, v)                       # noqa: E701
        elif f == 2: dt = v                                   # noqa: E701
        elif f == 8: name = v.decode()                        # noqa: E701
        elif f == 9: raw = v                                  # noqa: E701
        elif f == 4: fdata += _floats(wt, v)                  # noqa: E701
        elif f == 5: i32 += _ints(wt, v)                      # noqa: E701
        elif f == 7: i64 += _ints(wt, v)                      # noqa: E701
    np_dt = ONNX_TO_NP[dt]
    if raw is not None:
        arr = np.frombuffer(raw, dtype=np_dt).copy()
    elif fdata:
        arr = np.array(fdata, dtype=np_dt)
    elif i64:
        arr = np.array(i64, dtype=np_dt)
    else:
        arr = np.array(i32, dtype=np_dt)
    return name, arr.reshape(dims)


def enc_attr(name: str, value) -> bytes:
    b = f_str(1, name)
    if isinstance(value, bool) or isinstance(value, (int, np.integer)):
        return b + f_varint(3, int(value)) + f_varint(20, A_INT)
    if isinstance(value, float):
        return b + f_float(2, value) + f_varint(20, A_FLOAT)
    if isinstance(value, str):
        return b + f_bytes(4, value.encode()) + f_varint(20, A_STRING)
    if isinstance(value, np.ndarray):
        return b + f_bytes(5, enc_tensor("", value)) + f_varint(20, A_TENSOR)
    value = list(value)
    if value and isinstance(value[0], float):
        return b + f_packed_floats(7, value) + f_varint(20, A_FLOATS)
    return b + f_packed_ints(8, value) + f_varint(20, A_INTS)


def dec_attr(buf: bytes):
    name, typ, vals = "", 0, {"floats": [], "ints": [], "strings": []}
    for f, wt, v in parse(buf):
        if f == 1: name = v.decode()                          # noqa: E701
        elif f == 20: typ = v                                 # noqa: E701
        elif f == 2: vals["f"] = struct.unpack("<f", v)[0]    # noqa: E701
        elif f == 3: vals["i"] = _signed(v)                   # noqa: E701
        elif f == 4: vals["s"] = v.decode()                   # noqa: E701
        elif f == 5: vals["t"] = dec_tensor(v)[1]             # noqa: E701
        elif f == 7: vals["floats"] += _floats(wt, v)         # noqa: E701
        elif f == 8: vals["ints"] += _ints(wt, v)             # noqa: E701
        elif f == 9: vals["strings"].append(v.decode())       # noqa: E701
    key = {A_FLOAT: "f", A_INT: "i", A_STRING: "s", A_TENSOR: "t", A_FLOATS: "floats", A_INTS: "ints", A_STRINGS: "strings"}.get(typ)
    if key is None:      # type omitted by the producer: take whatever is present
        key = next((k for k in ("i", "f", "s", "t") if k in vals), "ints" if vals["ints"] else "floats")
    return name, vals.get(key)


def enc_node(op_type: str, inputs, outputs, name: str = "", attrs: Dict = None, domain: str = "") -> bytes:
    b = b"".join(f_str(1, i) for i in inputs) + b"".join(f_str(2, o) for o in outputs) + f_str(3, name) + f_str(4, op_type)
    for k, v in (attrs or {}).items():
        b += f_bytes(5, enc_attr(k, v))
    if domain:
        b += f_str(7, domain)
    return b


def dec_node(buf: bytes) -> dict:
    n = {"input": [], "output": [], "name": "", "op_type": "", "attrs": {}, "domain": ""}
    for f, wt, v in parse(buf):
        if f == 1: n["input"].append(v.decode())              # noqa: E701
        elif f == 2: n["output"].append(v.decode())           # noqa: E701
        elif f == 3: n["name"] = v.decode()                   # noqa: E701
        elif f == 4: n["op_type"] = v.decode()                # noqa: E701
        elif f == 5:
            k, val = dec_attr(v)
            n["attrs"][k] = val
        elif f == 7: n["domain"] = v.decode()                 # noqa: E701
    return n


def enc_value_info(name: str, elem_type: int, shape) -> bytes:
    dims = b"".join(f_bytes(1, f_varint(1, d) if isinstance(d, (int, np.integer)) else f_str(2, str(d))) for d in shape)
    tensor_type = f_varint(1, elem_type) + f_bytes(2, dims)
    return f_str(1, name) + f_bytes(2, f_bytes(1, tensor_type))


def dec_value_info(buf: bytes) -> dict:
    out = {"name": "", "elem_type": FLOAT, "shape": []}
    for f, wt, v in parse(buf):
        if f == 1:
            out["name"] = v.decode()
        elif f == 2:
            for f2, _, v2 in parse(v):
                if f2 != 1:
                    continue
                for f3, wt3, v3 in parse(v2):
                    if f3 == 1:
                        out["elem_type"] = v3
                    elif f3 == 2:
                        for f4, _, v4 in parse(v3):
                            if f4 == 1:
                                d = None
                                for f5, wt5, v5 in parse(v4):
                                    d = _signed(v5) if f5 == 1 else v5.decode()
                                out["shape"].append(d)
    return out


def enc_model(graph_name: str, nodes: List[bytes], initializers: List[bytes], inputs: List[bytes], outputs: List[bytes], opset: int = 20,
              producer: str = "hetu_b200") -> bytes:
    g = b"".join(f_bytes(1, n) for n in nodes) + f_str(2, graph_name) + b"".join(f_bytes(5, t) for t in initializers) + \
        b"".join(f_bytes(11, i) for i in inputs) + b"".join(f_bytes(12, o) for o in outputs)
    return f_varint(1, 8) + f_str(2, producer) + f_str(3, "0.1") + f_bytes(7, g) + f_bytes(8, f_str(1, "") + f_varint(2, opset))


def dec_model(buf: bytes) -> dict:
    m = {"ir_version": 0, "producer": "", "opset": {}, "graph": {"name": "", "nodes": [], "initializers": {}, "inputs": [], "outputs": []}}
    for f, wt, v in parse(buf):
        if f == 1: m["ir_version"] = v                        # noqa: E701
        elif f == 2: m["producer"] = v.decode()               # noqa: E701
        elif f == 8:
            dom, ver = "", 0
            for f2, _, v2 in parse(v):
                if f2 == 1: dom = v2.decode()                 # noqa: E701
                elif f2 == 2: ver = v2                        # noqa: E701
            m["opset"][dom] = ver
        elif f == 7:
            g = m["graph"]
            for f2, _, v2 in parse(v):
                if f2 == 1: g["nodes"].append(dec_node(v2))   # noqa: E701
                elif f2 == 2: g["name"] = v2.decode()         # noqa: E701
                elif f2 == 5:
                    name, arr = dec_tensor(v2)
                    g["initializers"][name] = arr
                elif f2 == 11: g["inputs"].append(dec_value_info(v2))    # noqa: E701
                elif f2 == 12: g["outputs"].append(dec_value_info(v2))   # noqa: E701
    return m
