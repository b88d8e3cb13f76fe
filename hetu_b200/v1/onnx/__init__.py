"""ONNX interchange for v1 graphs: `hetu2onnx.export(outputs, path)` / `onnx2hetu.load(path)`.  The protobuf wire format is
encoded / decoded by `proto.py`, so neither the `onnx` package nor compiled descriptors are required.
(ref: hetu/v1/python/hetu/onnx/{hetu2onnx,onnx2hetu,graph,handler}.py, onnx_opset/, X2hetu/)"""
from . import hetu2onnx, onnx2hetu, proto  # noqa: F401
from .hetu2onnx import export  # noqa: F401
from .onnx2hetu import load  # noqa: F401
