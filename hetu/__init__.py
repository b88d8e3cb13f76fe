"""`import hetu` compatibility alias: the reference's package name resolves to hetu_b200 (same API surface), including
its sub-packages (`hetu.nn`, `hetu.engine`, `hetu.data`, `hetu.models`, `hetu.peft`, `hetu.rpc`, `hetu.utils...`)."""
import importlib
import sys

import hetu_b200 as _impl
from hetu_b200 import *  # noqa: F401,F403

_SUBS = ["nn", "ops", "optim", "models", "engine", "data", "peft", "rpc", "utils", "utils.parallel", "utils.checkpoint", "distributed",
         "core", "parallel", "planner", "v1"]
for _name in _SUBS:
    try:
        _m = importlib.import_module(f"hetu_b200.{_name}")
    except Exception:   # noqa: BLE001 -- optional sub-package not importable in this environment
        continue
    sys.modules[f"hetu.{_name}"] = _m
    if "." not in _name:
        globals()[_name] = _m


def __getattr__(name):
    return getattr(_impl, name)


__version__ = _impl.__version__
