"""local file helpers used by the checkpoint code (ref: python/hetu/utils/file_utils.py; HDFS paths are not available here)"""
from __future__ import annotations

import json
import os
import shutil


def ensure_dir(path: str) -> str:
    os.makedirs(path, exist_ok=True)
    return path


def read_json(path: str):
    with open(path) as f:
        return json.load(f)


def write_json(obj, path: str, indent: int = 2):
    ensure_dir(os.path.dirname(os.path.abspath(path)))
    with open(path, "w") as f:
        json.dump(obj, f, indent=indent)


def remove_path(path: str):
    if os.path.isdir(path):
        shutil.rmtree(path, ignore_errors=True)
    elif os.path.exists(path):
        os.remove(path)


def is_remote_path(path: str) -> bool:
    return path.startswith(("hdfs://", "s3://"))
