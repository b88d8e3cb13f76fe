"""Minimal stand-in for the hydra command line the reference's entry points use:
    python -m hetu.<module> --config-path DIR --config-name NAME key.sub=value ...
-> the YAML file DIR/NAME(.yaml) with dotted `key=value` overrides applied, as an attribute-accessible dict."""
from __future__ import annotations

import argparse
import os
from typing import Any, List, Optional, Sequence

import yaml


class Cfg(dict):
    """dict with attribute access (OmegaConf-like), recursive"""

    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError as e:
            raise AttributeError(k) from e
        return Cfg(v) if isinstance(v, dict) and not isinstance(v, Cfg) else v

    def get_path(self, dotted: str, default=None):
        cur: Any = self
        for k in dotted.split("."):
            if not isinstance(cur, dict) or k not in cur:
                return default
            cur = cur[k]
        return Cfg(cur) if isinstance(cur, dict) else cur


def apply_overrides(raw: dict, overrides: Sequence[str]) -> dict:
    for ov in overrides:
        k, v = ov.lstrip("+").split("=", 1)
        d = raw
        keys = k.split(".")
        for kk in keys[:-1]:
            d = d.setdefault(kk, {})
        d[keys[-1]] = yaml.safe_load(v)
    return raw


def load(argv: Optional[List[str]] = None, default_path: Optional[str] = None, default_name: str = "config") -> Cfg:
    ap = argparse.ArgumentParser(add_help=True)
    ap.add_argument("--config-path", "-cp", default=default_path)
    ap.add_argument("--config-name", "-cn", default=default_name)
    ap.add_argument("overrides", nargs="*")
    a = ap.parse_args(argv)
    raw = {}
    if a.config_path is not None:
        base = os.path.join(a.config_path, a.config_name)
        path = next((p for p in (base, base + ".yaml", base + ".yml") if os.path.isfile(p)), None)
        if path is None:
            raise FileNotFoundError(f"no config {a.config_name}(.yaml) under {a.config_path}")
        with open(path) as f:
            raw = yaml.safe_load(f) or {}
    return Cfg(apply_overrides(raw, a.overrides))
