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
    """`a.b=v` sets (creating the path), `+a.b=v` adds, `~a.b` deletes, values are YAML scalars / lists / dicts"""
    for ov in overrides:
        if ov.startswith("~"):
            keys = ov[1:].split("=", 1)[0].split(".")
            d = raw
            for kk in keys[:-1]:
                d = d.get(kk, {}) if isinstance(d, dict) else {}
            if isinstance(d, dict):
                d.pop(keys[-1], None)
            continue
        k, v = ov.lstrip("+").split("=", 1)
        d = raw
        keys = k.split(".")
        for kk in keys[:-1]:
            nxt = d.get(kk)
            if not isinstance(nxt, dict):
                nxt = d[kk] = {}
            d = nxt
        d[keys[-1]] = yaml.safe_load(v)
    return raw


def _deep_merge(base: dict, over: dict) -> dict:
    for k, v in over.items():
        if isinstance(v, dict) and isinstance(base.get(k), dict):
            _deep_merge(base[k], v)
        else:
            base[k] = v
    return base


def _find(config_dir: str, name: str) -> Optional[str]:
    base = os.path.join(config_dir, name)
    return next((p for p in (base, base + ".yaml", base + ".yml") if os.path.isfile(p)), None)


def compose(config_dir: str, name: str, group_choice: Optional[dict] = None, _stack=()) -> dict:
    """hydra-style composition: the `defaults` list of a config names other files merged BEFORE it -- plain names (same
    directory), `group: option` entries (file `group/option.yaml`, placed under key `group` unless it carries `# @package _global_`
    on its first line), `_self_` to position this file's own content.  A `group=option` command-line choice replaces the option."""
    path = _find(config_dir, name)
    if path is None:
        raise FileNotFoundError(f"no config {name}(.yaml) under {config_dir}")
    if path in _stack:
        raise ValueError(f"defaults cycle through {path}")
    with open(path) as f:
        text = f.read()
    own = yaml.safe_load(text) or {}
    defaults = own.pop("defaults", None)
    if not defaults:
        return own
    out: dict = {}
    seen_self = False
    for entry in defaults:
        if entry == "_self_":
            _deep_merge(out, own)
            seen_self = True
        elif isinstance(entry, str):
            _deep_merge(out, compose(config_dir, entry, group_choice, _stack + (path,)))
        elif isinstance(entry, dict):
            for group, option in entry.items():
                group = str(group).replace("override ", "").lstrip("/")
                option = (group_choice or {}).get(group, option)
                if option in (None, "null"):
                    continue
                sub_path = _find(os.path.join(config_dir, group), str(option))
                if sub_path is None:
                    raise FileNotFoundError(f"no option {option} in config group {group} under {config_dir}")
                with open(sub_path) as f:
                    first = f.readline()
                sub = compose(os.path.join(config_dir, group), str(option), group_choice, _stack + (path,))
                if "@package _global_" in first:
                    _deep_merge(out, sub)
                else:
                    node = out
                    for part in group.split("/"):
                        node = node.setdefault(part, {})
                    _deep_merge(node, sub)
    if not seen_self:
        _deep_merge(out, own)
    return out


_INTERP = None


def resolve(raw: dict) -> dict:
    """OmegaConf-style interpolation: `${a.b}` (absolute path, whole value keeps its type, inside a string it is formatted),
    `${oc.env:VAR}` / `${oc.env:VAR,default}`; nested references are followed, cycles are an error"""
    import re
    global _INTERP
    _INTERP = _INTERP or re.compile(r"\$\{([^${}]+)\}")

    def lookup(path: str, trail):
        if path.startswith("oc.env:"):
            var, _, dflt = path[len("oc.env:"):].partition(",")
            if var in os.environ:
                return yaml.safe_load(os.environ[var])
            if _:
                return yaml.safe_load(dflt)
            raise KeyError(f"environment variable {var} is not set and the interpolation has no default")
        if path in trail:
            raise ValueError("interpolation cycle: " + " -> ".join(trail + (path,)))
        cur: Any = raw
        for k in path.split("."):
            if isinstance(cur, list):
                cur = cur[int(k)]
            elif isinstance(cur, dict) and k in cur:
                cur = cur[k]
            else:
                raise KeyError(f"interpolation ${{{path}}}: no such key")
        return walk(cur, trail + (path,))

    def walk(v, trail):
        if isinstance(v, dict):
            return {k: walk(x, trail) for k, x in v.items()}
        if isinstance(v, list):
            return [walk(x, trail) for x in v]
        if isinstance(v, str) and "${" in v:
            m = _INTERP.fullmatch(v.strip())
            if m:
                return lookup(m.group(1).strip(), trail)
            return _INTERP.sub(lambda mm: str(lookup(mm.group(1).strip(), trail)), v)
        return v
    return walk(raw, ())


def merge_dataclass(cls, values: dict, strict: bool = True):
    """structured config: fill a dataclass from a dict with type coercion of scalars; unknown keys are an error in strict mode"""
    import dataclasses
    import typing
    hints = typing.get_type_hints(cls)
    names = {f.name for f in dataclasses.fields(cls)}
    unknown = [k for k in values if k not in names]
    if unknown and strict:
        raise KeyError(f"{cls.__name__} has no field(s) {unknown}; known: {sorted(names)}")
    kw = {}
    for k, v in values.items():
        if k not in names:
            continue
        t = hints.get(k)
        origin = typing.get_origin(t)
        if origin is typing.Union:
            args = [x for x in typing.get_args(t) if x is not type(None)]
            t = args[0] if len(args) == 1 else None
        if v is not None and t in (int, float, str, bool) and not isinstance(v, t):
            if t is bool and isinstance(v, str):
                v = v.strip().lower() in ("1", "true", "yes", "on")
            elif t is int and isinstance(v, float) and v != int(v):
                raise TypeError(f"{cls.__name__}.{k}: {v!r} is not an integer")
            else:
                v = t(v)
        elif v is not None and dataclasses.is_dataclass(t) and isinstance(v, dict):
            v = merge_dataclass(t, v, strict)
        kw[k] = v
    return cls(**kw)


def to_yaml(cfg: dict) -> str:
    """the resolved configuration as YAML (what hydra writes to .hydra/config.yaml)"""
    def plain(v):
        if isinstance(v, dict):
            return {k: plain(x) for k, x in v.items()}
        if isinstance(v, (list, tuple)):
            return [plain(x) for x in v]
        return v
    return yaml.safe_dump(plain(cfg), sort_keys=False)


def load(argv: Optional[List[str]] = None, default_path: Optional[str] = None, default_name: str = "config") -> Cfg:
    ap = argparse.ArgumentParser(add_help=True)
    ap.add_argument("--config-path", "-cp", default=default_path)
    ap.add_argument("--config-name", "-cn", default=default_name)
    ap.add_argument("--cfg", choices=["job"], default=None, help="print the composed, resolved config and exit")
    ap.add_argument("overrides", nargs="*")
    a = ap.parse_args(argv)
    raw: dict = {}
    if a.config_path is not None:
        # `group=option` picks a config-group file when DIR/group/ exists; everything else is a value override
        choices, values = {}, []
        for ov in a.overrides:
            k, _, v = ov.partition("=")
            if _ and not ov.startswith(("+", "~")) and "." not in k and os.path.isdir(os.path.join(a.config_path, k)):
                choices[k] = v
            else:
                values.append(ov)
        raw = compose(a.config_path, a.config_name, choices)
        raw = apply_overrides(raw, values)
    else:
        raw = apply_overrides(raw, a.overrides)
    cfg = Cfg(resolve(raw))
    if a.cfg == "job":
        print(to_yaml(cfg))
        raise SystemExit(0)
    return cfg
