"""Op manifest tooling.  The reference generates its Python op bindings from `codegen/ops.yml` (one C++ parser per entry); here ops
live in a C++ registry reached through one generic `make_op`, so nothing has to be generated to USE an op.  What remains useful is the
manifest itself: parse the reference's yml format, check a manifest against the registry / the Python wrappers, and emit `.pyi` stubs
for editors.  `python -m hetu_b200.tools.gen_py_ops --input ops.yml --output-dir out/`
(ref: python/hetu/_binding/codegen/{gen_py_ops,args_bridge}.py)"""
from __future__ import annotations

import argparse
import inspect
import os
import re
from typing import Dict, List, Optional, Tuple

import yaml

_PY_TYPE = {"Tensor": "Tensor", "TensorList": "Sequence[Tensor]", "float": "float", "double": "float", "int": "int", "int64_t": "int",
            "bool": "bool", "std::string": "str", "string": "str", "HTShape": "Sequence[int]", "HTAxes": "Sequence[int]",
            "HTStride": "Sequence[int]", "IntSymbol": "IntSymbol", "SyShape": "Sequence[IntSymbol]", "DataType": "str",
            "DistributedStatesHierarchy": "Sequence", "List[int]": "Sequence[int]", "std::vector<int64_t>": "Sequence[int]"}


class ArgType:
    """one parsed argument of a manifest entry: C++-side type name, Python annotation, name, default (None = required)"""

    def __init__(self, type_str: str, name: str, default: Optional[str] = None):
        self.type_str, self.name, self.default = type_str.strip(), name.strip(), default
        self.py_type = _PY_TYPE.get(self.type_str.replace("const ", "").replace("&", "").strip(), "Any")
        self.optional = default is not None

    def signature(self) -> str:
        d = {"None": "None", "true": "True", "false": "False"}.get(self.default, self.default)
        return f"{self.name}: {self.py_type}" + (f" = {d}" if self.optional else "")

    def __repr__(self):
        return f"ArgType({self.type_str} {self.name}{'=' + self.default if self.optional else ''})"


def parse_args(args: str, kernel_or_operator: str = "", self_arg_name: Optional[str] = None, ret_type_str: Optional[str] = None) -> List[ArgType]:
    """'Tensor input, HTAxes axes=None, bool keepdims=false' -> [ArgType...]; commas inside [] / <> / () do not split"""
    out, depth, cur = [], 0, ""
    for ch in args or "":
        depth += ch in "[<(" 
        depth -= ch in "]>)"
        if ch == "," and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    parsed = []
    for a in out:
        a = a.strip()
        default = None
        if "=" in a:
            a, default = (x.strip() for x in a.split("=", 1))
        type_str, _, name = a.rpartition(" ")
        parsed.append(ArgType(type_str or "Any", name, default))
    return parsed


def load_manifest(input_file: str) -> List[Dict]:
    entries = yaml.safe_load(open(input_file)) or []
    for e in entries:
        e["parsed_args"] = parse_args(e.get("args", ""), e.get("op", ""), e.get("self"))
    return entries


def check_manifest(entries: List[Dict]) -> Dict[str, List[str]]:
    """which python names of the manifest this framework lacks, and which keyword names its wrappers would reject"""
    import hetu_b200 as ht
    missing, bad_kwargs = [], []
    for e in entries:
        fn = getattr(ht, e["name"], None)
        if fn is None:
            missing.append(e["name"])
            continue
        try:
            params = inspect.signature(fn).parameters
        except (TypeError, ValueError):
            continue
        if any(p.kind == p.VAR_KEYWORD for p in params.values()):
            continue
        bad = [a.name for a in e["parsed_args"] if a.name not in params]
        if bad:
            bad_kwargs.append(f'{e["name"]}({", ".join(bad)})')
        if e.get("self") and not hasattr(ht.Tensor, e["name"]):
            missing.append(f'Tensor.{e["name"]}')
    return {"missing": sorted(set(missing)), "bad_kwargs": bad_kwargs}


def gen_stubs(entries: List[Dict]) -> str:
    """`.pyi` text: one overload per manifest entry (module-level function; Tensor method when the entry has a `self` argument)"""
    by_name: Dict[str, List[Dict]] = {}
    for e in entries:
        by_name.setdefault(e["name"], []).append(e)
    lines = ["from typing import Any, Sequence, overload", "from hetu_b200 import IntSymbol, Tensor", ""]
    methods: List[Tuple[str, str]] = []
    for name, es in by_name.items():
        for e in es:
            sig = ", ".join(a.signature() for a in e["parsed_args"])
            if len(es) > 1:
                lines.append("@overload")
            lines.append(f"def {name}({sig}{', ' if sig else ''}**op_meta: Any) -> Tensor: ...   # {e.get('op', '')}")
            if e.get("self"):
                rest = ", ".join(a.signature() for a in e["parsed_args"] if a.name != e["self"])
                methods.append((name, f"    def {name}(self{', ' + rest if rest else ''}, **op_meta: Any) -> Tensor: ..."))
        lines.append("")
    if methods:
        lines.append("class _TensorOps:")
        seen = set()
        for name, text in methods:
            if (name, text) not in seen:
                seen.add((name, text))
                lines.append(text)
    return "\n".join(lines) + "\n"


def dump_registry(path: Optional[str] = None) -> List[Dict]:
    """the C++ op registry as a manifest (name + the python wrapper's signature when there is one)"""
    import hetu_b200 as ht
    from hetu_b200 import ops as pyops
    out = []
    for name in sorted(ht._C.list_ops()):
        fn = getattr(pyops, name, None)
        try:
            sig = str(inspect.signature(fn)) if fn is not None else ""
        except (TypeError, ValueError):
            sig = ""
        out.append({"op": name, "python": sig})
    if path:
        with open(path, "w") as f:
            yaml.safe_dump(out, f, sort_keys=False)
    return out


def gen_ops(input_file: str, output_dir: str) -> Dict:
    """manifest -> `<output_dir>/ops.pyi` + `<output_dir>/registry.yml`; returns the check report"""
    os.makedirs(output_dir, exist_ok=True)
    entries = load_manifest(input_file)
    with open(os.path.join(output_dir, "ops.pyi"), "w") as f:
        f.write(gen_stubs(entries))
    dump_registry(os.path.join(output_dir, "registry.yml"))
    report = check_manifest(entries)
    report["entries"] = len(entries)
    return report


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--input", type=str, required=True, help="path to a manifest in the reference's ops.yml format")
    ap.add_argument("--output-dir", type=str, required=True)
    a = ap.parse_args(argv)
    rep = gen_ops(a.input, a.output_dir)
    print(f'{rep["entries"]} entries; missing: {rep["missing"] or "none"}; keyword mismatches: {rep["bad_kwargs"] or "none"}')
    return 0 if not rep["missing"] else 1


if __name__ == "__main__":
    raise SystemExit(main())
