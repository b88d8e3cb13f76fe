"""Graph inspection / visualisation: a graph as a plain dict (ops, tensors, edges, placement, subgraphs), Graphviz DOT with module
clusters and forward / backward / communication colouring, and a self-contained HTML page (collapsible module tree + op table with
shapes and DistributedStates) that needs no server.
(ref: hetu/v1/python/graphboard -- the reference's TensorBoard-like viewer; here static files)"""
from __future__ import annotations

import html
import json
from typing import Dict, List, Optional


def graph_to_dict(g, strategy: Optional[int] = None) -> Dict:
    """{"ops": [{id, type, name, inputs: [tensor ids], outputs: [tensor ids], is_bwd, subgraph, placement, attrs}],
        "tensors": {id: {name, shape, dtype, producer, ds}}}"""
    ops, tensors = [], {}
    for i in range(g.num_ops):
        try:
            info = g.op_info(i)
        except Exception:      # noqa: BLE001 -- pruned ops of define-by-run graphs
            continue

        def tid(t):
            if t.id not in tensors:
                ds = None
                try:
                    d = t.get_ds(strategy or 0) if hasattr(t, "get_ds") else None
                    ds = {"states": dict(d.states), "order": list(d.order), "device_num": d.device_num} if d is not None else None
                except Exception:      # noqa: BLE001
                    ds = None
                tensors[t.id] = {"name": t.name, "shape": list(t.shape), "dtype": str(t.dtype), "producer": None, "ds": ds}
            return t.id
        ins, outs = [tid(t) for t in info["inputs"]], [tid(t) for t in info["outputs"]]
        for o in outs:
            tensors[o]["producer"] = i
        placement = info.get("placement")
        try:
            devs = [d.index for d in placement.devices] if placement is not None and not placement.empty else []
        except Exception:      # noqa: BLE001
            devs = []
        ops.append({"id": i, "type": info["type"], "name": info["name"], "inputs": ins, "outputs": outs, "is_bwd": bool(info["is_bwd"]),
                    "subgraph": info.get("subgraph") or "", "placement": devs,
                    "attrs": {k: v for k, v in info["attrs"].items() if isinstance(v, (int, float, str, bool, list))}})
    return {"name": g.name, "ops": ops, "tensors": tensors}


_COLORS = {"comm": "#f4b183", "bwd": "#9dc3e6", "update": "#c5e0b4", "leaf": "#d9d9d9", "fwd": "#fff2cc"}
_COMM = ("comm", "all_reduce", "all_gather", "reduce_scatter", "all_to_all", "hall_to_all", "broadcast_comm", "grouped_all_reduce",
         "grouped_reduce_scatter", "grouped_all_gather", "parallel_attn", "parallel_attn_bwd")


def _kind(op) -> str:
    if op["type"] in _COMM:
        return "comm"
    if op["type"] in ("variable", "placeholder", "const_tensor"):
        return "leaf"
    if op["type"].endswith("_update") or op["type"] == "group":
        return "update"
    return "bwd" if op["is_bwd"] else "fwd"


def to_dot(g, strategy: Optional[int] = None, max_ops: int = 5000) -> str:
    """Graphviz source: one node per op (label type + name + output shape), module subgraphs as clusters"""
    d = graph_to_dict(g, strategy)
    ops = d["ops"][:max_ops]
    lines = [f'digraph "{d["name"]}" {{', "  rankdir=TB; node [shape=box, style=filled, fontname=Helvetica, fontsize=10];"]
    by_sub: Dict[str, List] = {}
    for op in ops:
        by_sub.setdefault(op["subgraph"], []).append(op)
    for si, (sub, members) in enumerate(sorted(by_sub.items())):
        indent = "  "
        if sub:
            lines.append(f'  subgraph "cluster_{si}" {{ label="{sub}"; color="#888888";')
            indent = "    "
        for op in members:
            shape = d["tensors"][op["outputs"][0]]["shape"] if op["outputs"] else []
            label = f'{op["type"]}\\n{op["name"]}\\n{shape}'
            lines.append(f'{indent}op{op["id"]} [label="{label}", fillcolor="{_COLORS[_kind(op)]}"];')
        if sub:
            lines.append("  }")
    shown = {op["id"] for op in ops}
    for op in ops:
        for t in op["inputs"]:
            p = d["tensors"][t]["producer"]
            if p is not None and p in shown:
                lines.append(f'  op{p} -> op{op["id"]};')
    lines.append("}")
    return "\n".join(lines)


def to_html(g, path: str, strategy: Optional[int] = None, title: Optional[str] = None) -> str:
    """write a standalone page: summary counts, op-type histogram, collapsible module tree, searchable op table"""
    d = graph_to_dict(g, strategy)
    counts: Dict[str, int] = {}
    for op in d["ops"]:
        counts[op["type"]] = counts.get(op["type"], 0) + 1
    kinds = {k: sum(1 for op in d["ops"] if _kind(op) == k) for k in _COLORS}
    tree: Dict = {}
    for op in d["ops"]:
        node = tree
        for part in [p for p in op["subgraph"].split(".") if p]:
            node = node.setdefault(part, {})
        node.setdefault("__ops__", []).append(op["id"])

    def render(node, name="(graph)"):
        n_ops = len(node.get("__ops__", []))
        kids = "".join(render(v, k) for k, v in sorted(node.items()) if k != "__ops__")
        return f"<details><summary>{html.escape(name)} <small>({n_ops} ops)</small></summary>{kids}</details>"
    rows = []
    for op in d["ops"]:
        outs = ", ".join(f'{d["tensors"][t]["shape"]} {d["tensors"][t]["dtype"]}' for t in op["outputs"])
        ds = next((d["tensors"][t]["ds"] for t in op["outputs"] if d["tensors"][t]["ds"]), None)
        rows.append(f'<tr class="{_kind(op)}"><td>{op["id"]}</td><td>{html.escape(op["type"])}</td><td>{html.escape(op["name"])}</td>'
                    f'<td>{html.escape(op["subgraph"])}</td><td>{html.escape(outs)}</td><td>{html.escape(json.dumps(ds) if ds else "")}</td>'
                    f'<td>{op["placement"]}</td></tr>')
    css = "".join(f"tr.{k}{{background:{c}}}" for k, c in _COLORS.items())
    page = f"""<!doctype html><html><head><meta charset="utf-8"><title>{html.escape(title or d['name'])}</title>
<style>body{{font-family:Helvetica,Arial,sans-serif;margin:1.5em}}table{{border-collapse:collapse;font-size:12px}}td,th{{border:1px solid #bbb;padding:2px 6px}}{css}
details{{margin-left:1em}}input{{margin:.5em 0;padding:.3em;width:30em}}</style></head><body>
<h2>{html.escape(title or d['name'])}</h2>
<p>{len(d['ops'])} ops, {len(d['tensors'])} tensors &mdash; {', '.join(f'{k}: {v}' for k, v in kinds.items())}</p>
<p>{', '.join(f'{html.escape(k)} x{v}' for k, v in sorted(counts.items(), key=lambda kv: -kv[1]))}</p>
<h3>modules</h3>{render(tree)}
<h3>ops</h3><input id="q" placeholder="filter by type / name / module" oninput="f()">
<table id="t"><tr><th>id</th><th>type</th><th>name</th><th>module</th><th>outputs</th><th>distributed states</th><th>devices</th></tr>{''.join(rows)}</table>
<script>function f(){{var q=document.getElementById('q').value.toLowerCase();var r=document.getElementById('t').rows;
for(var i=1;i<r.length;i++)r[i].style.display=r[i].innerText.toLowerCase().indexOf(q)<0?'none':'';}}</script>
<script type="application/json" id="graph">{json.dumps(d)}</script></body></html>"""
    with open(path, "w") as f:
        f.write(page)
    return path
