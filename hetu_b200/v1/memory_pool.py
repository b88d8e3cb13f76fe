"""Static buffer-reuse planning over a v1 graph: walk the nodes in execution order, return a node's buffer to a (shape, dtype) pool
when its last consumer has run, and let later nodes of the same shape take it.  The plan ({node -> node whose buffer it reuses}) and
the peak it implies are what the v1 searches use to reject strategies that do not fit; at run time this framework's allocator does
the equivalent dynamically.  (ref: hetu/v1/python/hetu/memory_pool.py HetuMemoryPool.compute_memory_reuse_plan :32, memory_plan :87)"""
from __future__ import annotations

from collections import defaultdict
from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np

_PERSISTENT_TYPES = {"placeholder", "variable", "all_reduce", "pipeline_recv", "comm"}
_INPLACE_TYPES = {"reshape", "view", "detach", "stop_gradient", "contiguous_view", "group"}
_ITEMSIZE = {"float32": 4, "float16": 2, "bfloat16": 2, "int64": 8, "int32": 4, "int8": 1, "uint8": 1, "bool": 1, "float64": 8}


class _Node:
    __slots__ = ("id", "type", "inputs", "shape", "dtype", "inplace")

    def __init__(self, id, type, inputs, shape, dtype):     # noqa: A002
        self.id, self.type, self.inputs, self.shape, self.dtype = id, type, inputs, tuple(shape), dtype
        self.inplace = type in _INPLACE_TYPES


def nodes_of_graph(graph) -> List[_Node]:
    """the graph's ops in creation (= a valid execution) order, one entry per output tensor"""
    out = []
    for i in range(graph.num_ops):
        try:
            info = graph.op_info(i)
        except Exception:      # noqa: BLE001 -- pruned op
            continue
        for t in info["outputs"]:
            out.append(_Node(t.id, info["type"], [x.id for x in info["inputs"]], list(t.shape), str(t.dtype)))
    return out


class HetuMemoryPool:
    def compute_memory_reuse_plan(self, computing_nodes: Sequence, node_to_shape: Optional[Dict] = None, eval_node_list: Iterable = ()) -> Dict:
        """`computing_nodes`: objects with `.id .type .inputs .shape .dtype .inplace` in execution order (see nodes_of_graph) --
        -> {node id: id of the earlier node whose buffer it takes over}"""
        keep = {getattr(n, "id", n) for n in eval_node_list}
        by_id = {n.id: n for n in computing_nodes}
        for n in computing_nodes:
            if n.type in _PERSISTENT_TYPES:
                keep.add(n.id)
                if n.type in ("all_reduce", "pipeline_send"):
                    keep.update(n.inputs)
        outdeg = {n.id: 0 for n in computing_nodes}
        for n in computing_nodes:
            for i in n.inputs:
                if i in outdeg:
                    outdeg[i] += 1
        pool, reuse = defaultdict(list), {}

        def release(i):
            if i not in by_id:
                return
            outdeg[i] -= 1
            if outdeg[i] > 0 or i in keep:
                return
            node = by_id[i]
            if node.inplace:                       # a view dies with its last reader: that frees the buffer it aliases
                for j in node.inputs:
                    release(j)
            else:
                pool[(node.shape, node.dtype)].append(reuse.get(i, i))

        for n in computing_nodes:
            if n.inplace:
                continue
            key = (n.shape, n.dtype)
            if n.id not in keep and pool[key]:
                reuse[n.id] = pool[key].pop()
            for i in n.inputs:
                release(i)
        return reuse

    @staticmethod
    def nbytes(node) -> int:
        return int(np.prod(node.shape or (1,))) * _ITEMSIZE.get(node.dtype, 4)

    def memory_plan(self, computing_nodes: Sequence, eval_node_list: Iterable = ()) -> Dict:
        """-> {"reuse": plan, "allocated_bytes": bytes of distinct buffers, "naive_bytes": one buffer per node, "buffers": count}"""
        reuse = self.compute_memory_reuse_plan(computing_nodes, None, eval_node_list)
        owners = [n for n in computing_nodes if not n.inplace and n.id not in reuse]
        return {"reuse": reuse, "allocated_bytes": sum(self.nbytes(n) for n in owners),
                "naive_bytes": sum(self.nbytes(n) for n in computing_nodes if not n.inplace), "buffers": len(owners)}

    def plan_graph(self, graph, fetches: Iterable = ()) -> Dict:
        return self.memory_plan(nodes_of_graph(graph), fetches)

    def test_memory(self, devices, task_graph: Dict, capacity_bytes: Optional[int] = None) -> bool:
        """does every device's share fit?  task_graph: {device: [nodes]}; capacity defaults to 180 GB (B200 HBM3e)"""
        cap = capacity_bytes if capacity_bytes is not None else 180 * 2 ** 30
        return all(self.memory_plan(task_graph.get(d, []))["allocated_bytes"] <= cap for d in devices)
