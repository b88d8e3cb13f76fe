"""Graph neural networks of the v1 examples: GCN / GraphSage layers over a sparse normalised adjacency (COO spmm op) and
the 1.5-D partitioning of DistGCN (ref: hetu/v1/examples/gnn, hetu/v1/python/hetu/gpu_ops/DistGCN_15d.py): the adjacency is
split into `p / c` row blocks, every block is replicated `c` times, each replica multiplies its block with 1/c of the
feature rows and the partial products are summed inside the replica group."""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

from .. import ops
from ..nn import Linear, Module, ModuleList


def normalise_adjacency(edges: np.ndarray, num_nodes: int) -> Tuple[np.ndarray, np.ndarray]:
    """symmetric normalisation D^-1/2 (A + I) D^-1/2 of an edge list [2, E] -> (indices [2, nnz], values [nnz])"""
    e = np.concatenate([edges, edges[::-1], np.tile(np.arange(num_nodes), (2, 1))], 1)
    e = np.unique(e, axis=1)
    deg = np.bincount(e[0], minlength=num_nodes).astype(np.float32)
    v = 1.0 / np.sqrt(deg[e[0]] * deg[e[1]])
    return e.astype(np.int64), v.astype(np.float32)


class GCNLayer(Module):
    def __init__(self, in_features, out_features, activation="relu", name="gcn"):
        super().__init__()
        self.lin = Linear(in_features, out_features, name=f"{name}_lin")
        self.activation = activation

    def forward(self, indices, values, x, num_nodes):
        h = ops.spmm(indices, values, self.lin(x), num_nodes)       # A (X W): project first, the sparse product is cheaper
        return getattr(ops, self.activation)(h) if self.activation else h


class GraphSageLayer(Module):
    """mean aggregator: concat(x, mean_neighbours(x)) W"""

    def __init__(self, in_features, out_features, activation="relu", name="sage"):
        super().__init__()
        self.lin = Linear(2 * in_features, out_features, name=f"{name}_lin")
        self.activation = activation

    def forward(self, indices, mean_values, x, num_nodes):
        h = self.lin(ops.concat([x, ops.spmm(indices, mean_values, x, num_nodes)], 1))
        return getattr(ops, self.activation)(h) if self.activation else h


class GCN(Module):
    def __init__(self, in_features, hidden, num_classes, num_layers=2):
        super().__init__()
        dims = [in_features] + [hidden] * (num_layers - 1) + [num_classes]
        self.layers = ModuleList([GCNLayer(a, b, "relu" if i < num_layers - 1 else None, name=f"gcn{i}")
                                  for i, (a, b) in enumerate(zip(dims[:-1], dims[1:]))])

    def forward(self, indices, values, x, num_nodes, labels=None, mask=None):
        for l in self.layers:
            x = l(indices, values, x, num_nodes)
        if labels is None:
            return x
        loss = ops.softmax_cross_entropy_sparse(x, labels, ignored_index=-1, reduction="mean")
        return loss, x


def partition_15d(num_nodes: int, p: int, c: int) -> List[dict]:
    """1.5-D layout for p devices with replication factor c: device (i, j) (i < p / c, j < c) owns adjacency row block i and
    multiplies it with feature row chunk j of every column block; -> per device {"rows": (lo, hi), "col_chunk": j,
    "replica_group": [...], "row_group": [...]}"""
    assert p % c == 0
    nb = p // c
    per = (num_nodes + nb - 1) // nb
    out = []
    for d in range(p):
        i, j = d // c, d % c
        out.append({"device": d, "rows": (i * per, min(num_nodes, (i + 1) * per)), "col_chunk": j,
                    "replica_group": [i * c + k for k in range(c)], "row_group": [k * c + j for k in range(nb)]})
    return out


def dist_gcn_15d_forward(indices: np.ndarray, values: np.ndarray, x: np.ndarray, w: np.ndarray, p: int, c: int) -> np.ndarray:
    """reference execution of one 1.5-D GCN layer (all devices simulated in-process): every device multiplies its
    adjacency block restricted to its column chunk with the matching feature rows; replicas sum their partial products"""
    n = x.shape[0]
    parts = partition_15d(n, p, c)
    xw = x @ w
    out = np.zeros((n, w.shape[1]), np.float32)
    col_per = (n + c - 1) // c
    for d in parts:
        lo, hi = d["rows"]
        clo, chi = d["col_chunk"] * col_per, min(n, (d["col_chunk"] + 1) * col_per)
        sel = (indices[0] >= lo) & (indices[0] < hi) & (indices[1] >= clo) & (indices[1] < chi)
        np.add.at(out, indices[0][sel], values[sel, None] * xw[indices[1][sel]])      # partial product, reduced over the replica group
    return out
