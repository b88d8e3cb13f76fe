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


class ProcessGroupCollectives:
    """the collectives DistGCN15D needs, over the framework's communication runtime (NCCL / gloo process groups)"""

    def __init__(self, rank: int):
        from .. import _C
        self._C, self.rank = _C, rank

    def all_gather(self, x, ranks):
        return self._C.comm_all_gather(x.contiguous(), list(ranks), 0)

    def all_reduce(self, x, ranks):
        return self._C.comm_all_reduce(x.contiguous(), list(ranks), "sum")


class ThreadCollectives:
    """the same interface for p ranks living as threads of ONE process (a rendezvous table + barriers): lets the partitioned algorithm
    run and be checked without any network transport.  `ThreadCollectives.world(p)` -> one endpoint per rank."""

    class _Shared:
        def __init__(self, p):
            import threading
            self.p, self.lock, self.slots, self.barriers, self.threading = p, threading.Lock(), {}, {}, threading

        def barrier(self, key):
            with self.lock:
                if key not in self.barriers:
                    self.barriers[key] = self.threading.Barrier(len(key))
            return self.barriers[key]

    def __init__(self, shared, rank):
        self.shared, self.rank = shared, rank

    @classmethod
    def world(cls, p: int):
        shared = cls._Shared(p)
        return [cls(shared, r) for r in range(p)]

    def _exchange(self, x, ranks):
        key = tuple(ranks)
        sh = self.shared
        with sh.lock:
            sh.slots.setdefault(key, {})[self.rank] = x
        b = sh.barrier(key)
        b.wait(timeout=60)                       # everybody has deposited
        parts = [sh.slots[key][r] for r in key]
        b.wait(timeout=60)                       # everybody has read: the slot may be reused by the next call
        return parts

    def all_gather(self, x, ranks):
        import torch
        return torch.cat(self._exchange(x.clone(), ranks), 0)

    def all_reduce(self, x, ranks):
        parts = self._exchange(x.clone(), ranks)
        out = parts[0].clone()
        for t in parts[1:]:
            out = out + t
        return out


class DistGCN15D:
    """Full-graph GCN training with the 1.5-D partitioning, one instance per rank (p ranks, replication factor c, nb = p / c row
    blocks).  Rank d = (i, j) stores the adjacency block A[rows_i, :] restricted to the column chunk j (node chunks q with q % c == j)
    and the feature rows of block i.  One layer  Z = A (H W):
      forward   HW_i = H_i W (local GEMM);  all-gather HW over the ranks of the same replica index j (one per row block) -> the rows this
                rank's column chunk needs;  partial Z_i^(j) = A[i, chunk j] HW[chunk j];  all-reduce over the replica group -> Z_i
      backward  the normalised adjacency is symmetric, so dHW = A dZ uses the same communication pattern;  dW = H_i^T dHW_i summed over
                the row blocks (all-reduce over the data-parallel ranks of one replica index);  dH_i = dHW_i W^T
    Weights are replicated; every rank applies the same SGD step.  Communication per layer and rank: n / c rows gathered + n / nb rows
    reduced, against n rows for the 1-D layout (c = 1).
    (ref: hetu/v1/python/hetu/gpu_ops/DistGCN_15d.py, hetu/v1/examples/gnn/run_dist.py; CAGNET's 1.5-D algorithm)"""

    def __init__(self, indices: np.ndarray, values: np.ndarray, features: np.ndarray, labels: np.ndarray, dims: List[int], rank: int, p: int, c: int,
                 lr: float = 0.1, seed: int = 0, train_mask: np.ndarray = None, comm: str = "groups", collectives=None):
        """comm = "groups": gathers run inside the row group and reductions inside the replica group (the 1.5-D traffic pattern);
        comm = "world": the same data movement expressed with world-wide collectives only (gather everything, keep the chunk;
        reduce a zero-padded buffer) -- more bytes, but it avoids alternating collectives of overlapping process groups, which the
        gloo transport of some PyTorch builds dead-locks on intermittently (reproducible with plain torch.distributed)"""
        import torch
        self.torch = torch
        self.coll = collectives if collectives is not None else ProcessGroupCollectives(rank)
        self.rank, self.p, self.c, self.nb, self.lr = rank, p, c, p // c, lr
        n = features.shape[0]
        self.n = n
        part = partition_15d(n, p, c)[rank]
        self.i, self.j = rank // c, rank % c
        self.rows = part["rows"]
        self.replica_group, self.row_group = part["replica_group"], part["row_group"]
        per = (n + self.nb - 1) // self.nb
        self.per = per
        # column chunk j = the row blocks q with q % c == j (their feature rows arrive by the all-gather over row_group restricted to
        # those blocks); keep the block-local COO of A[rows_i, cols in chunk j]
        self.my_blocks = [q for q in range(self.nb) if q % c == self.j]
        lo, hi = self.rows
        col_block = indices[1] // per
        sel = (indices[0] >= lo) & (indices[0] < hi) & np.isin(col_block, self.my_blocks)
        # columns are re-indexed into the concatenation of the chunk's blocks (in block order)
        offset = {q: k * per for k, q in enumerate(self.my_blocks)}
        cols = np.array([offset[int(b)] + int(cc) - int(b) * per for b, cc in zip(col_block[sel], indices[1][sel])], dtype=np.int64)
        self.a_idx = torch.as_tensor(np.stack([indices[0][sel] - lo, cols]))
        self.a_val = torch.as_tensor(values[sel].astype(np.float32))
        self.a_shape = (hi - lo, len(self.my_blocks) * per)
        pad = per - (hi - lo)
        self.h0 = torch.nn.functional.pad(torch.as_tensor(features[lo:hi].astype(np.float32)), (0, 0, 0, pad))      # blocks padded to `per` rows
        self.labels = torch.as_tensor(labels[lo:hi].astype(np.int64))
        mask = np.ones(n, bool) if train_mask is None else train_mask
        self.mask = torch.as_tensor(mask[lo:hi])
        self.num_train = int(mask.sum())
        g = torch.Generator().manual_seed(seed)
        self.weights = [torch.randn(a, b, generator=g) * (1.0 / np.sqrt(a)) for a, b in zip(dims[:-1], dims[1:])]
        # ranks that share this replica index and together hold every row block exactly once: the gather / weight-gradient group
        self.dp_group = [q * c + self.j for q in range(self.nb)] if c > 1 else list(range(p))
        self.bytes_moved = 0
        assert comm in ("groups", "world")
        self.comm = comm
        self.world = list(range(p))

    # -- one distributed sparse product  out_i = A[i, :] X  given the local rows X_i (padded to `per` rows)
    def _spmm(self, x_local):
        torch, C = self.torch, self.coll
        # gather the blocks of this column chunk: all ranks of the row_group (same j, every row block) exchange their rows, the chunk
        # keeps the blocks q % c == j -- with c > 1 only n / c of the rows travel to each rank
        need = [q * self.c + self.j for q in self.my_blocks]            # owner of block q inside the replica index j ... every (q, j) holds block q
        if self.comm == "world" and self.p > 1:
            everything = C.all_gather(x_local, self.world).reshape(self.p, self.per, -1)
            blocks = everything[self.row_group]                          # the rows the row group would have exchanged
        else:
            gathered = C.all_gather(x_local, self.row_group) if len(self.row_group) > 1 else x_local
            blocks = gathered.reshape(len(self.row_group), self.per, -1)
        if need:
            pick = torch.cat([blocks[self.row_group.index(r)] for r in need], 0)
            self.bytes_moved += pick.numel() * 4
            a = torch.sparse_coo_tensor(self.a_idx, self.a_val, self.a_shape)
            partial = torch.sparse.mm(a, pick)
        else:                                          # more replicas than row blocks: this replica owns no column chunk
            partial = torch.zeros(self.rows[1] - self.rows[0], x_local.shape[1])
        partial = torch.nn.functional.pad(partial, (0, 0, 0, self.per - partial.shape[0]))
        if len(self.replica_group) > 1:
            if self.comm == "world":
                slots = torch.zeros(self.nb, self.per, partial.shape[1])
                slots[self.i] = partial                                   # replicas of a row block add up in its slot
                partial = C.all_reduce(slots, self.world)[self.i]
            else:
                partial = C.all_reduce(partial, self.replica_group)
            self.bytes_moved += partial.numel() * 4
        return partial

    def step(self) -> float:
        """one full-batch training step -> global mean cross-entropy over the training nodes"""
        torch, C = self.torch, self.coll
        hs, zs = [self.h0], []
        h = self.h0
        for li, w in enumerate(self.weights):
            z = self._spmm(h @ w)
            zs.append(z)
            h = torch.relu(z) if li < len(self.weights) - 1 else z
            hs.append(h)
        rows = self.rows[1] - self.rows[0]
        logits = h[:rows]
        logp = torch.log_softmax(logits, -1)
        m = self.mask.float()
        loss_sum = -(logp[torch.arange(rows), self.labels] * m).sum()
        # every replica of a row block computes the same loss: count each block once (replica 0)
        t = torch.stack([loss_sum if self.j == 0 else torch.zeros(())]).reshape(1)
        if self.p > 1:
            t = C.all_reduce(t, list(range(self.p)))
        loss = float(t[0]) / max(self.num_train, 1)
        # backward
        dz = (torch.softmax(logits, -1) - torch.nn.functional.one_hot(self.labels, logits.shape[1]).float()) * m[:, None] / max(self.num_train, 1)
        dz = torch.nn.functional.pad(dz, (0, 0, 0, self.per - rows))
        grads = [None] * len(self.weights)
        for li in reversed(range(len(self.weights))):
            if li < len(self.weights) - 1:
                dz = dz * (zs[li] > 0).float()
            dhw = self._spmm(dz)                                       # A^T dZ = A dZ (symmetric normalisation)
            gw = hs[li].t() @ dhw
            if len(self.dp_group) > 1:
                if self.comm == "world" and self.c > 1:
                    gw = C.all_reduce(gw, self.world) / float(self.c)      # every replica index holds the same sum
                else:
                    gw = C.all_reduce(gw, self.dp_group)
            grads[li] = gw
            dz = dhw @ self.weights[li].t()
        for w, g in zip(self.weights, grads):
            w -= self.lr * g
        return loss
