from __future__ import annotations

import math
from typing import Sequence

import numpy as np

from ... import ops
from ...core import constant_initializer, from_numpy, normal_initializer, ones_initializer, parallel_parameter, zeros_initializer
from ...nn import Linear, Module, ModuleList


def _table(rows, dim, name, std=None):
    init = normal_initializer(0.0, std if std is not None else 1.0 / math.sqrt(dim))
    return parallel_parameter(init, [rows, dim], None, requires_grad=True, name=name)


class _Base(Module):
    def __init__(self, num_embeddings, dim):
        super().__init__()
        self.num_embeddings, self.dim = num_embeddings, dim

    def _flat(self, ids):
        n = 1
        for s in ids.shape:
            n *= s
        return ops.reshape(ids, [n]), list(ids.shape)

    def _shape_out(self, e, shape):
        return ops.reshape(e, shape + [self.dim])

    def num_parameters(self) -> int:
        return int(sum(np.prod(p.shape) for _, p in self.named_parameters()))

    def compression_ratio(self) -> float:
        return self.num_embeddings * self.dim / max(self.num_parameters(), 1)


class HashEmbedding(_Base):
    """hashing trick: ids share `buckets` rows (optionally the sum of several independent hashes)"""

    def __init__(self, num_embeddings, dim, buckets, num_hashes=1, name="hash_emb"):
        super().__init__(num_embeddings, dim)
        self.buckets, self.num_hashes = buckets, num_hashes
        self.weight = _table(buckets, dim, f"{name}_weight")

    def forward(self, ids):
        f, shape = self._flat(ids)
        e = None
        for h in range(self.num_hashes):
            v = ops.embedding_lookup(self.weight, ops.hash_ids(f, self.buckets, a=1000003 + 7919 * h, b=12345 + h))
            e = v if e is None else e + v
        return self._shape_out(e, shape)


class CompositionalEmbedding(_Base):
    """quotient-remainder trick: e = Q[id // m] (*|+) R[id % m] with m ~ sqrt(N)"""

    def __init__(self, num_embeddings, dim, op="mul", name="qr_emb"):
        super().__init__(num_embeddings, dim)
        self.m = int(math.ceil(math.sqrt(num_embeddings)))
        self.op = op
        self.q = _table((num_embeddings + self.m - 1) // self.m, dim, f"{name}_q")
        self.r = _table(self.m, dim, f"{name}_r")

    def forward(self, ids):
        f, shape = self._flat(ids)
        a = ops.embedding_lookup(self.q, ops.floor_divide(f, self.m))
        b = ops.embedding_lookup(self.r, ops.remainder(f, self.m))
        return self._shape_out(a * b if self.op == "mul" else a + b, shape)


class TensorTrainEmbedding(_Base):
    """TT-Rec: the table is a product of 3 cores G1[n1, d1*r] G2[n2, r*d2*r] G3[n3, r*d3], N <= n1 n2 n3, dim = d1 d2 d3"""

    def __init__(self, num_embeddings, dim, rank=8, name="tt_emb"):
        super().__init__(num_embeddings, dim)
        n = int(math.ceil(num_embeddings ** (1.0 / 3)))
        self.ns = [n, n, (num_embeddings + n * n - 1) // (n * n)]
        d = int(round(dim ** (1.0 / 3)))
        self.ds = [d, d, dim // (d * d)]
        assert self.ds[0] * self.ds[1] * self.ds[2] == dim, "dim must factor into three integers (e.g. 64 = 4*4*4, 27, 125)"
        self.rank = rank
        s = 1.0 / math.sqrt(rank)
        self.g1 = _table(self.ns[0], self.ds[0] * rank, f"{name}_g1", s)
        self.g2 = _table(self.ns[1], rank * self.ds[1] * rank, f"{name}_g2", s)
        self.g3 = _table(self.ns[2], rank * self.ds[2], f"{name}_g3", s)

    def forward(self, ids):
        f, shape = self._flat(ids)
        n1, n2 = self.ns[0], self.ns[1]
        i1 = ops.remainder(f, n1)
        i2 = ops.remainder(ops.floor_divide(f, n1), n2)
        i3 = ops.floor_divide(f, n1 * n2)
        b = f.shape[0]
        a = ops.reshape(ops.embedding_lookup(self.g1, i1), [b, self.ds[0], self.rank])
        m = ops.reshape(ops.embedding_lookup(self.g2, i2), [b, self.rank, self.ds[1] * self.rank])
        c = ops.reshape(ops.embedding_lookup(self.g3, i3), [b, self.rank, self.ds[2]])
        am = ops.reshape(ops.bmm(a, m), [b, self.ds[0] * self.ds[1], self.rank])
        return self._shape_out(ops.reshape(ops.bmm(am, c), [b, self.dim]), shape)


class DeepHashEmbedding(_Base):
    """DHE: k hash values of the id, normalised to [-1, 1], decoded by an MLP -- no table at all"""

    def __init__(self, num_embeddings, dim, num_hashes=128, hidden: Sequence[int] = (256, 256), name="dhe"):
        super().__init__(num_embeddings, dim)
        self.k, self.mod = num_hashes, 1000003
        dims = [num_hashes] + list(hidden) + [dim]
        self.mlp = ModuleList([Linear(a, b, name=f"{name}_mlp{i}") for i, (a, b) in enumerate(zip(dims[:-1], dims[1:]))])

    def forward(self, ids):
        f, shape = self._flat(ids)
        feats = [ops.reshape(ops.hash_ids(f, self.mod, a=1000003 + 104729 * h, b=7 + 13 * h), [f.shape[0], 1]) for h in range(self.k)]
        x = ops.cast(ops.concat(feats, 1), "float32") * (2.0 / self.mod) - 1.0
        for i, l in enumerate(self.mlp):
            x = l(x, act="relu" if i < len(self.mlp) - 1 else "none")
        return self._shape_out(x, shape)


class RobeEmbedding(_Base):
    """ROBE-Z: one shared parameter array; an embedding is `dim / Z` chunks of Z consecutive cells at hashed offsets"""

    def __init__(self, num_embeddings, dim, array_size, chunk=8, name="robe"):
        super().__init__(num_embeddings, dim)
        assert dim % chunk == 0
        self.size, self.chunk = array_size // chunk, chunk
        self.weight = _table(self.size, chunk, f"{name}_array", 1.0 / math.sqrt(dim))

    def forward(self, ids):
        f, shape = self._flat(ids)
        parts = [ops.embedding_lookup(self.weight, ops.hash_ids(f, self.size, a=1000003 + 15485863 * c, b=c)) for c in range(self.dim // self.chunk)]
        return self._shape_out(ops.concat(parts, 1), shape)


class ProductQuantizedEmbedding(_Base):
    """DPQ (inference form): each id stores `subspaces` small codes; vectors are concatenated codebook entries.  Training
    uses the soft assignment over codebooks (softmax of query . codebook), so codes stay differentiable."""

    def __init__(self, num_embeddings, dim, subspaces=4, codes=256, name="dpq"):
        super().__init__(num_embeddings, dim)
        assert dim % subspaces == 0
        self.S, self.K, self.sub = subspaces, codes, dim // subspaces
        self.query = _table(num_embeddings, dim, f"{name}_query")          # training-time query table (dropped after export)
        self.codebooks = ModuleList()
        self.books = [_table(codes, self.sub, f"{name}_book{s}") for s in range(subspaces)]
        for s, b in enumerate(self.books):
            self.register_parameter(f"book{s}", b)

    def forward(self, ids):
        f, shape = self._flat(ids)
        q = ops.embedding_lookup(self.query, f)
        outs = []
        for s in range(self.S):
            qs = ops.slice(q, [0, s * self.sub], [f.shape[0], self.sub])
            att = ops.softmax(ops.linear(qs, self.books[s], None, trans_b=True) * (1.0 / math.sqrt(self.sub)), -1)
            outs.append(ops.matmul(att, self.books[s]))
        return self._shape_out(ops.concat(outs, 1), shape)

    def num_parameters(self):     # exported size: codes (log2 K bits each) + codebooks
        return int(self.num_embeddings * self.S * math.log2(self.K) / 32 + self.S * self.K * self.sub)


class MGQEmbedding(ProductQuantizedEmbedding):
    """MGQE: frequent ids may use all K codes, rare ids only the first K_rare (multi-granular quantisation)"""

    def __init__(self, num_embeddings, dim, subspaces=4, codes=256, rare_codes=64, hot_threshold=None, name="mgqe"):
        super().__init__(num_embeddings, dim, subspaces, codes, name)
        self.rare_codes, self.hot_threshold = rare_codes, hot_threshold if hot_threshold is not None else num_embeddings // 10

    def forward(self, ids):
        f, shape = self._flat(ids)
        q = ops.embedding_lookup(self.query, f)
        hot = ops.less(f, self.hot_threshold)            # ids are sorted by frequency (hot ids first)
        outs = []
        for s in range(self.S):
            qs = ops.slice(q, [0, s * self.sub], [f.shape[0], self.sub])
            logits = ops.linear(qs, self.books[s], None, trans_b=True) * (1.0 / math.sqrt(self.sub))
            rare_mask = from_numpy(np.concatenate([np.zeros(self.rare_codes), np.full(self.K - self.rare_codes, -1e9)]).astype(np.float32))
            logits = logits + ops.reshape(1.0 - hot, [f.shape[0], 1]) * rare_mask
            outs.append(ops.matmul(ops.softmax(logits, -1), self.books[s]))
        return self._shape_out(ops.concat(outs, 1), shape)


class MixedDimEmbedding(_Base):
    """MDE: frequency-sorted id blocks get smaller dimensions, projected up to `dim`"""

    def __init__(self, num_embeddings, dim, block_bounds: Sequence[int] = None, alpha=0.3, name="mde"):
        super().__init__(num_embeddings, dim)
        bounds = list(block_bounds) if block_bounds else [num_embeddings // 100, num_embeddings // 10, num_embeddings]
        self.bounds = [b for b in bounds if b > 0]
        self.tables, self.projs, lo = [], ModuleList(), 0
        for i, hi in enumerate(self.bounds):
            d = max(2, int(dim * ((lo + 1) / self.bounds[0]) ** (-alpha))) if i else dim
            d = min(dim, d)
            t = _table(hi - lo, d, f"{name}_t{i}")
            self.register_parameter(f"table{i}", t)
            self.tables.append((lo, hi, d, t))
            self.projs.append(Linear(d, dim, bias=False, name=f"{name}_p{i}"))
            lo = hi

    def forward(self, ids):
        f, shape = self._flat(ids)
        out = None
        for (lo, hi, d, t), p in zip(self.tables, self.projs):
            inside = ops.greater_equal(f, lo) * ops.less(f, hi) if lo else ops.less(f, hi)
            local = ops.clamp(f - lo, 0, hi - lo - 1)
            v = p(ops.embedding_lookup(t, local)) * ops.reshape(inside, [f.shape[0], 1])
            out = v if out is None else out + v
        return self._shape_out(out, shape)


class AutoDimEmbedding(_Base):
    """AutoDim-style search: candidate dimensions share the table prefix; a learnable softmax over candidates weights them"""

    def __init__(self, num_embeddings, dim, candidates: Sequence[int] = (2, 4, 8, 16), name="autodim"):
        super().__init__(num_embeddings, dim)
        self.cands = [c for c in candidates if c <= dim] + ([dim] if dim not in candidates else [])
        self.weight = _table(num_embeddings, dim, f"{name}_weight")
        self.alpha = parallel_parameter(zeros_initializer(), [len(self.cands)], None, requires_grad=True, name=f"{name}_alpha")
        self.projs = ModuleList([Linear(c, dim, bias=False, name=f"{name}_proj{c}") for c in self.cands])

    def forward(self, ids):
        f, shape = self._flat(ids)
        e = ops.embedding_lookup(self.weight, f)
        w = ops.softmax(ops.reshape(self.alpha, [1, len(self.cands)]), -1)
        out = None
        for i, (c, p) in enumerate(zip(self.cands, self.projs)):
            v = p(ops.slice(e, [0, 0], [f.shape[0], c])) * ops.slice(w, [0, i], [1, 1])
            out = v if out is None else out + v
        return self._shape_out(out, shape)

    def selected_dim(self, alpha_values) -> int:
        return self.cands[int(np.argmax(alpha_values))]


class PrunedEmbedding(_Base):
    """PEP: learnable soft threshold  e = sign(w) relu(|w| - sigmoid(s))  (per-dimension thresholds)"""

    def __init__(self, num_embeddings, dim, init_threshold=-6.0, name="pep"):
        super().__init__(num_embeddings, dim)
        self.weight = _table(num_embeddings, dim, f"{name}_weight")
        self.s = parallel_parameter(constant_initializer(init_threshold), [1, dim], None, requires_grad=True, name=f"{name}_s")

    def forward(self, ids):
        f, shape = self._flat(ids)
        w = ops.embedding_lookup(self.weight, f)
        mag = ops.relu(ops.abs(w) - ops.sigmoid(self.s))
        return self._shape_out(ops.tanh(w * 1e4) * mag, shape)          # tanh(1e4 w) ~ sign(w), keeps the graph differentiable


class DeepLightEmbedding(_Base):
    """DeepLight: magnitude pruning with a sparsity that ramps up during training (mask refreshed by `update_mask`)"""

    def __init__(self, num_embeddings, dim, target_sparsity=0.9, name="deeplight"):
        super().__init__(num_embeddings, dim)
        self.weight = _table(num_embeddings, dim, f"{name}_weight")
        self.mask = parallel_parameter(ones_initializer(), [num_embeddings, dim], None, requires_grad=False, name=f"{name}_mask")
        self.target = target_sparsity

    def sparsity_at(self, step, ramp=100.0):
        return self.target * (1.0 - 0.99 ** (step / ramp))

    def update_mask(self, weight_values: np.ndarray, step: int) -> np.ndarray:
        k = int(self.sparsity_at(step) * weight_values.size)
        thr = np.partition(np.abs(weight_values).reshape(-1), k)[k] if k > 0 else -1.0
        return (np.abs(weight_values) > thr).astype(np.float32)

    def forward(self, ids):
        f, shape = self._flat(ids)
        return self._shape_out(ops.embedding_lookup(self.weight, f) * ops.embedding_lookup(self.mask, f), shape)


class OptEmbedEmbedding(_Base):
    """OptEmbed: row mask from a learnable per-row norm threshold + a sampled dimension mask during supernet training"""

    def __init__(self, num_embeddings, dim, name="optembed"):
        super().__init__(num_embeddings, dim)
        self.weight = _table(num_embeddings, dim, f"{name}_weight")
        self.threshold = parallel_parameter(zeros_initializer(), [1], None, requires_grad=True, name=f"{name}_threshold")
        self.dim_mask = None

    def set_dim_mask(self, keep: int):
        self.dim_mask = from_numpy(np.concatenate([np.ones(keep), np.zeros(self.dim - keep)]).astype(np.float32).reshape(1, self.dim))

    def forward(self, ids):
        f, shape = self._flat(ids)
        w = ops.embedding_lookup(self.weight, f)
        norm = ops.sum(ops.abs(w), [1], True)
        gate = ops.sigmoid((norm - self.threshold) * 50.0)               # smooth step: rows below the threshold are dropped
        e = w * gate
        if self.dim_mask is not None:
            e = e * self.dim_mask
        return self._shape_out(e, shape)


class ALPTEmbedding(_Base):
    """ALPT: the table is trained in low precision with a learnable per-row step size (fake-quantised forward)"""

    def __init__(self, num_embeddings, dim, bits=8, name="alpt"):
        super().__init__(num_embeddings, dim)
        self.weight = _table(num_embeddings, dim, f"{name}_weight")
        self.step = parallel_parameter(constant_initializer(0.01), [num_embeddings, 1], None, requires_grad=True, name=f"{name}_step")
        self.qmax = float(2 ** (bits - 1) - 1)

    def forward(self, ids):
        f, shape = self._flat(ids)
        w, s = ops.embedding_lookup(self.weight, f), ops.abs(ops.embedding_lookup(self.step, f)) + 1e-8
        q = ops.clamp(w / s, -self.qmax, self.qmax)
        q = q + ops.stop_gradient(ops.round(q) - q) if hasattr(ops, "stop_gradient") else ops.round(q)
        return self._shape_out(q * s, shape)


class AdaptiveEmbedding(_Base):
    """AdaptEmb: the `hot` most frequent ids own exclusive rows, the rest share a hashed table"""

    def __init__(self, num_embeddings, dim, hot, buckets, name="adapt"):
        super().__init__(num_embeddings, dim)
        self.hot, self.buckets = hot, buckets
        self.hot_table = _table(hot, dim, f"{name}_hot")
        self.cold_table = _table(buckets, dim, f"{name}_cold")

    def forward(self, ids):
        f, shape = self._flat(ids)
        is_hot = ops.reshape(ops.less(f, self.hot), [f.shape[0], 1])
        h = ops.embedding_lookup(self.hot_table, ops.clamp(f, 0, self.hot - 1))
        c = ops.embedding_lookup(self.cold_table, ops.hash_ids(f, self.buckets))
        return self._shape_out(h * is_hot + c * (1.0 - is_hot), shape)


class CafeEmbedding(AdaptiveEmbedding):
    """CAFE: hot ids are discovered online with a HotSketch (gradient-norm importance); `observe` updates the sketch and
    `remap` turns raw ids into (hot slot | cold) ids fed to the AdaptiveEmbedding forward"""

    def __init__(self, num_embeddings, dim, hot, buckets, decay=0.98, name="cafe"):
        super().__init__(num_embeddings, dim, hot, buckets, name)
        self.score, self.slot_of, self.decay = {}, {}, decay

    def observe(self, ids: np.ndarray, importance: np.ndarray):
        for k in self.score:
            self.score[k] *= self.decay
        for i, v in zip(ids.reshape(-1).tolist(), importance.reshape(-1).tolist()):
            self.score[i] = self.score.get(i, 0.0) + float(v)
        top = sorted(self.score, key=self.score.get, reverse=True)[: self.hot]
        keep = {i: s for i, s in self.slot_of.items() if i in top}
        free = [s for s in range(self.hot) if s not in set(keep.values())]
        for i in top:
            if i not in keep and free:
                keep[i] = free.pop()
        self.slot_of = keep

    def remap(self, ids: np.ndarray) -> np.ndarray:
        flat = ids.reshape(-1)
        out = np.array([self.slot_of.get(int(i), self.hot + int(i)) for i in flat], dtype=np.int64)
        return out.reshape(ids.shape)


class AutoSrhEmbedding(_Base):
    """AutoSrh: rows are grouped (by frequency) into `nsplit` groups, every group has a learnable per-dimension gate alpha; after the
    search the small gates are pruned and the table retrained with the mask frozen (`retrain()`), which is what makes it sparse
    (ref: methods/layers/autosrh.py AutoSrhEmbedding / AutoSrhRetrainEmbedding)"""

    def __init__(self, num_embeddings, dim, nsplit=4, group_indices=None, name="autosrh", frozen_alpha=None, keep_rate=None):
        super().__init__(num_embeddings, dim)
        self.nsplit = int(nsplit)
        self.weight = _table(num_embeddings, dim, f"{name}_weight")
        groups = np.asarray(group_indices if group_indices is not None else np.arange(num_embeddings) * nsplit // max(num_embeddings, 1), np.int64)
        assert groups.shape == (num_embeddings,) and groups.max() < nsplit
        self.group_np = groups
        self.groups = from_numpy(groups.reshape(-1, 1).astype(np.float32))
        self.alpha = parallel_parameter(ones_initializer(), [self.nsplit, dim], None, requires_grad=True, name=f"{name}_alpha")
        self.frozen_mask = None
        if frozen_alpha is not None:
            self.retrain(frozen_alpha, float(keep_rate if keep_rate is not None else 0.5))

    def forward(self, ids):
        f, shape = self._flat(ids)
        w = ops.embedding_lookup(self.weight, f)
        gid = ops.cast(ops.reshape(ops.embedding_lookup(self.groups, f), [-1]), "int64")
        alpha = self.alpha if self.frozen_mask is None else self.frozen_mask
        return self._shape_out(w * ops.embedding_lookup(alpha, gid), shape)

    def l1_penalty(self):
        """the sparsity pressure of the search phase: sum |alpha|"""
        return ops.sum(ops.abs(self.alpha))

    def retrain(self, alpha_values: np.ndarray, keep_rate: float):
        """end of the search: keep the `keep_rate` largest gates (by magnitude), freeze them as a 0/1 mask"""
        a = np.abs(np.asarray(alpha_values, np.float32))
        k = max(int(round(a.size * keep_rate)), 1)
        thr = np.sort(a.reshape(-1))[-k]
        self.mask_np = (a >= thr).astype(np.float32)
        self.frozen_mask = from_numpy(self.mask_np)
        return float(self.mask_np.mean())

    def num_parameters(self) -> int:
        if self.frozen_mask is None:
            return super().num_parameters()
        mask = self.mask_np
        return int(sum(int((self.group_np == g).sum()) * int(mask[g].sum()) for g in range(self.nsplit)))


class DedupEmbedding(_Base):
    """block-level de-duplication of a trained table: blocks of `nemb_per_block` consecutive rows that are (near-)identical are stored
    once; `remap[block]` names the stored block (ref: methods/layers/deduplication.py; the LSH grouping of methods/scheduler)"""

    def __init__(self, num_embeddings, dim, table: np.ndarray = None, nemb_per_block=4, tolerance=0.0, trainable=True, name="dedup"):
        super().__init__(num_embeddings, dim)
        assert table is not None and tuple(table.shape) == (num_embeddings, dim), "DedupEmbedding compresses an existing table"
        self.block = int(nemb_per_block)
        nblocks = -(-num_embeddings // self.block)
        pad = np.zeros((nblocks * self.block, dim), np.float32)
        pad[:num_embeddings] = table
        blocks = pad.reshape(nblocks, self.block * dim)
        remap, stored = self.group_blocks(blocks, tolerance)
        self.remap_np = remap
        self.remap = from_numpy(remap.reshape(-1, 1).astype(np.float32))
        data = np.stack(stored).reshape(len(stored) * self.block, dim)
        from ...core import provided_initializer
        self.weight = parallel_parameter(provided_initializer(data), list(data.shape), None, requires_grad=bool(trainable), name=f"{name}_weight")

    @staticmethod
    def group_blocks(blocks: np.ndarray, tolerance: float):
        """-> (remap [nblocks], stored blocks): blocks whose entries fall in the same cells of a grid of width 2 * tolerance share one
        stored copy (the first seen); tolerance 0 merges exact duplicates only"""
        keys = blocks if tolerance <= 0 else np.round(blocks / (2.0 * tolerance))
        _, first, inverse = np.unique(np.ascontiguousarray(keys), axis=0, return_index=True, return_inverse=True)
        order = np.argsort(first)                                     # stored blocks keep their order of first appearance
        rank = np.empty_like(order)
        rank[order] = np.arange(order.size)
        return rank[inverse.reshape(-1)].astype(np.int64), [blocks[i] for i in first[order]]

    def forward(self, ids):
        f, shape = self._flat(ids)
        fl = ops.cast(f, "float32")
        blk = ops.floor(fl / float(self.block))
        within = fl - blk * float(self.block)
        stored_blk = ops.reshape(ops.embedding_lookup(self.remap, ops.cast(blk, "int64")), [-1])
        real = ops.cast(stored_blk * float(self.block) + within, "int64")
        return self._shape_out(ops.embedding_lookup(self.weight, real), shape)


class QuantizedEmbedding(_Base):
    """post-training / quantisation-aware uniform quantisation to `digit` bits: one (scale, zero point) for the table, or a pair per
    row (`use_qparam`); the forward fake-quantises with a straight-through gradient so the table keeps training
    (ref: methods/layers/quantize.py + src/ops/QuantizeEmbedding.cu)"""

    def __init__(self, num_embeddings, dim, digit=8, scale=0.01, middle=0.0, use_qparam=False, name="quant"):
        super().__init__(num_embeddings, dim)
        assert digit in (8, 16)
        self.digit, self.scale, self.middle, self.use_qparam = int(digit), float(scale), float(middle), bool(use_qparam)
        self.weight = _table(num_embeddings, dim, f"{name}_weight")

    def forward(self, ids):
        f, shape = self._flat(ids)
        w = ops.embedding_lookup(self.weight, f)
        levels = float(2 ** self.digit - 1)
        if self.use_qparam:
            lo, hi = ops.min(w, [-1], True), ops.max(w, [-1], True)
            step = (hi - lo) / levels
            step = step + ops.equal(step, 0.0)
            q = ops.floor((w - lo) / step + 0.5) * step + lo
        else:
            half = float(2 ** (self.digit - 1))
            q = ops.clamp(ops.floor((w - self.middle) / self.scale + 0.5), -half, half - 1.0) * self.scale + self.middle
        return self._shape_out(w + ops.stop_gradient(q - w), shape)

    def num_parameters(self) -> int:
        """in fp32 words: `digit` bits per entry (+ two floats per row with per-row parameters)"""
        return int(math.ceil(self.num_embeddings * self.dim * self.digit / 32.0)) + (2 * self.num_embeddings if self.use_qparam else 2)


class SparseEmbedding(_Base):
    """inference form of a pruned table: only the non-zero entries are kept (CSR); lookups gather rows from the sparse matrix
    (ref: methods/layers/sparse.py -- DeepLight / PEP / AutoSrh export to it)"""

    def __init__(self, num_embeddings, dim, table: np.ndarray = None, form="csr", name="sparse"):
        super().__init__(num_embeddings, dim)
        assert table is not None and tuple(table.shape) == (num_embeddings, dim), "SparseEmbedding is built from a pruned table"
        self.form = form
        t = np.asarray(table, np.float32)
        self.nnz = int((t != 0).sum())
        self.indptr = np.concatenate([[0], np.cumsum((t != 0).sum(1))]).astype(np.int64)
        self.indices = np.nonzero(t)[1].astype(np.int64)
        self.values = t[t != 0].astype(np.float32)
        self._dense = from_numpy(t)                                  # rows are rebuilt from (indptr, indices, values) in `rows()`

    def rows(self, ids: np.ndarray) -> np.ndarray:
        """host-side CSR gather (what the serving path does)"""
        out = np.zeros((ids.size, self.dim), np.float32)
        for k, i in enumerate(np.asarray(ids).reshape(-1)):
            lo, hi = self.indptr[i], self.indptr[i + 1]
            out[k, self.indices[lo:hi]] = self.values[lo:hi]
        return out.reshape(tuple(np.asarray(ids).shape) + (self.dim,))

    def forward(self, ids):
        f, shape = self._flat(ids)
        return self._shape_out(ops.embedding_lookup(self._dense, f), shape)

    def num_parameters(self) -> int:
        return int(2 * self.nnz + self.num_embeddings + 1)           # values + column indices + row pointers


METHODS = {"hash": HashEmbedding, "compo": CompositionalEmbedding, "qr": CompositionalEmbedding, "tt": TensorTrainEmbedding,
           "dhe": DeepHashEmbedding, "robe": RobeEmbedding, "dpq": ProductQuantizedEmbedding, "mgqe": MGQEmbedding, "mde": MixedDimEmbedding,
           "autodim": AutoDimEmbedding, "pep": PrunedEmbedding, "deeplight": DeepLightEmbedding, "optembed": OptEmbedEmbedding,
           "alpt": ALPTEmbedding, "adapt": AdaptiveEmbedding, "cafe": CafeEmbedding, "autosrh": AutoSrhEmbedding, "dedup": DedupEmbedding,
           "quantize": QuantizedEmbedding, "sparse": SparseEmbedding}


def build_compressed_embedding(method: str, num_embeddings: int, dim: int, **kw):
    return METHODS[method.lower()](num_embeddings, dim, **kw)
