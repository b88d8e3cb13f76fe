"""Click-through-rate models of the v1 examples: Wide&Deep, DeepFM, Deep&Cross (ref: hetu/v1/examples/ctr/models/
{wdl_criteo,deepfm_criteo,dcn_criteo}.py).  Sparse features share one embedding table addressed by offset ids
[batch, fields]; the table can be a plain parameter, a compressed embedding (tools.emb_compress) or -- in PS / Hybrid
mode -- a CacheSparseTable fed through `embedded=` (the HET path)."""
from __future__ import annotations

from typing import Optional, Sequence

from .. import ops
from ..nn import Embedding, Linear, Module, ModuleList


class _CTRBase(Module):
    def __init__(self, num_embeddings, embedding_dim, num_fields, num_dense, embedding: Optional[Module] = None, dtype="float32"):
        super().__init__()
        self.num_fields, self.dim, self.num_dense = num_fields, embedding_dim, num_dense
        self.embedding = embedding if embedding is not None else Embedding(num_embeddings, embedding_dim, dtype=dtype, name="ctr_embedding")

    def embed(self, sparse_ids, embedded=None):
        """-> [B, fields, dim]"""
        if embedded is not None:
            return embedded
        b = sparse_ids.shape[0]
        e = self.embedding(ops.reshape(sparse_ids, [b * self.num_fields]))
        return ops.reshape(e, [b, self.num_fields, self.dim])

    @staticmethod
    def loss(logit, label):
        return ops.mean(ops.binary_cross_entropy(ops.sigmoid(logit), label, reduction="none"), [0, 1])


class WDL(_CTRBase):
    """wide (linear over dense features) + deep (MLP over [embeddings, dense])"""

    def __init__(self, num_embeddings, embedding_dim=16, num_fields=26, num_dense=13, hidden: Sequence[int] = (256, 256, 256), **kw):
        super().__init__(num_embeddings, embedding_dim, num_fields, num_dense, **kw)
        self.wide = Linear(num_dense, 1, name="wdl_wide")
        dims = [num_fields * embedding_dim + num_dense] + list(hidden)
        self.deep = ModuleList([Linear(a, b, name=f"wdl_deep{i}") for i, (a, b) in enumerate(zip(dims[:-1], dims[1:]))])
        self.out = Linear(dims[-1], 1, name="wdl_out")

    def forward(self, dense, sparse_ids, label=None, embedded=None):
        b = dense.shape[0]
        e = ops.reshape(self.embed(sparse_ids, embedded), [b, self.num_fields * self.dim])
        h = ops.concat([e, dense], 1)
        for l in self.deep:
            h = l(h, act="relu")
        logit = self.out(h) + self.wide(dense)
        return logit if label is None else (self.loss(logit, label), logit)


class DeepFM(_CTRBase):
    """first-order weights + factorisation-machine second-order term + deep MLP, all over the same embeddings"""

    def __init__(self, num_embeddings, embedding_dim=16, num_fields=26, num_dense=13, hidden: Sequence[int] = (256, 256), **kw):
        super().__init__(num_embeddings, embedding_dim, num_fields, num_dense, **kw)
        self.first = Embedding(num_embeddings, 1, name="deepfm_first")
        self.dense_first = Linear(num_dense, 1, name="deepfm_dense_first")
        dims = [num_fields * embedding_dim + num_dense] + list(hidden)
        self.deep = ModuleList([Linear(a, b, name=f"deepfm_deep{i}") for i, (a, b) in enumerate(zip(dims[:-1], dims[1:]))])
        self.out = Linear(dims[-1], 1, name="deepfm_out")

    def forward(self, dense, sparse_ids, label=None, embedded=None):
        b = dense.shape[0]
        e = self.embed(sparse_ids, embedded)                                    # [B, F, D]
        first = ops.sum(ops.reshape(self.first(ops.reshape(sparse_ids, [b * self.num_fields])), [b, self.num_fields]), [1], True)
        s = ops.sum(e, [1])                                                      # (sum v)^2 - sum v^2
        second = ops.sum(s * s - ops.sum(e * e, [1]), [1], True) * 0.5
        h = ops.concat([ops.reshape(e, [b, self.num_fields * self.dim]), dense], 1)
        for l in self.deep:
            h = l(h, act="relu")
        logit = self.out(h) + first + second + self.dense_first(dense)
        return logit if label is None else (self.loss(logit, label), logit)


class DCN(_CTRBase):
    """cross network x_{l+1} = x_0 (x_l . w_l) + b_l + x_l in parallel with a deep MLP"""

    def __init__(self, num_embeddings, embedding_dim=16, num_fields=26, num_dense=13, num_cross=3, hidden: Sequence[int] = (256, 256), **kw):
        super().__init__(num_embeddings, embedding_dim, num_fields, num_dense, **kw)
        d = num_fields * embedding_dim + num_dense
        self.cross = ModuleList([Linear(d, 1, name=f"dcn_cross{i}") for i in range(num_cross)])
        self.cross_bias = ModuleList([Linear(1, d, bias=False, name=f"dcn_cross_b{i}") for i in range(num_cross)])
        dims = [d] + list(hidden)
        self.deep = ModuleList([Linear(a, b, name=f"dcn_deep{i}") for i, (a, b) in enumerate(zip(dims[:-1], dims[1:]))])
        self.out = Linear(d + dims[-1], 1, name="dcn_out")

    def forward(self, dense, sparse_ids, label=None, embedded=None):
        b = dense.shape[0]
        x0 = ops.concat([ops.reshape(self.embed(sparse_ids, embedded), [b, self.num_fields * self.dim]), dense], 1)
        x = x0
        for w in self.cross:
            x = x0 * w(x) + x           # w(x): [B, 1] (x_l . w_l + b_l), broadcast over the feature dim
        h = x0
        for l in self.deep:
            h = l(h, act="relu")
        logit = self.out(ops.concat([x, h], 1))
        return logit if label is None else (self.loss(logit, label), logit)
