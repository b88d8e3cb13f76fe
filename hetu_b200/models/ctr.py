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


class DeepCrossing(_CTRBase):
    """Deep Crossing (DC): embeddings + dense features through a stack of residual units h = relu(h + W2 relu(W1 h))
    (ref: hetu/v1/examples/ctr/models/dc_criteo.py)"""

    def __init__(self, num_embeddings, embedding_dim=16, num_fields=26, num_dense=13, num_units=3, unit_hidden=256, **kw):
        super().__init__(num_embeddings, embedding_dim, num_fields, num_dense, **kw)
        d = num_fields * embedding_dim + num_dense
        self.w1 = ModuleList([Linear(d, unit_hidden, name=f"dc_u{i}_w1") for i in range(num_units)])
        self.w2 = ModuleList([Linear(unit_hidden, d, name=f"dc_u{i}_w2") for i in range(num_units)])
        self.out = Linear(d, 1, name="dc_out")

    def forward(self, dense, sparse_ids, label=None, embedded=None):
        b = dense.shape[0]
        h = ops.concat([ops.reshape(self.embed(sparse_ids, embedded), [b, self.num_fields * self.dim]), dense], 1)
        for a, c in zip(self.w1, self.w2):
            h = ops.relu(h + c(a(h, act="relu")))
        logit = self.out(h)
        return logit if label is None else (self.loss(logit, label), logit)


class NCF(Module):
    """Neural collaborative filtering (NeuMF): a GMF branch (element-wise product of user / item factors) and an MLP branch over
    concatenated embeddings, fused by one linear layer; implicit-feedback BCE loss
    (ref: hetu/v1/examples/rec -- hetu_ncf.py)"""

    def __init__(self, num_users, num_items, factors=8, mlp_layers: Sequence[int] = (64, 32, 16, 8)):
        super().__init__()
        self.user_gmf, self.item_gmf = Embedding(num_users, factors, name="ncf_user_gmf"), Embedding(num_items, factors, name="ncf_item_gmf")
        half = mlp_layers[0] // 2
        self.user_mlp, self.item_mlp = Embedding(num_users, half, name="ncf_user_mlp"), Embedding(num_items, half, name="ncf_item_mlp")
        self.mlp = ModuleList([Linear(a, b, name=f"ncf_mlp{i}") for i, (a, b) in enumerate(zip(mlp_layers[:-1], mlp_layers[1:]))])
        self.out = Linear(factors + mlp_layers[-1], 1, name="ncf_out")

    def forward(self, users, items, label=None):
        gmf = self.user_gmf(users) * self.item_gmf(items)
        h = ops.concat([self.user_mlp(users), self.item_mlp(items)], 1)
        for l in self.mlp:
            h = l(h, act="relu")
        logit = self.out(ops.concat([gmf, h], 1))
        if label is None:
            return logit
        return ops.binary_cross_entropy(ops.sigmoid(logit), label, reduction="mean"), logit
