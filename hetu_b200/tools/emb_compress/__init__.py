"""Embedding memory compression methods (ref: tools/EmbeddingMemoryCompression/methods/layers/*.py -- hash, compositional
(QR), tensor-train, deep hash (DHE), ROBE, MGQE / DPQ product quantisation, mixed-dimension (MDE), AutoDim-style
dimension search, pruning (PEP / DeepLight / OptEmbed masks), ALPT low-precision training, AdaptEmb / CAFE frequency-aware
hot-cold tables).  Every module maps integer ids [...] -> vectors [..., dim] and can replace `nn.Embedding` in the CTR models;
`compression_ratio()` reports parameters relative to a full table."""
from .methods import (HashEmbedding, CompositionalEmbedding, TensorTrainEmbedding, DeepHashEmbedding, RobeEmbedding,  # noqa: F401
                      ProductQuantizedEmbedding, MGQEmbedding, MixedDimEmbedding, AutoDimEmbedding, PrunedEmbedding, DeepLightEmbedding,
                      OptEmbedEmbedding, ALPTEmbedding, AdaptiveEmbedding, CafeEmbedding, METHODS, build_compressed_embedding)
