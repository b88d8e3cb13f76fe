"""Llama family: RMSNorm (sequence-parallel aware), rotary positions (CP-slot offsets), GQA, SwiGLU MLP,
selective recompute, optional context-parallel ring attention.
(ref: python/hetu/models/llama/llama_model.py:10-492, llama_config.py)
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional

from ... import ops
from ...core import IntSymbol
from ...core import recompute as recompute_ctx
from ...nn import (HtMultiColumnParallelLinear, HtMultiParallelRMSNorm, HtMultiQKVColumnParallelLinear, HtMultiRowParallelLinear,
                   HtMultiVocabParallelEmbedding, Module, ModuleList)
from ...nn.parallel import get_multi_ds_parallel_config
from ...ops_extra import attn_packed, rotary_packed
from ..parallel_config import generate_ds_parallel_config
from ..utils.pretrained import PreTrainedConfig, PreTrainedModel


@dataclass
class LlamaConfig(PreTrainedConfig):
    # (fp8=True runs the four projection GEMMs of every block in e4m3 with per-row (activations) and per-column (weights) scales, see ops.linear_fp8)
    vocab_size: int = 32000
    hidden_size: int = 4096
    intermediate_size: int = 11008
    num_hidden_layers: int = 32
    num_attention_heads: int = 32
    num_key_value_heads: Optional[int] = None
    max_position_embeddings: int = 4096
    rms_norm_eps: float = 1e-6
    rope_theta: float = 10000.0
    initializer_range: float = 0.02
    sequence_parallel: bool = True
    dtype: str = "float32"
    cp_ranks: tuple = ()            # context-parallel ring (global ranks); empty = no CP
    recompute_layers: tuple = ()
    fp8: bool = False               # projection GEMMs in row- / column-scaled e4m3 (fwd + dgrad), weight gradients in bf16

    @property
    def kv_heads(self):
        return self.num_key_value_heads or self.num_attention_heads

    @staticmethod
    def llama2_7b(**kw):
        return LlamaConfig(hidden_size=4096, intermediate_size=11008, num_hidden_layers=32, num_attention_heads=32, **kw)

    @staticmethod
    def llama3_8b(**kw):
        return LlamaConfig(vocab_size=128256, hidden_size=4096, intermediate_size=14336, num_hidden_layers=32,
                           num_attention_heads=32, num_key_value_heads=8, rope_theta=500000.0, **kw)


class LlamaAttention(Module):
    def __init__(self, config: LlamaConfig, ds_parallel_configs, layer_idx, name="attn"):
        super().__init__()
        self.config = config
        h = config.hidden_size
        self.num_heads, self.kv_heads = config.num_attention_heads, config.kv_heads
        self.head_dim = h // self.num_heads
        std = config.initializer_range
        self.qkv_dense = HtMultiQKVColumnParallelLinear(h, self.head_dim, self.num_heads, self.kv_heads,
                                                        get_multi_ds_parallel_config(ds_parallel_configs, "qkv", layer_idx),
                                                        bias=False, dtype=config.dtype, name=f"{name}_qkv", init_std=std)
        self.dense = HtMultiRowParallelLinear(h, h, get_multi_ds_parallel_config(ds_parallel_configs, "dense", layer_idx),
                                              sequence_parallel=config.sequence_parallel, bias=False, dtype=config.dtype,
                                              name=f"{name}_dense", init_std=std / math.sqrt(2.0 * config.num_hidden_layers))
        self.qkv_dense.fp8 = self.dense.fp8 = config.fp8

    def forward(self, x, seq_len, residual=None, pos_offset=0, position_ids=None, cu_seqlens=None):
        tp = self.qkv_dense.tp[0]
        assert self.kv_heads % tp == 0, "tensor parallel degree must divide the number of kv heads"
        hq, hkv, d = self.num_heads // tp, self.kv_heads // tp, self.head_dim
        rep = hq // hkv
        # kv-head-major packed layout [g: q x rep, k, v]: any tp | kv_heads owns whole groups (strategy independent weights)
        qkv = self.qkv_dense(x)
        qkv = rotary_packed(qkv, seq_len, hq, hkv, d, positions=position_ids, base=self.config.rope_theta, pos_offset=pos_offset, layout="hqkv")
        if self.config.cp_ranks and len(self.config.cp_ranks) > 1:
            s = seq_len            # int or IntSymbol: the per-ring-member sequence length may change every step
            sym = (lambda t, shp: t.set_symbolic_shape([v if isinstance(v, IntSymbol) else IntSymbol(int(v)) for v in shp])) \
                if isinstance(s, IntSymbol) else (lambda t, shp: None)
            # the reshape gradients restore their input's shape: give the intermediates symbolic shapes so that follows `s`
            sym(qkv, [-1, (hq + 2 * hkv) * d])
            g5 = ops.reshape(qkv, [-1, s, hkv, rep + 2, d])
            q, k, v = ops.split(g5, [rep, 1, 1], dim=3)
            for t, n in ((q, rep), (k, 1), (v, 1)):
                sym(t, [-1, s, hkv, n, d])
            q = ops.reshape(q, [-1, s, hq, d])
            k = ops.reshape(k, [-1, s, hkv, d])
            v = ops.reshape(v, [-1, s, hkv, d])
            a = ops.parallel_attn(q, k, v, self.config.cp_ranks, is_causal=True, cu_seqlens=cu_seqlens)   # packed rows: per-document masks
            sym(a, [-1, s, hq, d])
            a = ops.reshape(a, [-1, hq * d])
        else:
            a = attn_packed(qkv, seq_len, hq, hkv, d, is_causal=True, layout="hqkv", cu_seqlens=cu_seqlens)
        return self.dense(a, residual=residual)


class LlamaMLP(Module):
    def __init__(self, config: LlamaConfig, ds_parallel_configs, layer_idx, name="mlp"):
        super().__init__()
        h, f = config.hidden_size, config.intermediate_size
        std = config.initializer_range
        # gate and up projections are fused in one column-parallel GEMM with interleaved rows (gate_0, up_0, gate_1, ...):
        # [T, 2f/tp] -> swiglu -> [T, f/tp], the same global weight under every tensor-parallel degree
        self.dense_h_to_4h = HtMultiColumnParallelLinear(h, 2 * f, get_multi_ds_parallel_config(ds_parallel_configs, "dense_h_to_4h", layer_idx),
                                                         bias=False, gather_output=False, dtype=config.dtype, name=f"{name}_gate_up",
                                                         init_std=std)
        self.dense_4h_to_h = HtMultiRowParallelLinear(f, h, get_multi_ds_parallel_config(ds_parallel_configs, "dense_4h_to_h", layer_idx),
                                                      sequence_parallel=config.sequence_parallel, bias=False, dtype=config.dtype,
                                                      name=f"{name}_down", init_std=std / math.sqrt(2.0 * config.num_hidden_layers))
        self.dense_h_to_4h.fp8 = self.dense_4h_to_h.fp8 = config.fp8

    def forward(self, x, residual=None):
        return self.dense_4h_to_h(ops.swiglu(self.dense_h_to_4h(x), interleaved=True), residual=residual)


class LlamaBlock(Module):
    def __init__(self, config: LlamaConfig, ds_parallel_configs, layer_idx):
        super().__init__()
        self.layer_idx = layer_idx
        sp = config.sequence_parallel
        self.rmsnorm_1 = HtMultiParallelRMSNorm(config.hidden_size, get_multi_ds_parallel_config(ds_parallel_configs, "layernorm1", layer_idx),
                                                sequence_parallel=sp, eps=config.rms_norm_eps, dtype=config.dtype,
                                                name=f"rmsnorm1_block{layer_idx}")
        self.attn = LlamaAttention(config, ds_parallel_configs, layer_idx, name=f"attn_block{layer_idx}")
        self.rmsnorm_2 = HtMultiParallelRMSNorm(config.hidden_size, get_multi_ds_parallel_config(ds_parallel_configs, "layernorm2", layer_idx),
                                                sequence_parallel=sp, eps=config.rms_norm_eps, dtype=config.dtype,
                                                name=f"rmsnorm2_block{layer_idx}")
        self.mlp = LlamaMLP(config, ds_parallel_configs, layer_idx, name=f"mlp_block{layer_idx}")

    def forward(self, x, seq_len, pos_offset=0, position_ids=None, cu_seqlens=None):
        n1 = self.rmsnorm_1
        # pipeline-stage entry (P2P) / relocation when this block's (tp, dp) differs from its predecessor's
        x = n1._adapt(x, n1._all_split0() if n1.sequence_parallel else self.attn.qkv_dense.ds_split0_dup())
        x = self.attn(self.rmsnorm_1(x), seq_len, residual=x, pos_offset=pos_offset, position_ids=position_ids, cu_seqlens=cu_seqlens)
        return self.mlp(self.rmsnorm_2(x), residual=x)


class LlamaModel(Module):
    def __init__(self, config: LlamaConfig, ds_parallel_configs):
        super().__init__()
        self.config = config
        self.wte = HtMultiVocabParallelEmbedding(config.vocab_size, config.hidden_size,
                                                 get_multi_ds_parallel_config(ds_parallel_configs, "wte"), dtype=config.dtype,
                                                 name="wte", init_std=config.initializer_range)
        self.h = ModuleList([LlamaBlock(config, ds_parallel_configs, i) for i in range(config.num_hidden_layers)])
        self.rmsnorm_f = HtMultiParallelRMSNorm(config.hidden_size, get_multi_ds_parallel_config(ds_parallel_configs, "layernorm_final"),
                                                sequence_parallel=config.sequence_parallel, eps=config.rms_norm_eps,
                                                dtype=config.dtype, name="rmsnorm_final")

    def forward(self, input_ids, seq_len, pos_offset=0, position_ids=None, cu_seqlens=None):
        from ..gpt.gpt_model import _placement
        x = self.wte(input_ids, sequence_parallel=self.config.sequence_parallel)
        for i, blk in enumerate(self.h):
            blk.ln_1 = blk.rmsnorm_1  # placement helper reads .ln_1
            with _placement(blk):
                if i in self.config.recompute_layers:
                    with recompute_ctx([True]):
                        x = blk(x, seq_len, pos_offset, position_ids, cu_seqlens)
                else:
                    x = blk(x, seq_len, pos_offset, position_ids, cu_seqlens)
        return self.rmsnorm_f(x)


class LlamaLMHeadModel(Module, PreTrainedModel):
    config_class = LlamaConfig

    def __init__(self, config: LlamaConfig, ds_parallel_configs: Optional[List[dict]] = None, num_gpus: int = 1):
        super().__init__()
        if ds_parallel_configs is None:
            ds_parallel_configs = [generate_ds_parallel_config(config.num_hidden_layers, num_gpus, num_gpus, 1, 1, model="llama")]
        self.config = config
        self.ds_parallel_configs = ds_parallel_configs
        self.transformer = LlamaModel(config, ds_parallel_configs)
        self.lm_head = HtMultiColumnParallelLinear(config.hidden_size, config.vocab_size,
                                                   get_multi_ds_parallel_config(ds_parallel_configs, "lm_head"), bias=False,
                                                   gather_output=False, dtype=config.dtype, name="lm_head",
                                                   init_std=config.initializer_range)

    def forward(self, input_ids, position_ids=None, labels=None, seq_len=None, pos_offset=0, cu_seqlens=None):
        """position_ids (optional): explicit rotary positions, needed when a batch row packs several documents;
        cu_seqlens: their boundaries (variable-length attention)"""
        hidden = self.transformer(input_ids, seq_len, pos_offset, position_ids if cu_seqlens is not None else None, cu_seqlens)
        logits = self.lm_head(hidden)
        if labels is None:
            return logits
        if any(t > 1 for t in self.lm_head.tp):
            return ops.vocab_parallel_cross_entropy(logits, labels, ignored_index=-1, reduction="mean")
        return ops.softmax_cross_entropy_sparse(logits, labels, ignored_index=-1, reduction="mean")
