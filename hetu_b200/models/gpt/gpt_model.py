"""GPT-2 family (learned positions, LayerNorm, GELU MLP, tied LM head, vocab-parallel cross entropy) over the
parallel modules; activations are kept as [tokens, hidden] so that sequence parallelism is a plain dim-0 split.
(ref: python/hetu/models/gpt/gpt_model.py:22-394, gpt_config.py)
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional

from ... import ops
from ...nn import (HtMultiColumnParallelLinear, HtMultiParallelEmbedding, HtMultiParallelLayerNorm,
                   HtMultiQKVColumnParallelLinear, HtMultiRowParallelLinear, HtMultiVocabParallelEmbedding, Module, ModuleList)
from ...nn.parallel import get_multi_ds_parallel_config
from ...ops_extra import attn_packed
from ..parallel_config import generate_ds_parallel_config
from ..utils.pretrained import PreTrainedConfig, PreTrainedModel


@dataclass
class GPTConfig(PreTrainedConfig):
    vocab_size: int = 50304
    n_positions: int = 1024
    n_embd: int = 768
    n_layer: int = 12
    n_head: int = 12
    n_inner: Optional[int] = None
    activation_function: str = "gelu"
    resid_pdrop: float = 0.0
    embd_pdrop: float = 0.0
    attn_pdrop: float = 0.0
    layer_norm_epsilon: float = 1e-5
    initializer_range: float = 0.02
    use_flash_attn: bool = True
    sequence_parallel: bool = False
    dtype: str = "float32"
    tie_word_embeddings: bool = True

    @property
    def hidden_size(self):
        return self.n_embd

    @property
    def ffn_hidden_size(self):
        return self.n_inner or 4 * self.n_embd

    @staticmethod
    def gpt2_1p3b(**kw):
        """the 1.3B configuration used by the flagship benchmark: 24 layers x 2048 hidden x 16 heads"""
        return GPTConfig(n_embd=2048, n_layer=24, n_head=16, n_positions=1024, vocab_size=50304, **kw)

    @staticmethod
    def gpt2_small(**kw):
        return GPTConfig(n_embd=768, n_layer=12, n_head=12, **kw)

    def num_parameters(self):
        h, f = self.n_embd, self.ffn_hidden_size
        per_layer = 4 * h * h + 2 * h * f + 4 * h + f + 2 * h + 2 * h + 3 * h
        return self.n_layer * per_layer + self.vocab_size * h + self.n_positions * h + 2 * h


class GPTAttention(Module):
    def __init__(self, config: GPTConfig, ds_parallel_configs, layer_idx, name="attn"):
        super().__init__()
        self.config = config
        h = config.n_embd
        self.num_heads, self.head_dim = config.n_head, h // config.n_head
        std = config.initializer_range
        self.qkv_dense = HtMultiQKVColumnParallelLinear(
            h, self.head_dim, self.num_heads, self.num_heads, get_multi_ds_parallel_config(ds_parallel_configs, "qkv", layer_idx),
            bias=True, dtype=config.dtype, name=f"{name}_qkv", init_std=std)
        self.dense = HtMultiRowParallelLinear(
            h, h, get_multi_ds_parallel_config(ds_parallel_configs, "dense", layer_idx), sequence_parallel=config.sequence_parallel,
            bias=True, dtype=config.dtype, name=f"{name}_dense", init_std=std / math.sqrt(2.0 * config.n_layer))

    def forward(self, x, seq_len, residual=None, cu_seqlens=None):
        tp = self.qkv_dense.tp[0]
        qkv = self.qkv_dense(x)                                     # [T, 3h/tp]
        a = attn_packed(qkv, seq_len, self.num_heads // tp, self.num_heads // tp, self.head_dim, is_causal=True, layout="hqkv",
                        cu_seqlens=cu_seqlens)
        return self.dense(a, residual=residual)


class GPTMLP(Module):
    def __init__(self, config: GPTConfig, ds_parallel_configs, layer_idx, name="mlp"):
        super().__init__()
        h, f = config.n_embd, config.ffn_hidden_size
        std = config.initializer_range
        self.act = config.activation_function
        self.dense_h_to_4h = HtMultiColumnParallelLinear(
            h, f, get_multi_ds_parallel_config(ds_parallel_configs, "dense_h_to_4h", layer_idx), bias=True, gather_output=False,
            dtype=config.dtype, name=f"{name}_h_to_4h", init_std=std)
        self.dense_4h_to_h = HtMultiRowParallelLinear(
            f, h, get_multi_ds_parallel_config(ds_parallel_configs, "dense_4h_to_h", layer_idx),
            sequence_parallel=config.sequence_parallel, bias=True, dtype=config.dtype, name=f"{name}_4h_to_h",
            init_std=std / math.sqrt(2.0 * config.n_layer))

    def forward(self, x, residual=None):
        hmid = self.dense_h_to_4h(x, act=self.act)      # bias + activation fused into the GEMM epilogue
        return self.dense_4h_to_h(hmid, residual=residual)


class GPTBlock(Module):
    def __init__(self, config: GPTConfig, ds_parallel_configs, layer_idx):
        super().__init__()
        self.layer_idx = layer_idx
        sp = config.sequence_parallel
        self.ln_1 = HtMultiParallelLayerNorm(config.n_embd, get_multi_ds_parallel_config(ds_parallel_configs, "layernorm1", layer_idx),
                                             sequence_parallel=sp, eps=config.layer_norm_epsilon, dtype=config.dtype,
                                             name=f"ln1_block{layer_idx}")
        self.attn = GPTAttention(config, ds_parallel_configs, layer_idx, name=f"attn_block{layer_idx}")
        self.ln_2 = HtMultiParallelLayerNorm(config.n_embd, get_multi_ds_parallel_config(ds_parallel_configs, "layernorm2", layer_idx),
                                             sequence_parallel=sp, eps=config.layer_norm_epsilon, dtype=config.dtype,
                                             name=f"ln2_block{layer_idx}")
        self.mlp = GPTMLP(config, ds_parallel_configs, layer_idx, name=f"mlp_block{layer_idx}")

    def forward(self, x, seq_len, cu_seqlens=None):
        # entering a new pipeline stage: receive the residual stream once (P2P), both branches then use the local copy; a
        # block with another (tp, dp) than its predecessor (Galvatron layer-wise strategies) relocates the activations here
        x = self.ln_1._adapt(x, self.ln_1._all_split0() if self.ln_1.sequence_parallel else self.attn.qkv_dense.ds_split0_dup())
        x = self.attn(self.ln_1(x), seq_len, residual=x, cu_seqlens=cu_seqlens)    # residual add fused into the row-parallel GEMM epilogue
        x = self.mlp(self.ln_2(x), residual=x)
        return x


class GPTModel(Module):
    def __init__(self, config: GPTConfig, ds_parallel_configs):
        super().__init__()
        self.config = config
        self.dtype = config.dtype
        std = config.initializer_range
        self.wte = HtMultiVocabParallelEmbedding(config.vocab_size, config.n_embd,
                                                 get_multi_ds_parallel_config(ds_parallel_configs, "wte"), dtype=config.dtype,
                                                 name="wte", init_std=std)
        self.wpe = HtMultiParallelEmbedding(config.n_positions, config.n_embd, get_multi_ds_parallel_config(ds_parallel_configs, "wpe"),
                                            dtype=config.dtype, name="wpe", init_std=std)
        self.h = ModuleList([GPTBlock(config, ds_parallel_configs, i) for i in range(config.n_layer)])
        self.ln_f = HtMultiParallelLayerNorm(config.n_embd, get_multi_ds_parallel_config(ds_parallel_configs, "layernorm_final"),
                                             sequence_parallel=config.sequence_parallel, eps=config.layer_norm_epsilon,
                                             dtype=config.dtype, name="ln_final")

    def forward(self, input_ids, position_ids, seq_len, cu_seqlens=None):
        """input_ids / position_ids: flattened [tokens]; cu_seqlens: document boundaries of a packed batch"""
        pe = self.wpe(position_ids)
        if self.config.sequence_parallel and any(t > 1 for t in self.wte.tp):
            pe = self.wte._adapt(pe, self.wte.ds_split0())      # local slice: keep this rank's token shard
        x = self.wte(input_ids, sequence_parallel=self.config.sequence_parallel) + pe
        if self.config.embd_pdrop > 0:
            x = ops.dropout(x, self.config.embd_pdrop)
        for blk in self.h:
            with _placement(blk):
                x = blk(x, seq_len, cu_seqlens)
        return self.ln_f(x)


class _placement:
    """ops of a block are placed on the block's device groups (pipeline stages)"""

    def __init__(self, blk):
        self.dgs = blk.ln_1.device_group_unions

    def __enter__(self):
        from ...core import cur_graph
        self.g = cur_graph()
        self.g.push_ctx(device_group_hierarchy=[list(u) for u in self.dgs])

    def __exit__(self, *a):
        self.g.pop_ctx()
        return False


class GPTLMHeadModel(Module, PreTrainedModel):
    config_class = GPTConfig

    def __init__(self, config: GPTConfig, ds_parallel_configs: Optional[List[dict]] = None, num_gpus: int = 1):
        super().__init__()
        if ds_parallel_configs is None:
            ds_parallel_configs = [generate_ds_parallel_config(config.n_layer, num_gpus, num_gpus, 1, 1)]
        self.config = config
        self.ds_parallel_configs = ds_parallel_configs
        self.transformer = GPTModel(config, ds_parallel_configs)
        if config.tie_word_embeddings:
            self.lm_head = None
        else:
            self.lm_head = HtMultiColumnParallelLinear(config.n_embd, config.vocab_size,
                                                       get_multi_ds_parallel_config(ds_parallel_configs, "lm_head"), bias=False,
                                                       gather_output=False, dtype=config.dtype, name="lm_head",
                                                       init_std=config.initializer_range)

    def forward(self, input_ids, position_ids=None, labels=None, seq_len=None, cu_seqlens=None):
        hidden = self.transformer(input_ids, position_ids, seq_len, cu_seqlens)
        wte = self.transformer.wte
        if self.lm_head is None:
            last = self.transformer.ln_f.device_group_unions
            # layouts of the head follow the LAST stage (its tensor-parallel degree may differ from the first stage's, e.g. a
            # re-planned pipeline whose stage lost a device); on a single stage that is the embedding's own layout
            head = wte if wte.device_group_unions == last else self.transformer.h[-1].attn.qkv_dense
            if not hidden.check_ds_hierarchy_equal(head.ds_split0_dup()):      # sequence-parallel hidden -> all-gather
                hidden = ops.comm(hidden, head.ds_split0_dup(), device_group_hierarchy=last)
            table = wte.embedding_table
            if wte.device_group_unions != last:
                # tied head under pipeline parallelism: the table travels first stage -> last stage (and its gradient
                # back) through a comm op ("share_weight_comm" in the reference), re-sharded when the tp degrees differ
                table = ops.comm(table, head.ds_dup_split0(), device_group_hierarchy=last, name="share_weight_comm")
            logits = ops.linear(hidden, table, None, trans_b=True, device_group_hierarchy=last, name="lm_head")
            head_tp = head.tp
        else:
            logits = self.lm_head(hidden)
        if labels is None:
            return logits
        if any(t > 1 for t in (head_tp if self.lm_head is None else wte.tp)):     # any strategy with a vocab split needs the vocab-parallel loss (it degenerates for tp = 1)
            loss = ops.vocab_parallel_cross_entropy(logits, labels, ignored_index=-1, reduction="mean")
        else:
            loss = ops.softmax_cross_entropy_sparse(logits, labels, ignored_index=-1, reduction="mean")
        return loss
