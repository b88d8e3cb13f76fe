"""HetuMoE: gates (Top-k GShard, k-Top-1, Hash, BASE balance assignment, SAM) -> capacity-based dispatch into
expert-major buffers -> all-to-all over the expert-parallel group -> local experts -> all-to-all -> gated combine.
(ref: hetu/v1/python/hetu/layers/{moe_layer,TopGate,KTop1Gate,HashGate,BalanceGate,SAMGate}.py,
 hetu/v1/src/ops/LayoutTransform.cu, hetu/v1/examples/moe)
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Sequence


from ... import ops
from ...core import normal_initializer, parallel_parameter, zeros_initializer
from ...nn import Module, ModuleList
from ..gpt.gpt_model import GPTConfig


@dataclass
class MoEConfig(GPTConfig):
    num_experts: int = 8
    top_k: int = 1
    capacity_factor: float = 1.0
    gate_type: str = "topk"          # topk | ktop1 | hash | balance | sam
    aux_loss_weight: float = 0.01
    moe_every: int = 2               # every n-th block uses an MoE MLP
    ep_ranks: tuple = ()             # expert-parallel group (global ranks); empty = all experts local

    @staticmethod
    def gpt_moe_350m_8e(**kw):
        """GPT-MoE 350M x 8 experts (BASELINE config #4): 24 layers x 1024 hidden x 16 heads, MoE every other layer"""
        return MoEConfig(n_embd=1024, n_layer=24, n_head=16, num_experts=8, top_k=1, **kw)


def balance_loss(logits, idx, num_experts):
    """GShard / Switch load-balancing loss  E * sum_e mean_t(softmax(logits))[t, e] * ce[e]  with ce[e] the fraction of
    tokens whose first choice is expert e.  `ce` comes from the (integer) routing decision and is a constant for autograd;
    the mean gate probability is built from differentiable ops, so the router weights receive the balancing gradient
    (ref: hetu/v1/python/hetu/layers/TopGate.py:28-45 builds it from the softmax gates the same way).  The aux output of the
    `moe_gate` op itself is computed inside a non-differentiable op and only serves as a monitoring value."""
    tokens = logits.shape[0]
    probs = ops.softmax(ops.data_transfer(logits, "float32"), -1)                      # [T, E], differentiable
    me = ops.reduce(probs, "mean", [0])                                                 # [E]
    first = ops.reshape(ops.slice(idx, [0, 0], [tokens, 1]), [tokens])
    ce = ops.reduce(ops.onehot(ops.data_transfer(first, "int64"), num_experts), "mean", [0])   # [E], constant
    return ops.reduce(me * ops.data_transfer(ce, "float32"), "sum", [0]) * float(num_experts)


class TopKGate(Module):
    """softmax -> top-k -> capacity (GShard ordering) ; returns (gates, idx, loc, aux_loss)"""

    def __init__(self, d_model, num_experts, k=1, capacity_factor=1.0, dtype="float32", name="gate"):
        super().__init__()
        self.num_experts, self.k, self.capacity_factor = num_experts, k, capacity_factor
        self.wg = parallel_parameter(normal_initializer(0.0, 0.02), [num_experts, d_model], None, dtype=dtype, requires_grad=True,
                                     name=f"{name}_wg")

    def capacity(self, tokens):
        return int(math.ceil(self.k * tokens / self.num_experts * self.capacity_factor))

    def forward(self, x):
        logits = ops.linear(x, self.wg, None, trans_b=True)
        cap = self.capacity(x.shape[0])
        _, idx, loc, _ = ops.moe_gate(logits, self.k, cap)
        gates = ops.moe_gate_values(logits, idx, loc)     # differentiable w.r.t. the router weights
        return gates, idx, loc, balance_loss(logits, idx, self.num_experts), cap


class KTop1Gate(TopKGate):
    """k independent top-1 routers over k disjoint expert groups (each token visits one expert per group)"""

    def forward(self, x):
        logits = ops.linear(x, self.wg, None, trans_b=True)
        groups = ops.split(logits, self.k, dim=1)
        e_per = self.num_experts // self.k
        cap = int(math.ceil(x.shape[0] / e_per * self.capacity_factor))
        gates, idxs, locs, aux_total = [], [], [], None
        for gi, lg in enumerate(groups):
            _, idx, loc, _ = ops.moe_gate(lg, 1, cap)
            aux = balance_loss(lg, idx, e_per)
            gates.append(ops.moe_gate_values(lg, idx, loc))
            idxs.append(idx if gi == 0 else _shift_idx(idx, gi * e_per))
            locs.append(loc)
            aux_total = aux if aux_total is None else aux_total + aux
        return ops.concat(gates, 1), ops.concat(idxs, 1), ops.concat(locs, 1), aux_total, cap


def _shift_idx(idx, off):
    """expert indices of group `gi` are offset by gi * experts_per_group (small integers: exact through float32)"""
    return ops.data_transfer(ops.add(ops.data_transfer(idx, "float32"), float(off)), "int32")


class HashGate(Module):
    """static hash routing: expert = token_id mod E (no learned router, gates = 1)"""

    def __init__(self, d_model, num_experts, capacity_factor=1.0, **kw):
        super().__init__()
        self.num_experts, self.k, self.capacity_factor = num_experts, 1, capacity_factor

    def forward(self, x, token_ids=None):
        tokens = x.shape[0]
        cap = int(math.ceil(tokens / self.num_experts * self.capacity_factor))
        if token_ids is None:
            token_ids = ops.arange(0, tokens, 1, "int64")
        onehot = ops.onehot(ops.data_transfer(_mod(token_ids, self.num_experts), "int64"), self.num_experts)
        logits = ops.data_transfer(onehot * 30.0, x.dtype)
        gates, idx, loc, aux = ops.moe_gate(logits, 1, cap)
        return None, idx, loc, None, cap


def _mod(t, n):
    tf = ops.data_transfer(t, "float32")
    return tf - ops.floor(tf / float(n)) * float(n)


class BalanceGate(TopKGate):
    """BASE layers: balanced assignment -- every expert receives exactly ceil(tokens / E) tokens, no token is dropped and
    no auxiliary loss is needed.  The assignment runs on the device (`moe_balance_assign`, csrc/kernels/moe.cu): rounds in
    which every unplaced token proposes to its best expert with room and over-subscribed experts keep their highest
    scoring proposers.  (ref: hetu/v1/python/hetu/layers/BalanceGate.py, gpu_ops/BalanceAssignment.py -- an auction with a
    host round trip per iteration)"""

    def forward(self, x):
        logits = ops.linear(x, self.wg, None, trans_b=True)
        tokens = x.shape[0]
        cap = int(math.ceil(tokens / self.num_experts))
        _, idx, loc, _ = ops.make_op("moe_balance_assign", [logits], {"capacity": cap})
        gates = ops.moe_gate_values(logits, idx, loc)
        return gates, idx, loc, None, cap


class SAMGate(TopKGate):
    """Switch-and-Mixture: first pick the best expert *group* (sum of probabilities), then top-k inside it, so all k
    experts of a token live on one device (one all-to-all hop instead of k)."""

    def __init__(self, d_model, num_experts, k=2, num_groups=None, capacity_factor=1.0, dtype="float32", name="samgate"):
        super().__init__(d_model, num_experts, k, capacity_factor, dtype, name)
        self.num_groups = num_groups or max(num_experts // max(k, 1), 1)

    def forward(self, x):
        logits = ops.linear(x, self.wg, None, trans_b=True)
        e, g = self.num_experts, self.num_groups
        per = e // g
        probs = ops.softmax(ops.data_transfer(logits, "float32"), -1)
        group_score = ops.reduce(ops.reshape(probs, [x.shape[0], g, per]), "sum", [2])          # sam_group_sum
        best = ops.reduce(group_score, "max", [1], keepdims=True)                                  # sam_max
        mask = ops.data_transfer(ops.reshape(ops.broadcast(ops.reshape(
            ops.data_transfer(_ge(group_score, best), "float32"), [x.shape[0], g, 1]), [x.shape[0], g, per]), [x.shape[0], e]),
            logits.dtype)
        masked = logits * mask + (mask - 1.0) * 1e4
        cap = self.capacity(x.shape[0])
        _, idx, loc, _ = ops.moe_gate(masked, self.k, cap)
        gates = ops.moe_gate_values(masked, idx, loc)
        return gates, idx, loc, balance_loss(logits, idx, self.num_experts), cap


def _ge(a, b):
    return ops.bool_op(ops.relu(a - b + 1e-9)) if hasattr(ops, "bool_op") else ops.make_op("bool_op", [ops.relu(a - b + 1e-9)])[0]


class Expert(Module):
    """E_local feed-forward experts evaluated over the expert-major buffer [E_local, C', H].  Parameters are named by the
    GLOBAL expert index, so expert e has the same weights whichever rank / expert-parallel degree holds it."""

    def __init__(self, d_model, d_ff, num_local_experts, act="gelu", dtype="float32", name="expert", expert_offset=0,
                 device_group=None):
        super().__init__()
        self.n, self.act = num_local_experts, act
        std = 0.02
        dgh = [[device_group]] if device_group is not None else None

        def param(init, shape, nm):
            return parallel_parameter(init, shape, None, dtype=dtype, requires_grad=True, name=nm, device_group_hierarchy=dgh)
        gl = [expert_offset + i for i in range(num_local_experts)]
        self.w1 = [param(normal_initializer(0.0, std), [d_ff, d_model], f"{name}{g}_w1") for g in gl]
        self.b1 = [param(zeros_initializer(), [d_ff], f"{name}{g}_b1") for g in gl]
        self.w2 = [param(normal_initializer(0.0, std), [d_model, d_ff], f"{name}{g}_w2") for g in gl]
        self.b2 = [param(zeros_initializer(), [d_model], f"{name}{g}_b2") for g in gl]
        for i, g in enumerate(gl):
            for nm, lst in (("w1", self.w1), ("b1", self.b1), ("w2", self.w2), ("b2", self.b2)):
                self.register_parameter(f"{nm}_{g}", lst[i])

    def forward(self, x):
        """x [E_local, C', H] -> same shape"""
        parts = ops.split(x, self.n, dim=0) if self.n > 1 else [x]
        outs = []
        for i, p in enumerate(parts):
            t = ops.reshape(p, [p.shape[1] * p.shape[0], p.shape[2]])
            hmid = ops.linear(t, self.w1[i], self.b1[i], act=self.act)
            o = ops.linear(hmid, self.w2[i], self.b2[i])
            outs.append(ops.reshape(o, [1, p.shape[1] * p.shape[0], p.shape[2]]))
        return ops.concat(outs, 0) if len(outs) > 1 else outs[0]


class MoELayer(Module):
    def __init__(self, d_model, d_ff, num_experts, k=1, capacity_factor=1.0, gate_type="topk", ep_ranks: Sequence[int] = (),
                 act="gelu", dtype="float32", name="moe", gate_ds=None):
        super().__init__()
        from ... import distributed
        self.ep_ranks = tuple(ep_ranks)
        self.ep = max(len(self.ep_ranks), 1)
        assert num_experts % self.ep == 0, "experts must divide evenly over the expert-parallel group"
        self.num_experts, self.num_local = num_experts, num_experts // self.ep
        me = distributed.rank()
        pos = self.ep_ranks.index(me) if me in self.ep_ranks else 0
        gate_cls = {"topk": TopKGate, "ktop1": KTop1Gate, "hash": HashGate, "balance": BalanceGate, "sam": SAMGate}[gate_type]
        self.gate = gate_cls(d_model, num_experts, k=k, capacity_factor=capacity_factor, dtype=dtype, name=f"{name}_gate") \
            if gate_type != "hash" else HashGate(d_model, num_experts, capacity_factor)
        if gate_ds is not None and hasattr(self.gate, "wg"):
            # router weights are replicated over the data-parallel group: re-create them with that layout so their
            # gradient is all-reduced like every other replicated parameter
            ds_h, dg_h = gate_ds
            self.gate.wg = parallel_parameter(normal_initializer(0.0, 0.02), list(self.gate.wg.shape), ds_h, dtype=dtype, requires_grad=True,
                                              device_group_hierarchy=dg_h, name=f"{name}_gate_wg")
        self.experts = Expert(d_model, d_ff, self.num_local, act=act, dtype=dtype, name=f"{name}_expert", expert_offset=pos * self.num_local)
        self.l_aux = None

    def forward(self, x):
        """x [T_local, H] -> [T_local, H]"""
        gates, idx, loc, aux, cap = self.gate(x)
        self.l_aux = aux
        ranks = self.ep_ranks if self.ep > 1 else ()
        # layout_transform (+ the dispatch all-to-all fused in: token rows go straight to their expert's rank over
        # NVLink peer memory) -> [E_local, ep * C, H]
        disp = ops.moe_dispatch(x, idx, loc, self.num_experts, cap, ep_ranks=ranks)
        out = self.experts(disp)
        # reverse_layout_transform (+ the combine all-to-all fused in: expert outputs are read from their rank)
        return ops.moe_combine(out, idx, loc, gates, ep_ranks=ranks)


class GPTMoELMHeadModel(Module):
    """GPT whose MLP is an MoE layer every `moe_every` blocks.  Attention / norms / embeddings are the data-parallel
    (replicated) GPT modules over `ds_parallel_configs`; the experts are sharded over `config.ep_ranks` (usually the same
    ranks: expert parallelism rides on the data-parallel group, HetuMoE style)."""

    def __init__(self, config: MoEConfig, ds_parallel_configs=None, num_gpus: int = 1):
        super().__init__()
        from ...nn import HtMultiParallelLayerNorm, HtMultiParallelEmbedding, HtMultiVocabParallelEmbedding
        from ...nn.parallel import get_multi_ds_parallel_config
        from ..gpt.gpt_model import GPTAttention, GPTMLP
        from ..parallel_config import generate_ds_parallel_config
        if ds_parallel_configs is None:
            ds_parallel_configs = [generate_ds_parallel_config(config.n_layer, num_gpus, num_gpus, 1, 1, zero=False)]
        self.config, self.ds_parallel_configs = config, ds_parallel_configs
        dsc, h = ds_parallel_configs, config.n_embd
        std = config.initializer_range
        self.wte = HtMultiVocabParallelEmbedding(config.vocab_size, h, get_multi_ds_parallel_config(dsc, "wte"), dtype=config.dtype, name="wte",
                                                 init_std=std)
        self.wpe = HtMultiParallelEmbedding(config.n_positions, h, get_multi_ds_parallel_config(dsc, "wpe"), dtype=config.dtype, name="wpe",
                                            init_std=std)
        self.blocks = ModuleList()
        for i in range(config.n_layer):
            blk = Module()
            blk.ln_1 = HtMultiParallelLayerNorm(h, get_multi_ds_parallel_config(dsc, "layernorm1", i), eps=config.layer_norm_epsilon,
                                                dtype=config.dtype, name=f"ln1_block{i}")
            blk.attn = GPTAttention(config, dsc, i, name=f"attn_block{i}")
            blk.ln_2 = HtMultiParallelLayerNorm(h, get_multi_ds_parallel_config(dsc, "layernorm2", i), eps=config.layer_norm_epsilon,
                                                dtype=config.dtype, name=f"ln2_block{i}")
            if (i + 1) % config.moe_every == 0:
                blk.moe = MoELayer(h, config.ffn_hidden_size, config.num_experts, config.top_k, config.capacity_factor, config.gate_type,
                                   config.ep_ranks, act=config.activation_function, dtype=config.dtype, name=f"moe_{i}",
                                   gate_ds=(blk.ln_2.ds_dup(), blk.ln_2.device_group_unions))
                blk.mlp = None
            else:
                blk.moe = None
                blk.mlp = GPTMLP(config, dsc, i, name=f"mlp_block{i}")
            self.blocks.append(blk)
        self.ln_f = HtMultiParallelLayerNorm(h, get_multi_ds_parallel_config(dsc, "layernorm_final"), eps=config.layer_norm_epsilon,
                                             dtype=config.dtype, name="ln_final")

    def forward(self, input_ids, position_ids, labels=None, seq_len=None):
        cfg = self.config
        x = self.wte(input_ids) + self.wpe(position_ids)
        aux_total = None
        for blk in self.blocks:
            x = blk.attn(blk.ln_1(x), seq_len, residual=x)
            hln = blk.ln_2(x)
            if blk.moe is not None:
                x = x + blk.moe(hln)
                if blk.moe.l_aux is not None:
                    aux_total = blk.moe.l_aux if aux_total is None else aux_total + blk.moe.l_aux
            else:
                x = blk.mlp(hln, residual=x)
        x = self.ln_f(x)
        logits = ops.linear(x, self.wte.embedding_table, None, trans_b=True, name="lm_head")
        if labels is None:
            return logits
        loss = ops.softmax_cross_entropy_sparse(logits, labels, ignored_index=-1, reduction="mean")
        if aux_total is not None and cfg.aux_loss_weight > 0:
            loss = loss + aux_total * cfg.aux_loss_weight
        return loss


MoELMHeadModel = GPTMoELMHeadModel
