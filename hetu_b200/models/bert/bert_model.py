"""BERT family: embeddings (word + position + token type), post-LayerNorm encoder, pooler, and the pre-training (MLM + NSP),
masked-LM and sequence-classification heads.  Batches are [B, S]; padded batches pass `attention_mask` [B, S] (1 = token, 0 = pad):
without a mask the fused non-causal flash kernels run, with one the composed masked-attention path does.  Data parallelism comes
from the graph's DistributedStates (replicated parameters, split batch), like the CTR / GNN models.
(ref: hetu/v1/examples/nlp/bert/hetu_bert.py BertModel / BertForPreTraining / BertForSequenceClassification, bert_config.py)"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict

from ... import ops
from ...nn import Dropout, Embedding, LayerNorm, Linear, Module, ModuleList
from ..utils.pretrained import PreTrainedConfig, PreTrainedModel


@dataclass
class BertConfig(PreTrainedConfig):
    vocab_size: int = 30522
    hidden_size: int = 768
    num_hidden_layers: int = 12
    num_attention_heads: int = 12
    intermediate_size: int = 3072
    hidden_act: str = "gelu"
    hidden_dropout_prob: float = 0.1
    attention_probs_dropout_prob: float = 0.1
    max_position_embeddings: int = 512
    type_vocab_size: int = 2
    layer_norm_eps: float = 1e-12
    num_labels: int = 2

    @staticmethod
    def base(**kw):
        return BertConfig(**kw)

    @staticmethod
    def large(**kw):
        return BertConfig(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096, **kw)


def _masked_attention(q, k, v, key_mask, p_dropout):
    """q/k/v [B, S, H, D], key_mask [B, S] (1 keep / 0 pad) -> [B, S, H, D]: scores + additive mask -> softmax -> dropout -> PV"""
    b, s, h, d = q.shape
    qt = ops.reshape(ops.transpose(q, [0, 2, 1, 3]), [b * h, s, d])
    kt = ops.reshape(ops.transpose(k, [0, 2, 1, 3]), [b * h, s, d])
    vt = ops.reshape(ops.transpose(v, [0, 2, 1, 3]), [b * h, s, d])
    scores = ops.bmm(qt, ops.transpose(kt, [0, 2, 1])) * (1.0 / float(d) ** 0.5)
    bias = (ops.reshape(key_mask, [b, 1, 1, s]) - 1.0) * 1e30                      # 0 for tokens, -1e30 for padding
    scores = ops.reshape(ops.reshape(scores, [b, h, s, s]) + bias, [b * h, s, s])
    probs = ops.softmax(scores, -1)
    if p_dropout > 0:
        probs = ops.dropout(probs, p_dropout)
    out = ops.bmm(probs, vt)
    return ops.transpose(ops.reshape(out, [b, h, s, d]), [0, 2, 1, 3])


class BertEmbeddings(Module):
    def __init__(self, c: BertConfig):
        super().__init__()
        self.word_embeddings = Embedding(c.vocab_size, c.hidden_size, name="bert_word_embeddings")
        self.position_embeddings = Embedding(c.max_position_embeddings, c.hidden_size, name="bert_position_embeddings")
        self.token_type_embeddings = Embedding(c.type_vocab_size, c.hidden_size, name="bert_token_type_embeddings")
        self.LayerNorm = LayerNorm(c.hidden_size, eps=c.layer_norm_eps, name="bert_emb_ln")
        self.dropout = Dropout(c.hidden_dropout_prob)

    def forward(self, input_ids, token_type_ids, position_ids):
        x = self.word_embeddings(input_ids) + self.position_embeddings(position_ids) + self.token_type_embeddings(token_type_ids)
        return self.dropout(self.LayerNorm(x))


class BertLayer(Module):
    def __init__(self, c: BertConfig, i: int):
        super().__init__()
        self.c, h = c, c.hidden_size
        self.query = Linear(h, h, name=f"bert_l{i}_query")
        self.key = Linear(h, h, name=f"bert_l{i}_key")
        self.value = Linear(h, h, name=f"bert_l{i}_value")
        self.attn_out = Linear(h, h, name=f"bert_l{i}_attn_out")
        self.attn_ln = LayerNorm(h, eps=c.layer_norm_eps, name=f"bert_l{i}_attn_ln")
        self.intermediate = Linear(h, c.intermediate_size, name=f"bert_l{i}_intermediate")
        self.output = Linear(c.intermediate_size, h, name=f"bert_l{i}_output")
        self.out_ln = LayerNorm(h, eps=c.layer_norm_eps, name=f"bert_l{i}_out_ln")
        self.dropout = Dropout(c.hidden_dropout_prob)

    def forward(self, x, batch, seq, key_mask=None):
        c = self.c
        nh, hd = c.num_attention_heads, c.hidden_size // c.num_attention_heads
        shape = [batch, seq, nh, hd]
        q, k, v = (ops.reshape(lin(x), shape) for lin in (self.query, self.key, self.value))
        p = float(c.attention_probs_dropout_prob) if self.training else 0.0
        if key_mask is None:
            a = ops.attn(q, k, v, p_dropout=p, is_causal=False)
        else:
            a = _masked_attention(q, k, v, key_mask, p)
        a = ops.reshape(a, [batch * seq, c.hidden_size])
        x = self.attn_ln(self.dropout(self.attn_out(a)) + x)
        h = self.intermediate(x, act="gelu" if c.hidden_act == "gelu" else c.hidden_act)
        return self.out_ln(self.dropout(self.output(h)) + x)


class BertModel(Module):
    """-> (sequence_output [B * S, hidden], pooled_output [B, hidden])"""

    def __init__(self, config: BertConfig, add_pooling_layer: bool = True):
        super().__init__()
        self.config = config
        self.embeddings = BertEmbeddings(config)
        self.layers = ModuleList([BertLayer(config, i) for i in range(config.num_hidden_layers)])
        self.pooler = Linear(config.hidden_size, config.hidden_size, name="bert_pooler") if add_pooling_layer else None

    def forward(self, input_ids, token_type_ids=None, attention_mask=None, position_ids=None):
        b, s = input_ids.shape
        flat = ops.reshape(input_ids, [b * s])
        if token_type_ids is None:
            token_type_ids = ops.zeros_like(flat)
        else:
            token_type_ids = ops.reshape(token_type_ids, [b * s])
        if position_ids is None:
            import numpy as np
            from ...core import from_numpy
            position_ids = from_numpy(np.tile(np.arange(s, dtype=np.int64), b))
        else:
            position_ids = ops.reshape(position_ids, [b * s])
        x = self.embeddings(flat, token_type_ids, position_ids)
        for layer in self.layers:
            x = layer(x, b, s, attention_mask)
        pooled = None
        if self.pooler is not None:
            first = ops.reshape(ops.slice(ops.reshape(x, [b, s, self.config.hidden_size]), [0, 0, 0], [b, 1, self.config.hidden_size]),
                                [b, self.config.hidden_size])
            pooled = ops.tanh(self.pooler(first))
        return x, pooled


class _MLMHead(Module):
    def __init__(self, c: BertConfig, word_embeddings: Embedding):
        super().__init__()
        self.transform = Linear(c.hidden_size, c.hidden_size, name="bert_mlm_transform")
        self.LayerNorm = LayerNorm(c.hidden_size, eps=c.layer_norm_eps, name="bert_mlm_ln")
        self.decoder_weight = word_embeddings.weight                 # tied to the input embedding
        from ...core import parallel_parameter, zeros_initializer
        self.decoder_bias = parallel_parameter(zeros_initializer(), [c.vocab_size], None, requires_grad=True, name="bert_mlm_decoder_bias")

    def forward(self, x):
        h = self.LayerNorm(self.transform(x, act="gelu"))
        return ops.linear(h, self.decoder_weight, self.decoder_bias, trans_b=True)


class BertForMaskedLM(Module, PreTrainedModel):
    config_class = BertConfig

    def __init__(self, config: BertConfig):
        super().__init__()
        self.config = config
        self.bert = BertModel(config, add_pooling_layer=False)
        self.cls = _MLMHead(config, self.bert.embeddings.word_embeddings)

    def forward(self, input_ids, token_type_ids=None, attention_mask=None, labels=None):
        seq, _ = self.bert(input_ids, token_type_ids, attention_mask)
        logits = self.cls(seq)
        if labels is None:
            return logits
        b, s = input_ids.shape
        return ops.softmax_cross_entropy_sparse(logits, ops.reshape(labels, [b * s]), ignored_index=-100, reduction="mean"), logits


class BertForPreTraining(Module, PreTrainedModel):
    config_class = BertConfig

    """masked-LM + next-sentence heads; loss = MLM cross entropy (labels -100 ignored) + NSP cross entropy"""

    def __init__(self, config: BertConfig):
        super().__init__()
        self.config = config
        self.bert = BertModel(config)
        self.cls = _MLMHead(config, self.bert.embeddings.word_embeddings)
        self.seq_relationship = Linear(config.hidden_size, 2, name="bert_nsp")

    def forward(self, input_ids, token_type_ids=None, attention_mask=None, masked_lm_labels=None, next_sentence_label=None):
        seq, pooled = self.bert(input_ids, token_type_ids, attention_mask)
        mlm, nsp = self.cls(seq), self.seq_relationship(pooled)
        if masked_lm_labels is None:
            return mlm, nsp
        b, s = input_ids.shape
        loss = ops.softmax_cross_entropy_sparse(mlm, ops.reshape(masked_lm_labels, [b * s]), ignored_index=-100, reduction="mean")
        if next_sentence_label is not None:
            loss = loss + ops.softmax_cross_entropy_sparse(nsp, ops.reshape(next_sentence_label, [b]), reduction="mean")
        return loss, mlm, nsp


class BertForSequenceClassification(Module, PreTrainedModel):
    config_class = BertConfig

    def __init__(self, config: BertConfig):
        super().__init__()
        self.config = config
        self.bert = BertModel(config)
        self.dropout = Dropout(config.hidden_dropout_prob)
        self.classifier = Linear(config.hidden_size, config.num_labels, name="bert_classifier")

    def forward(self, input_ids, token_type_ids=None, attention_mask=None, labels=None):
        _, pooled = self.bert(input_ids, token_type_ids, attention_mask)
        logits = self.classifier(self.dropout(pooled))
        if labels is None:
            return logits
        return ops.softmax_cross_entropy_sparse(logits, labels, reduction="mean"), logits


def convert_bert_hf_to_ht(hf_state: Dict, num_layers: int, prefix: str = "bert.") -> Dict:
    """HuggingFace BertModel / BertForPreTraining state dict -> the names of the classes above (`bert.` prefix for the heads' owner)"""
    src = "bert." if any(k.startswith("bert.") for k in hf_state) else ""
    out = {}
    e = src + "embeddings."
    for n in ("word_embeddings", "position_embeddings", "token_type_embeddings"):
        out[f"{prefix}embeddings.{n}.weight"] = hf_state[e + n + ".weight"]
    out[f"{prefix}embeddings.LayerNorm.weight"], out[f"{prefix}embeddings.LayerNorm.bias"] = hf_state[e + "LayerNorm.weight"], hf_state[e + "LayerNorm.bias"]
    ren = {"attention.self.query": "query", "attention.self.key": "key", "attention.self.value": "value", "attention.output.dense": "attn_out",
           "attention.output.LayerNorm": "attn_ln", "intermediate.dense": "intermediate", "output.dense": "output", "output.LayerNorm": "out_ln"}
    for i in range(num_layers):
        for a, b in ren.items():
            for t in ("weight", "bias"):
                out[f"{prefix}layers.{i}.{b}.{t}"] = hf_state[f"{src}encoder.layer.{i}.{a}.{t}"]
    if src + "pooler.dense.weight" in hf_state:
        out[f"{prefix}pooler.weight"], out[f"{prefix}pooler.bias"] = hf_state[src + "pooler.dense.weight"], hf_state[src + "pooler.dense.bias"]
    if "cls.predictions.transform.dense.weight" in hf_state:
        out["cls.transform.weight"], out["cls.transform.bias"] = hf_state["cls.predictions.transform.dense.weight"], hf_state["cls.predictions.transform.dense.bias"]
        out["cls.LayerNorm.weight"], out["cls.LayerNorm.bias"] = hf_state["cls.predictions.transform.LayerNorm.weight"], hf_state["cls.predictions.transform.LayerNorm.bias"]
        out["cls.decoder_bias"] = hf_state["cls.predictions.bias"]
    if "cls.seq_relationship.weight" in hf_state:
        out["seq_relationship.weight"], out["seq_relationship.bias"] = hf_state["cls.seq_relationship.weight"], hf_state["cls.seq_relationship.bias"]
    if "classifier.weight" in hf_state:
        out["classifier.weight"], out["classifier.bias"] = hf_state["classifier.weight"], hf_state["classifier.bias"]
    return out
