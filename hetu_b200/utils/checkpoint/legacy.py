"""Legacy pickle checkpoints and importers from Megatron-LM / HuggingFace layouts
(ref: python/hetu/utils/checkpoint/{save_checkpoint,load_checkpoint}.py, models/utils/converter/convert_llama_hf_to_ht.py)."""
from __future__ import annotations

import os
import pickle
from typing import Dict

import torch


def save_checkpoint(model, path: str, optimizer=None, step: int = 0):
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "wb") as f:
        pickle.dump({"state_dict": {k: v.cpu() for k, v in model.state_dict().items()}, "step": step}, f)


def load_checkpoint(model, path: str, strict: bool = False):
    with open(path, "rb") as f:
        ckpt = pickle.load(f)
    model.load_state_dict(ckpt["state_dict"], strict=strict)
    return ckpt.get("step", 0)


def _interleave_qkv(q, k, v, num_heads, num_kv_heads, head_dim, layout):
    """HF stores q | k | v blocked; our fused projection is kv-head-major [g: q x rep, k, v] ('hqkv'; per-head [q k v]
    when num_heads == num_kv_heads) so that any tensor-parallel degree dividing the kv heads owns whole groups."""
    if layout == "hqkv":
        rep = num_heads // num_kv_heads
        qh = q.reshape(num_kv_heads, rep, head_dim, -1)
        kh = k.reshape(num_kv_heads, 1, head_dim, -1)
        vh = v.reshape(num_kv_heads, 1, head_dim, -1)
        return torch.cat([qh, kh, vh], 1).reshape((num_heads + 2 * num_kv_heads) * head_dim, -1)
    return torch.cat([q, k, v], 0)


def _interleave_gate_up(gate, up):
    """rows (gate_0, up_0, gate_1, up_1, ...): the layout of the fused gate/up projection (swiglu interleaved=True)"""
    return torch.stack([gate, up], 1).reshape(2 * gate.shape[0], -1)


def convert_llama_hf_to_ht(hf_state: Dict[str, torch.Tensor], num_layers: int, num_heads: int, num_kv_heads: int) -> Dict[str, torch.Tensor]:
    """HuggingFace LlamaForCausalLM state dict -> hetu_b200.models.LlamaLMHeadModel names.
    HF's rotary uses the half-split convention natively, so no q/k permutation is needed."""
    out = {}
    hd = hf_state["model.layers.0.self_attn.q_proj.weight"].shape[0] // num_heads
    out["transformer.wte.embedding_table"] = hf_state["model.embed_tokens.weight"]
    out["transformer.rmsnorm_f.weight"] = hf_state["model.norm.weight"]
    out["lm_head.weight"] = hf_state.get("lm_head.weight", hf_state["model.embed_tokens.weight"])
    for i in range(num_layers):
        p = f"model.layers.{i}."
        o = f"transformer.h.{i}."
        out[o + "rmsnorm_1.weight"] = hf_state[p + "input_layernorm.weight"]
        out[o + "rmsnorm_2.weight"] = hf_state[p + "post_attention_layernorm.weight"]
        out[o + "attn.qkv_dense.weight"] = _interleave_qkv(hf_state[p + "self_attn.q_proj.weight"], hf_state[p + "self_attn.k_proj.weight"],
                                                            hf_state[p + "self_attn.v_proj.weight"], num_heads, num_kv_heads, hd, "hqkv")
        out[o + "attn.dense.weight"] = hf_state[p + "self_attn.o_proj.weight"]
        out[o + "mlp.dense_h_to_4h.weight"] = _interleave_gate_up(hf_state[p + "mlp.gate_proj.weight"], hf_state[p + "mlp.up_proj.weight"])
        out[o + "mlp.dense_4h_to_h.weight"] = hf_state[p + "mlp.down_proj.weight"]
    return out


def load_checkpoint_from_megatron(model, path: str, num_heads: int, strict: bool = False):
    """Megatron-LM GPT checkpoint (model_optim_rng.pt, tp=1) -> GPTLMHeadModel.  Megatron's fused qkv is already
    per-head interleaved, which is our 'hqkv' layout."""
    ckpt = torch.load(path, map_location="cpu", weights_only=False)
    lm = ckpt["model"]["language_model"]
    enc = lm.get("encoder", lm.get("transformer"))
    sd = {"transformer.wte.embedding_table": lm["embedding"]["word_embeddings"]["weight"],
          "transformer.wpe.embedding_table": lm["embedding"]["position_embeddings"]["weight"]}
    ren = {"input_layernorm": "ln_1", "post_attention_layernorm": "ln_2", "self_attention.query_key_value": "attn.qkv_dense",
           "attention.query_key_value": "attn.qkv_dense", "self_attention.dense": "attn.dense", "attention.dense": "attn.dense",
           "mlp.dense_h_to_4h": "mlp.dense_h_to_4h", "mlp.dense_4h_to_h": "mlp.dense_4h_to_h"}
    for k, v in enc.items():
        if k.startswith("final_layernorm"):
            sd["transformer.ln_f." + k.split(".")[-1]] = v
            continue
        parts = k.split(".")
        if parts[0] != "layers":
            continue
        i, rest = parts[1], ".".join(parts[2:-1])
        if rest in ren:
            sd[f"transformer.h.{i}.{ren[rest]}.{parts[-1]}"] = v
    model.load_state_dict(sd, strict=strict)
    return ckpt.get("iteration", 0)
