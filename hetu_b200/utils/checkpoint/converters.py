"""State-dict converters between HuggingFace checkpoints and this framework's parameter names / fused layouts, both directions, and
a checkpoint inspector (ref: examples/hetero/gpt_hf_to_ht.py, gpt_hf_to_hf.py, gpt_mt_to_ht.py, examine_ckpt.py,
python/hetu/models/utils/converter/convert_llama_hf_to_ht.py).

Layouts: HF GPT-2 stores linear weights as Conv1D [in, out] and the attention projection as q | k | v blocks; here weights are
[out, in] and the fused qkv projection is head-major ('hqkv': per kv head its query heads, then k, then v) so that any tensor-
parallel degree dividing the kv heads owns whole groups."""
from __future__ import annotations

import json
import os
from typing import Dict

import torch

from .legacy import _interleave_gate_up, _interleave_qkv, convert_llama_hf_to_ht  # noqa: F401  (re-exported)


def _split_hqkv(w: torch.Tensor, num_heads: int, num_kv_heads: int, head_dim: int):
    """inverse of legacy._interleave_qkv('hqkv'): fused rows -> (q, k, v) blocks"""
    rep = num_heads // num_kv_heads
    tail = w.shape[1:]
    g = w.reshape(num_kv_heads, rep + 2, head_dim, *tail)
    q = g[:, :rep].reshape(num_heads * head_dim, *tail)
    k = g[:, rep].reshape(num_kv_heads * head_dim, *tail)
    v = g[:, rep + 1].reshape(num_kv_heads * head_dim, *tail)
    return q, k, v


def convert_gpt2_hf_to_ht(hf_state: Dict[str, torch.Tensor], num_layers: int, num_heads: int) -> Dict[str, torch.Tensor]:
    """HuggingFace GPT2LMHeadModel state dict -> hetu_b200.models.GPTLMHeadModel"""
    pre = "transformer." if "transformer.wte.weight" in hf_state else ""
    out = {"transformer.wte.embedding_table": hf_state[pre + "wte.weight"], "transformer.wpe.embedding_table": hf_state[pre + "wpe.weight"],
           "transformer.ln_f.weight": hf_state[pre + "ln_f.weight"], "transformer.ln_f.bias": hf_state[pre + "ln_f.bias"]}
    hidden = hf_state[pre + "wte.weight"].shape[1]
    hd = hidden // num_heads
    for i in range(num_layers):
        p, o = f"{pre}h.{i}.", f"transformer.h.{i}."
        for ln in ("ln_1", "ln_2"):
            out[o + ln + ".weight"], out[o + ln + ".bias"] = hf_state[p + ln + ".weight"], hf_state[p + ln + ".bias"]
        w = hf_state[p + "attn.c_attn.weight"].t().contiguous()              # [3h, h], rows q | k | v
        b = hf_state[p + "attn.c_attn.bias"]
        q, k, v = w.split(hidden, 0)
        bq, bk, bv = b.split(hidden, 0)
        out[o + "attn.qkv_dense.weight"] = _interleave_qkv(q, k, v, num_heads, num_heads, hd, "hqkv")
        out[o + "attn.qkv_dense.bias"] = _interleave_qkv(bq.unsqueeze(1), bk.unsqueeze(1), bv.unsqueeze(1), num_heads, num_heads, hd, "hqkv").squeeze(1)
        out[o + "attn.dense.weight"] = hf_state[p + "attn.c_proj.weight"].t().contiguous()
        out[o + "attn.dense.bias"] = hf_state[p + "attn.c_proj.bias"]
        out[o + "mlp.dense_h_to_4h.weight"] = hf_state[p + "mlp.c_fc.weight"].t().contiguous()
        out[o + "mlp.dense_h_to_4h.bias"] = hf_state[p + "mlp.c_fc.bias"]
        out[o + "mlp.dense_4h_to_h.weight"] = hf_state[p + "mlp.c_proj.weight"].t().contiguous()
        out[o + "mlp.dense_4h_to_h.bias"] = hf_state[p + "mlp.c_proj.bias"]
    return out


def convert_gpt2_ht_to_hf(ht_state: Dict[str, torch.Tensor], num_layers: int, num_heads: int) -> Dict[str, torch.Tensor]:
    """GPTLMHeadModel state dict -> HuggingFace GPT2LMHeadModel (tied lm_head)"""
    out = {"transformer.wte.weight": ht_state["transformer.wte.embedding_table"], "transformer.wpe.weight": ht_state["transformer.wpe.embedding_table"],
           "transformer.ln_f.weight": ht_state["transformer.ln_f.weight"], "transformer.ln_f.bias": ht_state["transformer.ln_f.bias"],
           "lm_head.weight": ht_state.get("lm_head.weight", ht_state["transformer.wte.embedding_table"])}
    hidden = out["transformer.wte.weight"].shape[1]
    hd = hidden // num_heads
    for i in range(num_layers):
        o, p = f"transformer.h.{i}.", f"transformer.h.{i}."
        for ln in ("ln_1", "ln_2"):
            out[o + ln + ".weight"], out[o + ln + ".bias"] = ht_state[p + ln + ".weight"], ht_state[p + ln + ".bias"]
        q, k, v = _split_hqkv(ht_state[p + "attn.qkv_dense.weight"], num_heads, num_heads, hd)
        bq, bk, bv = _split_hqkv(ht_state[p + "attn.qkv_dense.bias"].unsqueeze(1), num_heads, num_heads, hd)
        out[o + "attn.c_attn.weight"] = torch.cat([q, k, v], 0).t().contiguous()
        out[o + "attn.c_attn.bias"] = torch.cat([bq, bk, bv], 0).squeeze(1)
        out[o + "attn.c_proj.weight"] = ht_state[p + "attn.dense.weight"].t().contiguous()
        out[o + "attn.c_proj.bias"] = ht_state[p + "attn.dense.bias"]
        out[o + "mlp.c_fc.weight"] = ht_state[p + "mlp.dense_h_to_4h.weight"].t().contiguous()
        out[o + "mlp.c_fc.bias"] = ht_state[p + "mlp.dense_h_to_4h.bias"]
        out[o + "mlp.c_proj.weight"] = ht_state[p + "mlp.dense_4h_to_h.weight"].t().contiguous()
        out[o + "mlp.c_proj.bias"] = ht_state[p + "mlp.dense_4h_to_h.bias"]
    return out


def convert_llama_ht_to_hf(ht_state: Dict[str, torch.Tensor], num_layers: int, num_heads: int, num_kv_heads: int) -> Dict[str, torch.Tensor]:
    """LlamaLMHeadModel state dict -> HuggingFace LlamaForCausalLM (inverse of convert_llama_hf_to_ht)"""
    out = {"model.embed_tokens.weight": ht_state["transformer.wte.embedding_table"], "model.norm.weight": ht_state["transformer.rmsnorm_f.weight"],
           "lm_head.weight": ht_state.get("lm_head.weight", ht_state["transformer.wte.embedding_table"])}
    fused = ht_state["transformer.h.0.attn.qkv_dense.weight"]
    hd = fused.shape[0] // (num_heads + 2 * num_kv_heads)
    for i in range(num_layers):
        o, p = f"model.layers.{i}.", f"transformer.h.{i}."
        out[o + "input_layernorm.weight"] = ht_state[p + "rmsnorm_1.weight"]
        out[o + "post_attention_layernorm.weight"] = ht_state[p + "rmsnorm_2.weight"]
        q, k, v = _split_hqkv(ht_state[p + "attn.qkv_dense.weight"], num_heads, num_kv_heads, hd)
        out[o + "self_attn.q_proj.weight"], out[o + "self_attn.k_proj.weight"], out[o + "self_attn.v_proj.weight"] = q, k, v
        out[o + "self_attn.o_proj.weight"] = ht_state[p + "attn.dense.weight"]
        gu = ht_state[p + "mlp.dense_h_to_4h.weight"]
        out[o + "mlp.gate_proj.weight"], out[o + "mlp.up_proj.weight"] = gu[0::2].contiguous(), gu[1::2].contiguous()
        out[o + "mlp.down_proj.weight"] = ht_state[p + "mlp.dense_4h_to_h.weight"]
    return out


def examine_checkpoint(path: str) -> Dict:
    """what is inside a split checkpoint directory: shard files, tensors with global shape / dtype / layout, optimizer states,
    completeness marker (ref: examples/hetero/examine_ckpt.py)"""
    from .ht_safetensors import WEIGHTS_FORMAT, WEIGHTS_NAME, load_file
    files = sorted(f for f in os.listdir(path) if f.startswith(WEIGHTS_NAME) and f.endswith(WEIGHTS_FORMAT))
    metas = sorted(f for f in os.listdir(path) if f.startswith("param_states-") and f.endswith(".json"))
    tensors: Dict[str, Dict] = {}
    for m in metas:
        for name, meta in json.load(open(os.path.join(path, m))).items():
            t = tensors.setdefault(name, {"writers": 0})
            t["writers"] += 1
            for k in ("global_shape", "dtype", "states", "order", "device_num", "device_group"):
                if k in meta and k not in t:
                    t[k] = meta[k]
    blocks, nbytes = 0, 0
    for f in files:
        for k, v in load_file(os.path.join(path, f)).items():
            blocks += 1
            nbytes += v.numel() * v.element_size()
    is_state = lambda n: n.endswith(("_mean", "_variance", "_step"))     # noqa: E731
    return {"path": path, "complete": os.path.exists(os.path.join(path, "COMPLETE")), "shard_files": files, "blocks": blocks, "bytes": nbytes,
            "parameters": {n: t for n, t in tensors.items() if not is_state(n)},
            "optimizer_states": sorted(n for n in tensors if is_state(n))}


def main(argv=None):
    import argparse
    ap = argparse.ArgumentParser(description="inspect a split checkpoint directory")
    ap.add_argument("path")
    a = ap.parse_args(argv)
    info = examine_checkpoint(a.path)
    print(f"{info['path']}: {'COMPLETE' if info['complete'] else 'INCOMPLETE'}, {len(info['shard_files'])} shard files, {info['blocks']} blocks, "
          f"{info['bytes'] / 2**20:.1f} MiB, {len(info['parameters'])} parameters, {len(info['optimizer_states'])} optimizer-state tensors")
    for n, t in sorted(info["parameters"].items()):
        print(f"  {n:60s} {str(t.get('global_shape')):>18s} {t.get('dtype', '?'):>9s} split {t.get('states')} over {t.get('device_num')} device(s), {t['writers']} writer(s)")


if __name__ == "__main__":
    main()
