"""Fused / packed variants used by the model zoo (attention on packed QKV, packed rotary)."""
from __future__ import annotations

from .core import IntSymbol, make_op
from .ops import _meta


def _seq(seq_len):
    if isinstance(seq_len, IntSymbol):
        return {"seq_len": int(seq_len.get_data())}, [seq_len]
    return {"seq_len": int(seq_len)}, []


def attn_packed(qkv, seq_len, num_heads, num_kv_heads=None, head_dim=None, is_causal=True, softmax_scale=-1.0, layout="qkv",
                cu_seqlens=None, **kw):
    """qkv [T, (H + 2*Hkv) * D] -> o [T, H*D]; q/k/v are read through strides, no copies.
    cu_seqlens (int32 [n + 1], cumulative document boundaries over the T tokens; trailing repeats of T are ignored) turns
    this into variable-length attention over a packed batch: every document attends only to itself."""
    num_kv_heads = num_kv_heads or num_heads
    head_dim = head_dim or qkv.shape[-1] // (num_heads + 2 * num_kv_heads)
    a, sy = _seq(seq_len)
    a.update({"num_heads": int(num_heads), "num_kv_heads": int(num_kv_heads), "head_dim": int(head_dim), "causal": bool(is_causal),
              "softmax_scale": float(softmax_scale if softmax_scale > 0 else 0.0), "layout": str(layout)})
    ins = [qkv] if cu_seqlens is None else [qkv, cu_seqlens]
    return make_op("attn_packed", ins, a, sy_shape=sy, **_meta(kw))[0]


def rotary_packed(qkv, seq_len, num_heads, num_kv_heads=None, head_dim=None, positions=None, base=10000.0, pos_offset=0,
                  layout="qkv", **kw):
    num_kv_heads = num_kv_heads or num_heads
    head_dim = head_dim or qkv.shape[-1] // (num_heads + 2 * num_kv_heads)
    a, sy = _seq(seq_len)
    a.update({"num_heads": int(num_heads), "num_kv_heads": int(num_kv_heads), "head_dim": int(head_dim), "base": float(base),
              "pos_offset": int(pos_offset), "layout": str(layout)})
    ins = [qkv] if positions is None else [qkv, positions]
    return make_op("rotary_packed", ins, a, sy_shape=sy, **_meta(kw))[0]
