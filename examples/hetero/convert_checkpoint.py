"""Convert checkpoints between HuggingFace and this framework, and inspect a split checkpoint.

    # HuggingFace GPT-2 / Llama state dict (a .pt / .bin / .safetensors file or a directory with one) -> framework names + fused layouts
    python examples/hetero/convert_checkpoint.py hf2ht --model gpt2  --layers 12 --heads 12 --src gpt2/pytorch_model.bin --dst gpt2_ht.pt
    python examples/hetero/convert_checkpoint.py hf2ht --model llama --layers 32 --heads 32 --kv-heads 32 --src llama-2-7b/ --dst llama_ht.pt
    # and back (e.g. to evaluate a model trained here with HuggingFace tooling)
    python examples/hetero/convert_checkpoint.py ht2hf --model llama --layers 32 --heads 32 --kv-heads 32 --src llama_ht.pt --dst llama_hf.pt
    # what is inside a split checkpoint written by ModelSaver / temp_save_split
    python examples/hetero/convert_checkpoint.py examine ckpt/step100

The converted `.pt` file is a plain state dict: `model.load_state_dict(torch.load(path))` inside a graph of ANY parallel strategy
places every rank's shard (the loaders slice by the parameter's DistributedStates).
(ref: examples/hetero/gpt_hf_to_ht.py, gpt_hf_to_hf.py, gpt_mt_to_ht.py, examine_ckpt.py)"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from hetu_b200.utils.checkpoint import (convert_gpt2_hf_to_ht, convert_gpt2_ht_to_hf, convert_llama_hf_to_ht, convert_llama_ht_to_hf,  # noqa: E402
                                        examine_checkpoint)


def read_state(path: str):
    if os.path.isdir(path):
        cands = sorted(f for f in os.listdir(path) if f.endswith((".safetensors", ".bin", ".pt")))
        assert cands, f"no weight file under {path}"
        state = {}
        for f in cands:
            state.update(read_state(os.path.join(path, f)))
        return state
    if path.endswith(".safetensors"):
        from hetu_b200.utils.checkpoint import load_file
        return load_file(path)
    obj = torch.load(path, map_location="cpu", weights_only=False)
    return obj.get("state_dict", obj) if isinstance(obj, dict) else obj


ap = argparse.ArgumentParser()
sub = ap.add_subparsers(dest="cmd", required=True)
for name in ("hf2ht", "ht2hf"):
    p = sub.add_parser(name)
    p.add_argument("--model", choices=["gpt2", "llama"], required=True)
    p.add_argument("--layers", type=int, required=True)
    p.add_argument("--heads", type=int, required=True)
    p.add_argument("--kv-heads", type=int, default=None)
    p.add_argument("--src", required=True)
    p.add_argument("--dst", required=True)
ex = sub.add_parser("examine")
ex.add_argument("path")
a = ap.parse_args()
if a.cmd == "examine":
    from hetu_b200.utils.checkpoint.converters import main as examine_main
    examine_main([a.path])
    sys.exit(0)
state = read_state(a.src)
kv = a.kv_heads or a.heads
if a.cmd == "hf2ht":
    out = convert_gpt2_hf_to_ht(state, a.layers, a.heads) if a.model == "gpt2" else convert_llama_hf_to_ht(state, a.layers, a.heads, kv)
else:
    out = convert_gpt2_ht_to_hf(state, a.layers, a.heads) if a.model == "gpt2" else convert_llama_ht_to_hf(state, a.layers, a.heads, kv)
os.makedirs(os.path.dirname(os.path.abspath(a.dst)), exist_ok=True)
torch.save({k: v.contiguous() for k, v in out.items()}, a.dst)
print(f"{a.cmd}: {len(state)} tensors -> {len(out)} tensors, {sum(v.numel() for v in out.values()) / 1e6:.2f} M parameters, written to {a.dst}")
