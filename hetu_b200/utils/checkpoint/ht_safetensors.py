"""Sharded safetensors checkpoints in the reference's on-disk format.

    <dir>/hetu_pytorch_model-{i}-of-{N}.safetensors      one file per device (i = rank + 1)
    <dir>/param_states-{g}-of-{G}.json                   layout of every tensor: {device_num, order, states, device_group,
                                                          split_group}
Every parameter -- and each optimizer state `<param>_mean | _variance | _step | _master` -- is cut into a virtual grid of
TEMP_SPLITS (8) blocks along each of its first SPLIT_DIMS (2) dims; a device stores the blocks covered by its shard under
keys `<name>_split_<k>` with k = i1 * TEMP_SPLITS + i0 (block i_d along dim d).  Because the grid is independent of the
strategy that wrote it, a job running under ANY other (dp, tp, pp) strategy can reassemble exactly the slice it needs.
(ref: python/hetu/utils/checkpoint/ht_safetensors.py:17-20, 905-1074 temp_save_split, 1147-1414 temp_load_split)
"""
from __future__ import annotations

import glob
import json
import os
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence, Tuple

import torch
from safetensors import safe_open
from safetensors.torch import load_file as _st_load
from safetensors.torch import save_file as _st_save

WEIGHTS_NAME = "hetu_pytorch_model"
WEIGHTS_FORMAT = ".safetensors"
TEMP_SPLITS = 8
SPLIT_DIMS = 2


def save_file(tensors: Dict[str, torch.Tensor], filename: str, metadata: Optional[Dict[str, str]] = None):
    os.makedirs(os.path.dirname(os.path.abspath(filename)), exist_ok=True)
    _st_save({k: v.detach().contiguous().cpu() for k, v in tensors.items()}, filename, metadata=metadata or {"format": "pt"})


def load_file(filename: str, device="cpu") -> Dict[str, torch.Tensor]:
    return _st_load(filename, device=device)


# ----------------------------------------------------------------------------- virtual block grid
def _grid(global_shape: Sequence[int], states: Dict[int, int]) -> List[int]:
    """number of virtual blocks along each of the first SPLIT_DIMS dims"""
    out = []
    for d in range(min(len(global_shape), SPLIT_DIMS)):
        split = int(states.get(d, 1))
        n = min(max(split, TEMP_SPLITS), int(global_shape[d]))
        while n > 1 and (global_shape[d] % n != 0 or n % split != 0):
            n -= 1
        out.append(max(n, split))
    return out


def _block_index(idx: Sequence[int]) -> int:
    k = 0
    for d in range(len(idx) - 1, -1, -1):       # last grid dim is the most significant digit (reference order)
        k = k * TEMP_SPLITS + idx[d]
    return k


def split_keys_for_shard(name: str, global_shape: Sequence[int], states: Dict[int, int], begin: Sequence[int], size: Sequence[int]):
    """[(key, global block slice per grid dim)] for the blocks intersecting the shard [begin, begin+size)"""
    grid = _grid(global_shape, states)
    ranges = []
    for d, n in enumerate(grid):
        blk = global_shape[d] // n
        lo, hi = begin[d] // blk, (begin[d] + size[d] - 1) // blk
        ranges.append([(i, i * blk, blk) for i in range(lo, hi + 1)])
    out = []

    def rec(d, idx, sl):
        if d == len(ranges):
            out.append((f"{name}_split_{_block_index(idx)}", list(sl)))
            return
        for i, start, blk in ranges[d]:
            rec(d + 1, idx + [i], sl + [(start, blk)])
    rec(0, [], [])
    return out


def _shard_of(ds, device_index: int, global_shape: Sequence[int]):
    if ds is None or ds.device_num <= 1:
        return [0] * len(global_shape), list(global_shape)
    return ds.local_slice(list(global_shape), device_index)


def _tensor_blocks(name: str, local: torch.Tensor, global_shape, ds, device_index) -> Tuple[Dict[str, torch.Tensor], dict]:
    states = {int(k): int(v) for k, v in (ds.states.items() if ds is not None else {}.items()) if v > 1}
    begin, size = _shard_of(ds, device_index, global_shape)
    blocks = OrderedDict()
    keys = []
    for key, sl in split_keys_for_shard(name, global_shape, states, begin, size):
        piece = local
        for d, (start, blk) in enumerate(sl):
            piece = piece.narrow(d, start - begin[d], blk)
        blocks[key] = piece.contiguous()
        keys.append(int(key.rsplit("_", 1)[1]))
    meta = {"device_num": ds.device_num if ds is not None else 1, "order": list(ds.order) if ds is not None else [],
            "states": states, "split_group": keys, "global_shape": list(global_shape), "dtype": str(local.dtype).replace("torch.", "")}
    return blocks, meta


# ----------------------------------------------------------------------------- split save / load
def temp_save_split(model, optimizer, filename: str, config=None, local_device=None, save_dtype=None, force_contiguous=False,
                    only_lora: bool = False, metadata: Optional[Dict[str, str]] = None, step: Optional[int] = None):
    """Save this rank's shards of `model` (+ optimizer states) under directory `filename`."""
    return write_split_state(filename, collect_split_state(model, optimizer, save_dtype=save_dtype, only_lora=only_lora, metadata=metadata,
                                                           step=step))


def collect_split_state(model, optimizer, save_dtype=None, only_lora: bool = False, metadata: Optional[Dict[str, str]] = None,
                        step: Optional[int] = None, snapshot=None) -> dict:
    """Phase 1 of a split save, on the training thread: this rank's blocks as HOST tensors.  `snapshot(t) -> host tensor`
    replaces the default `.cpu()` (asynchronous saving must own a private copy: pinned / shared memory)."""
    from ...core import _graphs_by_id
    from ...distributed import rank, world_size
    r, n = rank(), world_size()
    to_host = snapshot or (lambda t: t.cpu())
    tensors: Dict[str, torch.Tensor] = OrderedDict()
    ds_json = {}
    seen = set()
    for key, p in model.named_parameters():
        if only_lora != ("lora" in key):
            continue
        if p.id in seen:
            continue
        seen.add(p.id)
        g = _graphs_by_id[p.graph_id]
        group = p.device_group
        ranks = group.indices() if not group.empty else [r]
        if r not in ranks or not g.has_param(p):
            continue
        didx = ranks.index(r)
        items = [(key, p)]
        if optimizer is not None:
            for sname, st in optimizer.get_states(p).items():
                items.append((f"{key}_{sname}", st))
        for name, t in items:
            if not g.has_param(t):
                continue
            data = g.get_param(t).detach()
            if save_dtype is not None and data.is_floating_point() and t is p:
                from ...core import to_torch_dtype
                data = data.to(to_torch_dtype(save_dtype))
            ds = t.get_ds(g.cur_strategy)
            gshape = list(ds.global_shape(list(data.shape))) if ds is not None else list(data.shape)
            if ds is not None and list(ds.local_shape(gshape)) != list(data.shape):
                ds, gshape = None, list(data.shape)      # flat-sharded state: stored as an opaque local tensor
            blocks, meta = _tensor_blocks(name, to_host(data), gshape, ds, didx if ds is not None else 0)
            meta["device_group"] = ranks
            tensors.update(blocks)
            ds_json[name] = meta
    md = dict(metadata or {})
    md.setdefault("format", "pt")
    if step is not None:
        md["step"] = str(step)
    return {"tensors": tensors, "ds_json": ds_json, "metadata": md, "rank": r, "world": n}


def write_split_state(filename: str, state: dict):
    """Phase 2: write the collected blocks (no framework state is touched: safe on a background thread / forked process)"""
    os.makedirs(filename, exist_ok=True)
    r, n = state["rank"], state["world"]
    save_file(state["tensors"], os.path.join(filename, f"{WEIGHTS_NAME}-{r + 1}-of-{n}{WEIGHTS_FORMAT}"), state["metadata"])
    with open(os.path.join(filename, f"param_states-{r + 1}-of-{n}.json"), "w") as f:
        json.dump(state["ds_json"], f)
    return list(state["tensors"].keys())


class _SplitIndex:
    """key -> file for every block of a checkpoint directory"""

    def __init__(self, path: str):
        self.files = sorted(glob.glob(os.path.join(path, f"{WEIGHTS_NAME}-*-of-*{WEIGHTS_FORMAT}")))
        if not self.files:
            raise FileNotFoundError(f"no {WEIGHTS_NAME}-*{WEIGHTS_FORMAT} under {path}")
        self.where: Dict[str, str] = {}
        self.meta: Dict[str, dict] = {}
        for f in self.files:
            with safe_open(f, framework="pt") as h:
                for k in h.keys():
                    self.where.setdefault(k, f)
        for j in sorted(glob.glob(os.path.join(path, "param_states-*-of-*.json"))):
            with open(j) as fh:
                for k, v in json.load(fh).items():
                    self.meta.setdefault(k, v)
        self._open: Dict[str, object] = {}

    def get(self, key: str) -> torch.Tensor:
        f = self.where[key]
        if f not in self._open:
            self._open[f] = safe_open(f, framework="pt")
        return self._open[f].get_tensor(key)


def assemble_from_splits(index: _SplitIndex, name: str, global_shape, begin, size) -> torch.Tensor:
    """rebuild the slice [begin, begin+size) of tensor `name` from whatever blocks the checkpoint holds"""
    meta = index.meta.get(name)
    if meta is None:
        raise KeyError(f"{name} not in checkpoint")
    saved_states = {int(k): int(v) for k, v in meta["states"].items()}
    out = None
    for key, sl in split_keys_for_shard(name, global_shape, saved_states, begin, size):
        blk = index.get(key)
        if out is None:
            out = torch.empty(list(size), dtype=blk.dtype)
        dst, src = out, blk
        for d, (start, n) in enumerate(sl):
            lo, hi = max(start, begin[d]), min(start + n, begin[d] + size[d])
            dst = dst.narrow(d, lo - begin[d], hi - lo)
            src = src.narrow(d, lo - start, hi - lo)
        dst.copy_(src)
    return out


def temp_load_split(model, optimizer, filename: str, config=None, local_device=None, only_lora: bool = False, strict: bool = True):
    """Load a split checkpoint into `model` / `optimizer` under the CURRENT strategy (any layout may have written it)."""
    from ...core import _graphs_by_id
    from ...distributed import rank
    index = _SplitIndex(filename)
    r = rank()
    loaded, missing = [], []
    for key, p in model.named_parameters():
        if only_lora != ("lora" in key):
            continue
        g = _graphs_by_id[p.graph_id]
        group = p.device_group
        ranks = group.indices() if not group.empty else [r]
        if r not in ranks:
            continue
        didx = ranks.index(r)
        items = [(key, p)]
        if optimizer is not None:
            for sname, st in optimizer.get_states(p).items():
                items.append((f"{key}_{sname}", st))
        for name, t in items:
            if name not in index.meta:
                missing.append(name)
                continue
            gshape = index.meta[name]["global_shape"]
            ds = t.get_ds(g.cur_strategy)
            cur = g.get_param(t)
            if ds is not None and list(ds.local_shape(gshape)) == list(cur.shape):
                begin, size = _shard_of(ds, didx, gshape)
            else:
                begin, size = [0] * len(gshape), list(gshape)
            val = assemble_from_splits(index, name, gshape, begin, size)
            if list(val.shape) != list(cur.shape):
                missing.append(name)      # e.g. flat-sharded optimizer state written under another dp degree
                continue
            g.set_param(t, val.to(cur.dtype))
            loaded.append(name)
    if strict and any(m in dict(model.named_parameters()) for m in missing):
        raise KeyError(f"checkpoint misses parameters: {missing}")
    return loaded, missing


# ----------------------------------------------------------------------------- whole-tensor variants
def save_model(model, filename: str, metadata=None):
    """single-file save of the full (local) state dict"""
    save_file(model.state_dict(), filename if filename.endswith(WEIGHTS_FORMAT) else filename + WEIGHTS_FORMAT, metadata)


def load_model(model, filename: str, strict: bool = True):
    sd = load_file(filename if filename.endswith(WEIGHTS_FORMAT) else filename + WEIGHTS_FORMAT)
    return model.load_state_dict(sd, strict=strict)


def temp_save(model, optimizer, filename: str, **kw):
    from ...distributed import rank, world_size
    os.makedirs(filename, exist_ok=True)
    sd = dict(model.state_dict())
    if optimizer is not None:
        from ...core import _graphs_by_id
        for key, p in model.named_parameters():
            for sname, st in optimizer.get_states(p).items():
                sd[f"{key}_{sname}"] = _graphs_by_id[st.graph_id].get_param(st)
    save_file(sd, os.path.join(filename, f"{WEIGHTS_NAME}-{rank() + 1}-of-{world_size()}{WEIGHTS_FORMAT}"))


def temp_load(model, optimizer, filename: str, strict: bool = False, **kw):
    from ...distributed import rank, world_size
    sd = load_file(os.path.join(filename, f"{WEIGHTS_NAME}-{rank() + 1}-of-{world_size()}{WEIGHTS_FORMAT}"))
    own = dict(model.named_parameters())
    model.load_state_dict({k: v for k, v in sd.items() if k in own}, strict=strict)
    if optimizer is not None:
        from ...core import _graphs_by_id
        for key, p in own.items():
            for sname, st in optimizer.get_states(p).items():
                if f"{key}_{sname}" in sd:
                    _graphs_by_id[st.graph_id].set_param(st, sd[f"{key}_{sname}"])
