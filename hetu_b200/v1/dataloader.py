"""v1 data feeding: `dataloader_op([Dataloader(train_x, bs, 'train'), Dataloader(valid_x, bs, 'validate')])` is a graph
node whose value the Executor fills from the loader registered under the executor's current name
(ref: hetu/v1/python/hetu/dataloader.py: Dataloader, DataloaderOp, dataloader_op; data-parallel sharding by rank)."""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence, Union

import numpy as np

from .. import core


class Dataloader:
    def __init__(self, raw_data: Union[np.ndarray, Callable[[], np.ndarray]], batch_size: int, name: str = "default", func: Optional[Callable] = None,
                 shuffle: bool = False, drop_last: bool = True, dtype=np.float32, seed: int = 0):
        self.func = func
        self.raw_data = np.asarray(raw_data() if callable(raw_data) else raw_data, dtype=dtype)
        if func is not None:
            self.raw_data = np.asarray(func(self.raw_data), dtype=dtype)
        self.batch_size, self.name, self.shuffle, self.drop_last = int(batch_size), name, shuffle, drop_last
        self.dp_rank, self.dp_nrank = 0, 1
        self.rng = np.random.RandomState(seed)
        self._prepare()

    def set_dp_rank(self, dp_rank: int, dp_nrank: int):
        """every data-parallel worker reads its own contiguous part"""
        self.dp_rank, self.dp_nrank = int(dp_rank), int(dp_nrank)
        self._prepare()

    def _prepare(self):
        n = len(self.raw_data)
        per = n // self.dp_nrank
        lo = self.dp_rank * per
        self.part = self.raw_data[lo:lo + per] if self.dp_nrank > 1 else self.raw_data
        m = len(self.part)
        self.batch_num = m // self.batch_size if self.drop_last else -(-m // self.batch_size)
        assert self.batch_num > 0, "batch size larger than the (per-worker) dataset"
        self.order = np.arange(m)
        if self.shuffle:
            self.rng.shuffle(self.order)
        self.index = 0

    @property
    def shape(self):
        return (self.batch_size,) + tuple(self.part.shape[1:])

    def get_arr(self) -> np.ndarray:
        lo = self.index * self.batch_size
        idx = self.order[lo:lo + self.batch_size]
        self.index += 1
        if self.index >= self.batch_num:
            self.index = 0
            if self.shuffle:
                self.rng.shuffle(self.order)
        return self.part[idx]

    def get_cur_shape(self):
        lo = self.index * self.batch_size
        return (len(self.order[lo:lo + self.batch_size]),) + tuple(self.part.shape[1:])


class DataloaderOp:
    """graph placeholder + the loaders that can feed it"""

    def __init__(self, dataloaders: Sequence[Dataloader], dtype="float32"):
        from .executor import _g
        _g()
        self.dataloaders: Dict[str, Dataloader] = {d.name: d for d in dataloaders}
        first = dataloaders[0]
        np_int = np.issubdtype(first.raw_data.dtype, np.integer)
        self.node = core.placeholder("int64" if np_int else dtype, list(first.shape), name=f"dataloader_{'_'.join(self.dataloaders)}")
        _REGISTRY[self.node.id] = self

    def get_batch_num(self, name): return self.dataloaders[name].batch_num                       # noqa: E704
    def get_arr(self, name): return self.dataloaders[name].get_arr()                               # noqa: E704
    def set_dp_rank(self, r, n): [d.set_dp_rank(r, n) for d in self.dataloaders.values()]          # noqa: E704


_REGISTRY: Dict[int, DataloaderOp] = {}


def dataloader_op(dataloaders: Sequence, dtype="float32"):
    """-> graph node; accepts Dataloader objects or [raw_data, batch_size, name] triples"""
    dls = [d if isinstance(d, Dataloader) else Dataloader(*d) for d in dataloaders]
    return DataloaderOp(dls, dtype).node


def loaders_of(nodes) -> List[DataloaderOp]:
    """the DataloaderOps among the transitive inputs the executor has to feed (all registered ones of the v1 graph)"""
    return list(_REGISTRY.values())
