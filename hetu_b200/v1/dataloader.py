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


class BatchIndices:
    """the order in which batches are visited, shared by loaders that must stay aligned (features and labels): reshuffled every
    time batch 0 is requested after a full pass (ref: dataloader.py:10)"""

    def __init__(self, batch_num: int, need_shuffle: bool = False, seed: int = 0):
        self.batch_num, self.need_shuffle = int(batch_num), bool(need_shuffle)
        self.all_batch_indices = np.arange(self.batch_num)
        self.last_key = self.batch_num - 1
        self.rng = np.random.RandomState(seed)

    def shuffle(self):
        self.rng.shuffle(self.all_batch_indices)

    def assert_attr(self, dataloader):
        assert self.batch_num == dataloader.batch_num and self.need_shuffle == dataloader.shuffle

    def __getitem__(self, key):
        if key == 0 and self.last_key != key:
            assert self.last_key == self.batch_num - 1, "a pass must finish before the next one starts"
            if self.need_shuffle:
                self.shuffle()
        self.last_key = key
        return int(self.all_batch_indices[key])


class RawData:
    """one logical array over several chunks (arrays / memmaps with equal trailing dims): rows are addressed globally and gathered
    from the chunk that holds them, so a dataset larger than memory can stay memory-mapped (ref: dataloader.py:34)"""

    def __init__(self, raw_data, dtype=np.float32, func=None):
        self.dtype, self.func = dtype, (func or (lambda x: x))
        chunks = raw_data if isinstance(raw_data, (list, tuple)) else [raw_data]
        self.raw_data = [self._init_array(c) for c in chunks]
        self._shape = list(self.raw_data[0].shape)
        self._offsets = [0, self._shape[0]]
        for d in self.raw_data[1:]:
            assert list(d.shape[1:]) == self._shape[1:], "chunks must agree on every dimension but the first"
            self._shape[0] += d.shape[0]
            self._offsets.append(self._shape[0])

    def _init_array(self, a):
        a = self.func(a)
        if not isinstance(a, (np.memmap, np.ndarray)):
            return np.array(a, dtype=self.dtype)
        return a if a.dtype == self.dtype else a.astype(self.dtype)

    shape = property(lambda self: tuple(self._shape))

    def __len__(self):
        return self._shape[0]

    def __getitem__(self, key):
        if isinstance(key, (int, np.integer)):
            c = int(np.searchsorted(self._offsets, key, side="right")) - 1
            return self.raw_data[c][key - self._offsets[c]]
        if isinstance(key, slice):
            key = np.arange(*key.indices(len(self)))
        key = np.asarray(key)
        out = np.empty((key.size,) + tuple(self._shape[1:]), self.dtype)
        which = np.searchsorted(self._offsets, key, side="right") - 1
        for c in np.unique(which):
            m = which == c
            out[m] = self.raw_data[c][key[m] - self._offsets[c]]
        return out


class GNNDataLoaderOp:
    """double-buffered graph feeder: `GNNDataLoaderOp.step(next_graph)` rotates (current <- next); every instance turns the current
    graph into its array through `handler` (features, labels, adjacency ...) (ref: dataloader.py:253)"""
    graph = None
    nxt_graph = None

    def __init__(self, handler, ctx=None, shape=None, dtype="float32", name="GNNDataloaderOp"):
        from .executor import _g
        _g()
        self.handler, self.name = handler, name
        self.node = core.placeholder(dtype, list(shape) if shape is not None else [1], name=name)
        self.dataloaders = {}
        _GNN_REGISTRY[self.node.id] = self

    desc = property(lambda self: self.name)
    def get_batch_num(self, name): return None                                  # noqa: E704
    def get_arr(self, name=None): return self.handler(type(self).graph)         # noqa: E704
    def get_next_arr(self, name=None): return self.handler(type(self).nxt_graph)   # noqa: E704

    @classmethod
    def step(cls, graph):
        cls.graph, cls.nxt_graph = cls.nxt_graph, graph


_GNN_REGISTRY: Dict[int, GNNDataLoaderOp] = {}
