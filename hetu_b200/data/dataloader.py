"""DataLoader with SAMPLE / TOKEN load levels, data-parallel sharding, resume from consumed samples, background
prefetch into pinned host memory (ref: python/hetu/data/dataloader.py, hetu/graph/data/dataloader.{h,cc},
utils/parallel/data.py parallel_data_provider)."""
from __future__ import annotations

import queue
import threading
from typing import Iterator, List

import numpy as np
import torch


def parallel_data_provider(global_data, ds, device_index: int):
    """slice a global array according to a DistributedStates (what rank `device_index` feeds)"""
    arr = np.asarray(global_data)
    begin, size = ds.local_slice(list(arr.shape), device_index)
    sl = tuple(slice(b, b + s) for b, s in zip(begin, size))
    return arr[sl]


class DataLoader:
    def __init__(self, dataset, global_batch_size: int = 0, global_token_num: int = 0, load_level: str = "SAMPLE", shuffle: bool = False,
                 seed: int = 0, dp_rank: int = 0, dp_size: int = 1, drop_last: bool = True, prefetch: int = 2, collate_fn=None):
        assert load_level in ("SAMPLE", "TOKEN")
        self.ds, self.gbs, self.gtn, self.level = dataset, global_batch_size, global_token_num, load_level
        self.shuffle, self.seed, self.dp_rank, self.dp_size, self.drop_last = shuffle, seed, dp_rank, dp_size, drop_last
        self.prefetch, self.collate_fn = prefetch, collate_fn
        self.consumed = 0
        self.epoch = 0
        # GLOBAL number of samples (all data-parallel ranks together) of this epoch that the batches handed to the
        # consumer so far cover -- identical on every rank whatever its dp slice, and not ahead of the consumer when the
        # prefetch thread is; this is the value `restart()` expects back after a failure / re-plan
        self.consumed_yielded = 0

    def restart(self, consumed_samples: int):
        """resume after a failure / elastic re-plan: skip what the previous incarnation already consumed"""
        self.epoch, self.consumed = divmod(consumed_samples, max(len(self.ds), 1))
        self.consumed_yielded = self.consumed

    def _order(self):
        idx = np.arange(len(self.ds))
        if self.shuffle:
            np.random.RandomState(self.seed + self.epoch).shuffle(idx)
        return idx

    def _batches(self) -> Iterator[List[int]]:
        idx = self._order()
        pos = self.consumed
        while pos < len(idx):
            if self.level == "SAMPLE":
                b = idx[pos:pos + self.gbs]
                if len(b) < self.gbs and self.drop_last:
                    break
            else:  # TOKEN: fill up to a token budget
                b, tok = [], 0
                while pos + len(b) < len(idx):
                    n = len(self.ds[int(idx[pos + len(b)])])
                    if b and tok + n > self.gtn:
                        break
                    b.append(idx[pos + len(b)])
                    tok += n
                b = np.asarray(b)
                if len(b) == 0:
                    break
            pos += len(b)
            self.consumed = pos
            yield [int(i) for i in b[self.dp_rank::self.dp_size]], pos
        self.epoch += 1
        self.consumed = 0

    def __iter__(self):
        def produce(q):
            for ids, pos in self._batches():
                samples = [self.ds[i] for i in ids]
                q.put((self.collate_fn(samples) if self.collate_fn else samples, pos))
            q.put(None)
        if self.prefetch <= 0:
            for ids, pos in self._batches():
                samples = [self.ds[i] for i in ids]
                self.consumed_yielded = pos
                yield self.collate_fn(samples) if self.collate_fn else samples
            self.consumed_yielded = 0
            return
        q: "queue.Queue" = queue.Queue(maxsize=self.prefetch)
        t = threading.Thread(target=produce, args=(q,), daemon=True)
        t.start()
        while True:
            item = q.get()
            if item is None:
                break
            self.consumed_yielded = item[1]
            yield item[0]
        self.consumed_yielded = 0


def build_data_loader(dataset, consumed_samples: int, global_batch_size: int = 0, global_token_num: int = 0,
                      load_level: str = "SAMPLE", **kw) -> DataLoader:
    dl = DataLoader(dataset, global_batch_size, global_token_num, load_level, **kw)
    dl.restart(consumed_samples)
    return dl


def pinned(t: np.ndarray) -> torch.Tensor:
    x = torch.as_tensor(t)
    return x.pin_memory() if torch.cuda.is_available() else x
