"""Pre-tokenised corpora for pre-training: an append-only builder for the flat token file + document index, a JSONL -> indexed
dataset preprocessor (CLI), fixed-length GPT samples cut from the shuffled document stream, weighted blending of several corpora
and train / valid / test splitting.
(ref: examples/hydraulis/data_utils/{indexed_dataset.py MMapIndexedDataset(+Builder), hetuDataset.py, blendedDataset.py,
blendedHetuDatasetBuilder.py, llama_dataset.py}, examples/hetero/data_utils/{create_web_dataset.py, gpt_seq_dataset.py} --
capability parity, own format: `<prefix>.bin` raw little-endian tokens, `<prefix>.idx.npy` int64 document offsets, `<prefix>.meta.json`)"""
from __future__ import annotations

import json
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np

from .dataset import IndexedTokenDataset


class IndexedDatasetBuilder:
    """stream documents into `<prefix>.bin` without holding the corpus in memory"""

    def __init__(self, prefix: str, vocab_size: Optional[int] = None, dtype=None):
        self.prefix = prefix
        self.dtype = np.dtype(dtype) if dtype is not None else np.dtype(np.uint16 if (vocab_size is not None and vocab_size < 65500) else np.int32)
        os.makedirs(os.path.dirname(os.path.abspath(prefix)) or ".", exist_ok=True)
        self._f = open(prefix + ".bin", "wb")
        self._offsets = [0]

    def add_document(self, tokens: Sequence[int]):
        arr = np.asarray(tokens, dtype=self.dtype)
        self._f.write(arr.tobytes(order="C"))
        self._offsets.append(self._offsets[-1] + int(arr.size))

    def merge(self, other_prefix: str):
        """append an already built dataset (shards written by parallel workers)"""
        offs = np.load(other_prefix + ".idx.npy")
        meta = json.load(open(other_prefix + ".meta.json")) if os.path.exists(other_prefix + ".meta.json") else {}
        assert np.dtype(meta.get("dtype", str(self.dtype))) == self.dtype, "cannot merge datasets of different token widths"
        with open(other_prefix + ".bin", "rb") as f:
            while True:
                chunk = f.read(64 << 20)
                if not chunk:
                    break
                self._f.write(chunk)
        base = self._offsets[-1]
        self._offsets.extend(int(base + o) for o in offs[1:])

    def finalize(self) -> "IndexedTokenDataset":
        self._f.close()
        np.save(self.prefix + ".idx.npy", np.asarray(self._offsets, dtype=np.int64))
        lens = np.diff(self._offsets)
        json.dump({"dtype": str(self.dtype), "documents": len(lens), "tokens": int(self._offsets[-1]),
                   "length_percentiles": {str(p): int(np.percentile(lens, p)) for p in (50, 90, 99, 100)} if len(lens) else {}},
                  open(self.prefix + ".meta.json", "w"))
        return open_indexed(self.prefix)


def open_indexed(prefix: str) -> IndexedTokenDataset:
    meta = json.load(open(prefix + ".meta.json")) if os.path.exists(prefix + ".meta.json") else {}
    return IndexedTokenDataset(prefix, dtype=np.dtype(meta.get("dtype", "int32")))


def _tokenize_shard(args):
    path, lo, hi, key, tok_spec, out_prefix, eod = args
    from .tokenizers import build_tokenizer
    tok = build_tokenizer(**tok_spec)
    b = IndexedDatasetBuilder(out_prefix, vocab_size=getattr(tok, "vocab_size", None))
    with open(path, "rb") as f:
        f.seek(lo)
        if lo:
            f.readline()                              # the partial line belongs to the previous shard
        while f.tell() <= hi:
            line = f.readline()
            if not line:
                break
            line = line.strip()
            if not line:
                continue
            rec = json.loads(line)
            text = rec[key] if isinstance(rec, dict) else rec
            ids = list(tok.encode(text))
            if eod is not None:
                ids.append(eod)
            if ids:
                b.add_document(ids)
    b.finalize()
    return out_prefix


def preprocess_jsonl(input_path: str, output_prefix: str, tokenizer: Optional[dict] = None, key: str = "text", workers: int = 1,
                     append_eod: bool = True) -> IndexedTokenDataset:
    """tokenise a JSONL corpus (one {"text": ...} per line) into an indexed dataset; `workers` processes split the file by byte
    ranges and their shards are merged in order"""
    from .tokenizers import build_tokenizer
    tok_spec = dict(tokenizer or {"tokenizer_type": "byte"})
    tok = build_tokenizer(**tok_spec)
    eod = getattr(tok, "eos_id", None) if append_eod else None
    size = os.path.getsize(input_path)
    workers = max(1, min(int(workers), max(size // (1 << 16), 1)))
    bounds = [size * i // workers for i in range(workers + 1)]
    jobs = [(input_path, bounds[i], bounds[i + 1] - 1 if i + 1 < workers else size, key, tok_spec, f"{output_prefix}.shard{i}", eod) for i in range(workers)]
    if workers == 1:
        shards = [_tokenize_shard(jobs[0])]
    else:
        import multiprocessing as mp
        with mp.get_context("spawn").Pool(workers) as pool:
            shards = pool.map(_tokenize_shard, jobs)
    out = IndexedDatasetBuilder(output_prefix, vocab_size=getattr(tok, "vocab_size", None))
    for s in shards:
        out.merge(s)
        for ext in (".bin", ".idx.npy", ".meta.json"):
            os.remove(s + ext)
    return out.finalize()


class GPTSampleDataset:
    """fixed-length language-model samples from a document corpus: every epoch the documents are shuffled and concatenated, the
    stream is cut into `seq_length + 1` token windows (a window may span documents; the extra token is the label of the last
    position), and the windows are visited in a shuffled order.  Sample i is a pure function of (seed, i), so a resumed job asks
    for the next index and gets the same data."""

    def __init__(self, indexed: IndexedTokenDataset, seq_length: int, num_samples: Optional[int] = None, seed: int = 1234, documents=None):
        self.ds, self.seq, self.seed = indexed, int(seq_length), int(seed)
        self.docs = np.asarray(documents if documents is not None else np.arange(len(indexed)), dtype=np.int64)
        lens = np.asarray([indexed.offsets[d + 1] - indexed.offsets[d] for d in self.docs], dtype=np.int64)
        self.tokens_per_epoch = int(lens.sum())
        assert self.tokens_per_epoch > self.seq, "corpus shorter than one sample"
        self.samples_per_epoch = (self.tokens_per_epoch - 1) // self.seq
        self.num_samples = int(num_samples) if num_samples is not None else self.samples_per_epoch
        self._lens = lens
        self._epoch_cache = {}

    def __len__(self):
        return self.num_samples

    def _epoch(self, e: int):
        if e not in self._epoch_cache:
            rng = np.random.RandomState((self.seed + 7919 * e) % (2 ** 31))
            order = rng.permutation(len(self.docs))
            starts = np.concatenate([[0], np.cumsum(self._lens[order])])       # token offset of every document in the epoch stream
            sample_order = rng.permutation(self.samples_per_epoch)
            self._epoch_cache = {e: (order, starts, sample_order)}             # one epoch resident
        return self._epoch_cache[e]

    def __getitem__(self, i: int) -> np.ndarray:
        if i < 0 or i >= self.num_samples:
            raise IndexError(i)
        e, j = divmod(i, self.samples_per_epoch)
        order, starts, sample_order = self._epoch(e)
        lo = int(sample_order[j]) * self.seq
        hi = lo + self.seq + 1
        d = int(np.searchsorted(starts, lo, side="right") - 1)
        out = np.empty(self.seq + 1, dtype=np.int64)
        filled = 0
        while filled < self.seq + 1:
            doc = self.ds[int(self.docs[order[d]])]
            a = lo + filled - int(starts[d])
            take = min(len(doc) - a, self.seq + 1 - filled)
            out[filled:filled + take] = doc[a:a + take]
            filled += take
            d += 1
        return out


def blending_indices(weights: Sequence[float], size: int) -> Tuple[np.ndarray, np.ndarray]:
    """deterministic weighted interleave: at every position pick the dataset whose realised share lags its target the most
    -> (dataset index per position, index inside that dataset per position)"""
    w = np.asarray(weights, dtype=np.float64)
    w = w / w.sum()
    which = np.empty(size, dtype=np.int64)
    inner = np.empty(size, dtype=np.int64)
    counts = np.zeros(len(w), dtype=np.int64)
    for i in range(size):
        k = int(np.argmax(w * (i + 1) - counts))
        which[i], inner[i] = k, counts[k]
        counts[k] += 1
    return which, inner


class BlendedDataset:
    """several corpora mixed by weight (component datasets wrap around when a weight asks for more samples than they hold)"""

    def __init__(self, datasets: Sequence, weights: Sequence[float], size: int):
        assert len(datasets) == len(weights) and len(datasets) > 0
        self.datasets, self.size = list(datasets), int(size)
        self.which, self.inner = blending_indices(weights, self.size)

    def __len__(self):
        return self.size

    def __getitem__(self, i):
        d = self.datasets[int(self.which[i])]
        return d[int(self.inner[i]) % len(d)]


def split_documents(num_documents: int, split: str = "969,30,1") -> List[np.ndarray]:
    """'train,valid,test' proportions -> contiguous document index ranges (every document in exactly one split)"""
    parts = [float(x) for x in split.replace("/", ",").split(",")]
    parts += [0.0] * (3 - len(parts))
    total = sum(parts)
    bounds = [0]
    for p in parts:
        bounds.append(bounds[-1] + int(round(p / total * num_documents)))
    bounds[-1] = num_documents
    return [np.arange(bounds[i], max(bounds[i], bounds[i + 1])) for i in range(3)]


def main(argv=None):
    import argparse
    ap = argparse.ArgumentParser(description="tokenise a JSONL corpus into <prefix>.bin / .idx.npy / .meta.json")
    ap.add_argument("--input", required=True)
    ap.add_argument("--output-prefix", required=True)
    ap.add_argument("--json-key", default="text")
    ap.add_argument("--tokenizer-type", default="byte")
    ap.add_argument("--vocab-file", default=None)
    ap.add_argument("--merge-file", default=None)
    ap.add_argument("--tokenizer-path", default=None)
    ap.add_argument("--workers", type=int, default=1)
    ap.add_argument("--no-append-eod", action="store_true")
    a = ap.parse_args(argv)
    spec = {"tokenizer_type": a.tokenizer_type}
    for k, v in (("vocab_file", a.vocab_file), ("merge_file", a.merge_file), ("name_or_path", a.tokenizer_path)):
        if v:
            spec[k] = v
    ds = preprocess_jsonl(a.input, a.output_prefix, spec, a.json_key, a.workers, not a.no_append_eod)
    meta = json.load(open(a.output_prefix + ".meta.json"))
    print(f"{a.output_prefix}: {len(ds)} documents, {meta['tokens']} tokens ({meta['dtype']}), length percentiles {meta['length_percentiles']}")


if __name__ == "__main__":
    main()
