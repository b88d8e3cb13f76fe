"""Training harness for the embedding-compression methods: one entry point that sizes a method for a target compression rate,
trains a CTR model with it on a Criteo-shaped stream, runs the method's schedule (mask refresh, supernet sampling + search,
search -> retrain stages, online hot-id discovery) and reports AUC / log-loss / achieved compression.

    python -m hetu_b200.tools.emb_compress.trainer --method deeplight --compress-rate 0.1 --model deepfm

(ref: tools/EmbeddingMemoryCompression/run_compressed.py and methods/scheduler/{base,compressor,multistage,hash,compo,
tensortrain,dhe,robe,dpq,mgqe,md,autodim,optembed,pep,deeplight,alpt,adapt}.py -- capability parity, own design: the
reference drives v1 executors with one scheduler class per method; here one `CompressionTrainer` owns the graph and a small
per-method `Schedule` object contributes the hooks)
"""
from __future__ import annotations

import argparse
import json
import math
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Tuple

import numpy as np
import torch

from ... import core
from ...models.ctr import DCN, WDL, DeepFM
from ...optim import AdamOptimizer
from ...v1.metrics import auc as _auc
from .methods import AutoDimEmbedding, CafeEmbedding, DeepLightEmbedding, OptEmbedEmbedding, PrunedEmbedding, build_compressed_embedding


# ----------------------------------------------------------------------------------------------------------------- data
class SyntheticCTR:
    """Criteo-shaped stream with a planted signal: Zipf-distributed categorical ids per field, the label is a logistic function
    of per-id latent weights plus a linear dense term, so compressing the table measurably costs AUC."""

    def __init__(self, num_embeddings=20000, num_fields=8, num_dense=4, seed=0, zipf=1.2):
        self.N, self.F, self.D = num_embeddings, num_fields, num_dense
        r = np.random.RandomState(seed)
        self.id_weight = r.randn(num_embeddings).astype(np.float32) * 1.2
        self.dense_weight = r.randn(num_dense).astype(np.float32) * 0.5
        self.zipf = zipf
        self.field_offset = (np.arange(num_fields) * (num_embeddings // num_fields)).astype(np.int64)
        self.field_size = num_embeddings // num_fields

    def batch(self, batch_size: int, rng: np.random.RandomState) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        dense = rng.randn(batch_size, self.D).astype(np.float32)
        local = (rng.zipf(self.zipf, (batch_size, self.F)) - 1) % self.field_size
        sparse = (local + self.field_offset[None, :]).astype(np.int64)
        logit = self.id_weight[sparse].sum(1) / math.sqrt(self.F) + dense @ self.dense_weight
        label = (rng.rand(batch_size) < 1.0 / (1.0 + np.exp(-logit))).astype(np.float32).reshape(-1, 1)
        return dense, sparse, label

    def frequency(self, samples: int = 50000, seed: int = 1) -> np.ndarray:
        rng = np.random.RandomState(seed)
        _, sparse, _ = self.batch(samples // self.F + 1, rng)
        return np.bincount(sparse.reshape(-1), minlength=self.N).astype(np.float64)


# ------------------------------------------------------------------------------------------------------ sizing for a budget
def plan_for_rate(method: str, num_embeddings: int, dim: int, rate: float) -> Dict:
    """constructor arguments that bring `method` to (about) `rate` x the parameters of the full table
    (ref: methods/scheduler/*.py `_get_*` sizing helpers driven by --compress_rate)"""
    budget = max(int(num_embeddings * dim * rate), dim)
    m = method.lower()
    if m == "hash":
        return {"buckets": max(budget // dim, 1)}
    if m in ("compo", "qr"):
        return {}                                                  # two ~sqrt(N) tables: the rate is fixed by N
    if m == "tt":
        rank = max(2, int(math.sqrt(budget / (3.0 * max(num_embeddings ** (1 / 3), 1) * max(dim ** (1 / 3), 1)))))
        return {"rank": min(rank, 64)}
    if m == "dhe":
        hidden = max(16, int(math.sqrt(budget / 3)))
        return {"num_hashes": max(16, min(1024, hidden)), "hidden": (hidden, hidden)}
    if m == "robe":
        return {"array_size": budget}
    if m in ("dpq", "mgqe"):
        return {"subspaces": 4 if dim % 4 == 0 else 1, "codes": int(min(256, max(4, budget // dim)))}
    if m == "mde":
        hot = max(1, int(num_embeddings * rate * 0.5))
        return {"block_bounds": [hot, min(num_embeddings, hot * 8), num_embeddings], "alpha": 0.45}
    if m == "autodim":
        return {"candidates": sorted({max(2, dim // 8), max(2, dim // 4), max(2, dim // 2), dim})}
    if m == "deeplight":
        return {"target_sparsity": 1.0 - rate}
    if m == "alpt":
        return {"bits": 8 if rate >= 0.25 else 4}
    if m in ("adapt", "cafe"):
        hot = max(1, int(budget * 0.7) // dim)
        return {"hot": hot, "buckets": max(1, (budget - hot * dim) // dim)}
    return {}                                                      # pep / optembed learn their own sparsity


# ------------------------------------------------------------------------------------------------------------- schedules
@dataclass
class Schedule:
    """per-method hooks around the common loop"""
    stages: int = 1
    feed_ids: Optional[Callable[[np.ndarray], np.ndarray]] = None            # id remapping before the lookup (cafe)
    before_step: Optional[Callable[["CompressionTrainer", int], None]] = None
    after_step: Optional[Callable[["CompressionTrainer", int, Dict], None]] = None
    between_stages: Optional[Callable[["CompressionTrainer"], Dict]] = None  # -> constructor kwargs of the stage-2 embedding
    stage2_method: Optional[str] = None
    report: Dict = field(default_factory=dict)


def _deeplight_schedule(every: int = 10) -> Schedule:
    def after(tr, step, _out):
        if step % every == 0:
            emb: DeepLightEmbedding = tr.embedding
            w = tr.value_of(emb.weight)
            mask = emb.update_mask(w, step)
            tr.assign(emb.mask, mask)
            tr.assign(emb.weight, w * mask)                       # pruned weights stay zero between refreshes
            tr.schedule.report["sparsity"] = float(1.0 - mask.mean())
    return Schedule(after_step=after)


def _optembed_schedule() -> Schedule:
    rng = np.random.RandomState(0)

    def before(tr, step):                                            # supernet training: a random dimension mask per step
        emb: OptEmbedEmbedding = tr.embedding
        keep = int(rng.randint(max(1, emb.dim // 4), emb.dim + 1))
        tr.assign(tr.dim_mask_var, np.concatenate([np.ones(keep), np.zeros(emb.dim - keep)]).astype(np.float32).reshape(1, emb.dim))

    def search(tr) -> Dict:                                          # evolutionary search reduced to its essence: score candidates
        emb: OptEmbedEmbedding = tr.embedding
        best, scores = None, {}
        for keep in sorted({emb.dim, emb.dim * 3 // 4, emb.dim // 2, max(1, emb.dim // 4)}, reverse=True):
            tr.assign(tr.dim_mask_var, np.concatenate([np.ones(keep), np.zeros(emb.dim - keep)]).astype(np.float32).reshape(1, emb.dim))
            m = tr.evaluate(batches=4)
            scores[keep] = m["auc"]
            if best is None or m["auc"] >= scores[best] - 0.003:    # the smallest width within 0.003 AUC of the best seen
                best = keep
        w = tr.value_of(emb.weight)
        thr = float(tr.value_of(emb.threshold).reshape(-1)[0])
        rows_kept = float((np.abs(w).sum(1) > thr).mean())
        tr.schedule.report.update({"searched_dim": best, "candidate_auc": scores, "rows_kept": rows_kept})
        tr.assign(tr.dim_mask_var, np.concatenate([np.ones(best), np.zeros(emb.dim - best)]).astype(np.float32).reshape(1, emb.dim))
        return {}
    s = Schedule(before_step=before)
    s.between_stages = search
    return s


def _autodim_schedule() -> Schedule:
    def pick(tr) -> Dict:
        emb: AutoDimEmbedding = tr.embedding
        d = emb.selected_dim(tr.value_of(emb.alpha))
        tr.schedule.report["selected_dim"] = int(d)
        n = emb.num_embeddings
        return {"block_bounds": [n], "alpha": 0.0, "_dim_override": int(d)}
    return Schedule(stages=2, between_stages=pick, stage2_method="mde")


def _pep_schedule() -> Schedule:
    def derive(tr) -> Dict:
        emb: PrunedEmbedding = tr.embedding
        w, s = tr.value_of(emb.weight), tr.value_of(emb.s)
        mask = (np.abs(w) > 1.0 / (1.0 + np.exp(-s))).astype(np.float32)
        tr.schedule.report["sparsity"] = float(1.0 - mask.mean())
        tr.saved_mask = mask
        return {"target_sparsity": float(1.0 - mask.mean())}

    def freeze(tr, step, _out):                                      # stage 2: retrain the surviving weights only
        if tr.stage == 1 and step == 0 and getattr(tr, "saved_mask", None) is not None:
            tr.assign(tr.embedding.mask, tr.saved_mask)
    s = Schedule(stages=2, between_stages=derive, stage2_method="deeplight")
    s.after_step = freeze
    return s


_POST_TRAINING = ("dedup", "sparse", "quantize")     # compress a table that was trained at full size (methods/scheduler/{compressor,switchinference}.py)


def _post_training_schedule(method: str, rate: float, kwargs: Dict) -> Schedule:
    def compress(tr) -> Dict:
        table = tr.value_of(tr.model.embedding.weight)
        if method == "sparse":                                       # magnitude pruning down to the budget of a CSR table (2 words / entry)
            keep = max(int((table.size * rate - table.shape[0] - 1) / 2.0), 1)      # values + column ids, after the row pointers
            thr = np.sort(np.abs(table).reshape(-1))[-keep]
            tr.schedule.report["sparsity"] = float((np.abs(table) < thr).mean())
            return {"table": table * (np.abs(table) >= thr)}
        if method == "dedup":                                        # widen the grid until the stored blocks fit the budget
            from .methods import DedupEmbedding
            block = int(kwargs.get("nemb_per_block", 4))
            nblocks = -(-table.shape[0] // block)
            pad = np.zeros((nblocks * block, table.shape[1]), np.float32)
            pad[:table.shape[0]] = table
            blocks = pad.reshape(nblocks, -1)
            lo, hi = 0.0, float(np.abs(table).max()) + 1e-6
            for _ in range(20):
                mid = 0.5 * (lo + hi)
                if len(DedupEmbedding.group_blocks(blocks, mid)[1]) > rate * nblocks:
                    lo = mid
                else:
                    hi = mid
            tr.schedule.report["tolerance"] = hi
            return {"table": table, "nemb_per_block": block, "tolerance": hi}
        tr.schedule.report["digit"] = int(kwargs.get("digit", 8))
        span = float(np.abs(table).max()) or 1.0
        tr.trained_table = table
        return {"digit": int(kwargs.get("digit", 8)), "scale": 2.0 * span / (2 ** int(kwargs.get("digit", 8)) - 1), "middle": 0.0}
    s = Schedule(stages=2, between_stages=compress, stage2_method=method)
    return s


def _autosrh_schedule(rate: float) -> Schedule:
    def prune(tr) -> Dict:
        emb = tr.embedding
        alpha = tr.value_of(emb.alpha)
        tr.schedule.report["alpha_abs_mean"] = float(np.abs(alpha).mean())
        return {"nsplit": emb.nsplit, "group_indices": emb.group_np, "frozen_alpha": alpha, "keep_rate": rate}
    return Schedule(stages=2, between_stages=prune, stage2_method="autosrh")


def _cafe_schedule() -> Schedule:
    def after(tr, step, out):
        emb: CafeEmbedding = tr.embedding
        ids = out["raw_ids"]
        emb.observe(ids, np.ones_like(ids, dtype=np.float32))        # importance = frequency (gradient norms need the eager path)
    s = Schedule(after_step=after)
    return s


def _adapt_schedule(data: SyntheticCTR, hot: int) -> Schedule:
    order = np.argsort(-data.frequency())
    rank_of = np.empty(data.N, np.int64)
    rank_of[order] = np.arange(data.N)                               # frequency rank: AdaptEmb's hot ids are the `hot` most frequent
    return Schedule(feed_ids=lambda ids: rank_of[ids])


# ---------------------------------------------------------------------------------------------------------------- trainer
class CompressionTrainer:
    def __init__(self, method: Optional[str], model: str = "deepfm", num_embeddings: int = 20000, dim: int = 16, num_fields: int = 8,
                 num_dense: int = 4, batch_size: int = 256, lr: float = 0.01, compress_rate: float = 0.1, seed: int = 0,
                 method_kwargs: Optional[Dict] = None):
        self.method = method.lower() if method else None
        self.model_name, self.N, self.dim, self.F, self.ND, self.B, self.lr = model, num_embeddings, dim, num_fields, num_dense, batch_size, lr
        self.data = SyntheticCTR(num_embeddings, num_fields, num_dense, seed)
        self.rate = compress_rate
        self.kwargs = dict(plan_for_rate(self.method, num_embeddings, dim, compress_rate)) if self.method else {}
        self.kwargs.update(method_kwargs or {})
        self.schedule = self._schedule_for(self.method)
        self.stage = 0
        self.history: List[Dict] = []
        self._build(self.method, self.kwargs)

    def _schedule_for(self, m: Optional[str]) -> Schedule:
        if m == "deeplight":
            return _deeplight_schedule()
        if m == "optembed":
            return _optembed_schedule()
        if m == "autodim":
            return _autodim_schedule()
        if m == "pep":
            return _pep_schedule()
        if m == "cafe":
            s = _cafe_schedule()
            return s
        if m == "adapt":
            return _adapt_schedule(self.data, self.kwargs.get("hot", 1))
        if m == "autosrh":
            return _autosrh_schedule(float(self.kwargs.pop("keep_rate", getattr(self, "rate", None) or 0.5)))
        if m in _POST_TRAINING:
            return _post_training_schedule(m, float(getattr(self, "rate", None) or 0.5), self.kwargs)
        return Schedule()

    # -- graph ----------------------------------------------------------------------------------------------------
    def _build(self, method: Optional[str], kwargs: Dict):
        kw = dict(kwargs)
        dim = int(kw.pop("_dim_override", self.dim))
        if self.stage == 0 and method in _POST_TRAINING:             # these compress a trained table: stage 1 trains the full one
            method, kw = None, {}
        with core.graph("define_and_run", create_new=True) as g:
            self.embedding = build_compressed_embedding(method, self.N, dim, **kw) if method else None
            if method == "optembed":
                self.dim_mask_var = core.parallel_parameter(core.ones_initializer(), [1, dim], None, requires_grad=False, name="optembed_dim_mask")
                self.embedding.dim_mask = self.dim_mask_var
            cls = {"wdl": WDL, "deepfm": DeepFM, "dcn": DCN}[self.model_name]
            self.model = cls(self.N, dim, num_fields=self.F, num_dense=self.ND, embedding=self.embedding)
            self.dense = core.placeholder("float32", [self.B, self.ND], name="dense")
            self.sparse = core.placeholder("int64", [self.B, self.F], name="sparse")
            self.label = core.placeholder("float32", [self.B, 1], name="label")
            self.loss, self.logit = self.model(self.dense, self.sparse, self.label)
            self.train_op = AdamOptimizer(lr=self.lr).minimize(self.loss)
        self.graph = g
        if method == "cafe":
            self.schedule.feed_ids = self.embedding.remap

    def value_of(self, var) -> np.ndarray:
        return np.asarray(self.graph.get_param(var).float().cpu().numpy())

    def assign(self, var, value: np.ndarray):
        self.graph.set_param(var, torch.as_tensor(np.asarray(value, np.float32)))

    # -- loop -----------------------------------------------------------------------------------------------------
    def _feed(self, rng) -> Tuple[Dict, np.ndarray, np.ndarray]:
        dense, sparse, label = self.data.batch(self.B, rng)
        ids = self.schedule.feed_ids(sparse) if self.schedule.feed_ids else sparse
        return {self.dense: torch.as_tensor(dense), self.sparse: torch.as_tensor(ids), self.label: torch.as_tensor(label)}, sparse, label

    def train(self, steps: int, seed: int = 100, log_every: int = 0) -> List[float]:
        rng = np.random.RandomState(seed + 17 * self.stage)
        losses = []
        for step in range(steps):
            if self.schedule.before_step:
                self.schedule.before_step(self, step)
            feed, raw_ids, _ = self._feed(rng)
            out = self.graph.run(self.loss, [self.loss, self.train_op], feed)
            losses.append(float(out[0]))
            if self.schedule.after_step:
                self.schedule.after_step(self, step, {"loss": losses[-1], "raw_ids": raw_ids})
            if log_every and step % log_every == 0:
                print(f"[{self.method or 'full'} stage {self.stage}] step {step} loss {losses[-1]:.4f}", flush=True)
        return losses

    def evaluate(self, batches: int = 8, seed: int = 999) -> Dict:
        rng = np.random.RandomState(seed)
        ys, ps, ls = [], [], []
        for _ in range(batches):
            feed, _, label = self._feed(rng)
            out = self.graph.run(self.loss, [self.loss, self.logit], feed)
            ls.append(float(out[0]))
            ps.append(1.0 / (1.0 + np.exp(-np.asarray(out[1].float().cpu().numpy()).reshape(-1))))
            ys.append(label.reshape(-1))
        y, p = np.concatenate(ys), np.concatenate(ps)
        return {"auc": float(_auc(y, p)), "logloss": float(np.mean(ls))}

    def compression(self) -> Dict:
        full = self.N * self.dim
        if self.embedding is None:
            return {"parameters": full, "ratio": 1.0, "effective_ratio": 1.0}
        params = self.embedding.num_parameters()
        eff = float(params)
        rep = self.schedule.report
        if "sparsity" in rep:                                       # pruned tables are stored sparse: count the survivors
            eff = full * (1.0 - rep["sparsity"])
        if self.method == "optembed" and "searched_dim" in rep:
            eff = full * rep.get("rows_kept", 1.0) * rep["searched_dim"] / self.dim
        if self.method == "alpt":
            eff = full * self.kwargs.get("bits", 8) / 32.0 + self.N
        return {"parameters": int(params), "ratio": full / max(params, 1), "effective_ratio": full / max(eff, 1.0)}

    def run(self, steps: int = 200, eval_batches: int = 8, log_every: int = 0) -> Dict:
        """the whole schedule of the method: stage 1 (train / search), optional hand-over, stage 2 (retrain)"""
        sched = self.schedule
        losses = self.train(steps, log_every=log_every)
        result = {"method": self.method or "full", "model": self.model_name, "stage1_loss": [losses[0], losses[-1]]}
        if sched.between_stages is not None:
            kw2 = sched.between_stages(self)
            if sched.stages == 2:
                result["stage1"] = {**self.evaluate(eval_batches), **self.compression()}
                self.stage = 1
                carried = sched
                self._build(sched.stage2_method, kw2)
                self.method_stage2 = sched.stage2_method
                table = getattr(self, "trained_table", None)         # post-training quantisation starts from the trained values
                if table is not None and hasattr(self.embedding, "weight") and tuple(self.embedding.weight.shape) == tuple(table.shape):
                    self.assign(self.embedding.weight, table)
                self.schedule = _deeplight_schedule(every=10 ** 9) if sched.stage2_method == "deeplight" else Schedule()
                self.schedule.report = carried.report                # keep what stage 1 found
                if carried.after_step is not None and sched.stage2_method == "deeplight":
                    inner = carried.after_step
                    self.schedule.after_step = inner                 # installs the stage-1 mask at step 0 of stage 2
                losses2 = self.train(steps, log_every=log_every)
                result["stage2_loss"] = [losses2[0], losses2[-1]]
        result.update(self.evaluate(eval_batches))
        result.update(self.compression())
        result["schedule"] = {k: v for k, v in self.schedule.report.items()}
        self.history.append(result)
        return result


def main(argv=None):
    ap = argparse.ArgumentParser(description="train a CTR model with a compressed embedding table")
    ap.add_argument("--method", default=None, help="hash compo tt dhe robe dpq mgqe mde autodim pep deeplight optembed alpt adapt cafe (default: full table)")
    ap.add_argument("--model", default="deepfm", choices=["wdl", "deepfm", "dcn"])
    ap.add_argument("--compress-rate", type=float, default=0.1)
    ap.add_argument("--num-embeddings", type=int, default=20000)
    ap.add_argument("--dim", type=int, default=16)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--batch-size", type=int, default=256)
    ap.add_argument("--lr", type=float, default=0.01)
    ap.add_argument("--log-every", type=int, default=50)
    a = ap.parse_args(argv)
    tr = CompressionTrainer(a.method, a.model, a.num_embeddings, a.dim, batch_size=a.batch_size, lr=a.lr, compress_rate=a.compress_rate)
    print(json.dumps(tr.run(a.steps, log_every=a.log_every)))


if __name__ == "__main__":
    main()
