"""Trainer: builds the define-and-run graph for a model under one or several parallel strategies, feeds it padded or
packed micro-batches, drives the optimizer, logging, profiling, checkpoints and hot strategy switching.
(ref: python/hetu/engine/trainer.py:66-828 -- Trainer.build/create_define_graph/prepare_feed_dict/train/save_model)
"""
from __future__ import annotations

import json
import os
import time
from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence

import numpy as np
import torch

from .. import core, distributed
from ..core import IntSymbol
from ..data import Bucket, build_data_loader, build_tokenizer
from ..utils.parallel import StrategyConfig, convert_strategy, parse_multi_ds_parallel_config, read_ds_parallel_config
from .data_collator import DataCollatorForLanguageModel
from .trainer_config import DataLoadLevel, TrainingConfig
from .wrapper import ModelWrapperFromConfig, OptimizerWrapper


@dataclass
class TrainerStates:
    """everything `build()` creates (ref: trainer.py:27 TrainerStates)"""
    graph: object = None
    model: object = None
    optimizer: object = None
    input_ids: object = None
    position_ids: object = None
    labels: object = None
    loss: object = None
    train_op: object = None
    seq_len_symbol: object = None
    cu_seqlens: object = None
    config: object = None


def _strategy_sizes(cfg: dict):
    """(dp, tp, pp) of a ds_parallel_config from its input leaf and block ranges"""
    inp = cfg["input"]
    dp = int(inp["split"].get("0", [1])[0]) if inp.get("split") else 1
    tp = int(inp["dup"][0]) if inp.get("dup") else 1
    stages = {tuple(map(tuple, b["layernorm1"]["device_group_union"])) for b in cfg["blocks"].values()}
    return dp, tp, max(len(stages), 1)


def _cp_degree(cfg: dict) -> int:
    """context-parallel degree of a strategy (the batch split of the input leaf counts dp * cp replicas)"""
    return max(int(cfg.get("cp", 1) or 1), 1)


def cp_rows(width: int, cp: int, cp_idx: int) -> np.ndarray:
    """token positions of one sequence owned by ring member `cp_idx`: chunk i and its mirror 2cp-1-i of 2cp equal chunks
    (the SYM split that balances causal attention work; ref: generate_cp_pack_data / ParallelAttention SYM pattern)"""
    chunks = np.split(np.arange(width), 2 * cp)
    return np.concatenate([chunks[cp_idx], chunks[2 * cp - 1 - cp_idx]])


class Trainer:
    def __init__(self, pretrain_config: TrainingConfig, model, tokenizer=None, optimizer=None, train_dataset=None,
                 data_collator: Optional[Callable] = None, **kwargs):
        self.pretrain_config = cfg = pretrain_config
        if cfg.packing and cfg.micro_batch_size:
            raise ValueError("micro_batch_size is only valid when packing is off")
        self.model_wrapper = model
        self.optimizer_wrapper = optimizer if not isinstance(optimizer, dict) else OptimizerWrapper(optimizer)
        if self.optimizer_wrapper is None:
            self.optimizer_wrapper = OptimizerWrapper({"type": "adam", "lr": cfg.learning_rate, "weight_decay": cfg.weight_decay,
                                                       "lr_warmup_steps": cfg.warmup_steps, "lr_decay_steps": cfg.steps,
                                                       "lr_decay_style": cfg.lr_decay_style, "min_lr": cfg.min_lr})
        if isinstance(tokenizer, dict):
            tokenizer = build_tokenizer(tokenizer.pop("type", "byte"), **tokenizer)
        self.tokenizer = tokenizer
        self.train_dataset = train_dataset
        self.data_collator = data_collator or DataCollatorForLanguageModel(tokenizer)
        self.precision = "bfloat16" if cfg.bf16 else "float32"
        self.epoch, self.consumed_samples, self.global_step = 0, 0, 0
        self.trainer_states: Optional[TrainerStates] = None
        self.is_model_built = False
        self.loss_history: List[float] = []
        self.eval_history: List[dict] = []
        self.eval_dataset = None
        self.step_times: List[float] = []
        self.callbacks: List[Callable] = list(kwargs.get("callbacks", []))
        self.max_docs_per_row = int(kwargs.get("max_docs_per_row", 64))

        dspc = kwargs.get("ds_parallel_configs")
        if dspc is None:
            sc: Optional[StrategyConfig] = cfg.ds_parallel
            if sc is not None and sc.ds_parallel_config_path and sc.ds_parallel_config_name:
                dspc = read_ds_parallel_config(os.path.join(sc.ds_parallel_config_path, sc.ds_parallel_config_name))
            elif sc is not None:
                dspc = [convert_strategy(sc, self._num_layers())]
            else:
                dspc = [convert_strategy(StrategyConfig(dp=max(distributed.world_size(), 1)), self._num_layers())]
        shares = kwargs.get("hetero_shares")
        if shares is None and cfg.ds_parallel is not None and getattr(cfg.ds_parallel, "micro_batch_num_list", None):
            shares = list(cfg.ds_parallel.micro_batch_num_list)
        self._set_strategies(dspc, shares)

    def _set_strategies(self, dspc, hetero_shares=None):
        # heterogeneous strategy (pipelines with different tp / stage counts, e.g. a Malleus plan): this rank trains the
        # homogeneous member-local graph of its own pipeline on its share of every global batch (engine/hetero.py)
        self.hetero, self.idle = None, False
        if len(dspc) == 1 and dspc[0].get("hetero") and "device_group_union" in dspc[0].get("input", {}):
            used = {d for grp in dspc[0]["input"]["device_group_union"] for d in grp} | \
                   {d for b in dspc[0]["blocks"].values() for grp in b["layernorm1"]["device_group_union"] for d in grp}
            if len(dspc[0]["input"]["device_group_union"]) > 1 or len(used) < max(distributed.world_size(), 1):
                from .hetero import HeteroSession
                self.hetero = HeteroSession(dspc[0], shares=hetero_shares)
                self.idle = self.hetero.idle
                dspc = [self.hetero.local_cfg]
        self.ds_parallel_configs = dspc
        self.num_strategy = len(dspc)
        self.cur_strategy_id = 0

    def rebuild(self, ds_parallel_configs, hetero_shares=None, ckpt_dir: Optional[str] = None):
        """Continue training under strategies that hot switching cannot reach (another device layout per pipeline, a
        heterogeneous plan, ...): parameters and optimizer states go through a strategy-independent split checkpoint, the
        graph is rebuilt under the new strategies and the state is re-sharded on load.  Step counters, consumed samples and
        the loss history carry over.  (the in-process form of the reference's re-plan + restart-from-checkpoint path:
        python/hetu/rpc/heturpc_elastic_server.py + ModelSaver)"""
        from ..nn.parallel import HETERO_PARAMS
        from ..utils.checkpoint import ModelSaver
        self.build()
        st = self.trainer_states
        saver = ModelSaver(ckpt_dir or os.path.join(self.pretrain_config.output_dir, "_rebuild"), save_copies=1)
        if self.idle:
            distributed.global_comm_barrier_rpc()             # the barrier inside the active ranks' saver.save()
        else:
            writer = min(d for p in self.hetero.pipelines for stg in p for d in stg) if self.hetero is not None else 0
            saver.save(st.model, st.optimizer, self.global_step, self.consumed_samples, self.loss_history[-1] if self.loss_history else float("nan"),
                       writer_rank=writer)
            saver.wait()
        distributed.global_comm_barrier_rpc()
        self.trainer_states, self.is_model_built = None, False
        HETERO_PARAMS.clear()
        wrapper_cfg = getattr(self.model_wrapper, "model_config", None)
        if wrapper_cfg is not None and hasattr(wrapper_cfg, "cp_ranks"):
            wrapper_cfg.cp_ranks = ()
        self._set_strategies(list(ds_parallel_configs), hetero_shares)
        self._iter_version = getattr(self, "_iter_version", 0) + 1
        self.build()
        if self.idle:
            return self
        st = self.trainer_states
        loaded = saver.load_latest(st.model, st.optimizer)
        assert loaded is not None and loaded[0] == self.global_step, "the rebuild checkpoint could not be read back"
        if hasattr(st.optimizer, "set_step"):
            st.optimizer.set_step(self.global_step)       # scheduled lr / wd of the next step, not the fresh optimizer's
        return self

    # ------------------------------------------------------------------ build
    def _num_layers(self):
        mc = getattr(self.model_wrapper, "model_config", None)
        if mc is None and isinstance(self.model_wrapper, ModelWrapperFromConfig):
            c = self.model_wrapper.config
            return int(c.get("n_layer", c.get("num_hidden_layers", 12)))
        for k in ("n_layer", "num_hidden_layers"):
            if hasattr(mc, k):
                return int(getattr(mc, k))
        return 12

    def get_train_data_loader(self):
        cfg = self.pretrain_config
        dp, _, _ = _strategy_sizes(self.ds_parallel_configs[self.cur_strategy_id])
        dp //= _cp_degree(self.ds_parallel_configs[self.cur_strategy_id])
        level = cfg.data_load_level.value if isinstance(cfg.data_load_level, DataLoadLevel) else str(cfg.data_load_level)
        kw = dict(global_batch_size=cfg.global_load_size) if level == "SAMPLE" else dict(global_token_num=cfg.global_load_size)
        return build_data_loader(self.train_dataset, self.consumed_samples, load_level=level, dp_rank=self._dp_rank(), dp_size=dp,
                                 seed=cfg.seed, collate_fn=self.data_collator, **kw)

    def _dp_rank(self):
        """data-parallel replica of this rank (devices are laid out [pp][dp][cp][tp]: ring members share their replica's data)"""
        dcp, tp, pp = _strategy_sizes(self.ds_parallel_configs[self.cur_strategy_id])
        cp = _cp_degree(self.ds_parallel_configs[self.cur_strategy_id])
        r = distributed.rank()
        return (r % (dcp * tp)) // tp // cp

    def _cp_index_and_ring(self, strategy_id: int = 0):
        """(position of this rank in its context-parallel ring, global ranks of the ring)"""
        dcp, tp, pp = _strategy_sizes(self.ds_parallel_configs[strategy_id])
        cp = _cp_degree(self.ds_parallel_configs[strategy_id])
        r = distributed.rank()
        stage, within = divmod(r, dcp * tp)
        dcp_idx, tp_idx = divmod(within, tp)
        d, c = divmod(dcp_idx, cp)
        return c, tuple(stage * dcp * tp + (d * cp + k) * tp + tp_idx for k in range(cp))

    def build(self):
        if self.is_model_built:
            return self.trainer_states
        if self.idle:
            # this rank has no work under the current plan: join the collective group creation, build nothing
            from ..nn.parallel import HETERO_PARAMS
            HETERO_PARAMS.clear()
            self.hetero.precreate_groups()
            self.trainer_states, self.is_model_built = None, True
            return None
        self.trainer_states = self.create_define_graph()
        if self.hetero is not None:
            self.hetero.precreate_groups()        # collective over all ranks: tp groups of every stage + cross-pipeline groups
        self.is_model_built = True
        return self.trainer_states

    def create_define_graph(self) -> TrainerStates:
        cfg = self.pretrain_config
        st = TrainerStates()
        in_dsh, in_dgh = parse_multi_ds_parallel_config(self.ds_parallel_configs, "input")
        lb_dsh, lb_dgh = parse_multi_ds_parallel_config(self.ds_parallel_configs, "label")
        dp, _, _ = _strategy_sizes(self.ds_parallel_configs[0])
        seq = int(cfg.max_seq_length or 1024)
        mbs = int(cfg.micro_batch_size or 1)
        tokens = mbs * seq * dp
        ac = core.autocast("bfloat16") if cfg.bf16 else _null()
        with core.graph("define_and_run", create_new=True, num_strategy=self.num_strategy) as g, ac:
            st.seq_len_symbol = IntSymbol(seq)
            cp = _cp_degree(self.ds_parallel_configs[0])
            if cp > 1:
                mc = getattr(self.model_wrapper, "model_config", None) or getattr(self.model_wrapper, "config", None)
                if mc is None or not hasattr(mc, "cp_ranks"):
                    raise ValueError("this model has no context-parallel attention path (config.cp_ranks)")
                mc.cp_ranks = self._cp_index_and_ring(0)[1]
            if hasattr(self.model_wrapper, "create_model"):
                st.model = self.model_wrapper.create_model(self.ds_parallel_configs)
                st.config = getattr(self.model_wrapper, "model_config", None)
            else:
                st.model = self.model_wrapper
                st.config = getattr(st.model, "config", None)
            st.input_ids = core.parallel_placeholder("int64", [tokens], in_dsh, device_group_hierarchy=in_dgh, name="input_ids")
            st.position_ids = core.parallel_placeholder("int64", [tokens], in_dsh, device_group_hierarchy=in_dgh, name="position_ids")
            st.labels = core.parallel_placeholder("int64", [tokens], lb_dsh, device_group_hierarchy=lb_dgh, name="labels")
            if cfg.packing:
                # document boundaries of every packed row (padded with repeats of the row length): variable-length attention
                st.cu_seqlens = core.placeholder("int32", [self.max_docs_per_row + 1], name="cu_seqlens")
                st.loss = st.model(st.input_ids, st.position_ids, st.labels, seq_len=st.seq_len_symbol, cu_seqlens=st.cu_seqlens)
            else:
                st.loss = st.model(st.input_ids, st.position_ids, st.labels, seq_len=st.seq_len_symbol)
            st.optimizer = self.optimizer_wrapper.create_optimizer() if isinstance(self.optimizer_wrapper, OptimizerWrapper) \
                else self.optimizer_wrapper
            st.train_op = st.optimizer.minimize(st.loss)
            st.graph = g
        return st

    # ------------------------------------------------------------------ data
    def prepare_feed_dict(self, batch: Sequence, strategy_id: int = 0):
        """batch: [(inputs, labels)] of this data-parallel rank -> (feed_dict, num_micro_batches, seq_len, stats).
        Padding mode: micro-batches of [micro_batch_size, width]; packing mode: one packed row per micro-batch
        (documents aligned to `pack_alignment` tokens so attention tiles never straddle two documents)."""
        cfg = self.pretrain_config
        st = self.trainer_states
        max_len = int(cfg.max_seq_length or 1024)
        pad_id = getattr(self.tokenizer, "pad_id", 0) if self.tokenizer is not None else 0
        ins, labs = Bucket(pad_id, max_len, cfg.pack_alignment), Bucket(-1, max_len, cfg.pack_alignment)
        for x, y in batch:
            ins.add_data(x)
            labs.add_data(y)
        feeds_i, feeds_p, feeds_l = [], [], []
        if cfg.packing:
            order = np.argsort([-len(x) for x, _ in batch], kind="stable")
            rows_i = ins.pack_data()
            # pack labels with the identical assignment: re-pack (input, label) jointly
            rows = _pack_pairs([batch[i] for i in order], max_len, pad_id, cfg.pack_alignment)
            width = max(len(r[0]) for r in rows)
            cp = _cp_degree(self.ds_parallel_configs[strategy_id])
            unit = cfg.pack_alignment if cp == 1 else int(np.lcm(cfg.pack_alignment, 2 * cp))     # CP: 2 * cp equal chunks per row
            width = (width + unit - 1) // unit * unit
            feeds_cu = []
            for toks, lab, pos, cu in rows:
                n = len(toks)
                feeds_i.append(np.concatenate([toks, np.full(width - n, pad_id, np.int64)]))
                feeds_l.append(np.concatenate([lab, np.full(width - n, -1, np.int64)]))
                feeds_p.append(np.concatenate([pos, np.zeros(width - n, np.int64)]))
                cu = list(cu[: self.max_docs_per_row]) + [width]          # the padding tail is one more (label-masked) document
                feeds_cu.append(np.asarray(cu + [width] * (self.max_docs_per_row + 1 - len(cu)), np.int32))
            seq = width
            if cp > 1:
                # every ring member keeps its SYM chunks of each packed row; cu_seqlens stay in whole-row coordinates (the
                # ring attention masks documents across chunk and rank borders)
                cols = cp_rows(width, cp, self._cp_index_and_ring(strategy_id)[0])
                feeds_i, feeds_l, feeds_p = [a[cols] for a in feeds_i], [a[cols] for a in feeds_l], [a[cols] for a in feeds_p]
                seq = len(cols)
            stats = {"real_tokens": int(sum(len(x) for x, _ in batch)), "fed_tokens": int(width * len(rows)), "rows": len(rows_i)}
        else:
            mbs = int(cfg.micro_batch_size or len(batch))
            cp = _cp_degree(self.ds_parallel_configs[strategy_id])
            unit = cfg.pack_alignment if cp == 1 else int(np.lcm(cfg.pack_alignment, 2 * cp))
            if cfg.dynamic_micro_batch_padding:
                # rows sorted by length so that a micro-batch holds rows of similar length, each micro-batch padded to its own
                # (aligned) longest row: the executor re-infers shapes per micro-batch (graph.run int_symbol_dict lists)
                order = np.argsort([-len(x) for x, _ in batch], kind="stable")
                rows = [batch[i] for i in order]
                nmb = (len(rows) + mbs - 1) // mbs
                widths = []
                for m in range(nmb):
                    chunk = rows[m * mbs:(m + 1) * mbs]
                    w = min(max(len(x) for x, _ in chunk), max_len)
                    w = (w + unit - 1) // unit * unit
                    bi = np.full((mbs, w), pad_id, np.int64)
                    bl = np.full((mbs, w), -1, np.int64)
                    for r, (x, y) in enumerate(chunk):
                        n = min(len(x), w)
                        bi[r, :n], bl[r, :n] = x[:n], y[:n]
                    feeds_i.append(bi.reshape(-1)); feeds_l.append(bl.reshape(-1)); feeds_p.append(np.tile(np.arange(w), mbs))
                    widths.append(w)
                width = widths
            else:
                pi, pl = ins.pad_data(), labs.pad_data()
                width = pi.shape[1]
                nmb = (len(batch) + mbs - 1) // mbs
                for m in range(nmb):
                    bi, bl = pi[m * mbs:(m + 1) * mbs], pl[m * mbs:(m + 1) * mbs]
                    if len(bi) < mbs:   # ragged tail: pad with fully-masked rows so every micro-batch has the same shape
                        k = mbs - len(bi)
                        bi = np.concatenate([bi, np.full((k, width), pad_id, np.int64)])
                        bl = np.concatenate([bl, np.full((k, width), -1, np.int64)])
                    feeds_i.append(bi.reshape(-1))
                    feeds_l.append(bl.reshape(-1))
                    feeds_p.append(np.tile(np.arange(width), mbs))
            widths = width if isinstance(width, list) else [width] * nmb
            seq = list(widths) if isinstance(width, list) else width
            if cp > 1:
                # every ring member keeps its SYM chunks of each row (positions stay the original ones for the rotary embedding)
                ci = self._cp_index_and_ring(strategy_id)[0]
                for m, w in enumerate(widths):
                    assert w % (2 * cp) == 0, f"padded width {w} must be a multiple of 2 * cp = {2 * cp} (set pack_alignment accordingly)"
                    cols = cp_rows(w, cp, ci)
                    feeds_i[m], feeds_l[m], feeds_p[m] = (a.reshape(-1, w)[:, cols].reshape(-1) for a in (feeds_i[m], feeds_l[m], feeds_p[m]))
                seq = [w // cp for w in widths] if isinstance(width, list) else widths[0] // cp
            stats = {"real_tokens": int(sum(len(x) for x, _ in batch)), "fed_tokens": int(sum(widths) * mbs), "rows": len(batch)}
        to_t = (lambda a: torch.as_tensor(a).pin_memory()) if torch.cuda.is_available() else torch.as_tensor
        feed = {st.input_ids: [to_t(a) for a in feeds_i], st.position_ids: [to_t(a) for a in feeds_p],
                st.labels: [to_t(a) for a in feeds_l]}
        if cfg.packing and st.cu_seqlens is not None:
            feed[st.cu_seqlens] = [to_t(a) for a in feeds_cu]
        return feed, len(feeds_i), seq, stats

    def train_data_iterator(self):
        while True:
            loader = self.get_train_data_loader()
            got = False
            for batch in loader:
                got = True
                # global sample count covered so far (the loader's own book-keeping: identical on every data-parallel
                # rank under both load levels -- `len(batch)` is only this rank's slice)
                self.consumed_samples = int(loader.consumed_yielded)
                yield batch
            self.epoch += 1
            self.consumed_samples = 0
            if not got:
                raise RuntimeError("the training dataset produced no batch (global_load_size larger than the dataset?)")

    # ------------------------------------------------------------------ train
    def _train_step(self, batch, strategy_id=0):
        st = self.trainer_states
        scale = None
        if self.hetero is not None:
            # every rank loaded the whole global batch: keep this pipeline's share, weight its gradient by the share
            n_global = len(batch)
            batch = batch[self.hetero.batch_slice(n_global)]
            scale = self.hetero.grad_scale(len(batch), n_global)
        feed, nmb, seq, stats = self.prepare_feed_dict(batch, strategy_id)
        dp, _, _ = _strategy_sizes(self.ds_parallel_configs[strategy_id])
        # `seq` is one sequence length, or one per micro-batch (dynamic_micro_batch_padding)
        out = st.graph.run(st.loss, [st.loss, st.train_op], feed, int_symbol_dict={st.seq_len_symbol: seq}, num_micro_batches=nmb,
                           cur_strategy_id=strategy_id, grad_scale=scale if scale is not None else 1.0 / dp)
        loss = out[0]
        lv = float(loss.float().mean()) if loss is not None else None
        if self.hetero is not None and lv is not None and self.hetero.rank in self.hetero.last_stage_ranks:
            lv = self.hetero.reduce_loss(lv, len(batch), n_global)      # global mean from the per-pipeline means
        return lv, stats

    def train(self, steps: Optional[int] = None, strategy_schedule: Optional[Callable[[int], int]] = None, strategy_id: Optional[int] = None,
              cp_list=None, run_level=None):
        """run `steps` optimizer steps; `strategy_schedule(step) -> strategy id` enables hot switching between the
        strategies of ds_parallel_configs (HotSPa: e.g. by the step's max sequence length).  `strategy_id` (reference
        signature `train(cp_list, strategy_id, run_level)`) starts under that strategy; the context-parallel layout comes
        from the strategy's `cp` field, so `cp_list` is accepted for compatibility only; `run_level` wraps the loop in
        `hetu.run_level(...)` (e.g. "compute_only" for a forward-only pass over the data)."""
        if run_level is not None:
            with core.run_level(run_level):
                return self.train(steps, strategy_schedule, strategy_id, cp_list, None)
        cfg = self.pretrain_config
        self.build()
        if strategy_id is not None and int(strategy_id) != self.cur_strategy_id and not self.idle:
            self.trainer_states.graph.switch_strategy(int(strategy_id))
            self.cur_strategy_id = int(strategy_id)
        it, it_version = self.train_data_iterator(), getattr(self, "_iter_version", 0)
        steps = int(steps if steps is not None else cfg.steps)
        saver = None
        if cfg.save_interval:
            from ..utils.checkpoint import ModelSaver
            saver = ModelSaver(cfg.output_dir, save_interval=cfg.save_interval)
        prof = None
        for _ in range(steps):
            if self.idle:
                # keep the step counter (and every world-collective the active ranks execute) in lock step
                self.global_step += 1
                for cb in self.callbacks:
                    cb(self, None, {})
                if cfg.save_interval and self.global_step % cfg.save_interval == 0:
                    distributed.global_comm_barrier_rpc()
                continue
            if it_version != getattr(self, "_iter_version", 0):
                # the strategy was rebuilt by a callback: the data-parallel sharding of the loader changed with it
                it, it_version = self.train_data_iterator(), self._iter_version
            batch = next(it)
            sid = int(strategy_schedule(self.global_step)) if strategy_schedule else self.cur_strategy_id
            if sid != self.cur_strategy_id:
                self.trainer_states.graph.switch_strategy(sid)
                self.cur_strategy_id = sid
            if cfg.torch_profile and self.global_step == cfg.start_profile_step:
                prof = core.profiler(graph=self.trainer_states.graph)
                prof.__enter__()
            t0 = time.perf_counter()
            try:
                loss, stats = self._train_step(batch, sid)
            except RuntimeError as e:
                # a failed rank must not leave its peers blocked in a collective: record the error where the launcher /
                # elastic controller looks for it and (when asked) take the whole local worker group down
                # (ref: engine/trainer.py:316-322 -- logs/exception.txt + os.killpg)
                os.makedirs(os.path.join(cfg.output_dir, "logs"), exist_ok=True)
                with open(os.path.join(cfg.output_dir, "logs", "exception.txt"), "a") as f:
                    f.write(f"rank {distributed.rank()} step {self.global_step}: {type(e).__name__}: {e}\n")
                if os.environ.get("HETU_KILL_GROUP_ON_ERROR", "0") == "1":
                    import signal
                    os.killpg(os.getpgrp(), signal.SIGTERM)
                raise
            if torch.cuda.is_available():
                torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            self.step_times.append(dt)
            if loss is not None:
                self.loss_history.append(loss)
            if prof is not None and self.global_step == cfg.end_profile_step:
                prof.__exit__(None, None, None)
                os.makedirs(cfg.profile_save_path, exist_ok=True)
                with open(os.path.join(cfg.profile_save_path, f"trace_rank{distributed.rank()}.json"), "w") as f:
                    json.dump(prof.chrome_trace(pid=distributed.rank()), f)           # open in chrome://tracing / Perfetto
                with open(os.path.join(cfg.profile_save_path, f"summary_rank{distributed.rank()}.json"), "w") as f:
                    json.dump(prof.summary(), f)
                prof = None
            self.global_step += 1
            if hasattr(self.trainer_states.optimizer, "step_lr"):
                self.trainer_states.optimizer.step_lr()        # lr warm-up / decay and weight-decay schedule for the next step
            if loss is not None and cfg.log_interval and self.global_step % cfg.log_interval == 0 and distributed.rank() in self._loss_ranks():
                print(f"[trainer] step {self.global_step} loss {loss:.4f} time {dt * 1e3:.1f} ms "
                      f"tokens {stats['real_tokens']}/{stats['fed_tokens']} strategy {sid}", flush=True)
            for cb in self.callbacks:
                cb(self, loss, stats)
            if getattr(cfg, "eval_interval", 0) and getattr(self, "eval_dataset", None) is not None and self.global_step % cfg.eval_interval == 0:
                ev = self.evaluate(max_batches=getattr(cfg, "eval_iters", None) or None)
                if ev is not None and cfg.log_interval:
                    print(f"[trainer] step {self.global_step} eval loss {ev['loss']:.4f} ppl {ev['perplexity']:.2f} ({ev['tokens']} tokens)", flush=True)
            if saver is not None and self.global_step % cfg.save_interval == 0:
                writer = min(d for p in self.hetero.pipelines for stg in p for d in stg) if self.hetero is not None else 0
                saver.save(self.trainer_states.model, self.trainer_states.optimizer, self.global_step, self.consumed_samples,
                           loss if loss is not None else float("nan"), writer_rank=writer)
            if cfg.plot_loss and self.global_step % cfg.plot_update_freq == 0:
                self.plot_training_loss()
        if saver is not None:
            saver.wait()
        return self.loss_history

    # ------------------------------------------------------------------ evaluation
    def evaluate(self, eval_dataset=None, max_batches: Optional[int] = None, strategy_id: Optional[int] = None) -> Optional[dict]:
        """forward-only pass over `eval_dataset` (default: `self.eval_dataset`) under the current strategy: no optimizer step, no
        gradient ops are fetched.  -> {"loss": token-weighted mean, "perplexity", "batches", "tokens"} on the ranks that own the
        loss, None elsewhere.  The data-parallel ranks see disjoint slices; their means are combined by token count."""
        import math
        ds = eval_dataset if eval_dataset is not None else getattr(self, "eval_dataset", None)
        if ds is None:
            raise ValueError("evaluate() needs an eval_dataset")
        self.build()
        if self.idle:
            return None
        cfg = self.pretrain_config
        sid = self.cur_strategy_id if strategy_id is None else int(strategy_id)
        dp, _, _ = _strategy_sizes(self.ds_parallel_configs[sid])
        dp //= _cp_degree(self.ds_parallel_configs[sid])
        level = cfg.data_load_level.value if isinstance(cfg.data_load_level, DataLoadLevel) else str(cfg.data_load_level)
        kw = dict(global_batch_size=cfg.global_load_size) if level == "SAMPLE" else dict(global_token_num=cfg.global_load_size)
        loader = build_data_loader(ds, 0, load_level=level, dp_rank=self._dp_rank(), dp_size=dp, seed=cfg.seed, collate_fn=self.data_collator, **kw)
        st = self.trainer_states
        tot_loss, tot_tok, nb = 0.0, 0, 0
        for batch in loader:
            if max_batches is not None and nb >= max_batches:
                break
            if self.hetero is not None:
                batch = batch[self.hetero.batch_slice(len(batch))]
            feed, nmb, seq, stats = self.prepare_feed_dict(batch, sid)
            out = st.graph.run(st.loss, [st.loss], feed, int_symbol_dict={st.seq_len_symbol: seq}, num_micro_batches=nmb, cur_strategy_id=sid)
            nb += 1
            if out[0] is not None:
                n = max(int(stats.get("real_tokens", 1)), 1)
                tot_loss += float(out[0].float().mean()) * n
                tot_tok += n
        if tot_tok == 0:
            return None
        ranks = sorted(self._loss_ranks())
        if len(ranks) > 1 and distributed.world_size() > 1 and distributed.rank() in ranks:
            from .. import _C
            t = _C.comm_all_reduce(torch.tensor([tot_loss, float(tot_tok)], dtype=torch.float64), ranks, "sum")
            tot_loss, tot_tok = float(t[0]), int(t[1])
        if distributed.rank() not in ranks:
            return None
        loss = tot_loss / tot_tok
        res = {"loss": loss, "perplexity": math.exp(min(loss, 50.0)), "batches": nb, "tokens": tot_tok, "step": self.global_step}
        self.eval_history.append(res)
        return res

    def _loss_ranks(self):
        if self.hetero is not None:
            return {self.hetero.last_stage_ranks[0]}
        cfg = self.ds_parallel_configs[self.cur_strategy_id]
        return {g[0] for g in cfg["label"]["device_group_union"]}

    def save_model(self, output_dir: Optional[str] = None):
        from ..utils.checkpoint import ModelSaver
        sv = ModelSaver(output_dir or self.pretrain_config.output_dir)
        sv.save(self.trainer_states.model, self.trainer_states.optimizer, self.global_step, self.consumed_samples,
                self.loss_history[-1] if self.loss_history else float("nan"))
        sv.wait()

    def plot_training_loss(self):
        """loss curve as CSV + (when matplotlib exists) PNG under output_dir"""
        os.makedirs(self.pretrain_config.output_dir, exist_ok=True)
        with open(os.path.join(self.pretrain_config.output_dir, "loss.csv"), "w") as f:
            f.write("step,loss\n" + "\n".join(f"{i + 1},{v}" for i, v in enumerate(self.loss_history)))
        # dependency-free SVG of the same curve (train loss, plus evaluation points when there are any)
        if self.loss_history:
            W, H, pad = 720, 360, 40
            lo, hi = min(self.loss_history), max(self.loss_history)
            span = (hi - lo) or 1.0
            n = max(len(self.loss_history) - 1, 1)
            sx = lambda i: pad + (W - 2 * pad) * i / n                                   # noqa: E731
            sy = lambda v: H - pad - (H - 2 * pad) * (min(max(v, lo), hi) - lo) / span   # noqa: E731
            pts = " ".join(f"{sx(i):.1f},{sy(v):.1f}" for i, v in enumerate(self.loss_history))
            dots = "".join(f'<circle cx="{sx(min(e["step"] - 1, n)):.1f}" cy="{sy(e["loss"]):.1f}" r="3" fill="#d62728"/>' for e in self.eval_history)
            svg = (f'<svg xmlns="http://www.w3.org/2000/svg" width="{W}" height="{H}"><rect width="100%" height="100%" fill="white"/>'
                   f'<polyline fill="none" stroke="#1f77b4" stroke-width="1.5" points="{pts}"/>{dots}'
                   f'<text x="{pad}" y="{pad - 12}" font-size="12">loss {hi:.4f} (top) .. {lo:.4f} (bottom), {len(self.loss_history)} steps</text>'
                   f'<line x1="{pad}" y1="{H - pad}" x2="{W - pad}" y2="{H - pad}" stroke="black"/><line x1="{pad}" y1="{pad}" x2="{pad}" y2="{H - pad}" stroke="black"/></svg>')
            with open(os.path.join(self.pretrain_config.output_dir, "loss.svg"), "w") as f:
                f.write(svg)
        try:
            import matplotlib
            matplotlib.use("Agg")
            import matplotlib.pyplot as plt
            plt.figure()
            plt.plot(range(1, len(self.loss_history) + 1), self.loss_history)
            plt.xlabel("step")
            plt.ylabel("loss")
            plt.savefig(os.path.join(self.pretrain_config.output_dir, "loss.png"))
            plt.close()
        except Exception:
            pass


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def _pack_pairs(pairs, max_len, pad_id, alignment):
    """first-fit-decreasing packing of (inputs, labels) pairs -> [(tokens, labels, positions)] per packed row"""
    def al(n):
        return min((n + alignment - 1) // alignment * alignment, max_len)
    rows, space = [], []
    for x, y in pairs:
        x, y = np.asarray(x[:max_len], np.int64), np.asarray(y[:max_len], np.int64)
        need = al(len(x))
        for r in range(len(rows)):
            if space[r] >= need:
                break
        else:
            rows.append([[], [], []])
            space.append(max_len)
            r = len(rows) - 1
        pad = need - len(x)
        rows[r][0].append(np.concatenate([x, np.full(pad, pad_id, np.int64)]))
        rows[r][1].append(np.concatenate([y, np.full(pad, -1, np.int64)]))
        rows[r][2].append(np.concatenate([np.arange(len(x)), np.zeros(pad, np.int64)]))
        space[r] -= need
    out = []
    for a, b, c in rows:
        cu = np.concatenate([[0], np.cumsum([len(t) for t in a])])
        out.append((np.concatenate(a), np.concatenate(b), np.concatenate(c), cu))
    return out
