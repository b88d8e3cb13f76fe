"""Trainers that change the parallel strategy while training:
* HotSPaTrainer -- sequence-length buckets mapped to strategies, hot switching between them step by step
  (ref: examples/hotspa/llama_hot_switch_trainer.py, `--bucket_sizes "32768 16384 4096 0"`);
* MalleusTrainer -- periodically profiles per-device slow-down ratios, re-plans with the Malleus StrategyModel and moves
  to the new plan (ref: examples/hydraulis/llama_trainer.py MalleusTrainer, engine/strategy.py)."""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence


from .straggler import Straggler
from .strategy import StrategyModel, TrainerCtxs, TrainerStrategyArgs
from .trainer import Trainer


class HotSPaTrainer(Trainer):
    """ds_parallel_configs[i] serves sequences longer than bucket_sizes[i] (sizes descending, last = 0 catches the rest):
    every step runs under the strategy of its longest sequence; the executor re-shards parameters / optimizer states when the
    strategy changes (hot switch)"""

    def __init__(self, *args, bucket_sizes: Sequence[int] = (0,), **kwargs):
        super().__init__(*args, **kwargs)
        self.bucket_sizes = list(bucket_sizes)
        assert len(self.bucket_sizes) == self.num_strategy, "one bucket bound per strategy"
        assert self.bucket_sizes == sorted(self.bucket_sizes, reverse=True), "bucket sizes must be descending"
        self.switch_count = 0

    def strategy_for(self, max_len: int) -> int:
        for i, lo in enumerate(self.bucket_sizes):
            if max_len > lo or i == len(self.bucket_sizes) - 1:
                return i
        return len(self.bucket_sizes) - 1

    def _train_step(self, batch, strategy_id=0):
        sid = self.strategy_for(max(len(x) for x, _ in batch))
        if sid != self.cur_strategy_id:
            self.switch_count += 1
            self.cur_strategy_id = sid
        return super()._train_step(batch, sid)


class MalleusTrainer(Trainer):
    """every `replan_interval` steps: measure slow-down ratios (Straggler workload), solve the Malleus plan; when the plan
    is expressible as a homogeneous strategy (full tensor-parallel groups, identical pipelines) it is appended as a new
    strategy and training hot switches to it; genuinely heterogeneous plans (different tp degrees / stage counts per
    pipeline) are executed through engine.hetero.HeteroSession (member-local graphs + grouped gradient all-reduce)"""

    def __init__(self, *args, ctxs: Optional[TrainerCtxs] = None, strategy_args: Optional[TrainerStrategyArgs] = None, replan_interval: int = 50,
                 ratio_source: Optional[Callable[[], Dict[int, float]]] = None, auto_apply: bool = False, **kwargs):
        super().__init__(*args, **kwargs)
        self.auto_apply = auto_apply
        self.ctxs = ctxs or TrainerCtxs()
        self.strategy_args = strategy_args
        self.replan_interval, self.ratio_source = replan_interval, ratio_source
        self.last_model: Optional[StrategyModel] = None
        self.plans_log: List[dict] = []

    def measure(self) -> Dict[int, float]:
        if self.ratio_source is not None:
            return self.ratio_source()
        from .. import distributed
        return Straggler(max(distributed.world_size(), 1)).run_profile()

    def maybe_replan(self) -> Optional[dict]:
        if self.strategy_args is None or self.global_step == 0 or self.global_step % self.replan_interval:
            return None
        ratios = self.measure()
        model = StrategyModel(self.ctxs, self.strategy_args, ratios)
        if self.last_model is not None and model == self.last_model:
            return None                                  # nothing changed beyond the safe gap
        strategy, cfg = model.make_plans()
        self.last_model = model
        rec = {"step": self.global_step, "ratios": ratios, "hetero_layers": strategy.hetero_layers,
               "micro_batches": strategy.hetero_micro_batch_num_list, "executable": model.executable_config is not None,     # adoptable by in-place hot switching
               "hetero_session": model.executable_config is None,     # otherwise: rebuild through engine.hetero.HeteroSession
               "estimated_time": model.estimate_time(model.plans)}
        self.plans_log.append(rec)
        if self.auto_apply:
            rec["applied"] = self.apply_plan(model)
        return rec

    def apply_plan(self, model: StrategyModel) -> str:
        """move the running job onto the plan: identical pipelines with equal micro-batch counts are an ordinary
        (dp, tp, pp) strategy; anything else -- unequal batch shares, different tp degrees or stage counts -- runs through
        the heterogeneous member-local path.  Both go through `Trainer.rebuild` (split checkpoint, re-sharded on load)."""
        strategy = model.strategies
        shares = list(strategy.hetero_micro_batch_num_list)
        from .. import distributed
        used = {d for pl in model.plans for g in pl["groups"] for d in g.devices}
        covers_world = len(used) >= max(distributed.world_size(), 1)      # otherwise some ranks idle: member-local path
        if model.executable_config is not None and len(set(shares)) == 1 and covers_world:
            self.rebuild([model.executable_config])
            return "homogeneous"
        self.rebuild([model.ds_parallel_configs], hetero_shares=shares)
        return "hetero"

    def train(self, steps=None, strategy_schedule=None):
        self.callbacks.append(lambda trainer, loss, stats: trainer.maybe_replan())
        return super().train(steps, strategy_schedule)
