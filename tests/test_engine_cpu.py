"""Trainer / data / LoRA / Malleus planner on CPU (ref test model: tests/ci_test + examples/pretrain scripts)."""
import json
import os

import numpy as np
import pytest
import torch

import hetu_b200 as ht
from hetu_b200.data import (Bucket, ByteTokenizer, DataLoader, IndexedTokenDataset, JsonDataset, SyntheticDataset,
                            build_chat_sample, generate_cp_pack_data, pack_sequences)
from hetu_b200.engine import (ModelWrapper, OptimizerWrapper, SFTConfig, SFTTrainer, StrategyModel, Trainer, TrainerCtxs,
                              TrainerStrategyArgs, TrainingConfig)
from hetu_b200.engine.strategy import dispatch_sequences
from hetu_b200.models import GPTConfig, GPTLMHeadModel


def _mcfg():
    return GPTConfig(vocab_size=259, n_positions=64, n_embd=32, n_layer=2, n_head=4)


def test_pack_sequences_alignment_and_cu_seqlens():
    seqs = [np.arange(n) for n in (5, 30, 17, 64, 9)]
    rows = pack_sequences(seqs, 64, pad_id=-7, alignment=16)
    tot = 0
    for toks, cu in rows:
        assert len(toks) <= 64 and cu[0] == 0 and cu[-1] == len(toks)
        assert all(c % 16 == 0 for c in cu)
        tot += len(cu) - 1
    assert tot == len(seqs)
    b = Bucket(0, 64, 16)
    for s in seqs:
        b.add_data(s)
    b.pad_data()
    b.pack_data()
    st = b.padding_stats()
    assert st["pack_efficiency"] >= st["pad_efficiency"]


def test_cp_pack_sym_split_is_balanced_and_complete():
    toks = np.arange(64)
    cu = np.array([0, 32, 64], dtype=np.int32)
    parts = generate_cp_pack_data(toks, cu, cp=2, pattern="SYM")
    assert sorted(np.concatenate([p[0] for p in parts]).tolist()) == list(range(64))
    # causal work of rank r ~ sum of positions: symmetric split equalises it
    w = [p[0].sum() for p in parts]
    assert abs(w[0] - w[1]) <= 2


def test_dataloader_resume_and_dp_sharding(tmp_path):
    ds = SyntheticDataset(40, 100, 16, seed=3)
    a = DataLoader(ds, global_batch_size=8, dp_rank=0, dp_size=2, prefetch=0)
    b = DataLoader(ds, global_batch_size=8, dp_rank=1, dp_size=2, prefetch=0)
    ba, bb = next(iter(a)), next(iter(b))
    assert len(ba) == len(bb) == 4 and not any(np.array_equal(x, y) for x in ba for y in bb)
    full = [x for batch in DataLoader(ds, global_batch_size=8, prefetch=0) for x in batch]
    r = DataLoader(ds, global_batch_size=8, prefetch=2)
    r.restart(16)
    rest = [x for batch in r for x in batch]
    assert len(rest) == 24 and np.array_equal(rest[0], full[16])
    IndexedTokenDataset.build(str(tmp_path / "tok"), [[1, 2, 3], [4, 5]])
    it = IndexedTokenDataset(str(tmp_path / "tok"))
    assert len(it) == 2 and it[1].tolist() == [4, 5]
    p = tmp_path / "d.jsonl"
    p.write_text("\n".join(json.dumps({"text": "hello world %d" % i}) for i in range(3)))
    jd = JsonDataset(str(p), "text", ByteTokenizer(), max_seq_len=8)
    assert len(jd) == 3 and len(jd[0]) == 9


def test_chat_sample_masks_non_assistant_tokens():
    tok = ByteTokenizer()
    ids, labels = build_chat_sample([{"role": "user", "content": "hi"}, {"role": "assistant", "content": "yo"}], tok)
    assert len(ids) == len(labels) and (labels != -1).sum() == len("yo\n")


def test_trainer_padding_and_packing_run():
    ht.init_comm_group(1)
    ds = SyntheticDataset(64, 259, 32, min_seq_len=8, seed=1, length_distribution="uniform")
    cfg = TrainingConfig(packing=False, micro_batch_size=2, global_load_size=4, max_seq_length=32, steps=3, learning_rate=1e-2,
                         log_interval=0, pack_alignment=16)
    tr = Trainer(cfg, ModelWrapper(GPTLMHeadModel, _mcfg()), ByteTokenizer(), OptimizerWrapper({"type": "adam", "lr": 1e-2}), ds)
    losses = tr.train()
    assert len(losses) == 3 and all(np.isfinite(losses))
    cfg2 = TrainingConfig(packing=True, global_load_size=8, max_seq_length=64, steps=3, learning_rate=1e-2, pack_alignment=16, log_interval=0)
    tr2 = Trainer(cfg2, ModelWrapper(GPTLMHeadModel, _mcfg()), ByteTokenizer(), None, ds)
    assert len(tr2.train()) == 3


def test_trainer_overfits_fixed_text():
    ht.init_comm_group(1)

    class Fixed:
        def __len__(self):
            return 8

        def __getitem__(self, i):
            return np.asarray(ByteTokenizer().encode("the quick brown fox jumps")[:25], np.int64)
    cfg = TrainingConfig(packing=False, micro_batch_size=4, global_load_size=4, max_seq_length=32, steps=30, log_interval=0,
                         pack_alignment=8)
    tr = Trainer(cfg, ModelWrapper(GPTLMHeadModel, _mcfg()), ByteTokenizer(), OptimizerWrapper({"type": "adam", "lr": 1e-2}), Fixed())
    losses = tr.train()
    assert losses[-1] < 0.5 * losses[0]


def test_sft_trainer_with_lora_trains_only_adapters(tmp_path):
    ht.init_comm_group(1)
    from hetu_b200.peft import lora_state_dict, merge_lora_weights
    recs = [{"messages": [{"role": "user", "content": f"say {i}"}, {"role": "assistant", "content": f"number {i} it is"}]} for i in range(16)]
    scfg = SFTConfig(packing=False, micro_batch_size=4, global_load_size=4, max_seq_length=64, steps=6, learning_rate=1e-2, lora_rank=4,
                     log_interval=0, pack_alignment=16)
    st = SFTTrainer(scfg, ModelWrapper(GPTLMHeadModel, _mcfg()), ByteTokenizer(), None, recs)
    base_before = {k: v.clone() for k, v in st.build().model.state_dict().items() if "lora_" not in k}
    losses = st.train()
    assert losses[-1] < losses[0]
    m = st.trainer_states.model
    after = m.state_dict()
    for k, v in base_before.items():
        assert torch.equal(after[k], v), f"frozen weight {k} changed"
    assert len(lora_state_dict(m)) == 16
    merged = merge_lora_weights(m)
    assert not any("lora_" in k or ".base." in k for k in merged)


def test_malleus_planner_gives_stragglers_fewer_layers():
    ctx = TrainerCtxs(normal_layers=8, normal_mbn=8)
    old = TrainerStrategyArgs(dp=2, tp=2, pp=2, rank_to_device_mapping={i: i for i in range(8)})
    sr = {i: 1.0 for i in range(8)}
    sr[3] = 2.5
    m = StrategyModel(ctx, old, sr)
    st, cfg = m.make_plans()
    assert sum(st.hetero_micro_batch_num_list) == 16 and all(sum(l) == 16 for l in st.hetero_layers)
    slow = [(pl, i) for pl in m.plans for i, g in enumerate(pl["groups"]) if 3 in g.devices]
    pl, i = slow[0]
    assert pl["layers"][i] < 8      # the stage containing the straggler holds fewer layers
    assert cfg["hetero"] and len(cfg["blocks"]) == 16
    same = StrategyModel(ctx, old, dict(sr))
    assert same == m
    bal = dispatch_sequences([100, 4000, 300, 2000, 1500, 800], [1.0, 0.5])
    assert sorted(sum(bal, [])) == list(range(6))


def test_yaml_experiment_config_with_overrides():
    from hetu_b200.engine.config_loader import load_experiment
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples", "pretrain", "config", "gpt_small_dp2_tp2.yaml")
    e = load_experiment(path, ["trainer.steps=7", "ds_parallel.tp=4"])
    assert e["trainer"].steps == 7 and e["strategy"].tp == 4 and e["strategy"].world() == 8
    assert e["trainer"].ds_parallel is e["strategy"] and e["model"]["type"] == "gpt" and e["optimizer"]["weight_decay"] == 0.1


def test_hydraulis_dispatch_mixes_strategies_and_balances_replicas():
    from hetu_b200.engine.hydraulis import HydraulisPlanner, StrategyCost, dispatch_batch
    short = StrategyCost("dp8", 8, a=0.004, b=1e-6, max_seq=4096, switch_ms=1.0)          # fast per token, cannot hold long sequences
    long_ = StrategyCost("tp4dp2", 2, a=0.006, b=6e-7, max_seq=32768, switch_ms=1.0)
    rng = np.random.RandomState(0)
    lens = rng.choice([256, 512, 1024, 2048], 64).tolist() + [16384, 30000]
    p = dispatch_batch(lens, [short, long_])
    assert set(p["assignment"][-2:]) == {1}                      # the long sequences can only run on the long strategy
    n_short = sum(1 for a in p["assignment"] if a == 0)
    assert 0 < n_short < 64        # some short sequences fill the idle replica of the long strategy, the rest use the cheap one
    only_long = dispatch_batch(lens, [long_])
    assert p["makespan_ms"] < only_long["makespan_ms"]
    loads = [sum(short.seq_ms(lens[i]) for i in r) for r in p["per_strategy"][0]["replicas"]]
    assert max(loads) - min(loads) <= short.seq_ms(2048) + 1e-6   # LPT keeps the replicas within one sequence of each other
    fitted = StrategyCost.fit("x", 1, [(512, 6.0), (1024, 13.5), (2048, 33.0), (4096, 90.0)])
    assert abs(fitted.seq_ms(2048) - 33.0) < 1.0
    plans = []

    class Sink:
        def produce(self, item):
            plans.append(item)
    pl = HydraulisPlanner([short, long_], Sink())
    pl.plan(lens)
    pl.plan(lens[:10])
    assert [x["step"] for x in plans] == [0, 1] and plans[0]["strategies"] == ["dp8", "tp4dp2"]


def test_hotspa_trainer_picks_strategy_by_sequence_bucket_and_malleus_trainer_replans():
    from hetu_b200.engine import HotSPaTrainer, MalleusTrainer
    from hetu_b200.models import generate_ds_parallel_config
    ht.init_comm_group(1)
    ds = SyntheticDataset(64, 259, 48, min_seq_len=8, seed=2, length_distribution="uniform")
    dsc = [generate_ds_parallel_config(2, 1, 1, 1, 1), generate_ds_parallel_config(2, 1, 1, 1, 1)]
    cfg = TrainingConfig(packing=False, micro_batch_size=2, global_load_size=2, max_seq_length=48, steps=6, learning_rate=1e-2, log_interval=0,
                         pack_alignment=8)
    tr = HotSPaTrainer(cfg, ModelWrapper(GPTLMHeadModel, _mcfg()), ByteTokenizer(), None, ds, ds_parallel_configs=dsc, bucket_sizes=[24, 0])
    assert tr.strategy_for(40) == 0 and tr.strategy_for(10) == 1
    losses = tr.train()
    assert len(losses) == 6 and all(np.isfinite(losses))
    before = tr.switch_count
    short = [(np.arange(9), np.arange(9)), (np.arange(7), np.arange(7))]
    long_ = [(np.arange(40) % 200, np.arange(40) % 200), (np.arange(30), np.arange(30))]
    for b in (short, long_, short):
        l, _ = tr._train_step(b)
        assert np.isfinite(l)
    assert tr.switch_count - before >= 2 and tr.cur_strategy_id == 1
    ratios = {i: 1.0 for i in range(8)}
    ratios[5] = 2.0
    mt = MalleusTrainer(cfg, ModelWrapper(GPTLMHeadModel, _mcfg()), ByteTokenizer(), None, ds, ctxs=TrainerCtxs(normal_layers=8, normal_mbn=8),
                        strategy_args=TrainerStrategyArgs(dp=2, tp=2, pp=2, rank_to_device_mapping={i: i for i in range(8)}), replan_interval=2,
                        ratio_source=lambda: ratios)
    mt.train(steps=5)
    assert len(mt.plans_log) == 1 and mt.plans_log[0]["step"] == 2      # second measurement equals the first -> no new plan
    assert any(sum(l) == 16 for l in mt.plans_log[0]["hetero_layers"])


def test_hetero_session_localises_and_plans_gradient_sync():
    """a heterogeneous strategy (tp4 | tp2 | tp1 pipelines) seen from single ranks: local config, batch shares and the
    slice / group plan of the cross-pipeline gradient synchronisation"""
    from hetu_b200.engine.hetero import HeteroSession
    from hetu_b200.models.parallel_config import generate_hetero_ds_parallel_config
    from hetu_b200.nn import parallel as P
    pipes = [{"stages": [{"devices": [0, 1, 2, 3], "layers": [0, 1]}]},
             {"stages": [{"devices": [4, 5], "layers": [0, 0]}, {"devices": [6, 7], "layers": [1, 1]}]},
             {"stages": [{"devices": [8], "layers": [0, 1]}]}]
    cfg = generate_hetero_ds_parallel_config(2, pipes, zero=False)
    s = HeteroSession(cfg, rank=5, shares=[4, 2, 1])
    assert s.pipeline == 1 and s.num_pipelines == 3
    assert s.pipelines[1] == [[4, 5], [6, 7]] and s.first_stage_ranks == [0, 4, 8] and s.last_stage_ranks == [0, 6, 8]
    assert s.split_batch(14) == [8, 4, 2] and s.batch_slice(14) == slice(8, 12)
    assert sum(s.split_batch(9)) == 9 and min(s.split_batch(3)) == 1
    loc = s.local_cfg
    assert loc["blocks"]["blocks0"]["attn"]["qkv"]["device_group_union"] == [[4, 5]]
    assert loc["blocks"]["blocks1"]["attn"]["qkv"]["device_group_union"] == [[6, 7]]
    assert loc["blocks"]["blocks0"]["attn"]["qkv"]["split"] == {"0": [2]}
    assert loc["wpe"]["dup"] == [2] and loc["input"]["device_group_union"] == [[4, 5]]
    # gradient sync plan of a column-split parameter of layer 0: lcm(4, 2, 1) = 4 finest shards
    leaf = cfg["blocks"]["blocks0"]["attn"]["qkv"]
    P.HETERO_PARAMS.clear()
    P.HETERO_PARAMS["w"] = (leaf, 1, 0)
    spec = P.hetero_grad_sync_spec("w", [8, 16], 5)            # rank 5 = second tp rank of the tp2 pipeline: fine shards 2, 3
    assert spec["dim"] == 0 and spec["offsets"] == [0, 4] and spec["lengths"] == [4, 4]
    assert spec["groups"] == [[2, 5, 8], [3, 5, 8]] and spec["bcast_ranks"] == []
    P.HETERO_PARAMS["w"] = (leaf, 2, 0)
    spec = P.hetero_grad_sync_spec("w", [16, 16], 8)           # the tp1 pipeline holds all four
    assert spec["offsets"] == [0, 4, 8, 12] and spec["groups"] == [[0, 4, 8], [1, 4, 8], [2, 5, 8], [3, 5, 8]]
    # replicated parameter: leaders reduce, the tp group broadcasts
    P.HETERO_PARAMS["ln"] = (cfg["blocks"]["blocks0"]["layernorm1"], 0, None)
    lead, follower = P.hetero_grad_sync_spec("ln", [16], 0), P.hetero_grad_sync_spec("ln", [16], 2)
    assert lead["groups"] == [[0, 4, 8]] and lead["lengths"] == [16] and lead["bcast_ranks"] == [0, 1, 2, 3]
    assert follower["groups"] == [] and follower["bcast_ranks"] == [0, 1, 2, 3]
    assert [0, 4, 8] in P.all_hetero_groups() and [3, 5, 8] in P.all_hetero_groups()
    P.HETERO_PARAMS.clear()


def _lobra_setup():
    from hetu_b200.engine.lobra import LoraCostModel
    rng = np.random.RandomState(0)
    true = {1: (2e-7, 0.0, 4e-4, 0.0, 0.0, 0.05), 2: (1e-7, 0.0, 2.3e-4, 0.0, 0.0, 0.09), 4: (5e-8, 0.0, 1.4e-4, 0.0, 0.0, 0.12)}
    recs = []
    for tp, c in true.items():
        for mbs in (1, 2, 4, 8):
            for s in (256, 512, 1024, 2048, 4096, 8192):
                t = float(np.dot(c, LoraCostModel.features(mbs, s)))
                recs.append((tp, mbs, s, t * (1 + 0.01 * rng.randn())))
    cm = LoraCostModel.fit(recs)
    cands = [{"tp": 1, "pp": 1, "max_tokens": 2048, "throughput_per_gpu": 10.0}, {"tp": 2, "pp": 1, "max_tokens": 2048, "throughput_per_gpu": 8.0},
             {"tp": 2, "pp": 1, "max_tokens": 8192, "throughput_per_gpu": 8.5}, {"tp": 4, "pp": 1, "max_tokens": 8192, "throughput_per_gpu": 7.0},
             {"tp": 4, "pp": 1, "max_tokens": 16384, "throughput_per_gpu": 7.0}, {"tp": 2, "pp": 2, "max_tokens": 16384, "throughput_per_gpu": 7.5},
             {"tp": 1, "pp": 1, "max_tokens": 0}]
    tasks = [{256: 300, 512: 120, 1024: 20}, {512: 60, 2048: 30, 8192: 6}, {1024: 40, 4096: 10, 16384: 2}]
    return cm, true, cands, tasks


def test_lobra_cost_model_and_scheme_pool():
    from hetu_b200.engine.lobra import LoraCostModel, Scheme, optimized_scheme_pool
    cm, true, cands, _ = _lobra_setup()
    for tp, c in true.items():
        for mbs, s in ((3, 700), (6, 3000)):
            ref = float(np.dot(c, LoraCostModel.features(mbs, s)))
            assert abs(cm.layer_time(mbs, s, tp) - ref) < 0.03 * ref
    assert cm.layer_time(0, 512, 1) == 0.0
    assert cm.batch_time(2, 1024, Scheme(2, 2, 4096), 32) == pytest.approx(cm.layer_time(2, 1024, 2) * 16)
    pool = optimized_scheme_pool(cands)
    # per capacity the best-throughput scheme survives, plus cheaper alternatives; dominated / zero-capacity ones are dropped
    assert Scheme(1, 1, 2048, 10.0) in pool and Scheme(2, 1, 2048, 8.0) not in pool
    assert Scheme(2, 1, 8192, 8.5) in pool and Scheme(4, 1, 8192, 7.0) not in pool
    assert Scheme(2, 2, 16384, 7.5) in pool and all(s.max_tokens > 0 for s in pool)
    an = LoraCostModel.analytic(4096, 11008)
    assert an.layer_time(1, 4096, 8) < an.layer_time(1, 4096, 1) and an.layer_time(2, 512, 2) > an.layer_time(1, 512, 2)


def test_lobra_static_planners_and_dynamic_dispatch():
    from hetu_b200.engine import lobra as L
    cm, _, cands, tasks = _lobra_setup()
    total = L._PlannerCore.merge_tasks(tasks)
    kw = dict(cost_model=cm, num_layers=32, train_task_num=3, global_batch_size_list=[440, 96, 52], ngpus=16, scheme_candidates=cands)
    plans = {}
    for name, cls in (("group", L.GroupStaticPlanner), ("balance", L.BalanceStaticPlanner), ("prune", L.PruneStaticPlanner)):
        pl = cls(**kw)
        p = pl.schedule(tasks)
        plans[name] = (p, pl)
        assert p.gpus == 16 and np.isfinite(p.time) and p.time > 0
        got = {}
        for d in p.dispatch:
            for s, n in d.items():
                got[s] = got.get(s, 0) + n
        assert got == total                                               # every sequence is dispatched exactly once
        for j, d in enumerate(p.dispatch):                                # ... to a replica that can hold it
            assert all(s <= p.schemes[j].max_tokens for s in d) and (p.dp[j] > 0 or not d)
        for ti, t in enumerate(tasks):                                    # the per-task split adds up as well
            for s, n in t.items():
                assert sum(p.task_dispatch[ti][j].get(s, 0) for j in range(len(p.schemes))) == n
        assert any(sc.max_tokens >= 16384 and d for sc, d in zip(p.schemes, p.dp))
    g, b, pr = plans["group"][0], plans["balance"][0], plans["prune"][0]
    assert b.time <= g.time * 1.02 and pr.time <= b.time * 1.02
    assert plans["prune"][1].pruned > 0 and plans["prune"][1].evaluated < plans["balance"][1].evaluated
    # the heterogeneous mix beats the best homogeneous deployment (only the 16384-token schemes can hold every sequence)
    core = plans["balance"][1]
    homo = []
    for j, sc in enumerate(core.schemes):
        if sc.max_tokens >= 16384:
            dp = [0] * len(core.schemes)
            dp[j] = 16 // sc.ngpus
            homo.append(core.dispatch(dp, total, "balance").time)
    assert b.time < 0.9 * min(homo)
    assert len(b.strategy()) >= 2 and len(b.pipelines(32)) == sum(b.dp)
    assert sorted(d for p in b.pipelines(32) for st in p["stages"] for d in st["devices"]) == list(range(16))
    # dynamic dispatch of one step on the deployed mix
    strategy = b.strategy()
    mts = [sc.max_tokens for sc, d in zip(b.schemes, b.dp) if d]
    step = [{256: 40, 512: 11}, {512: 5, 2048: 4, 8192: 1}, {1024: 6, 16384: 1}]
    for cls in (L.GroupDynamicDispatcher, L.BalanceDynamicDispatcher):
        p = cls(cm, 32, strategy, mts, train_task_num=3).schedule(step)
        assert sum(n for d in p.dispatch for n in d.values()) == 68 and np.isfinite(p.time)
    bal = L.BalanceDynamicDispatcher(cm, 32, strategy, mts, 3).schedule(step)
    grp = L.GroupDynamicDispatcher(cm, 32, strategy, mts, 3).schedule(step)
    assert bal.time <= grp.time * 1.02


def test_lobra_batch_schedulers():
    from hetu_b200.engine import lobra as L
    cm, _, cands, _ = _lobra_setup()
    rng = np.random.RandomState(1)
    buckets = [256, 1024, 4096]
    lens = [[int(v) for v in rng.randint(10, 250, 20)] + [900, 700], [int(v) for v in rng.randint(300, 1000, 8)] + [3000, 4000]]
    batches = [[list(rng.randint(1, 100, n)) for n in task] for task in lens]
    dist = L.seq_distribution(batches, buckets)
    assert dist[0] == {256: 20, 1024: 2} and dist[1] == {1024: 8, 4096: 2}
    strategy, mts = [(2, 1, 1), (1, 2, 1)], [1024, 4096]
    plan = L.BalanceDynamicDispatcher(cm, 8, strategy, mts, 2).schedule(dist)
    per = L.global_batch_scheduler(batches, plan, buckets)
    assert len(per) == 2 and len(per[0]) == 2 and len(per[1]) == 1
    assert sum(len(r) for sch in per for r in sch) == 32
    want = sorted(int(sum(s)) for task in batches for s in task)
    assert sorted(int(sum(t)) for sch in per for rep in sch for _, t, _ in rep) == want
    for j, sch in enumerate(per):
        for rep in sch:
            mbs = L.greedy_local_batch_scheduler(rep, mts[j], 2)
            assert all(m.token_num() <= mts[j] and m.batch_data.shape == (m.batch_size, m.seq_length) for m in mbs)
            assert sum(m.batch_size for m in mbs) == len(rep)
            for m in mbs:                                                # task rows are contiguous: [offset, offset + size)
                assert sum(m.batch_size_list) == m.batch_size
                assert all(m.batch_offset_list[t] + m.batch_size_list[t] <= m.batch_size for t in m.task_id())
            packed = L.local_batch_pack_scheduler(rep, mts[j], 2)
            assert sum(cu[-1] for _, cu in packed) == sum(len(t) for _, t, _ in rep)
            assert all(cu[-1] <= mts[j] and m.batch_data.shape == (1, mts[j]) for m, cu in packed)
            assert len(packed) <= len(mbs)                               # packing never needs more micro-batches than padding


def test_trainer_writes_chrome_trace_and_records_step_exceptions(tmp_path):
    """torch_profile window -> one Chrome trace per rank (complete events on compute / attention / comm / optimizer tracks);
    a RuntimeError inside a step is appended to <output_dir>/logs/exception.txt before it propagates"""
    ht.init_comm_group(1)
    ds = SyntheticDataset(64, 259, 32, min_seq_len=8, seed=1, length_distribution="uniform")
    cfg = TrainingConfig(packing=False, micro_batch_size=2, global_load_size=4, max_seq_length=32, steps=4, learning_rate=1e-2, log_interval=0,
                         pack_alignment=16, torch_profile=True, start_profile_step=1, end_profile_step=2, profile_save_path=str(tmp_path / "trace"),
                         output_dir=str(tmp_path / "out"))
    tr = Trainer(cfg, ModelWrapper(GPTLMHeadModel, _mcfg()), ByteTokenizer(), OptimizerWrapper({"type": "adam", "lr": 1e-2}), ds)
    assert len(tr.train()) == 4
    trace = json.load(open(tmp_path / "trace" / "trace_rank0.json"))
    ev = [e for e in trace["traceEvents"] if e["ph"] == "X"]
    assert len(ev) > 20 and all(e["dur"] >= 0 and e["pid"] == 0 for e in ev)
    assert {e["cat"] for e in ev} >= {"compute", "attention"} and any(e["name"].startswith("linear") for e in ev)
    assert all(b["ts"] >= a["ts"] for a, b in zip(ev, ev[1:]))
    assert "breakdown" in json.load(open(tmp_path / "trace" / "summary_rank0.json"))

    def boom(batch, sid=0):
        raise RuntimeError("injected device failure")
    tr._train_step = boom
    with pytest.raises(RuntimeError, match="injected"):
        tr.train(steps=1)
    txt = (tmp_path / "out" / "logs" / "exception.txt").read_text()
    assert "rank 0 step 4" in txt and "injected device failure" in txt


def test_strategy_config_hetero_fields_become_a_heterogeneous_ds_config(tmp_path):
    """YAML `ds_parallel` with hetero_layers / hetero_tp / micro_batch_num_list / rank_to_device_mapping (the reference's hetero
    experiment configs) -> StrategyConfig -> heterogeneous ds_parallel_config -> HeteroSession pipelines + batch shares"""
    from hetu_b200.engine import HeteroSession, load_experiment
    from hetu_b200.utils.parallel import convert_strategy
    (tmp_path / "exp.yaml").write_text(
        "ds_parallel:\n  hetero: true\n  tp: 2\n  hetero_layers: [[3, 5], [8]]\n  hetero_tp: [2, 1]\n  micro_batch_num_list: [6, 2]\n"
        "  rank_to_device_mapping: {0: 0, 1: 1, 2: 2, 3: 3, 4: 7}\n  zero: false\n"
        "trainer:\n  steps: 2\n  global_load_size: 8\nmodel:\n  n_layer: 8\n")
    exp = load_experiment(str(tmp_path / "exp.yaml"), ["ds_parallel.micro_batch_num_list=[3, 1]"])
    sc = exp["strategy"]
    assert sc.hetero and sc.hetero_layers == [[3, 5], [8]] and sc.micro_batch_num_list == [3, 1] and sc.hetero_tp == [2, 1]
    cfg = convert_strategy(sc, 8)
    assert cfg["hetero"] and cfg["blocks"]["blocks2"]["attn"]["qkv"]["device_group_union"] == [[0, 1], [7]]
    assert cfg["blocks"]["blocks3"]["attn"]["qkv"]["device_group_union"] == [[2, 3], [7]]
    assert cfg["blocks"]["blocks3"]["attn"]["qkv"]["split"] == {"0": [2, 1]}
    s = HeteroSession(cfg, rank=7, shares=sc.micro_batch_num_list)
    assert s.pipelines == [[[0, 1], [2, 3]], [[7]]] and s.pipeline == 1 and s.split_batch(8) == [6, 2]
    with pytest.raises(AssertionError, match="covers"):
        sc.hetero_layers = [[3, 4], [8]]
        convert_strategy(sc, 8)


def test_pretrain_yaml_configs_load_and_train():
    """every experiment YAML under examples/pretrain/config parses into trainer / strategy / model sections; the context-
    parallel one trains through `build_trainer` on one process with cp overridden to 1"""
    import glob
    from hetu_b200.engine import build_trainer, load_experiment
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples", "pretrain", "config")
    files = sorted(glob.glob(os.path.join(root, "*.yaml")))
    assert {os.path.basename(f) for f in files} >= {"gpt_small_dp2_tp2.yaml", "llama_pack_tp.yaml", "llama_pad_cp.yaml", "llama_pack_cp.yaml", "gpt_hetero.yaml"}
    for f in files:
        exp = load_experiment(f)
        assert exp["model"]["type"] in ("gpt", "llama") and exp["trainer"].steps > 0 and exp["strategy"] is not None
    assert load_experiment(os.path.join(root, "gpt_hetero.yaml"))["strategy"].micro_batch_num_list == [3, 1]
    assert load_experiment(os.path.join(root, "llama_pad_cp.yaml"))["strategy"].cp == 2
    ht.init_comm_group(1)
    tr = build_trainer(os.path.join(root, "llama_pad_cp.yaml"), ["ds_parallel.cp=1", "trainer.steps=2", "trainer.log_interval=0"],
                       train_dataset=SyntheticDataset(32, 259, 64, seed=0, length_distribution="fixed"))
    losses = tr.train()
    assert len(losses) == 2 and all(np.isfinite(losses)) and tr.model_wrapper.model_config.num_key_value_heads == 2


def test_dynamic_micro_batch_padding_feeds_fewer_tokens_for_the_same_losses():
    """padding mode with dynamic_micro_batch_padding: rows sorted by length, every micro-batch padded to its own longest row (the
    executor takes one sequence length per micro-batch) -- same loss curve as padding everything to the step's longest row"""
    ht.init_comm_group(1)

    def run(dynamic):
        ht.set_seed(5)
        ds = SyntheticDataset(64, 259, 64, min_seq_len=6, seed=2, length_distribution="uniform")
        cfg = TrainingConfig(packing=False, micro_batch_size=2, global_load_size=8, max_seq_length=64, steps=3, learning_rate=1e-2, log_interval=0,
                             pack_alignment=8, dynamic_micro_batch_padding=dynamic)
        tr = Trainer(cfg, ModelWrapper(GPTLMHeadModel, _mcfg()), ByteTokenizer(), OptimizerWrapper({"type": "sgd", "lr": 0.1}), ds)
        fed = []
        tr.callbacks.append(lambda t, loss, stats: fed.append(stats["fed_tokens"]))
        return tr.train(), fed
    base, fed_base = run(False)
    dyn, fed_dyn = run(True)
    assert len(base) == len(dyn) == 3
    assert sum(fed_dyn) < 0.9 * sum(fed_base)
    # the loss of a step is the mean over micro-batches of the per-micro-batch token means; the micro-batch composition differs
    # (sorted rows), so compare the quantity that does not depend on it: training must track closely
    for a, b in zip(base, dyn):
        assert abs(a - b) < 0.05 * abs(a)


def test_multi_lora_routes_tokens_to_their_tasks_adapters():
    """multi-tenant LoRA (LobRA): one frozen base model, one adapter pair per task, a [tokens, tasks] one-hot mask routes every
    token; training on task-0 tokens only changes task-0 adapters, task-1 tokens still see the base model"""
    from hetu_b200.peft import get_peft_model
    from hetu_b200.peft.lora.config import LoraConfig
    ht.init_comm_group(1)
    ht.set_seed(9)
    rng = np.random.RandomState(0)
    S, B = 16, 4
    X = rng.randint(0, 64, (B, S))
    X[:, 1::2] = (X[:, 0::2] + 1) % 64
    L = np.roll(X, -1, axis=1)
    P = np.tile(np.arange(S), (B, 1))
    with ht.graph("define_and_run", create_new=True) as g:
        from hetu_b200.models import generate_ds_parallel_config
        base = GPTLMHeadModel(GPTConfig(vocab_size=64, n_positions=S, n_embd=32, n_layer=2, n_head=2), [generate_ds_parallel_config(2, 1, 1, 1, 1, zero=False)])
        model = get_peft_model(base, LoraConfig(rank=4, num_tasks=2, init_std=0.05))
        ids, pos, lab = (ht.placeholder("int64", [B * S], name=n) for n in ("ids", "pos", "lab"))
        mask = ht.placeholder("float32", [B * S, 2], name="task_mask")
        model.set_task_mask(mask)
        loss = model(ids, pos, lab, seq_len=S)
        train = ht.AdamOptimizer(lr=2e-2).minimize(loss)
    feed = {ids: torch.as_tensor(X.reshape(-1)), pos: torch.as_tensor(P.reshape(-1)), lab: torch.as_tensor(L.reshape(-1))}
    task0 = torch.zeros(B * S, 2); task0[:, 0] = 1.0
    task1 = torch.zeros(B * S, 2); task1[:, 1] = 1.0
    params = dict(model.named_parameters())
    trainable = [n for n, p in params.items() if p.requires_grad]
    assert trainable and all("lora" in n for n in trainable) and any("task0" in n for n in trainable) and any("task1" in n for n in trainable)
    before = {n: g.get_param(p).clone() for n, p in params.items()}
    base_loss = float(g.run(loss, [loss], {**feed, mask: task1})[0])
    losses = [float(g.run(loss, [loss, train], {**feed, mask: task0})[0]) for _ in range(25)]
    assert losses[-1] < 0.97 * losses[0]
    after = {n: g.get_param(p) for n, p in params.items()}
    changed = {n for n in params if not torch.equal(before[n], after[n])}
    assert changed and all("task0" in n for n in changed)                       # base weights and task-1 adapters untouched
    assert float(g.run(loss, [loss], {**feed, mask: task1})[0]) == pytest.approx(base_loss, rel=1e-5)   # task 1 still = base model
    assert float(g.run(loss, [loss], {**feed, mask: task0})[0]) < 0.97 * base_loss
    # mixed batch: the first half of the tokens belongs to task 0, the rest to task 1
    mixed = torch.cat([task0[:B * S // 2], task1[B * S // 2:]])
    lm = float(g.run(loss, [loss], {**feed, mask: mixed})[0])
    assert float(g.run(loss, [loss], {**feed, mask: task0})[0]) < lm < base_loss


def test_data_loader_reports_global_consumed_samples_for_every_dp_rank():
    """TOKEN load level: batches hold a variable number of samples; the resume counter must be the GLOBAL count and equal on
    all data-parallel ranks (each rank only sees its b[dp_rank::dp_size] slice), and must not run ahead with prefetching"""
    from hetu_b200.data.dataloader import DataLoader
    rng = np.random.RandomState(0)
    ds = [list(range(int(n))) for n in rng.randint(3, 20, size=57)]
    seen = {}
    for dp_rank in range(3):
        dl = DataLoader(ds, global_token_num=64, load_level="TOKEN", dp_rank=dp_rank, dp_size=3, prefetch=2)
        counts, mine = [], 0
        for batch in dl:
            mine += len(batch)
            counts.append(dl.consumed_yielded)
        seen[dp_rank] = (counts, mine)
    assert seen[0][0] == seen[1][0] == seen[2][0] and seen[0][0][-1] == 57
    assert sum(v[1] for v in seen.values()) == 57 and len({v[1] for v in seen.values()}) > 1     # slices differ, counter does not
    # restart from a mid-epoch count reproduces the tail of the stream
    dl = DataLoader(ds, global_token_num=64, load_level="TOKEN", prefetch=0)
    full = [b for b in dl]
    cut = seen[0][0][2]
    dl2 = DataLoader(ds, global_token_num=64, load_level="TOKEN", prefetch=0)
    dl2.restart(cut)
    assert [b for b in dl2] == full[3:]


def test_hydraulis_ilp_dispatch_is_optimal_and_respects_limits():
    """the MILP formulations of examples/hydraulis/strategy/dynamic_scip.py (solved with scipy/HiGHS): optimal on an instance
    small enough to brute-force, never worse than the greedy heuristic, honours max_seq and the token bounds"""
    import itertools
    from hetu_b200.engine.hydraulis import StrategyCost, batching_strategy_ilp, dispatch_batch, dispatch_batch_ilp
    long_s = StrategyCost("tp8", 1, a=1e-3, b=2e-8, max_seq=32768)
    short_s = StrategyCost("tp2", 1, a=2e-3, b=8e-8, max_seq=8192)
    mid_s = StrategyCost("tp4", 1, a=1.4e-3, b=4e-8, max_seq=16384)
    sts, pps = [long_s, mid_s, short_s], [2, 1, 1]
    seqs = [30000, 12000, 9000, 7000, 4000, 2500, 2000, 1000]
    r = dispatch_batch_ilp(seqs, sts, pps)
    assert all(seqs[i] <= sts[j].max_seq for i, j in enumerate(r["assignment"]))

    def makespan(assign):
        loads = [0.0] * 3
        for i, j in enumerate(assign):
            loads[j] += sts[j].seq_ms(seqs[i])
        return max(loads[j] + (pps[j] - 1) * sts[j].seq_ms(min(max(seqs), sts[j].max_seq)) for j in range(3))
    brute = min(makespan(a) for a in itertools.product(range(3), repeat=len(seqs))
                if all(seqs[i] <= sts[j].max_seq for i, j in enumerate(a)))
    assert abs(r["makespan_ms"] - brute) < 1e-6 * brute and abs(makespan(r["assignment"]) - brute) < 1e-6 * brute
    h = dispatch_batch(seqs, [long_s, mid_s, short_s], sequential=False)
    assert dispatch_batch_ilp(seqs, sts)["makespan_ms"] <= h["makespan_ms"] + 1e-9
    b = batching_strategy_ilp([4000, 3000, 2500, 2000, 1500, 1000, 800, 600], short_s, pp=2, max_tokens=8192, min_tokens=2048)
    toks = [sum([4000, 3000, 2500, 2000, 1500, 1000, 800, 600][i] for i in mb) for mb in b["micro_batches"]]
    assert all(2048 <= t <= 8192 for t in toks) and sorted(i for mb in b["micro_batches"] for i in mb) == list(range(8))
    assert b["e2e_ms"] == pytest.approx(b["max_micro_batch_ms"] * (2 - 1 + b["num_micro_batches"]))


def test_hydra_lite_composition_interpolation_overrides_and_structured_merge(tmp_path, monkeypatch):
    """ref: SURVEY 5.6 -- hydra/OmegaConf YAML tier: `defaults` composition with config groups, `${...}` interpolation incl.
    `oc.env`, dotted overrides (+add, ~delete, group=option), merge onto dataclasses with type checks"""
    import dataclasses
    from hetu_b200.utils import hydra_lite as H
    (tmp_path / "model").mkdir()
    (tmp_path / "ds_parallel").mkdir()
    (tmp_path / "base.yaml").write_text("trainer:\n  steps: 10\n  bf16: true\nseed: 1\n")
    (tmp_path / "model" / "gpt_small.yaml").write_text("type: gpt\nn_embd: 256\nn_layer: 4\n")
    (tmp_path / "model" / "llama_7b.yaml").write_text("type: llama\nn_embd: 4096\nn_layer: 32\n")
    (tmp_path / "ds_parallel" / "dp2_tp2.yaml").write_text("# @package _global_\nds_parallel:\n  dp: 2\n  tp: 2\n  pp: 1\n")
    (tmp_path / "exp.yaml").write_text(
        "defaults:\n  - base\n  - model: gpt_small\n  - ds_parallel: dp2_tp2\n  - _self_\n"
        "trainer:\n  steps: 20\n  max_seq_length: ${model.n_positions}\n  out_dir: runs/${model.type}_${trainer.steps}\n"
        "model:\n  n_positions: 512\n  hidden4: ${model.n_embd}\nworld: ${oc.env:HYDRA_LITE_WORLD,4}\n")
    cfg = H.load(["--config-path", str(tmp_path), "--config-name", "exp"])
    assert cfg.trainer.steps == 20 and cfg.trainer.bf16 is True and cfg.seed == 1                     # _self_ after base
    assert cfg.model.type == "gpt" and cfg.model.n_layer == 4 and cfg.ds_parallel.tp == 2             # group under its key / _global_ package
    assert cfg.trainer.max_seq_length == 512 and cfg.model.hidden4 == 256 and cfg.trainer.out_dir == "runs/gpt_20" and cfg.world == 4
    monkeypatch.setenv("HYDRA_LITE_WORLD", "8")
    cfg = H.load(["--config-path", str(tmp_path), "--config-name", "exp", "model=llama_7b", "trainer.steps=5", "+trainer.log_interval=2", "~seed",
                  "ds_parallel.recompute.layer_idxs=[[0,1],[2]]"])
    assert cfg.model.type == "llama" and cfg.model.hidden4 == 4096 and cfg.trainer.out_dir == "runs/llama_5" and cfg.world == 8
    assert cfg.trainer.log_interval == 2 and "seed" not in cfg and cfg.ds_parallel.recompute.layer_idxs == [[0, 1], [2]]
    assert "n_embd: 4096" in H.to_yaml(cfg)
    with pytest.raises(FileNotFoundError):
        H.load(["--config-path", str(tmp_path), "--config-name", "exp", "model=nope"])
    with pytest.raises(ValueError, match="cycle"):
        H.resolve({"a": "${b}", "b": "${a}"})
    with pytest.raises(KeyError):
        H.resolve({"a": "${missing.key}"})

    @dataclasses.dataclass
    class Inner:
        granularity: str = "full"
        num_layers: int = 0

    @dataclasses.dataclass
    class Strat:
        dp: int = 1
        tp: int = 1
        zero: bool = False
        lr: float = 1e-3
        recompute: Inner = dataclasses.field(default_factory=Inner)
    st = H.merge_dataclass(Strat, {"dp": "2", "zero": "true", "lr": 3, "recompute": {"num_layers": "4"}})
    assert (st.dp, st.tp, st.zero, st.lr, st.recompute.num_layers, st.recompute.granularity) == (2, 1, True, 3.0, 4, "full")
    with pytest.raises(KeyError, match="no field"):
        H.merge_dataclass(Strat, {"pipeline": 2})
    with pytest.raises(TypeError):
        H.merge_dataclass(Strat, {"dp": 2.5})


def test_indexed_corpus_builder_preprocessor_gpt_samples_blending_and_splits(tmp_path):
    """ref: examples/hydraulis/data_utils (MMapIndexedDataset + builder, HetuDataset sample mapping, BlendedDataset, split string)"""
    import json as _json
    from hetu_b200.data import (BlendedDataset, GPTSampleDataset, IndexedDatasetBuilder, blending_indices, open_indexed, preprocess_jsonl,
                                split_documents)
    rng = np.random.RandomState(0)
    docs = [rng.randint(5, 250, rng.randint(3, 40)).tolist() for _ in range(57)]
    b = IndexedDatasetBuilder(str(tmp_path / "a"), vocab_size=259)
    for d in docs[:30]:
        b.add_document(d)
    ds_a = b.finalize()
    b2 = IndexedDatasetBuilder(str(tmp_path / "b"), vocab_size=259)
    for d in docs[30:]:
        b2.add_document(d)
    b2.finalize()
    m = IndexedDatasetBuilder(str(tmp_path / "merged"), vocab_size=259)
    m.merge(str(tmp_path / "a")); m.merge(str(tmp_path / "b"))
    merged = m.finalize()
    assert len(ds_a) == 30 and len(merged) == 57 and all(merged[i].tolist() == docs[i] for i in range(57))
    assert merged.data.dtype == np.uint16 and _json.load(open(tmp_path / "merged.meta.json"))["tokens"] == sum(map(len, docs))
    # JSONL preprocessing with 2 worker processes == 1 worker, documents in file order, EOD appended
    lines = [{"text": "".join(chr(97 + (i + j) % 26) for j in range(5 + i % 9))} for i in range(400)]
    with open(tmp_path / "corpus.jsonl", "w") as f:
        for r in lines:
            f.write(_json.dumps(r) + "\n")
    one = preprocess_jsonl(str(tmp_path / "corpus.jsonl"), str(tmp_path / "tok1"), workers=1)
    two = preprocess_jsonl(str(tmp_path / "corpus.jsonl"), str(tmp_path / "tok2"), workers=2)
    assert len(one) == len(two) == 400 and all(np.array_equal(one[i], two[i]) for i in range(400))
    from hetu_b200.data import ByteTokenizer
    tok = ByteTokenizer()
    assert one[7].tolist() == list(tok.encode(lines[7]["text"])) + [tok.eos_id]
    assert not any(os.path.exists(str(tmp_path / f"tok2.shard{i}.bin")) for i in range(2))
    # GPT samples: windows of seq+1 tokens over the shuffled document stream; an epoch covers its token stream exactly once
    S = 16
    g = GPTSampleDataset(merged, S, seed=3)
    n = g.samples_per_epoch
    assert n == (sum(map(len, docs)) - 1) // S and len(g) == n
    samples = [g[i] for i in range(n)]
    assert all(len(s) == S + 1 for s in samples)
    order, starts, sample_order = g._epoch(0)
    stream = np.concatenate([np.asarray(docs[d]) for d in order])
    for i in range(n):
        lo = int(sample_order[i]) * S
        assert np.array_equal(samples[i], stream[lo:lo + S + 1])
    assert sorted(int(x) for x in sample_order) == list(range(n))
    again = GPTSampleDataset(open_indexed(str(tmp_path / "merged")), S, seed=3)
    assert np.array_equal(again[5], samples[5])                               # resumable: a pure function of (seed, index)
    two_epochs = GPTSampleDataset(merged, S, num_samples=2 * n, seed=3)
    assert not np.array_equal(two_epochs[n], two_epochs[0]) or n == 1         # the second epoch is reshuffled
    # blending: realised shares follow the weights at every prefix
    which, inner = blending_indices([0.7, 0.2, 0.1], 1000)
    for k, w in enumerate([0.7, 0.2, 0.1]):
        assert abs((which == k).mean() - w) < 0.002 and abs((which[:100] == k).mean() - w) < 0.02
        assert inner[which == k].tolist() == list(range(int((which == k).sum())))
    mix = BlendedDataset([g, GPTSampleDataset(ds_a, S, seed=1)], [0.75, 0.25], 40)
    assert len(mix) == 40 and len(mix[39]) == S + 1
    tr, va, te = split_documents(1000, "969,30,1")
    assert (len(tr), len(va), len(te)) == (969, 30, 1) and tr[-1] + 1 == va[0] and te[-1] == 999
    tr, va, te = split_documents(10, "8,2")
    assert (len(tr), len(va), len(te)) == (8, 2, 0)


def test_trainer_evaluate_and_periodic_evaluation():
    """forward-only evaluation: token-weighted loss / perplexity over an evaluation set, no parameter changes; `eval_interval`
    runs it inside the training loop and the evaluation loss falls as the model learns the pattern"""
    from hetu_b200.engine import ModelWrapper, OptimizerWrapper, Trainer, TrainingConfig
    from hetu_b200.models import GPTConfig, GPTLMHeadModel, generate_ds_parallel_config

    class Pattern:
        """sequences that count upwards modulo 50 from a per-sample start: learnable next-token structure"""

        def __init__(self, n, seq, offset=0):
            self.n, self.seq, self.offset = n, seq, offset

        def __len__(self):
            return self.n

        def __getitem__(self, i):
            return ((np.arange(self.seq + 1) + (i + self.offset) * 7) % 50 + 3).astype(np.int64)
    ht.set_seed(5)
    mcfg = GPTConfig(vocab_size=64, n_positions=16, n_embd=32, n_layer=2, n_head=2)
    cfg = TrainingConfig(packing=False, micro_batch_size=4, global_load_size=8, max_seq_length=16, steps=30, learning_rate=3e-3, log_interval=0,
                         pack_alignment=16, eval_interval=10, eval_iters=2)
    tr = Trainer(cfg, ModelWrapper(GPTLMHeadModel, mcfg), ByteTokenizer(), OptimizerWrapper({"type": "adam", "lr": 3e-3}), Pattern(256, 16),
                 ds_parallel_configs=[generate_ds_parallel_config(2, 1, 1, 1, 1, zero=False)])
    tr.eval_dataset = Pattern(32, 16, offset=1000)
    before = tr.evaluate()
    assert before["batches"] == 4 and before["tokens"] == 32 * 16 and abs(before["perplexity"] - np.exp(before["loss"])) < 1e-6
    w0 = tr.trainer_states.graph.get_param(next(iter(tr.trainer_states.model.parameters()))).clone()
    again = tr.evaluate(max_batches=4)
    assert abs(again["loss"] - before["loss"]) < 1e-6                                   # evaluation changes nothing
    assert torch.equal(tr.trainer_states.graph.get_param(next(iter(tr.trainer_states.model.parameters()))), w0)
    tr.train()
    periodic = [e for e in tr.eval_history if e["step"] in (10, 20, 30)]
    assert [e["step"] for e in periodic] == [10, 20, 30] and all(e["batches"] == 2 for e in periodic)
    after = tr.evaluate()
    assert after["loss"] < 0.7 * before["loss"] and periodic[-1]["loss"] < periodic[0]["loss"]


def test_bert_wordpiece_tokenizer_matches_the_huggingface_reference_implementation(tmp_path):
    """own WordPiece implementation (basic tokenisation + greedy longest match) against transformers.BertTokenizer on the same vocab"""
    transformers = pytest.importorskip("transformers")
    from hetu_b200.data.tokenizers import build_tokenizer
    words = ["the", "quick", "brown", "fox", "jump", "##s", "##ed", "##ing", "over", "lazy", "dog", "un", "##believ", "##able", "cafe", "na", "##ive", "2024",
             "hello", "world", "!", ",", ".", "?", "-", "'", "s", "中", "文", "token", "##izer", "##ization", "a", "b", "##c", "new", "york"]
    vocab = ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"] + words
    (tmp_path / "vocab.txt").write_text("\n".join(vocab) + "\n", encoding="utf-8")
    ours = build_tokenizer("bert", vocab_file=str(tmp_path / "vocab.txt"))
    ref = transformers.BertTokenizer(str(tmp_path / "vocab.txt"), do_lower_case=True)
    texts = ["The quick brown fox jumps over the lazy dog.", "Unbelievable!  Tokenization, tokenizer's naive café?", "hello 中文 world", "jumping jumped xyz-abc 2024",
             "New   York\tnew\nyork", "[MASK] the [SEP] fox"]
    for t in texts:
        assert ours.tokenize(t) == ref.tokenize(t), t
        assert ours.encode(t) == ref.encode(t), t
    a, b = "the quick fox", "jumps over the lazy dog"
    ids, types, mask = ours.encode_plus(a, b, max_length=12, padding=True)
    enc = ref(a, b, max_length=12, padding="max_length", truncation="longest_first")
    assert ids == enc["input_ids"] and types == enc["token_type_ids"] and mask == enc["attention_mask"]
    ids, types, mask = ours.encode_plus(a, b, max_length=7)
    enc = ref(a, b, max_length=7, truncation="longest_first")
    assert ids == enc["input_ids"] and types == enc["token_type_ids"]
    assert ours.decode(ours.encode("unbelievable tokenization")) == "unbelievable tokenization"
    assert ours.vocab_size == len(vocab) and ours.pad == 0 and ours.mask_id == 4


def test_gpt2_byte_level_bpe_matches_the_huggingface_reference_implementation(tmp_path):
    """own byte-level BPE (pre-tokenisation pattern, byte symbols, ranked merges) against transformers.GPT2Tokenizer on a vocabulary
    learned here from a tiny corpus"""
    transformers = pytest.importorskip("transformers")
    import json as _json
    from collections import Counter
    from hetu_b200.data.tokenizers import build_tokenizer
    from hetu_b200.data.tokenizers.bpe import _PATTERN, bytes_to_unicode
    import regex
    corpus = "the quick brown fox jumps over the lazy dog. The dog's tokenization isn't naive: 2024 tokens, tokenizer tokens! héllo wörld 中文 \n\n  spaced   out"
    be = bytes_to_unicode()
    words = Counter("".join(be[b] for b in w.encode("utf-8")) for w in regex.findall(_PATTERN, corpus * 3))
    vocab = {c: i for i, c in enumerate(sorted(set(be.values())))}
    splits = {w: list(w) for w in words}
    merges = []
    for _ in range(60):                                    # learn 60 merges: most frequent adjacent pair first
        pairs = Counter()
        for w, n in words.items():
            for p in zip(splits[w], splits[w][1:]):
                pairs[p] += n
        if not pairs:
            break
        (a, b), _n = max(sorted(pairs.items()), key=lambda kv: kv[1])
        merges.append((a, b))
        vocab.setdefault(a + b, len(vocab))
        for w in splits:
            s, out, i = splits[w], [], 0
            while i < len(s):
                if i < len(s) - 1 and s[i] == a and s[i + 1] == b:
                    out.append(a + b); i += 2
                else:
                    out.append(s[i]); i += 1
            splits[w] = out
    vocab["<|endoftext|>"] = len(vocab)
    (tmp_path / "vocab.json").write_text(_json.dumps(vocab), encoding="utf-8")
    (tmp_path / "merges.txt").write_text("#version: 0.2\n" + "\n".join(f"{a} {b}" for a, b in merges) + "\n", encoding="utf-8")
    ours = build_tokenizer("gpt2", vocab_file=str(tmp_path / "vocab.json"), merge_file=str(tmp_path / "merges.txt"))
    ref = transformers.GPT2Tokenizer(str(tmp_path / "vocab.json"), str(tmp_path / "merges.txt"))
    for t in (corpus, "the dog's tokens", "  leading spaces and a tab\there", "unseen ßtring with ünïcode 🙂", ""):
        assert ours.tokenize(t) == ref.tokenize(t), t
        assert ours.encode(t) == ref.encode(t), t
        assert ours.decode(ours.encode(t)) == t
    assert ours.encode("fox", add_special_tokens=True)[-1] == ours.eos_id == vocab["<|endoftext|>"] and ours.vocab_size == len(vocab)
