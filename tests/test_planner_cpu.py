"""Galvatron planner: C++ DP core vs brute force, memory-pressure behaviour of the search, plan emission."""
import pytest
import itertools

import numpy as np

import hetu_b200 as ht
from hetu_b200 import _C
from hetu_b200.planner import GalvatronSearchEngine, LayerProfile, galvatron_plan_to_ds_parallel_config


def test_dp_core_matches_brute_force():
    rng = np.random.RandomState(0)
    L, S, M = 5, 3, 14
    mem = rng.randint(1, 5, (L, S))
    intra = rng.rand(L, S)
    inter = rng.rand(L, S, S) * 0.3
    inter[0] = 0
    cost, picks, rem = _C.galvatron_dp(L, M + 1, S,   # table rows 0..max_mem-1: budget = max_mem - 1 units (as the reference)
                                         mem.reshape(-1).tolist(), intra.reshape(-1).tolist(), inter.reshape(-1).tolist())
    best = (1e18, None)
    for combo in itertools.product(range(S), repeat=L):
        m = sum(mem[i, c] for i, c in enumerate(combo))
        if m > M:
            continue
        c = sum(intra[i, s] for i, s in enumerate(combo)) + sum(inter[i, combo[i - 1], combo[i]] for i in range(1, L))
        best = min(best, (c, combo))
    assert abs(cost - best[0]) < 1e-9 and tuple(picks) == best[1]
    assert _C.galvatron_dp(L, 3, S, mem.reshape(-1).tolist(), intra.reshape(-1).tolist(), inter.reshape(-1).tolist())[1] == []


def test_search_prefers_plain_dp_when_memory_is_plentiful_and_shards_when_tight():
    lp = LayerProfile.transformer(4096, 11008, 4096, 32, swiglu=True)
    roomy = GalvatronSearchEngine(32, 8, lp, vocab=32000, hidden=4096, seq=4096, memory_mb=180 * 1024).search(batch_sizes=(16, 32))
    assert roomy is not None and roomy["pp"] == 1 and all(s.tp == 1 for s in roomy["strategies"])
    tight = GalvatronSearchEngine(32, 8, lp, vocab=32000, hidden=4096, seq=4096, memory_mb=24 * 1024).search(batch_sizes=(16, 32))
    assert tight is not None
    ss = tight["strategies"]
    assert tight["pp"] > 1 or any(s.tp > 1 or s.sdp == 3 or s.ckpt for s in ss)     # had to trade speed for memory
    assert tight["throughput_samples_per_s"] <= roomy["throughput_samples_per_s"] * 1.0001
    none = GalvatronSearchEngine(32, 8, lp, vocab=32000, hidden=4096, seq=4096, memory_mb=2 * 1024).search(batch_sizes=(16,))
    assert none is None


def test_plan_json_and_ds_parallel_config_roundtrip(tmp_path):
    lp = LayerProfile.transformer(2048, 8192, 1024, 16)
    eng = GalvatronSearchEngine(8, 4, lp, hidden=2048, seq=1024, memory_mb=40 * 1024)
    plan = eng.search(batch_sizes=(8, 16))
    js = GalvatronSearchEngine.to_json(plan)
    assert len(js["tp_sizes_enc"].split(",")) == 8 and js["pp_deg"] == plan["pp"]
    GalvatronSearchEngine.save(plan, str(tmp_path / "plan.json"))
    cfg = galvatron_plan_to_ds_parallel_config(plan, 4)
    assert len(cfg["blocks"]) == 8
    ds, dg = ht.nn.parallel.config2ds(cfg["blocks"]["blocks3"]["attn"]["qkv"])
    assert ds.get(0).device_num == plan["strategies"][3].tp * plan["strategies"][3].dp


def test_hetu_alias_package():
    import hetu
    from hetu.engine import Trainer   # noqa: F401
    from hetu.utils.parallel import read_ds_parallel_config   # noqa: F401
    import hetu.nn as nn
    assert hetu.AdamOptimizer is ht.AdamOptimizer and nn.Module is ht.nn.Module


def test_reference_module_paths_resolve_under_the_hetu_alias():
    """a user of the reference keeps their imports: every module path of python/hetu resolves (to the same module objects as
    hetu_b200.*) through the alias hook"""
    import importlib
    paths = ["hetu.logger", "hetu.context", "hetu.nn.parallel", "hetu.optim", "hetu.data.data_collator", "hetu.data.bucket", "hetu.data.dataloader",
             "hetu.data.dataset", "hetu.data.utils", "hetu.data.messages.message_template", "hetu.data.messages.prompt_template",
             "hetu.data.tokenizers.gpt2_tokenizer", "hetu.data.tokenizers.hf_tokenizer", "hetu.data.tokenizers.sentencepiece_tokenizer",
             "hetu.data.tokenizers.tiktoken_tokenizer", "hetu.data.tokenizers.tokenizer", "hetu.utils.common_utils", "hetu.utils.file_utils",
             "hetu.utils.data.dataloader", "hetu.utils.parallel.distributed", "hetu.utils.parallel.ds_config", "hetu.utils.parallel.generate_ds",
             "hetu.utils.parallel.read_ds", "hetu.utils.checkpoint.ht_safetensors", "hetu.utils.checkpoint.model_saver",
             "hetu.utils.checkpoint.load_checkpoint", "hetu.utils.checkpoint.save_checkpoint", "hetu.engine.trainer", "hetu.engine.trainer_config",
             "hetu.engine.sft_trainer", "hetu.engine.wrapper", "hetu.engine.parallel_config", "hetu.engine.straggler", "hetu.engine.strategy",
             "hetu.models.gpt.gpt_model", "hetu.models.gpt.gpt_config", "hetu.models.gpt.gpt_tokenizer", "hetu.models.gpt.generate_gpt_4d_config",
             "hetu.models.gpt.generate_gpt_hetero_4d_config", "hetu.models.llama.llama_model", "hetu.models.llama.llama_config",
             "hetu.models.llama.llama_tokenizer", "hetu.models.llama.generate_llama_4d_config", "hetu.models.llama.generate_llama_hetero_4d_config",
             "hetu.models.utils.converter.convert_llama_hf_to_ht", "hetu.peft.lora.layer", "hetu.peft.lora.model", "hetu.peft.lora.config",
             "hetu.rpc.pssh_start", "hetu.rpc.pssh_start_config", "hetu.rpc.pssh_start_elastic", "hetu.rpc.local_start", "hetu.rpc.pssh_workers",
             "hetu.rpc.elastic_arg_parser", "hetu.rpc.kv_store", "hetu.rpc.heturpc_elastic_server", "hetu.rpc.heturpc_polling_server",
             "hetu.rpc.heturpc_async_server"]
    # every python module of the reference resolves, its op-manifest tooling included: facades for the finer
    # grained reference layout -- nn.modules.*, nn.functional / init / parameter, optim.{optimizer,sgd}, engine.{utils,sft_config},
    # models.utils.*, data.tokenizers.{utils,pretrained_tokenizer}, rpc.kv_store.{client,server,const,producer_consumer}
    more = ["hetu._binding.codegen.gen_py_ops", "hetu._binding.codegen.args_bridge", "hetu.nn.modules", "hetu.nn.modules.linear", "hetu.nn.modules.activation", "hetu.nn.modules.parallel_multi_ds", "hetu.nn.modules.parallel_utils",
            "hetu.nn.modules.container", "hetu.nn.functional", "hetu.nn.init", "hetu.nn.parameter", "hetu.optim.optimizer", "hetu.optim.sgd",
            "hetu.engine.utils", "hetu.engine.sft_config", "hetu.models.utils.model_utils", "hetu.models.utils.config_utils", "hetu.models.utils.hub",
            "hetu.models.utils.common_utils", "hetu.data.tokenizers.utils", "hetu.data.tokenizers.pretrained_tokenizer", "hetu.rpc.kv_store.client",
            "hetu.rpc.kv_store.server", "hetu.rpc.kv_store.const", "hetu.rpc.kv_store.producer_consumer", "hetu.rpc.pssh_start_exp"]
    for p in more:
        importlib.import_module(p)
    from hetu.nn.modules.linear import Linear
    from hetu.nn.modules import HtMultiColumnParallelLinear, Module   # noqa: F401
    from hetu.optim.sgd import SGD
    assert Linear is ht.nn.Linear and SGD is ht.SGDOptimizer
    for p in paths:
        m = importlib.import_module(p)
        assert m is importlib.import_module("hetu_b200." + p[len("hetu."):]), p
    import hetu
    from hetu.engine.parallel_config import config_spread_zero
    from hetu.utils.parallel.ds_config import StrategyConfig
    assert config_spread_zero({"zero": True, "w": {"type": "variable"}})["w"]["zero"] is True and StrategyConfig().world() == 1
    assert hetu.logger is importlib.import_module("hetu_b200").logger and callable(hetu.logger.info)
    with pytest.raises(ModuleNotFoundError):
        importlib.import_module("hetu.no_such_module")


def test_ampelos_replans_after_device_loss_and_stragglers():
    """elastic re-planning (ref: python/hetu/engine/strategy_ampelos.py): dead devices shrink a tensor-parallel group, the
    planner may change the number of pipelines and their depths, every layer / micro-batch stays assigned, and the chosen
    plan is no slower than any other candidate it enumerated"""
    from hetu_b200.engine.strategy import TrainerCtxs, TrainerStrategyArgs
    from hetu_b200.engine.strategy_ampelos import AmpelosStrategyModel, partition_into_k_groups, replan_after_failure
    parts = partition_into_k_groups([5, 4, 3, 3, 2, 1], 2)
    assert sorted(sum([5, 4, 3, 3, 2, 1][i] for i in p) for p in parts) == [9, 9]
    ctxs = TrainerCtxs(normal_layers=8, normal_mbn=8, top_k=5, memory_bound=12)
    old = TrainerStrategyArgs(dp=2, tp=4, pp=2, hetero_layers=[[8, 8], [8, 8]], rank_to_device_mapping={r: r for r in range(16)})
    # healthy job whose stages can hold at most 8 layers: the homogeneous plan comes back (2 pipelines x 2 stages of tp 4,
    # 8 layers each, 8 micro-batches each); without a memory bound the planner would rather run 4 bubble-free pipelines
    m = AmpelosStrategyModel(TrainerCtxs(normal_layers=8, normal_mbn=8, memory_bound=8), old, {d: 1.0 for d in range(16)})
    st, cfg = m.make_plans()
    assert st.dp == 2 and st.hetero_layers == [[8, 8], [8, 8]] and st.hetero_micro_batch_num_list == [8, 8] and not st.unused_rank_list
    base_time = m.estimate_time()
    # device 5 dies, device 12 runs 2.5x slower
    m2 = replan_after_failure(ctxs, old, [d for d in range(16) if d != 5], {12: 2.5})
    st2, cfg2 = m2.make_plans()
    used = [d for pl in m2.plans for g in pl["groups"] for d in g.devices]
    assert 5 not in used and len(used) == len(set(used))
    assert all(sum(ls) == 16 for ls in st2.hetero_layers) and sum(st2.hetero_micro_batch_num_list) == 16
    assert all(len(ls) == n for ls, n in zip(st2.hetero_layers, st2.hetero_stages))
    assert m2.estimate_time() >= base_time                     # fewer / slower devices cannot be faster
    assert m2.candidates == sorted(m2.candidates) and len(m2.candidates) >= 2
    # the straggler never shares a group with healthy devices of its node
    for pl in m2.plans:
        for g in pl["groups"]:
            if 12 in g.devices:
                assert g.devices == [12] or all(m2._sr(d) >= ctxs.straggler_threshold for d in g.devices)
    # the emitted heterogeneous config is consumable
    assert "pipelines" in cfg2 or "devices" in cfg2 or isinstance(cfg2, dict)
    # a whole node lost: all work moves to the surviving node, still a valid plan
    m3 = replan_after_failure(ctxs, old, list(range(8)))
    st3, _ = m3.make_plans()
    assert all(d < 8 for pl in m3.plans for g in pl["groups"] for d in g.devices) and all(sum(ls) == 16 for ls in st3.hetero_layers)
    # memory bound: 10 layers per full-width stage cannot hold 16 layers on one stage -> at least 2 stages per pipeline
    tight = TrainerCtxs(normal_layers=8, normal_mbn=8, memory_bound=10)
    m4 = replan_after_failure(tight, old, list(range(8)))
    st4, _ = m4.make_plans()
    assert all(max(ls) <= 10 for ls in st4.hetero_layers)


def test_galvatron_runtime_validates_describes_builds_and_checks_a_plan(tmp_path):
    """ref: tools/Galvatron core/hybrid_parallel_config.py (check / print of a layer-wise plan) + hybrid_parallel_model.py"""
    import json
    from hetu_b200.planner import GalvatronRuntime, HardwareProfile
    from hetu_b200.models import GPTConfig, GPTLMHeadModel
    lp = LayerProfile.transformer(1024, 4096, 512, 8)
    js = {"pp_deg": 2, "tp_sizes_enc": "2,2,1,1", "tp_consecutive_flags": "1,1,1,1", "dp_types_enc": "0,0,1,1", "checkpoint": "0,1,0,0",
          "global_bsz": 16, "chunks": 4, "pp_division": "2,2", "default_dp_type": "zero2"}
    rt = GalvatronRuntime.from_json(js, num_gpus=4, layer=lp, hardware=HardwareProfile(), hidden=1024, seq=512)
    ss = rt.plan["strategies"]
    assert [(s.tp, s.dp, s.sdp, s.ckpt) for s in ss] == [(2, 1, 2, False), (2, 1, 2, True), (1, 2, 3, False), (1, 2, 3, False)]
    text = rt.describe()
    assert "pp 2" in text and "tp 2 (consecutive) +sp  dp 1 (zero-2)  recompute off" in text and "zero-3" in text and len(text.splitlines()) == 4
    groups = rt.comm_groups()
    assert groups[0]["tp"] == [[0, 1], [2, 3]] and groups[2]["dp"] == [[0, 1], [2, 3]] and groups[0]["pp"] == [[0, 2], [1, 3]]
    mem = rt.memory_per_stage_mb()
    assert len(mem) == 2 and all(m > 0 for m in mem)
    pred = rt.predicted_step_ms()
    assert pred > 0
    for ms in (pred * 1.3, pred * 1.2, pred * 1.25):
        rt.record_step(ms)
    rep = rt.cost_model_report()
    assert abs(rep["ratio"] - 1.25) < 1e-6 and rep["memory_per_stage_mb"] == mem
    # JSON round trip keeps the plan
    (tmp_path / "p.json").write_text(json.dumps(rt.to_json()))
    rt2 = GalvatronRuntime.from_json(str(tmp_path / "p.json"), num_gpus=4)
    assert [s.key() for s in rt2.plan["strategies"]] == [s.key() for s in ss] and rt2.plan["layer_split"] == [2, 2]
    # a plan that cannot run is rejected with the reason
    for bad, msg in (({**js, "tp_sizes_enc": "4,2,1,1"}, "tp 4"), ({**js, "pp_division": "3,2"}, "pp_division"), ({**js, "pp_deg": 3}, "pp_deg 3"),
                     ({**js, "global_bsz": 6}, "global batch")):
        with pytest.raises(ValueError, match=msg):
            GalvatronRuntime.from_json(bad, num_gpus=4)
    with pytest.raises(ValueError, match="budget"):
        GalvatronRuntime.from_json(js, num_gpus=4, layer=lp, memory_mb=1.0)
    # build: world 1 plan -> model + run arguments
    one = GalvatronRuntime.from_json({"pp_deg": 1, "tp_sizes_enc": "1,1", "tp_consecutive_flags": "1,1", "dp_types_enc": "0,0", "checkpoint": "0,1",
                                      "global_bsz": 4, "chunks": 2}, num_gpus=1)
    with ht.graph("define_and_run", create_new=True):
        model, cfg, run_kw = one.build(GPTLMHeadModel, GPTConfig(vocab_size=64, n_positions=16, n_embd=32, n_layer=2, n_head=2))
    assert run_kw == {"num_micro_batches": 2, "grad_scale": 1.0} and cfg["blocks"]["blocks1"]["recompute"] == [True]


def test_op_manifest_tooling_parses_checks_and_emits_stubs(tmp_path):
    """ref: python/hetu/_binding/codegen/{gen_py_ops,args_bridge}.py + ops.yml -- the manifest format is parsed, checked against the
    framework's Python surface, and turned into editor stubs; the C++ registry can be dumped as a manifest"""
    import yaml
    from hetu._binding.codegen.args_bridge import parse_args
    from hetu._binding.codegen.gen_py_ops import dump_registry, gen_ops
    a = parse_args("Tensor input, HTAxes axes=None, bool keepdims=false, List[int] pads=[0, 0]")
    assert [(x.type_str, x.name, x.default) for x in a] == [("Tensor", "input", None), ("HTAxes", "axes", "None"), ("bool", "keepdims", "false"),
                                                           ("List[int]", "pads", "[0, 0]")]
    assert a[2].signature() == "keepdims: bool = False" and a[0].py_type == "Tensor"
    manifest = [{"name": "add", "op": "AddElewiseOp", "args": "Tensor input, Tensor other", "self": "input"},
                {"name": "add", "op": "AddByConstOp", "args": "Tensor tensor, float value", "self": "tensor"},
                {"name": "relu", "op": "ReluOp", "args": "Tensor input", "self": "input"},
                {"name": "no_such_op_anywhere", "op": "NopeOp", "args": "Tensor x"}]
    src = tmp_path / "ops.yml"
    src.write_text(yaml.safe_dump(manifest))
    rep = gen_ops(str(src), str(tmp_path / "out"))
    assert rep["entries"] == 4 and rep["missing"] == ["no_such_op_anywhere"] and rep["bad_kwargs"] == []
    stub = (tmp_path / "out" / "ops.pyi").read_text()
    assert stub.count("@overload") == 2 and "def relu(input: Tensor, **op_meta: Any) -> Tensor: ...   # ReluOp" in stub
    assert "    def add(self, other: Tensor, **op_meta: Any) -> Tensor: ..." in stub
    compile(stub, "ops.pyi", "exec")
    reg = yaml.safe_load((tmp_path / "out" / "registry.yml").read_text())
    names = {r["op"] for r in reg}
    assert {"matmul", "adam_update", "rule_update", "pipeline_send", "pipeline_recv", "comm"} <= names and len(dump_registry()) == len(reg)
