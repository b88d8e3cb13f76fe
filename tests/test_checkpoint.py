"""Checkpoint format: virtual-block split keys, strategy-independent reassembly, ModelSaver rotation, cross-strategy
reload in a real multi-process job (save under tp=2, load under dp=2)."""
import json
import os

import numpy as np
import pytest
import torch

import hetu_b200 as ht
from hetu_b200.utils.checkpoint import ht_safetensors as hs
from hetu_b200.utils.checkpoint import ModelSaver, convert_llama_hf_to_ht
from dist_utils import run_workers

DS = ht.DistributedStates


def test_split_keys_follow_reference_numbering():
    # [64, 32] tensor, unsplit: 8 x 8 virtual blocks, key index = i1 * 8 + i0
    keys = hs.split_keys_for_shard("w", [64, 32], {}, [0, 0], [64, 32])
    assert len(keys) == 64
    assert keys[0][0] == "w_split_0" and keys[1][0] == "w_split_8" and keys[8][0] == "w_split_1"
    # a tp=2 shard of dim 0 owns block rows 4..7
    keys = hs.split_keys_for_shard("w", [64, 32], {0: 2}, [32, 0], [32, 32])
    idx = sorted(int(k.rsplit("_", 1)[1]) for k, _ in keys)
    assert idx == sorted(i1 * 8 + i0 for i1 in range(8) for i0 in range(4, 8))
    # 1-D tensors use a single digit
    assert [k for k, _ in hs.split_keys_for_shard("b", [16], {}, [0], [16])] == [f"b_split_{i}" for i in range(8)]


def test_save_under_tp4_reassemble_under_dp2_tp2(tmp_path):
    full = torch.arange(64 * 32, dtype=torch.float32).reshape(64, 32)
    src = DS(4, {0: 4}, [0])
    for r in range(4):
        b, s = src.local_slice([64, 32], r)
        local = full[b[0]:b[0] + s[0], b[1]:b[1] + s[1]]
        blocks, meta = hs._tensor_blocks("w", local, [64, 32], src, r)
        meta["device_group"] = [0, 1, 2, 3]
        hs.save_file(blocks, str(tmp_path / f"{hs.WEIGHTS_NAME}-{r + 1}-of-4{hs.WEIGHTS_FORMAT}"))
        json.dump({"w": meta}, open(tmp_path / f"param_states-{r + 1}-of-4.json", "w"))
    index = hs._SplitIndex(str(tmp_path))
    dst = DS(4, {-1: 2, 1: 2}, [-1, 1])       # now split along dim 1 instead
    for r in range(4):
        b, s = dst.local_slice([64, 32], r)
        got = hs.assemble_from_splits(index, "w", [64, 32], b, s)
        assert torch.equal(got, full[b[0]:b[0] + s[0], b[1]:b[1] + s[1]])


def test_model_saver_rotation_and_resume(tmp_path):
    with ht.graph("define_and_run", create_new=True) as g:
        m = ht.nn.Linear(8, 4, name="lin")
        x = ht.placeholder("float32", [2, 8], name="x")
        loss = ht.sum(m(x))
        opt = ht.AdamOptimizer(lr=0.1)
        train = opt.minimize(loss)
        saver = ModelSaver(str(tmp_path), save_copies=2, save_interval=1)
        X = np.ones((2, 8), np.float32)
        snaps = {}
        for step in range(1, 5):
            g.run(loss, [loss, train], {x: X})
            saver.save(m, opt, step, consumed_samples=step * 2, loss=0.0)
            snaps[step] = g.get_param(m.weight).clone()
        assert sorted(d for d in os.listdir(tmp_path) if d.startswith("step") and d[4:].isdigit()) == ["step3", "step4"]
        g.run(loss, [loss, train], {x: X})                       # drift away, then restore
        step, consumed = saver.load_latest(m, opt)
        assert (step, consumed) == (4, 8)
        assert torch.equal(g.get_param(m.weight), snaps[4])
        st = opt.get_states(m.weight)
        assert int(g.get_param(st["step"])) == 4


@pytest.mark.parametrize("mode", ["thread", "process"])
def test_async_model_saver_snapshots_before_training_continues(tmp_path, mode):
    """async_save: the snapshot is taken on the training thread (thread mode: private clone; process mode: shared-memory
    pool + forked writer), so optimizer steps issued right after save() cannot leak into the checkpoint"""
    with ht.graph("define_and_run", create_new=True) as g:
        m = ht.nn.Linear(64, 32, name=f"lin_{mode}")
        x = ht.placeholder("float32", [2, 64], name="x")
        loss = ht.sum(m(x))
        opt = ht.AdamOptimizer(lr=0.1)
        train = opt.minimize(loss)
        saver = ModelSaver(str(tmp_path), save_copies=3, async_save=True, async_mode=mode)
        X = np.ones((2, 64), np.float32)
        g.run(loss, [loss, train], {x: X})
        want_w = g.get_param(m.weight).clone()
        want_m = g.get_param(opt.get_states(m.weight)["mean"]).clone()
        saver.save(m, opt, 1, consumed_samples=2, loss=0.0)
        for _ in range(3):                                        # training goes on while the files are written
            g.run(loss, [loss, train], {x: X})
        assert not torch.equal(g.get_param(m.weight), want_w)
        saver.wait()
        assert saver._shm_blocks == [] and saver._child is None and (saver._queue is None or saver._queue.pending == 0)
        files = os.listdir(tmp_path / "step1")
        assert any(f.endswith(".safetensors") for f in files) and any(f.startswith("param_states") for f in files)
        assert saver.load_latest(m, opt) == (1, 2)
        assert torch.equal(g.get_param(m.weight), want_w)
        assert torch.equal(g.get_param(opt.get_states(m.weight)["mean"]), want_m)
        assert int(g.get_param(opt.get_states(m.weight)["step"])) == 1


def test_hf_llama_converter_shapes():
    H, KV, D, L, F, V = 4, 2, 8, 2, 48, 64
    hf = {"model.embed_tokens.weight": torch.randn(V, H * D), "model.norm.weight": torch.ones(H * D), "lm_head.weight": torch.randn(V, H * D)}
    for i in range(L):
        p = f"model.layers.{i}."
        hf.update({p + "input_layernorm.weight": torch.ones(H * D), p + "post_attention_layernorm.weight": torch.ones(H * D),
                   p + "self_attn.q_proj.weight": torch.randn(H * D, H * D), p + "self_attn.k_proj.weight": torch.randn(KV * D, H * D),
                   p + "self_attn.v_proj.weight": torch.randn(KV * D, H * D), p + "self_attn.o_proj.weight": torch.randn(H * D, H * D),
                   p + "mlp.gate_proj.weight": torch.randn(F, H * D), p + "mlp.up_proj.weight": torch.randn(F, H * D),
                   p + "mlp.down_proj.weight": torch.randn(H * D, F)})
    out = convert_llama_hf_to_ht(hf, L, H, KV)
    assert out["transformer.h.1.attn.qkv_dense.weight"].shape == ((H + 2 * KV) * D, H * D)
    assert out["transformer.h.0.mlp.dense_h_to_4h.weight"].shape == (2 * F, H * D)


WORKER = os.path.join(os.path.dirname(__file__), "workers", "ckpt_worker.py")


@pytest.mark.dist
def test_cross_strategy_reload(tmp_path):
    ok, outs = run_workers(WORKER, 2, ["save", 1, 2, str(tmp_path)])
    assert ok, "\n---\n".join(outs)
    saved = [json.loads(l[5:]) for o in outs for l in o.splitlines() if l.startswith("CKPT ")][0]
    ok, outs = run_workers(WORKER, 2, ["load", 2, 1, str(tmp_path)])
    assert ok, "\n---\n".join(outs)
    loaded = [json.loads(l[5:]) for o in outs for l in o.splitlines() if l.startswith("CKPT ")][0]
    assert abs(saved["next_loss"] - loaded["next_loss"]) < 2e-3 * max(1.0, abs(saved["next_loss"])), (saved, loaded)


def test_async_checkpoint_is_published_after_the_write_and_torn_copies_are_skipped(tmp_path):
    """the step-info row / COMPLETE marker / rotation happen only once the writer finished (next wait() or save()); a
    directory without the marker (crash mid-write) is never chosen by load_latest, which falls back to the previous copy"""
    with ht.graph("define_and_run", create_new=True) as g:
        m = ht.nn.Linear(16, 8, name="lin_pub")
        x = ht.placeholder("float32", [2, 16], name="x")
        loss = ht.sum(m(x))
        opt = ht.AdamOptimizer(lr=0.1, lr_warmup_steps=4, lr_decay_steps=20, lr_decay_style="linear", min_lr=0.01)
        train = opt.minimize(loss)
        saver = ModelSaver(str(tmp_path), save_copies=1, async_save=True)
        X = np.ones((2, 16), np.float32)
        g.run(loss, [loss, train], {x: X}); opt.step_lr()
        saver.save(m, opt, 1, consumed_samples=2, loss=0.0)
        assert saver.latest_step() is None or saver.latest_step() == 1     # row appears only with the finished files
        saver.wait()
        assert saver.latest_step() == 1 and os.path.exists(tmp_path / "step1" / "COMPLETE")
        w1 = g.get_param(m.weight).clone()
        g.run(loss, [loss, train], {x: X}); opt.step_lr()
        saver.save(m, opt, 2, consumed_samples=4, loss=0.0)
        # simulate a crash before publication: files of step2 may exist, but no marker and no row; step1 must survive
        saver._queue.wait(); saver._pending = None
        assert os.path.isdir(tmp_path / "step1")
        import csv
        with open(tmp_path / "step_info.csv", "a", newline="") as f:       # even a stray row must not resurrect a torn copy
            csv.writer(f).writerow([2, 4, 0.0, str(tmp_path / "step2"), 0.0])
        fresh = ModelSaver(str(tmp_path), save_copies=1)
        assert fresh.latest_step() == 1
        assert fresh.load_latest(m, opt) == (1, 2)
        assert torch.equal(g.get_param(m.weight), w1)
        # the schedule resumes at the scheduled rate of step 2, not at the fresh optimizer's first warm-up value
        assert opt.step_count == 1 and abs(opt.learning_rate - opt.scheduler.get_lr(2)) < 1e-12


def _fake_hdfs(bin_dir, store):
    """an `hdfs` executable that implements `hdfs dfs -mkdir -p / -put -f / -get [-f] / -ls / -test -e / -rm -r -f` on a local
    directory (`store` plays the namenode) -- enough to drive HdfsCliFS without a Hadoop installation"""
    script = f'''#!{__import__("sys").executable}
import os, shutil, sys
ROOT = {str(store)!r}
a = sys.argv[1:]
assert a[0] == "dfs", a
cmd, rest = a[1], [x for x in a[2:] if not x.startswith("-")]
m = lambda p: os.path.join(ROOT, p.replace("hdfs://nn", "").lstrip("/"))
if cmd == "-mkdir":
    os.makedirs(m(rest[0]), exist_ok=True)
elif cmd == "-put":
    src, dst = rest[0], m(rest[1])
    if os.path.isdir(src):
        shutil.rmtree(dst, ignore_errors=True); shutil.copytree(src, dst)
    else:
        os.makedirs(os.path.dirname(dst), exist_ok=True); shutil.copyfile(src, dst)
elif cmd == "-get":
    src, dst = m(rest[0]), rest[1]
    if not os.path.exists(src): sys.exit(1)
    shutil.copytree(src, dst) if os.path.isdir(src) else shutil.copyfile(src, dst)
elif cmd == "-ls":
    p = m(rest[0])
    if not os.path.isdir(p): sys.exit(1)
    names = sorted(os.listdir(p)); print(f"Found {{len(names)}} items")
    for n in names: print("drwxr-xr-x   - u g 0 2026-01-01 00:00 " + os.path.join(rest[0], n))
elif cmd == "-test":
    sys.exit(0 if os.path.exists(m(rest[0])) else 1)
elif cmd == "-rm":
    shutil.rmtree(m(rest[0]), ignore_errors=True)
else:
    sys.exit(2)
'''
    path = os.path.join(bin_dir, "hdfs")
    with open(path, "w") as f:
        f.write(script)
    os.chmod(path, 0o755)
    return path


@pytest.mark.parametrize("backend", ["arrow-file", "hdfs-cli"])
def test_model_saver_mirrors_to_a_remote_filesystem_and_restores_on_a_fresh_node(tmp_path, backend, monkeypatch):
    """ref: model_saver.py SAVER_DST.HDFS -- published steps are uploaded to the remote (pyarrow filesystem URI, or the `hdfs dfs`
    command line), rotation applies remotely, and a node with an empty local directory resumes from the mirror"""
    from hetu_b200.utils.checkpoint.remote_fs import HdfsCliFS, open_remote
    store = tmp_path / "remote_store"
    store.mkdir()
    if backend == "arrow-file":
        uri = "file://" + str(store / "ckpt")
    else:
        bin_dir = tmp_path / "bin"
        bin_dir.mkdir()
        monkeypatch.setenv("HETU_HDFS_BIN", _fake_hdfs(str(bin_dir), store))
        uri = "hdfs-cli://ckpt"
        assert isinstance(open_remote(uri), HdfsCliFS)
    local_a, local_b = tmp_path / "node_a", tmp_path / "node_b"
    with ht.graph("define_and_run", create_new=True) as g:
        m = ht.nn.Linear(8, 4, name=f"lin_remote_{backend.replace('-', '_')}")
        x = ht.placeholder("float32", [2, 8], name="x")
        loss = ht.sum(m(x))
        opt = ht.AdamOptimizer(lr=0.1)
        train = opt.minimize(loss)
        saver = ModelSaver(str(local_a), save_copies=2, save_interval=1, remote=uri)
        X = np.ones((2, 8), np.float32)
        snaps = {}
        for step in range(1, 5):
            g.run(loss, [loss, train], {x: X})
            saver.save(m, opt, step, consumed_samples=step * 2, loss=0.0)
            snaps[step] = g.get_param(m.weight).clone()
        fs = saver.remote()
        assert [d for d in fs.listdir("") if d.startswith("step") and d[4:].isdigit()] == ["step3", "step4"]       # rotated remotely too
        assert fs.exists("step4/COMPLETE") and fs.exists("step_info.csv") and len(fs.listdir("step4")) >= 3
        g.run(loss, [loss, train], {x: X})                                   # drift
        fresh = ModelSaver(str(local_b), save_copies=2, remote=uri)          # another node: nothing on its disk
        assert not os.path.exists(local_b / "step_info.csv")
        step, consumed = fresh.load_latest(m, opt)
        assert (step, consumed) == (4, 8) and torch.equal(g.get_param(m.weight), snaps[4])
        assert os.path.exists(local_b / "step4" / "COMPLETE")


def test_huggingface_gpt2_weights_round_trip_and_logits_match():
    """HF GPT2LMHeadModel (tiny, random) -> convert_gpt2_hf_to_ht -> GPTLMHeadModel: the logits of this framework equal the
    HuggingFace forward; converting back reproduces the HF state dict bit for bit (ref: examples/hetero/gpt_hf_to_ht.py, gpt_hf_to_hf.py)"""
    transformers = pytest.importorskip("transformers")
    from hetu_b200.models import GPTConfig, GPTLMHeadModel, generate_ds_parallel_config
    from hetu_b200.utils.checkpoint import convert_gpt2_hf_to_ht, convert_gpt2_ht_to_hf
    torch.manual_seed(0)
    L, H, NH, V, S = 2, 32, 4, 97, 12
    hf = transformers.GPT2LMHeadModel(transformers.GPT2Config(vocab_size=V, n_positions=S, n_embd=H, n_layer=L, n_head=NH, activation_function="gelu",
                                                              resid_pdrop=0.0, embd_pdrop=0.0, attn_pdrop=0.0)).eval()
    hf_sd = {k: v.detach().clone() for k, v in hf.state_dict().items() if not k.endswith((".attn.bias", ".attn.masked_bias"))}
    ht_sd = convert_gpt2_hf_to_ht(hf_sd, L, NH)
    ids = torch.randint(0, V, (2, S))
    with torch.no_grad():
        want = hf(ids).logits.reshape(2 * S, V)
    with ht.graph("define_and_run", create_new=True) as g:
        m = GPTLMHeadModel(GPTConfig(vocab_size=V, n_positions=S, n_embd=H, n_layer=L, n_head=NH), [generate_ds_parallel_config(L, 1, 1, 1, 1, zero=False)])
        x = ht.placeholder("int64", [2 * S], name="ids")
        p = ht.placeholder("int64", [2 * S], name="pos")
        logits = m(x, p, None, seq_len=S)
        missing = m.load_state_dict(ht_sd, strict=False)
        got = g.run(logits, [logits], {x: ids.reshape(-1), p: torch.arange(S).repeat(2)})[0]
    assert torch.allclose(got.float(), want, atol=2e-4, rtol=1e-4), float((got.float() - want).abs().max())
    back = convert_gpt2_ht_to_hf(ht_sd, L, NH)
    for k, v in hf_sd.items():
        assert torch.equal(back[k], v), k


def test_huggingface_llama_weights_round_trip_and_logits_match():
    transformers = pytest.importorskip("transformers")
    from hetu_b200.models import LlamaConfig, LlamaLMHeadModel, generate_ds_parallel_config
    from hetu_b200.utils.checkpoint import convert_llama_hf_to_ht, convert_llama_ht_to_hf
    torch.manual_seed(1)
    L, H, NH, KV, F, V, S = 2, 32, 4, 2, 64, 89, 10
    hf_cfg = transformers.LlamaConfig(vocab_size=V, hidden_size=H, intermediate_size=F, num_hidden_layers=L, num_attention_heads=NH, num_key_value_heads=KV,
                                      max_position_embeddings=S, rms_norm_eps=1e-6, tie_word_embeddings=False, attention_dropout=0.0)
    hf = transformers.LlamaForCausalLM(hf_cfg).eval()
    hf_sd = {k: v.detach().clone() for k, v in hf.state_dict().items() if "rotary_emb" not in k}
    ht_sd = convert_llama_hf_to_ht(hf_sd, L, NH, KV)
    ids = torch.randint(0, V, (2, S))
    with torch.no_grad():
        want = hf(ids).logits.reshape(2 * S, V)
    cfg = LlamaConfig(vocab_size=V, hidden_size=H, intermediate_size=F, num_hidden_layers=L, num_attention_heads=NH, num_key_value_heads=KV,
                      max_position_embeddings=S, rms_norm_eps=1e-6)
    with ht.graph("define_and_run", create_new=True) as g:
        m = LlamaLMHeadModel(cfg, [generate_ds_parallel_config(L, 1, 1, 1, 1, zero=False, model="llama")])
        x = ht.placeholder("int64", [2 * S], name="ids")
        p = ht.placeholder("int64", [2 * S], name="pos")
        logits = m(x, p, None, seq_len=S)
        m.load_state_dict(ht_sd, strict=False)
        got = g.run(logits, [logits], {x: ids.reshape(-1), p: torch.arange(S).repeat(2)})[0]
    assert torch.allclose(got.float(), want, atol=2e-4, rtol=1e-4), float((got.float() - want).abs().max())
    back = convert_llama_ht_to_hf(ht_sd, L, NH, KV)
    for k, v in hf_sd.items():
        assert torch.equal(back[k], v), k


def test_examine_checkpoint_lists_tensors_layouts_and_states(tmp_path, capsys):
    from hetu_b200.utils.checkpoint import examine_checkpoint
    from hetu_b200.utils.checkpoint.converters import main as examine_main
    with ht.graph("define_and_run", create_new=True) as g:
        m = ht.nn.Linear(8, 4, name="lin_examine")
        x = ht.placeholder("float32", [2, 8], name="x")
        loss = ht.sum(m(x))
        opt = ht.AdamOptimizer(lr=0.1)
        train = opt.minimize(loss)
        g.run(loss, [loss, train], {x: np.ones((2, 8), np.float32)})
        ModelSaver(str(tmp_path), save_copies=1).save(m, opt, 1, consumed_samples=2, loss=0.0)
    info = examine_checkpoint(str(tmp_path / "step1"))
    assert info["complete"] and len(info["shard_files"]) == 1 and info["blocks"] > 0 and info["bytes"] > 0
    w = next(t for n, t in info["parameters"].items() if n.endswith("weight"))
    assert w["global_shape"] == [4, 8] and w["dtype"] == "float32" and w["device_num"] == 1
    assert any(n.endswith("_mean") for n in info["optimizer_states"]) and any(n.endswith("_variance") for n in info["optimizer_states"])
    examine_main([str(tmp_path / "step1")])
    out = capsys.readouterr().out
    assert "COMPLETE" in out and "weight" in out and "[4, 8]" in out
