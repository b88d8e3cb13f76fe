"""Rendezvous / KV / heartbeat / elastic restart (ref: python/hetu/rpc/** behaviour)."""
import os
import subprocess
import sys
import threading
import time

import pytest

from hetu_b200.rpc import (DeviceClient, DeviceControllerServer, ElasticServer, ElasticStrategy, KeyValueStoreClient,
                           KeyValueStoreServer, ProducerConsumer, pssh_start, read_hosts_yaml)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("client_cls", ["python", "native", "mixed"])
def test_rendezvous_ranks_kv_barrier_consistent(client_cls):
    from hetu_b200.rpc import NativeDeviceClient
    port = _free_port()
    pick = {"python": lambda i: DeviceClient, "native": lambda i: NativeDeviceClient,
            "mixed": lambda i: NativeDeviceClient if i % 2 == 0 else DeviceClient}[client_cls]
    srv = DeviceControllerServer(3, "127.0.0.1", port).start()
    out = {}

    def worker(i):
        c = pick(i)(f"127.0.0.1:{port}", hostname="nodeA" if i < 2 else "nodeB", heartbeat_interval=0.1)
        r, l, w = c.connect()
        hosts = c.all_gather_hostnames()
        infos = c.exchange_device_info({"rank": r, "mem": 180})
        if r == 0:
            c.commit_nccl_id([0, 1, 2], 0, b"\x01\x02unique")
            c.put_json("plan", {"dp": 3})
            c.put_double("lr", 0.5)
            c.put_bytes("blob", b"abc")
        nid = c.get_nccl_id([2, 1, 0], 0)
        c.barrier()
        ok = c.consistent("graph-hash-1")
        bad = c.consistent(r, tag="differs")
        out[r] = (l, w, hosts, len(infos), nid, c.get_json("plan"), c.get_double("lr"), c.get_bytes("blob"), ok, bad)
        c.exit()
    ts = [threading.Thread(target=worker, args=(i,)) for i in range(3)]
    [t.start() for t in ts]
    [t.join(30) for t in ts]
    assert sorted(out) == [0, 1, 2]
    for r, (l, w, hosts, ninfo, nid, plan, lr, blob, ok, bad) in out.items():
        assert w == 3 and ninfo == 3 and nid == b"\x01\x02unique" and plan == {"dp": 3} and lr == 0.5 and blob == b"abc"
        assert ok is True and bad is False
        assert sorted(hosts) == ["nodeA", "nodeA", "nodeB"]
    assert sorted(out[r][0] for r in out if out[r][2][r] == "nodeA") == [0, 1]     # local device indices per host
    assert srv.all_exited()
    srv.shutdown()


def test_heartbeat_detects_dead_rank_and_stop_flag():
    port = _free_port()
    srv = DeviceControllerServer(2, "127.0.0.1", port, heartbeat_timeout=0.5).start()
    cs = [DeviceClient(f"127.0.0.1:{port}", heartbeat_interval=0.1) for _ in range(2)]
    ts = [threading.Thread(target=c.connect, kwargs={"start_heartbeat": i == 0}) for i, c in enumerate(cs)]
    [t.start() for t in ts]
    [t.join(10) for t in ts]
    time.sleep(1.0)
    silent = cs[1].rank
    assert srv.dead_ranks() == [silent]
    cs[0].worker_stop()
    time.sleep(0.3)
    assert cs[0].already_stop() and cs[1].already_stop()
    srv.shutdown()


def test_kv_store_producer_consumer():
    port = _free_port()
    srv = KeyValueStoreServer("127.0.0.1", port).start()
    prod = ProducerConsumer(KeyValueStoreClient(f"127.0.0.1:{port}"), "plans", max_ahead=2)
    cons = ProducerConsumer(KeyValueStoreClient(f"127.0.0.1:{port}"), "plans")
    got = []
    t = threading.Thread(target=lambda: [got.append(cons.consume(i)) for i in range(5)])
    t.start()
    for i in range(5):
        prod.produce({"step": i, "strategy": i % 2})
    t.join(20)
    assert [g["step"] for g in got] == list(range(5))
    d = KeyValueStoreClient(f"127.0.0.1:{port}").register_dict("cfg")
    d["a"] = [1, 2]
    assert d["a"] == [1, 2]
    srv.shutdown()


def test_elastic_server_replans_after_worker_failure(tmp_path):
    marker = tmp_path / "gen.log"
    script = tmp_path / "w.py"
    script.write_text(
        "import sys, time\n"
        "gen, rank, world = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])\n"
        f"open(r'{marker}', 'a').write(f'{{gen}} {{rank}} {{world}}\\n')\n"
        "if gen == 0 and rank == 3: sys.exit(7)\n"      # one worker dies in the first generation
        "time.sleep(0.3)\n")

    def launch(gen, plan, addr):
        return [subprocess.Popen([sys.executable, str(script), str(gen), str(r), str(plan["num_devices"])]) for r in range(plan["num_devices"])]
    es = ElasticServer(launch, 4, ElasticStrategy(tp=2, pp=1), port=_free_port(), max_restarts=2)
    assert es.run() == 0
    assert [g["plan"]["num_devices"] for g in es.generations] == [4, 2]     # 3 survivors -> dp1 x tp2
    lines = marker.read_text().split("\n")
    assert any(l.startswith("1 0 2") for l in lines)


def test_pssh_dry_run_and_hosts_yaml(tmp_path):
    y = tmp_path / "hosts.yaml"
    y.write_text("hosts:\n  - addr: 127.0.0.1\n    initial_workers: 2\n  - addr: node1\n    initial_workers: 1\n")
    hosts = read_hosts_yaml(str(y))
    assert hosts == [{"addr": "127.0.0.1", "workers": 2}, {"addr": "node1", "workers": 1}]
    lines = pssh_start("python train.py", hosts, 29511, dry_run=True)
    assert len(lines) == 3 and lines[2][0] == "ssh" and "RANK=2" in lines[2][-1] and "WORLD_SIZE=3" in lines[0][-1]


def test_native_client_json_scanner_typed_kv_and_heartbeat():
    """C++ client: raw-JSON pass-through (nesting, escapes, unicode), typed KV round trips, heart-beats on a second
    connection, stop flag propagation, error replies become exceptions"""
    import hetu_b200 as ht
    from hetu_b200.rpc import NativeDeviceClient
    C = ht._C
    obj = '{"a": [1, {"b": "x,}\\"]"}], "ok": true, "value": {"k": "v\\u00e9\\n", "n": -1.5e3}, "z": null}'
    assert C.json_field(obj, "ok") == "true" and C.json_field(obj, "z") == "null" and C.json_field(obj, "missing") is None
    assert C.json_field(obj, "a") == '[1, {"b": "x,}\\"]"}]'
    assert C.json_unquote(C.json_field(C.json_field(obj, "value"), "k")) == "v\u00e9\n"
    assert C.json_unquote(C.json_quote('q"\\\n\t\x01')) == 'q"\\\n\t\x01'
    port = _free_port()
    srv = DeviceControllerServer(1, "127.0.0.1", port, heartbeat_timeout=0.5).start()
    c = NativeDeviceClient(f"127.0.0.1:{port}", hostname="n0", heartbeat_interval=0.05)
    assert c.connect() == (0, 0, 1)
    c.put_int("i", -(1 << 40)); c.put_double("d", 0.1); c.put_double("whole", 3.0); c.put_string("s", 'he said "hi"\n\u00e9')
    blob = bytes(range(256)) * 3 + b"x"
    c.put_bytes("b", blob); c.put_json("j", {"nested": [1, 2, {"x": None}], "t": True})
    assert c.get_int("i") == -(1 << 40) and c.get_double("d") == 0.1 and c.get_double("whole") == 3.0
    assert c.get_string("s") == 'he said "hi"\n\u00e9' and c.get_bytes("b") == blob
    assert c.get_json("j") == {"nested": [1, 2, {"x": None}], "t": True}
    assert c.remove("j") is True and c.remove("j") is False
    assert c.consistent({"graph": [1, 2]}) is True
    with pytest.raises(RuntimeError, match="unknown method"):
        c.call("NoSuchMethod")
    time.sleep(0.4)
    assert c.heartbeats_sent >= 3 and srv.dead_ranks() == []
    assert c.already_stop() is False
    c.worker_stop()
    time.sleep(0.2)
    assert c.already_stop() is True
    c.exit()
    assert srv.all_exited()
    srv.shutdown()


def test_rpc_bootstrap_assigns_ranks_and_brings_up_the_process_group():
    """no RANK / WORLD_SIZE in the environment: the controller assigns ranks, rank 0 publishes the store address, every
    worker joins torch.distributed (gloo) and an all-reduce over the framework's comm runtime works"""
    port = _free_port()
    srv = DeviceControllerServer(2, "127.0.0.1", port).start()
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT")}
    env.update({"HETU_RENDEZVOUS": "rpc", "HETU_B200_FORCE_CPU": "1", "CUDA_VISIBLE_DEVICES": "", "OMP_NUM_THREADS": "1",
                "PYTHONPATH": ROOT + os.pathsep + env.get("PYTHONPATH", "")})
    worker = os.path.join(ROOT, "tests", "workers", "rpc_bootstrap_worker.py")
    procs = [subprocess.Popen([sys.executable, worker, f"127.0.0.1:{port}", "2"], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for _ in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n----\n".join(outs)
    lines = sorted(l for o in outs for l in o.splitlines() if l.startswith("BOOT"))
    assert lines == ["BOOT rank=0 local=0 world=2 sum=3.0", "BOOT rank=1 local=1 world=2 sum=3.0"], outs
    assert srv.all_exited()
    srv.shutdown()


def test_pssh_start_config_launches_workers_that_rendezvous_through_the_controller(tmp_path):
    """`python -m hetu.rpc.pssh_start_config rpc.command=...`: starts the controller, spawns the workers without ranks in their
    environment; they obtain ranks from the controller (native client) and form the process group.  Also the hydra-style
    YAML + override loader and the reference-named strategy generators."""
    from hetu_b200.models.gpt.generate_gpt_4d_config import main as gen_main
    from hetu_b200.rpc.pssh_start_config import main as launch_main
    port = _free_port()
    worker = os.path.join(ROOT, "tests", "workers", "rpc_bootstrap_worker.py")
    (tmp_path / "config.yaml").write_text(
        f"rpc:\n  num_gpus: 2\n  timeout: 240\n  server_port: {port}\n  log_path: {tmp_path}/logs\n  envs: {{HETU_B200_FORCE_CPU: 1, CUDA_VISIBLE_DEVICES: '', OMP_NUM_THREADS: 1, PYTHONPATH: {ROOT}}}\n"
        f"ds_parallel:\n  num_layers: 8\n  num_gpus: 8\n  dp: 2\n  tp: 2\n  pp: 2\n  zero: true\n  ds_parallel_config_path: {tmp_path}/ds\n  ds_parallel_config_name: s.json\n")
    rc = launch_main(["--config-path", str(tmp_path), "--config-name", "config", f"rpc.command={sys.executable} {worker} 127.0.0.1:{port} 2"])
    logs = sorted((tmp_path / "logs").glob("rank*.log"))
    text = "\n".join(p.read_text() for p in logs)
    assert rc == 0, text
    assert sorted(l for l in text.splitlines() if l.startswith("BOOT")) == ["BOOT rank=0 local=0 world=2 sum=3.0", "BOOT rank=1 local=1 world=2 sum=3.0"]
    out = gen_main(["--config-path", str(tmp_path), "ds_parallel.tp=4", "ds_parallel.dp=1"])
    import json
    cfg = json.load(open(out))
    assert cfg["blocks"]["blocks0-3"]["attn"]["qkv"]["split"] == {"0": [4]} and cfg["zero"] in (True, False)
    r = subprocess.run([sys.executable, "-m", "hetu.models.llama.generate_llama_hetero_4d_config", "--config-path", str(tmp_path),
                        "ds_parallel.hetero_layers=[[4,4],[8]]", "ds_parallel.hetero_tp=[2,1]", "ds_parallel.ds_parallel_config_name=h.json"],
                       env={**os.environ, "PYTHONPATH": ROOT, "HETU_B200_FORCE_CPU": "1"}, capture_output=True, text=True, cwd=str(tmp_path), timeout=240)
    assert r.returncode == 0, r.stderr
    h = json.load(open(tmp_path / "ds" / "h.json"))
    assert h["blocks"]["blocks0"]["attn"]["qkv"]["device_group_union"] == [[0, 1], [4]] and h["blocks"]["blocks5"]["attn"]["qkv"]["device_group_union"] == [[2, 3], [4]]


def test_elastic_node_detection_replan_and_rank_remapping():
    """ref: heturpc_elastic_server.py:497-859 -- survey the nodes (nvidia-smi), re-plan around the GPUs that are really there
    and make the rendezvous server hand out ranks according to the plan instead of the connection order"""
    from hetu_b200.rpc.elastic_server import ElasticStrategy, available_gpus, detect_node_info
    row = "{i}, NVIDIA B200, 183359, {free}, 100, 0"
    smi = {"n0": "\n".join(row.format(i=i, free=180000) for i in range(8)),
           "n1": "\n".join(row.format(i=i, free=180000 if i != 6 else 1000) for i in range(8) if i != 3)}

    def runner(node, cmd):
        if node == "n2":
            return None                                   # unreachable node
        return smi[node] if "nvidia-smi" in cmd else node + "-host\n"
    info = detect_node_info(["n0", "n1", "n2"], runner)
    assert info[2]["reachable"] is False and info[2]["gpu_info"] == [] and len(info[1]["gpu_info"]) == 7
    gpus, bound = available_gpus(info, min_free_fraction=0.5)
    assert bound == 183359 and [g["idx"] for g in gpus[1]["gpus"]] == [8, 9, 10, 12, 13, 15]      # 11 missing, 14 busy
    es = ElasticStrategy(tp=4, pp=2, dp=2, num_layers=16, global_micro_batches=16, memory_bound_layers=12, min_free_fraction=0.5)
    plan = es.plan_from_nodes({k: v for k, v in info.items() if k < 2})
    devs = sorted(plan["rank_to_device_mapping"].values())
    assert 11 not in devs and 14 not in devs and len(devs) == len(set(devs)) == plan["num_devices"]
    assert all(sum(ls) == 16 for ls in plan["hetero_layers"]) and sum(plan["micro_batch_num_list"]) == 16
    assert sorted(r for rs in plan["host_to_ranks"].values() for r in rs) == sorted(plan["rank_to_device_mapping"])
    cmd = es.replace_cmd("python train.py --dp 2 --tp 4 --pp 2 --num_gpus=16 --hetero_stages [2,2] --steps 100", plan)
    assert f"--dp {plan['dp']} " in cmd and f"--num_gpus={plan['num_devices']}" in cmd and "--hetero_stages " + str(plan["hetero_stages"]).replace(" ", "") in cmd
    assert "--steps 40" in es.renew_step(cmd, 40)
    # the server enforces the mapping: the k-th worker of a host gets the k-th planned rank / local GPU of that host
    small = {"dp": 1, "tp": 1, "pp": 1, "num_devices": 3}
    srv = DeviceControllerServer(3, port=_free_port(), host_to_ranks={"hA": [2, 0], "hB": [1]}, host_to_local={"hA": [5, 1], "hB": [7]}).start()
    try:
        assert srv.rpc_Connect("c0", "hA") == 2 and srv.rpc_Connect("c1", "hB") == 1 and srv.rpc_Connect("c2", "hA") == 0
        assert srv.rpc_GetRank("c0") == {"rank": 2, "local_device": 5, "world_size": 3}
        assert srv.rpc_GetRank("c2") == {"rank": 0, "local_device": 1, "world_size": 3}
        assert srv.rpc_GetRank("c1")["local_device"] == 7
        with pytest.raises(RuntimeError):
            srv.rpc_Connect("c3", "hB")                   # no rank left for that host in the plan
    finally:
        srv.shutdown()


def test_ps_scheduler_sharded_servers_key_ranges_barriers_heartbeats():
    """ref: ps-lite scheduler / postoffice -- 2 servers + 3 workers register with the scheduler, dense parameters are cut into
    per-server key ranges, sparse rows go to row % S, worker barriers run through the scheduler, a silent node shows up as
    dead, everybody checks out"""
    import os
    import threading
    import time

    import numpy as np

    from hetu_b200 import _C
    from hetu_b200.v1.ps import ShardedPSContext
    S, W = 2, 3
    sched = _C.PsScheduler(S, W, 0, "127.0.0.1")
    addr = f"127.0.0.1:{sched.port}"
    os.environ["HETU_PS_NUM_WORKERS"] = str(W)
    servers, workers, errors = [None] * S, [None] * W, []

    def start_server(i):
        try:
            servers[i] = ShardedPSContext.serve(addr)
        except Exception as e:      # noqa: BLE001
            errors.append(repr(e))

    def start_worker(i):
        try:
            workers[i] = ShardedPSContext(addr)
        except Exception as e:      # noqa: BLE001
            errors.append(repr(e))
    ts = [threading.Thread(target=start_server, args=(i,)) for i in range(S)] + [threading.Thread(target=start_worker, args=(i,)) for i in range(W)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(60)
    assert not errors and sched.registered == S + W, errors
    assert sorted(w.worker_id for w in workers) == [0, 1, 2] and sorted(c.rank for _, c in servers) == [0, 1]
    w0 = workers[0]
    assert w0.num_servers == S and w0.num_workers == W and [s.rank for s in w0.sched.servers] == [0, 1]
    assert w0.sched.key_ranges(11) == [0, 6, 11] and w0.sched.key_ranges(4) == [0, 2, 4]

    # dense: every worker pushes a gradient; SGD on the servers; each server holds only its slice
    init = np.arange(11, dtype=np.float32)
    w0.init_dense("w", init, opt="sgd", lr=0.5)
    for w in workers[1:]:
        w._dense_len["w"] = 11
    done = []

    def step(w):
        w.push("w", np.ones(11, np.float32) * (w.worker_id + 1))
        w.barrier()                                    # all pushes have landed once every worker passed the barrier
        done.append(w.pull("w"))
    ts = [threading.Thread(target=step, args=(w,)) for w in workers]
    for t in ts:
        t.start()
    for t in ts:
        t.join(60)
    assert len(done) == W
    for got in done:
        np.testing.assert_allclose(got, init - 0.5 * (1 + 2 + 3), rtol=1e-6)
    per_server = [w0.servers[s].pull_dense(w0.key("w")) for s in range(S)]
    assert [len(p) for p in per_server] == [6, 5]

    # sparse: rows split by row % S
    table = np.arange(10 * 4, dtype=np.float32).reshape(10, 4)
    w0.init_sparse("emb", table, opt="sgd", lr=1.0)
    np.testing.assert_array_equal(workers[1].sparse_pull("emb", [9, 0, 3, 4], 4), table[[9, 0, 3, 4]])
    workers[2].sparse_push("emb", [3, 4], np.ones((2, 4), np.float32))
    np.testing.assert_array_equal(workers[0].sparse_pull("emb", [3, 4, 5], 4), np.stack([table[3] - 1, table[4] - 1, table[5]]))

    # liveness: nobody heartbeats for 0.3 s except worker 0 and server 0 -> the others are reported
    time.sleep(0.35)
    w0.sched.heartbeat(); servers[0][1].heartbeat()
    dead = w0.dead_nodes(0.3)
    assert len(dead) == S + W - 2 and all(not (d.role == 1 and d.rank == w0.worker_id) for d in dead)
    for w in workers:
        w.finalize()
    for net, c in servers:
        c.finalize()
    assert sched.wait_finalized(10.0) and sched.finalized == S + W
    for net, _ in servers:
        net.stop()
    sched.stop()
