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


def test_rendezvous_ranks_kv_barrier_consistent():
    port = _free_port()
    srv = DeviceControllerServer(3, "127.0.0.1", port).start()
    out = {}

    def worker(i):
        c = DeviceClient(f"127.0.0.1:{port}", hostname="nodeA" if i < 2 else "nodeB", heartbeat_interval=0.1)
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
