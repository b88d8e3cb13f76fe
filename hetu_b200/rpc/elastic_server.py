"""Elastic controller: runs a training job as restartable generations.  It watches worker heartbeats, on a failure (or
a straggler report) tells every worker to stop, detects the surviving nodes / GPUs, asks an `ElasticStrategy` for a
new (dp, tp, pp) plan and relaunches the workers, which resume from the latest checkpoint.
(ref: python/hetu/rpc/heturpc_elastic_server.py -- heartbeat monitor :463-468, restart loop :785-795,
elastic_arg_parser.py)"""
from __future__ import annotations

import argparse
import math
import subprocess
import sys
import time
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence

from .server import DeviceControllerServer


DEVICES_PER_NODE = 8
NVIDIA_SMI_QUERY = ("nvidia-smi --query-gpu=index,name,memory.total,memory.free,memory.used,utilization.gpu "
                    "--format=csv,noheader,nounits")


def _local_runner(node: str, cmd: str, timeout: float = 10.0) -> Optional[str]:
    """run `cmd` on `node`: directly for this machine, over ssh (paramiko) otherwise; None when the node is unreachable"""
    if node in ("localhost", "127.0.0.1", ""):
        try:
            return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=timeout).stdout
        except (subprocess.TimeoutExpired, OSError):
            return None
    try:
        import paramiko
        cli = paramiko.SSHClient()
        cli.set_missing_host_key_policy(paramiko.AutoAddPolicy())
        cli.connect(node, timeout=timeout)
        _, out, _ = cli.exec_command(cmd, timeout=timeout)
        text = out.read().decode("utf-8")
        cli.close()
        return text
    except Exception:   # noqa: BLE001 -- any failure means "node not usable"
        return None


def detect_node_info(nodes: Sequence[str], runner: Callable[[str, str], Optional[str]] = _local_runner) -> Dict[int, dict]:
    """per node: reachability, hostname and one record per visible GPU (index, name, MiB total / free / used, utilisation).
    (ref: heturpc_elastic_server.py:497 detect_node_info -- ssh + nvidia-smi csv)"""
    info: Dict[int, dict] = {}
    for idx, node in enumerate(nodes):
        rec = {"addr": node, "hostname": node, "reachable": False, "gpu_info": []}
        text = runner(node, NVIDIA_SMI_QUERY)
        if text is not None:
            rec["reachable"] = True
            host = runner(node, "hostname")
            if host:
                rec["hostname"] = host.strip()
            for line in text.strip().splitlines():
                row = [c.strip() for c in line.split(",")]
                if len(row) != 6:
                    continue
                try:
                    di, name, tot, free, used, util = int(row[0]), row[1], int(row[2]), int(row[3]), int(row[4]), float(row[5])
                except ValueError:
                    continue
                rec["gpu_info"].append({"device_idx": di, "name": name, "total_memory": tot, "free_memory": free, "used_memory": used,
                                        "remain_percent": free / tot if tot else 0.0, "utilization": util})
        info[idx] = rec
    return info


def available_gpus(node_info: Dict[int, dict], min_free_fraction: float = 0.0, devices_per_node: int = DEVICES_PER_NODE):
    """-> ({node idx: {"addr", "hostname", "gpus": [{"idx": global id, "local_idx", "memory"}]}}, smallest total memory in MiB).
    A GPU is usable when more than `min_free_fraction` of its memory is free (somebody else's job may sit on it).
    Global device id = node index * devices_per_node + local index, the numbering the planners use.
    (ref: elastic_arg_parser.py:240 parse_gpu_info)"""
    out, bound = {}, math.inf
    for k, v in sorted(node_info.items()):
        gpus = []
        for g in v["gpu_info"]:
            if g["remain_percent"] > min_free_fraction:
                bound = min(bound, g["total_memory"])
                gpus.append({"idx": k * devices_per_node + g["device_idx"], "local_idx": g["device_idx"], "memory": g["free_memory"]})
        out[k] = {"addr": v["addr"], "hostname": v["hostname"], "ngpu": len(v["gpu_info"]), "gpus": gpus}
    return out, bound


@dataclass
class ElasticStrategy:
    """Chooses the strategy of the next generation.

    Without topology knowledge (`plan(num_devices)`): the largest homogeneous (dp, tp, pp) that fits the surviving device
    count, shrinking tp then pp when necessary.  With a node survey (`plan_from_nodes(node_info)`): the Ampelos planner
    (engine/strategy_ampelos.py) re-plans around the GPUs that are really present -- possibly a heterogeneous plan with
    narrower tensor-parallel groups, pipelines of different depth and uneven micro-batch counts -- and the result carries
    the rank -> physical device mapping the rendezvous server enforces (`host_to_ranks` / `host_to_local`).
    (ref: python/hetu/rpc/elastic_arg_parser.py:89-610 ElasticStrategy, heturpc_elastic_server.py:544 detect_gpu_nums)"""
    tp: int = 1
    pp: int = 1
    min_dp: int = 1
    dp: int = 0                     # dp of the ORIGINAL job (needed by plan_from_nodes; 0: derive from the first survey)
    num_layers: int = 0
    global_micro_batches: int = 0   # micro-batches per step of the whole job
    min_free_fraction: float = 0.0
    memory_bound_layers: float = math.inf
    devices_per_node: int = DEVICES_PER_NODE
    straggler_ratios: Optional[Dict[int, float]] = None

    def plan(self, num_devices: int) -> Optional[Dict[str, int]]:
        tp, pp = self.tp, self.pp
        while tp * pp > num_devices and tp > 1:
            tp //= 2
        while tp * pp > num_devices and pp > 1:
            pp //= 2
        dp = num_devices // (tp * pp)
        if dp < self.min_dp:
            return None
        return {"dp": dp, "tp": tp, "pp": pp, "num_devices": dp * tp * pp}

    def plan_from_nodes(self, node_info: Dict[int, dict]) -> Optional[dict]:
        from ..engine.strategy import TrainerCtxs, TrainerStrategyArgs
        from ..engine.strategy_ampelos import AmpelosStrategyModel
        gpus, _ = available_gpus(node_info, self.min_free_fraction, self.devices_per_node)
        alive = sorted(g["idx"] for v in gpus.values() for g in v["gpus"])
        if not alive:
            return None
        n_nodes = max(node_info) + 1
        total = n_nodes * self.devices_per_node
        dp = self.dp or max(total // (self.tp * self.pp), 1)
        if dp * self.tp * self.pp != total:          # the original job did not cover whole nodes: plan on the alive count only
            p = self.plan(len(alive))
            if p is None:
                return None
            p["rank_to_device_mapping"] = {r: alive[r] for r in range(p["num_devices"])}
            return self._attach_hosts(p, gpus)
        layers = self.num_layers or 8 * self.pp
        mbn = max(self.global_micro_batches // dp, 1) if self.global_micro_batches else 8
        ctxs = TrainerCtxs(normal_layers=max(layers // self.pp, 1), normal_mbn=mbn, memory_bound=self.memory_bound_layers)
        old = TrainerStrategyArgs(dp=dp, tp=self.tp, pp=self.pp, rank_to_device_mapping={r: r for r in range(total)})
        sr = {d: float((self.straggler_ratios or {}).get(d, 1.0)) for d in alive}
        model = AmpelosStrategyModel(ctxs, old, sr, {}, [], dead_devices=[d for d in range(total) if d not in set(alive)],
                                     devices_per_node=self.devices_per_node)
        try:
            st, cfg = model.make_plans()
        except AssertionError:
            return None
        active = [r for r in sorted(st.rank_to_device_mapping) if r not in st.unused_rank_list and r not in st.suspended_rank_list]
        plan = {"dp": st.dp, "tp": st.tp, "pp": st.pp, "num_devices": len(active), "hetero": True,
                "hetero_layers": st.hetero_layers, "hetero_stages": st.hetero_stages,
                "micro_batch_num_list": st.hetero_micro_batch_num_list, "unused_rank": sorted(st.unused_rank_list + st.suspended_rank_list),
                "rank_to_device_mapping": {r: st.rank_to_device_mapping[r] for r in active},
                "estimated_time": model.estimate_time(), "ds_parallel_config": cfg,
                "candidates": [c.describe() for c in model.candidates]}
        return self._attach_hosts(plan, gpus)

    def _attach_hosts(self, plan: dict, gpus: dict) -> dict:
        """which worker of which host plays which rank: ranks sorted by physical device so a host's k-th worker gets its
        k-th planned GPU (the server hands ranks out in connection order per host)"""
        by_dev = {g["idx"]: (v["hostname"], g["local_idx"]) for v in gpus.values() for g in v["gpus"]}
        host_to_ranks: Dict[str, list] = {}
        host_to_local: Dict[str, list] = {}
        for r, d in sorted(plan["rank_to_device_mapping"].items(), key=lambda kv: kv[1]):
            host, local = by_dev[d]
            host_to_ranks.setdefault(host, []).append(r)
            host_to_local.setdefault(host, []).append(local)
        plan["host_to_ranks"], plan["host_to_local"] = host_to_ranks, host_to_local
        return plan

    @staticmethod
    def replace_cmd(cmd: str, plan: dict) -> str:
        """rewrite `--dp/--tp/--pp/--num_gpus/--hetero_layers/--hetero_stages/--micro_batch_num_list/--unused_rank/
        --rank_to_device_mapping` of a worker command line for the new plan (ref: elastic_arg_parser.py:569 replace_cmd)"""
        def fmt(v):
            return str(v).replace(" ", "")
        args = {"dp": plan["dp"], "tp": plan["tp"], "pp": plan["pp"], "num_gpus": plan["num_devices"], "ngpus": plan["num_devices"]}
        for k in ("hetero_layers", "hetero_stages", "micro_batch_num_list", "unused_rank", "rank_to_device_mapping"):
            if k in plan:
                args[k] = fmt(plan[k])
        toks = cmd.split()
        out, i, seen = [], 0, set()
        while i < len(toks):
            t = toks[i]
            key = t[2:].split("=")[0].replace("-", "_") if t.startswith("--") else None
            if key in args:
                seen.add(key)
                if "=" in t:
                    out.append(f"{t.split('=')[0]}={args[key]}")
                else:
                    out += [t, str(args[key])]
                    i += 1           # drop the old value
            else:
                out.append(t)
            i += 1
        return " ".join(out)

    @staticmethod
    def renew_step(cmd: str, remaining_steps: int) -> str:
        """a restarted generation only runs the steps that are left (ref: elastic_arg_parser.py:591)"""
        toks = cmd.split()
        for i, t in enumerate(toks):
            if t in ("--steps", "--num_steps", "--train_steps") and i + 1 < len(toks):
                toks[i + 1] = str(int(remaining_steps))
        return " ".join(toks)


def elastic_arg_parser(argv: Optional[Sequence[str]] = None):
    ap = argparse.ArgumentParser("elastic launcher")
    ap.add_argument("--command", required=True, help="worker command; {rank} {world} {dp} {tp} {pp} {addr} are substituted")
    ap.add_argument("--ngpus", type=int, required=True)
    ap.add_argument("--tp", type=int, default=1)
    ap.add_argument("--pp", type=int, default=1)
    ap.add_argument("--port", type=int, default=23457)
    ap.add_argument("--max-restarts", type=int, default=3)
    ap.add_argument("--heartbeat-timeout", type=float, default=10.0)
    return ap.parse_args(argv)


class ElasticServer:
    def __init__(self, launch: Callable[[int, Dict[str, int], str], List[subprocess.Popen]], num_devices: int,
                 strategy: ElasticStrategy, host: str = "127.0.0.1", port: int = 23457, heartbeat_timeout: float = 10.0,
                 max_restarts: int = 3, device_probe: Optional[Callable[[], int]] = None, nodes: Optional[Sequence[str]] = None,
                 node_runner: Callable[[str, str], Optional[str]] = _local_runner):
        self.launch, self.num_devices, self.strategy = launch, num_devices, strategy
        self.host, self.port, self.hb_timeout, self.max_restarts = host, port, heartbeat_timeout, max_restarts
        self.device_probe = device_probe
        self.nodes, self.node_runner = list(nodes) if nodes else None, node_runner
        self.generations: List[Dict] = []

    def survey(self) -> Optional[Dict[int, dict]]:
        """node / GPU detection before every generation (only when the job was started with a host list)"""
        return detect_node_info(self.nodes, self.node_runner) if self.nodes else None

    def run(self) -> int:
        """-> 0 when a generation finishes cleanly, 1 when restarts are exhausted"""
        devices = self.num_devices
        for gen in range(self.max_restarts + 1):
            info = self.survey()
            plan = self.strategy.plan_from_nodes(info) if info is not None else self.strategy.plan(devices)
            if plan is None:
                return 1
            srv = DeviceControllerServer(plan["num_devices"], self.host, self.port + gen, self.hb_timeout,
                                         host_to_ranks=plan.get("host_to_ranks"), host_to_local=plan.get("host_to_local")).start()
            procs = self.launch(gen, plan, f"{self.host}:{self.port + gen}")
            self.generations.append({"gen": gen, "plan": plan})
            failed = False
            while True:
                codes = [p.poll() for p in procs]
                if all(c == 0 for c in codes):
                    break
                if any(c not in (None, 0) for c in codes) or srv.dead_ranks():
                    failed = True
                    break
                time.sleep(0.2)
            if not failed:
                srv.shutdown()
                return 0
            srv.rpc_WorkerStop()                       # survivors poll AlreadyStop and leave the step loop
            deadline = time.time() + 5.0
            for p in procs:
                try:
                    p.wait(timeout=max(0.1, deadline - time.time()))
                except subprocess.TimeoutExpired:
                    p.kill()
            srv.shutdown()
            lost = sum(1 for p in procs if p.returncode not in (0, None))
            devices = self.device_probe() if self.device_probe else max(devices - max(lost, 1), 0)
        return 1


def main(argv=None):
    a = elastic_arg_parser(argv)

    def launch(gen, plan, addr):
        procs = []
        for r in range(plan["num_devices"]):
            cmd = a.command.format(rank=r, world=plan["num_devices"], dp=plan["dp"], tp=plan["tp"], pp=plan["pp"], addr=addr, gen=gen)
            procs.append(subprocess.Popen(cmd, shell=True))
        return procs
    return ElasticServer(launch, a.ngpus, ElasticStrategy(a.tp, a.pp), port=a.port, heartbeat_timeout=a.heartbeat_timeout,
                         max_restarts=a.max_restarts).run()


if __name__ == "__main__":
    sys.exit(main())
