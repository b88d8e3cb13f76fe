"""Elastic controller: runs a training job as restartable generations.  It watches worker heartbeats, on a failure (or
a straggler report) tells every worker to stop, detects the surviving nodes / GPUs, asks an `ElasticStrategy` for a
new (dp, tp, pp) plan and relaunches the workers, which resume from the latest checkpoint.
(ref: python/hetu/rpc/heturpc_elastic_server.py -- heartbeat monitor :463-468, restart loop :785-795,
elastic_arg_parser.py)"""
from __future__ import annotations

import argparse
import subprocess
import sys
import time
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence

from .server import DeviceControllerServer


@dataclass
class ElasticStrategy:
    """picks the largest (dp, tp, pp) that fits the surviving device count, keeping tp and pp when possible"""
    tp: int = 1
    pp: int = 1
    min_dp: int = 1

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
                 max_restarts: int = 3, device_probe: Optional[Callable[[], int]] = None):
        self.launch, self.num_devices, self.strategy = launch, num_devices, strategy
        self.host, self.port, self.hb_timeout, self.max_restarts = host, port, heartbeat_timeout, max_restarts
        self.device_probe = device_probe
        self.generations: List[Dict] = []

    def run(self) -> int:
        """-> 0 when a generation finishes cleanly, 1 when restarts are exhausted"""
        devices = self.num_devices
        for gen in range(self.max_restarts + 1):
            plan = self.strategy.plan(devices)
            if plan is None:
                return 1
            srv = DeviceControllerServer(plan["num_devices"], self.host, self.port + gen, self.hb_timeout).start()
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
