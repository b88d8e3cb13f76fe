"""heturun: start the roles of a parameter-server job on this node from a YAML file

    python -m hetu.v1.launcher -c config.yaml python train.py --comm PS

    shared:            # environment of every role
      DMLC_PS_ROOT_PORT: 13100
    launch:
      worker: 4        # worker processes (HETU_PS_WORKER_ID = 0..3, HETU_PS_ADDRESS points at the server)
      server: 1        # parameter servers hosted by this launcher process (native transport, csrc/v1/ps_net.cc)
      scheduler: 1     # with more than one server the launcher also runs the native scheduler (csrc/v1/ps_scheduler.cc):
                       # servers and workers register there, workers get HETU_PS_SCHEDULER and shard parameters over the servers

(ref: hetu/v1/python/hetu/launcher.py `launch`, bin/heturun; ps-lite scheduler / server / worker roles)"""
from __future__ import annotations

import argparse
import os
import signal
import subprocess
import sys
from typing import Dict, List, Optional, Sequence

import yaml

from .ps import PSContext


def launch(command, settings, log_dir: Optional[str] = None, wait: bool = True, timeout: Optional[float] = None):
    if callable(command):                       # v1 form: launch(target, args) -- every role a process of this program
        return launch_roles(command, settings, timeout)
    shared = {str(k): str(v) for k, v in (settings.get("shared") or {}).items()}
    lc = settings.get("launch") or {}
    n_worker, n_server = int(lc.get("worker", 1)), int(lc.get("server", 1))
    port = int(shared.get("DMLC_PS_ROOT_PORT", 0))
    sched, shard_servers = None, []
    if n_server > 1:
        import threading

        from .. import _C
        from .ps import ShardedPSContext
        sched = _C.PsScheduler(n_server, n_worker, port, "0.0.0.0")
        addr = f"{shared.get('DMLC_PS_ROOT_URI', '127.0.0.1')}:{sched.port}"
        # registration blocks until the workers are up too: the server roles register from threads
        for _ in range(n_server):
            t = threading.Thread(target=lambda: shard_servers.append(ShardedPSContext.serve(addr, num_workers=n_worker, heartbeat_s=1.0)), daemon=True)
            t.start()
        server = None
    else:
        # a fresh store per job: the in-process registry is keyed by name and would otherwise hand a later job the parameters of
        # an earlier one launched from the same interpreter
        import uuid
        server = PSContext.serve(n_worker, port, name=f"heturun-{uuid.uuid4().hex[:8]}") if n_server > 0 else None
    procs: List[subprocess.Popen] = []
    for w in range(n_worker):
        env = dict(os.environ)
        env.update(shared)
        env.update({"DMLC_ROLE": "worker", "DMLC_NUM_WORKER": str(n_worker), "DMLC_NUM_SERVER": str(n_server), "HETU_PS_WORKER_ID": str(w),
                    "WORKER_ID": str(w)})
        if server is not None:
            env["HETU_PS_ADDRESS"] = f"{shared.get('DMLC_PS_ROOT_URI', '127.0.0.1')}:{server.port}"
        if sched is not None:
            env["HETU_PS_SCHEDULER"] = f"{shared.get('DMLC_PS_ROOT_URI', '127.0.0.1')}:{sched.port}"
        out = open(os.path.join(log_dir, f"worker{w}.log"), "w") if log_dir else None
        procs.append(subprocess.Popen(list(command), env=env, stdout=out, stderr=subprocess.STDOUT if out else None))

    def stop(*_):
        for p in procs:
            if p.poll() is None:
                p.kill()
    signal.signal(signal.SIGINT, stop)
    if not wait:
        return procs, (server if sched is None else sched)
    from ..rpc.launcher import _wait_all
    codes = _wait_all(procs, timeout)
    if server is not None:
        server.stop()
    if sched is not None:
        for net, client in shard_servers:
            client.finalize()
            net.stop()
        sched.stop()
    return codes


def main(argv=None):
    ap = argparse.ArgumentParser(prog="heturun")
    ap.add_argument("-c", "--config", required=True)
    ap.add_argument("--log-dir", default=None)
    ap.add_argument("command", nargs=argparse.REMAINDER)
    a = ap.parse_args(argv)
    with open(a.config) as f:
        settings = yaml.safe_load(f) or {}
    if a.log_dir:
        os.makedirs(a.log_dir, exist_ok=True)
    cmd = a.command[1:] if a.command and a.command[0] == "--" else a.command
    return max(launch(cmd, settings, a.log_dir))


if __name__ == "__main__":
    sys.exit(main())


# ---- the function-style launcher of v1 programs (ref: hetu/v1/python/hetu/launcher.py): every role is a process of this program
_procs: List = []


def signal_handler(signum=None, frame=None):
    print("SIGINT signal caught, stop Training")
    for p in _procs:
        if p.is_alive():
            p.kill()
    raise SystemExit(0)


def start_sched():
    os.environ["DMLC_ROLE"] = "scheduler"
    from . import runtime_api as api
    api.scheduler_init()
    api.scheduler_finish()


def start_server():
    os.environ["DMLC_ROLE"] = "server"
    from . import runtime_api as api
    api.server_init()
    api.server_finish()


def start_worker(target, args):
    os.environ["DMLC_ROLE"] = "worker"
    from . import runtime_api as api
    api.worker_init()
    target(args)
    api.worker_finish()


def launch_roles(target, args, join_timeout: Optional[float] = None):
    """`launch(target, args)` of v1: args.config names a yaml with `shared` (DMLC_* environment) and `launch` {worker, server,
    scheduler}; runs `target(args)` in every worker process next to the server / scheduler roles and waits for all of them.
    -> exit codes"""
    import multiprocessing as mp

    import yaml
    settings = yaml.safe_load(open(args.config).read())
    for k, v in (settings.get("shared") or {}).items():
        os.environ[str(k)] = str(v)
    lc = settings.get("launch") or {}
    ctx = mp.get_context("spawn")
    del _procs[:]
    if int(lc.get("scheduler", 0)) != 0:
        _procs.append(ctx.Process(target=start_sched))
    for _ in range(int(lc.get("server", 0))):
        _procs.append(ctx.Process(target=start_server))
    for _ in range(int(lc.get("worker", 1))):
        _procs.append(ctx.Process(target=start_worker, args=(target, args)))
    signal.signal(signal.SIGINT, signal_handler)
    for i, p in enumerate(_procs):
        p.start()
        if i == 0 and int(lc.get("scheduler", 0)) != 0:
            import time
            time.sleep(1.0)                     # the scheduler listens before the others dial it
    import time
    deadline = None if join_timeout is None else time.time() + float(join_timeout)
    while any(p.is_alive() for p in _procs) and (deadline is None or time.time() < deadline):
        if any(p.exitcode not in (None, 0) for p in _procs):      # a role died: the others would wait for it forever
            break
        time.sleep(0.2)
    codes = [p.exitcode for p in _procs]
    for p in _procs:
        if p.is_alive():
            p.kill()
    return codes
