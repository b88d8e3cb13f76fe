"""Launchers: N local worker processes (one per GPU) or parallel-ssh over the hosts of a YAML file
(ref: python/hetu/rpc/local_start.py, pssh_start.py, pssh_start_config.py, pssh_workers.py)."""
from __future__ import annotations

import os
import shlex
import subprocess
import sys
from typing import Dict, List, Optional, Sequence

import yaml


def read_hosts_yaml(path: str) -> List[Dict]:
    """hosts: [{addr: host0, initial_workers: 8}, ...]  (the reference's hostfile format)"""
    with open(path) as f:
        doc = yaml.safe_load(f)
    return [{"addr": h.get("addr", h.get("host", "127.0.0.1")), "workers": int(h.get("initial_workers", h.get("workers", 1)))}
            for h in doc.get("hosts", [])]


def _env(rank, local_rank, world, master_addr, master_port, extra, rendezvous_env=True):
    e = dict(os.environ)
    if rendezvous_env:
        e.update({"RANK": str(rank), "LOCAL_RANK": str(local_rank), "WORLD_SIZE": str(world), "MASTER_ADDR": master_addr,
                  "MASTER_PORT": str(master_port)})
    else:       # ranks come from the DeviceController (HETU_RENDEZVOUS=rpc)
        for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT"):
            e.pop(k, None)
    e["HETU_LOCAL_HOSTNAME"] = master_addr
    e.update(extra or {})
    return e


def local_start(command: Sequence[str], ngpus: int, master_port: int = 29500, env: Optional[Dict[str, str]] = None,
                log_dir: Optional[str] = None, wait: bool = True, rendezvous_env: bool = True, timeout: Optional[float] = None):
    """spawn `ngpus` copies of `command` on this node with torchrun-style env; returns exit codes (or the Popen list)"""
    procs = []
    for r in range(ngpus):
        out = open(os.path.join(log_dir, f"rank{r}.log"), "w") if log_dir else None
        procs.append(subprocess.Popen(list(command), env=_env(r, r, ngpus, "127.0.0.1", master_port, env, rendezvous_env), stdout=out,
                                      stderr=subprocess.STDOUT if out else None))
    if not wait:
        return procs
    return _wait_all(procs, timeout)


def _wait_all(procs, timeout: Optional[float] = None):
    """exit codes; with a timeout the stragglers are killed when it expires (exit code -9)"""
    import time
    deadline = None if timeout is None else time.time() + float(timeout)
    while any(p.poll() is None for p in procs):
        if deadline is not None and time.time() > deadline:
            for p in procs:
                if p.poll() is None:
                    p.kill()
            break
        time.sleep(0.1)
    return [p.wait() for p in procs]


def pssh_start(command: str, hosts: List[Dict], master_port: int = 29500, envs: Optional[Dict[str, str]] = None,
               env_script: Optional[str] = None, ssh_user: Optional[str] = None, dry_run: bool = False, rendezvous_env: bool = True,
               timeout: Optional[float] = None):
    """one ssh session per worker; hosts whose addr is local run without ssh.  Returns exit codes (or the command lines
    when dry_run)."""
    world = sum(h["workers"] for h in hosts)
    master = hosts[0]["addr"]
    lines, procs, rank = [], [], 0
    for h in hosts:
        for lr in range(h["workers"]):
            exports = " ".join(f"{k}={shlex.quote(str(v))}" for k, v in
                               {**({"RANK": rank, "LOCAL_RANK": lr, "WORLD_SIZE": world, "MASTER_ADDR": master, "MASTER_PORT": master_port}
                                   if rendezvous_env else {}), **(envs or {})}.items())
            inner = (f"source {shlex.quote(env_script)} && " if env_script else "") + f"env {exports} {command}"
            if h["addr"] in ("127.0.0.1", "localhost"):
                line = ["bash", "-c", inner]
            else:
                target = f"{ssh_user}@{h['addr']}" if ssh_user else h["addr"]
                line = ["ssh", "-o", "StrictHostKeyChecking=no", target, inner]
            lines.append(line)
            if not dry_run:
                procs.append(subprocess.Popen(line))
            rank += 1
    if dry_run:
        return lines
    return _wait_all(procs, timeout)


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--hosts", default=None, help="hosts YAML; omitted = local")
    ap.add_argument("--ngpus", type=int, default=1)
    ap.add_argument("--port", type=int, default=29500)
    ap.add_argument("command", nargs=argparse.REMAINDER)
    a = ap.parse_args()
    cmd = a.command[1:] if a.command and a.command[0] == "--" else a.command
    if a.hosts:
        sys.exit(max(pssh_start(" ".join(shlex.quote(c) for c in cmd), read_hosts_yaml(a.hosts), a.port)))
    sys.exit(max(local_start(cmd, a.ngpus, a.port)))
