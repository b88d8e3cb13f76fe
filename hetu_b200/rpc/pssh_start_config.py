"""python -m hetu.rpc.pssh_start_config --config-path DIR --config-name NAME "rpc.command='python3 train.py ...'"
starts the DeviceController for `rpc.num_gpus` workers, then launches the workers (local processes, or one ssh session per
worker over the hosts of `rpc.hosts` YAML) with HETU_RENDEZVOUS=rpc so they obtain their ranks from the controller.
(ref: python/hetu/rpc/pssh_start_config.py, pssh_start.py, local_start.py)"""
import os
import shlex
import sys

from ..utils import hydra_lite
from .launcher import local_start, pssh_start, read_hosts_yaml
from .server import DeviceControllerServer


def main(argv=None):
    c = hydra_lite.load(argv).rpc
    port = int(c.get("server_port", 23457))
    addr = c.get("server_addr", "127.0.0.1")
    n = int(c.get("num_gpus", c.get("ngpus", 1)))
    log_dir = c.get("log_path")
    if log_dir:
        os.makedirs(log_dir, exist_ok=True)
    srv = DeviceControllerServer(n, "0.0.0.0" if c.get("hosts") else "127.0.0.1", port).start()
    envs = {"HETU_RENDEZVOUS": "rpc", "HETU_RPC_SERVER": f"{addr}:{port}", **{str(k): str(v) for k, v in (c.get("envs") or {}).items()}}
    try:
        if c.get("hosts"):
            hosts = read_hosts_yaml(c.hosts)
            codes = pssh_start(c.command, hosts, master_port=port + 1, envs=envs, env_script=c.get("env_script"), ssh_user=c.get("ssh_user"),
                               rendezvous_env=False, timeout=c.get("timeout"))
        else:
            codes = local_start(shlex.split(c.command), n, master_port=port + 1, env=envs, log_dir=log_dir, rendezvous_env=False, timeout=c.get("timeout"))
    finally:
        srv.shutdown()
    return max(codes) if codes else 0


if __name__ == "__main__":
    sys.exit(main())
