"""Launch helpers for multi-process tests: spawn N workers of a script with a torchrun-style environment."""
import os
import socket
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def run_workers(script, world, args=(), force_cpu=True, timeout=600, env_extra=None):
    port = free_port()
    procs = []
    for r in range(world):
        env = dict(os.environ)
        env.update({"RANK": str(r), "WORLD_SIZE": str(world), "LOCAL_RANK": str(r), "MASTER_ADDR": "127.0.0.1",
                    "MASTER_PORT": str(port), "PYTHONPATH": ROOT + os.pathsep + env.get("PYTHONPATH", ""),
                    "OMP_NUM_THREADS": "1"})
        if force_cpu:
            env["HETU_B200_FORCE_CPU"] = "1"
            env["CUDA_VISIBLE_DEVICES"] = ""
        if env_extra:
            env.update(env_extra)
        procs.append(subprocess.Popen([sys.executable, script, *map(str, args)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = []
    ok = True
    for p in procs:
        try:
            o, _ = p.communicate(timeout=timeout)
        except subprocess.TimeoutExpired:
            p.kill()
            o, _ = p.communicate()
            ok = False
        outs.append(o)
        ok = ok and p.returncode == 0
    return ok, outs
