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


def _run_once(script, world, args, force_cpu, timeout, env_extra):
    import tempfile
    port = free_port()
    procs, logs = [], []
    for r in range(world):
        env = dict(os.environ)
        env.update({"RANK": str(r), "WORLD_SIZE": str(world), "LOCAL_RANK": str(r), "MASTER_ADDR": "127.0.0.1",
                    "MASTER_PORT": str(port), "PYTHONPATH": ROOT + os.pathsep + env.get("PYTHONPATH", ""),
                    "OMP_NUM_THREADS": "1", "HETU_PG_TIMEOUT_S": str(max(int(timeout) // 2, 60))})
        if force_cpu:
            env["HETU_B200_FORCE_CPU"] = "1"
            env["CUDA_VISIBLE_DEVICES"] = ""
        if env_extra:
            env.update(env_extra)
        # output goes to a file, not a pipe: a chatty worker (long tracebacks, NCCL_DEBUG) can never block on a full pipe
        log = tempfile.TemporaryFile(mode="w+", prefix=f"hb_worker{r}_")
        logs.append(log)
        procs.append(subprocess.Popen([sys.executable, script, *map(str, args)], env=env, stdout=log, stderr=subprocess.STDOUT, text=True))
    # one deadline for the whole group: a hung rank must not cost `timeout` per process
    import time
    deadline = time.time() + timeout
    timed_out = False
    while any(p.poll() is None for p in procs):
        if time.time() > deadline:
            timed_out = True
            for p in procs:
                if p.poll() is None:
                    p.kill()
            break
        time.sleep(0.1)
    outs = []
    for p, log in zip(procs, logs):
        p.wait()
        log.seek(0)
        outs.append(log.read())
        log.close()
    ok = (not timed_out) and all(p.returncode == 0 for p in procs)
    return ok, outs, timed_out


def run_workers(script, world, args=(), force_cpu=True, timeout=300, env_extra=None):
    """-> (ok, [stdout+stderr per rank]).  A run that TIMES OUT or fails to bind the rendezvous port (taken between
    `free_port()` and the workers' bind) is retried once on a fresh port; other failures are never retried."""
    ok, outs, timed_out = _run_once(script, world, args, force_cpu, timeout, env_extra)
    port_taken = (not ok) and any("Address already in use" in o or "EADDRINUSE" in o for o in outs)
    if (timed_out and force_cpu) or port_taken:       # GPU runs are not repeated after a timeout (they are long)
        ok, outs, timed_out = _run_once(script, world, args, force_cpu, timeout, env_extra)
        if timed_out:
            outs = [o + "\n[run_workers] timed out twice" for o in outs]
    return ok, outs
