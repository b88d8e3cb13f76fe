"""Hetu 1.x style parameter-server training, one program for every role (ref: hetu/v1/examples/ctr + python/hetu/launcher.py):

    python examples/v1/train_ps_roles.py --config examples/v1/local_ps.yml

`launch(train, args)` starts the scheduler, the servers and the workers as processes of this program; every worker runs `train`
after `worker_init()`.  The model is logistic regression on a planted-signal stream; the dense weight lives on the servers
(sharded over both), the workers push gradients and pull the fresh weight every step (BSP)."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


def train(args):
    import hetu_b200.v1 as ht
    ps = ht.get_worker_communicate()
    rank, nworkers = ps.worker_id, ps.num_workers
    rng = np.random.RandomState(100 + rank)
    true_w = np.random.RandomState(0).randn(args.dim).astype(np.float32)

    x = ht.placeholder_op("x", [args.batch, args.dim])
    y = ht.placeholder_op("y", [args.batch, 1])
    w = ht.Variable("w", value=np.zeros((args.dim, 1), np.float32))
    logit = ht.matmul_op(x, w)
    loss = ht.reduce_mean_op(ht.binarycrossentropywithlogits_op(logit, y), [0, 1])
    grad = ht.gradients(loss, [w])[0]
    # the gradient goes to the servers, which apply SGD; the executor pulls the new weight back before the next step
    push = ht.parameterServerCommunicate_op(grad, w, ht.optim.SGDOptimizer(args.lr))
    ex = ht.Executor([loss, push])
    logger = ht.HetuLogger(rank=rank, nrank=nworkers, echo=(rank == 0))
    for step in range(args.steps):
        xb = rng.randn(args.batch, args.dim).astype(np.float32)
        yb = (xb @ true_w > 0).astype(np.float32).reshape(-1, 1)
        out = ex.run(feed_dict={x: xb, y: yb}, convert_to_numpy_ret_vals=True)
        if step % 10 == 0 or step == args.steps - 1:
            logger.log("step", step)
            logger.log("loss", float(out[0]))
            logger.step()
    final = ex.graph.get_param(w).numpy().reshape(-1)
    cos = float(final @ true_w / (np.linalg.norm(final) * np.linalg.norm(true_w) + 1e-12))
    print(f"[worker {rank}] cosine(w, w*) = {cos:.3f} over {ps.num_servers} servers", flush=True)
    if args.out:
        with open(os.path.join(args.out, f"worker{rank}.txt"), "w") as f:
            f.write(f"{cos:.6f}\n")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default=os.path.join(os.path.dirname(os.path.abspath(__file__)), "local_ps.yml"))
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--dim", type=int, default=16)
    ap.add_argument("--lr", type=float, default=0.5)
    ap.add_argument("--out", default="")
    ap.add_argument("--port", type=int, default=None, help="override the scheduler port of the config (0: pick a free one)")
    a = ap.parse_args()
    if a.port is not None:
        import socket
        import tempfile

        import yaml
        cfg = yaml.safe_load(open(a.config))
        if a.port == 0:
            with socket.socket() as sock:
                sock.bind(("127.0.0.1", 0))
                a.port = sock.getsockname()[1]
        cfg["shared"]["DMLC_PS_ROOT_PORT"] = a.port
        tmp = tempfile.NamedTemporaryFile("w", suffix=".yml", delete=False)
        yaml.safe_dump(cfg, tmp)
        tmp.close()
        a.config = tmp.name
    from hetu_b200.v1 import launcher
    codes = launcher.launch(train, a, timeout=600)
    sys.exit(0 if all(c == 0 for c in codes) else 1)
