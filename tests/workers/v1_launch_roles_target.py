"""target of the function-style v1 launcher: runs inside every worker process after worker_init()"""
import json
import os

import numpy as np


def train(args):
    import hetu_b200.v1 as v1
    ps = v1.get_worker_communicate()
    if ps.worker_id == 0:
        ps.init_dense("w", np.zeros(4, np.float32), opt="sgd", lr=1.0)
    ps.barrier()
    ps.push("w", np.ones(4, np.float32))
    ps.barrier()
    w = ps.pull("w", [4])
    with open(os.path.join(args.out, f"worker{ps.worker_id}.json"), "w") as f:
        json.dump({"w": w.tolist(), "role": os.environ["DMLC_ROLE"]}, f)
