"""one program, three roles (as ps-lite programs are written): DMLC_ROLE picks scheduler / server / worker.  Workers push
gradients of a shared dense parameter and a sparse table through `get_worker_communicate()`."""
import json
import os

import numpy as np

import hetu_b200.v1 as v1

role = os.environ["DMLC_ROLE"]
if role == "scheduler":
    v1.scheduler_init()
    v1.scheduler_finish(120)
elif role == "server":
    v1.server_init()
    v1.server_finish(120)
else:
    v1.worker_init()
    ps = v1.get_worker_communicate()
    wid = ps.worker_id
    if wid == 0:
        ps.init_dense("w", np.zeros(10, np.float32), opt="sgd", lr=0.5)
        ps.init_sparse("emb", np.zeros((6, 2), np.float32), opt="sgd", lr=1.0)
    ps.barrier()
    for step in range(3):
        ps.push("w", np.full(10, float(wid + 1), np.float32))
        ps.barrier()
    ps.sparse_push("emb", [wid, 5], np.ones((2, 2), np.float32))
    ps.barrier()
    w = ps.pull("w", [10])
    emb = ps.sparse_pull("emb", [0, 1, 5], 2)
    print("ROLES " + json.dumps({"wid": wid, "w": float(w[0]), "same": bool(np.all(w == w[0])), "emb": emb.tolist(),
                                 "servers": ps.num_servers}), flush=True)
    v1.worker_finish()
