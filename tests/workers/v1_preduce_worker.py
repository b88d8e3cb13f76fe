"""scheduler / server / three workers; the workers also form a torch.distributed world.  Round 1: rank 2 is late, so ranks 0 and 1
reduce among themselves and rank 2 reduces alone; round 2: everybody is on time."""
import json
import os
import time

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
    from hetu_b200.v1.preduce import PartialReduce
    v1.worker_init()
    pr = PartialReduce()
    rank = pr.rank
    g = np.full(4, float(rank + 1), np.float32)
    if rank == 2:
        time.sleep(2.0)
    p1 = pr.get_partner(wait_time=500.0)
    pr.preduce(g, p1)
    first = g.copy()
    v1.get_worker_communicate().barrier()
    g = np.full(4, float(rank + 1), np.float32)
    p2 = pr.get_partner(max_worker=3, wait_time=20000.0)
    pr.preduce(g, p2)
    print("PREDUCE " + json.dumps({"rank": rank, "p1": list(p1), "first": first.tolist(), "p2": list(p2), "second": g.tolist()}), flush=True)
    v1.worker_finish()
