"""worker process of a sharded parameter-server job (scheduler + 2 servers started by heturun): BSP linear regression whose
weight vector is split over the servers' key ranges, plus an embedding table split by row % S"""
import numpy as np

from hetu_b200.v1.ps import ShardedPSContext, connect

ps = connect()
assert isinstance(ps, ShardedPSContext) and ps.num_servers == 2
w_true = np.linspace(-1, 1, 9).astype(np.float32)
rng = np.random.default_rng(ps.worker_id)
if ps.worker_id == 0:
    ps.init_dense("w", np.zeros(9, np.float32), opt="sgd", lr=0.1)
    ps.init_sparse("emb", np.zeros((8, 2), np.float32), opt="sgd", lr=1.0)
ps.barrier()
ps._dense_len["w"] = 9
for step in range(150):
    x = rng.standard_normal((16, 9)).astype(np.float32)
    w = ps.pull("w")
    g = x.T @ (x @ w - x @ w_true) / 16 / ps.num_workers
    ps.push("w", g)
    ps.barrier()
err = float(np.abs(ps.pull("w") - w_true).max())
ps.sparse_push("emb", [ps.worker_id, 7], -np.ones((2, 2), np.float32))
ps.barrier()
emb = ps.sparse_pull("emb", list(range(8)), 2)
print(f"PSSHARD worker={ps.worker_id} err={err:.4f} emb7={emb[7, 0]:.1f} own={emb[ps.worker_id, 0]:.1f} dead={len(ps.dead_nodes(30.0))}", flush=True)
ps.finalize()
