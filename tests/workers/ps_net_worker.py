"""worker process of a parameter-server job started by the heturun-style launcher: BSP linear regression through the remote
parameter store (dense push / pull, barrier), a sparse table with the HET cache, SSP clocks and partial reduce"""

import numpy as np

from hetu_b200.v1.ps import CacheSparseTable, PSContext

ps = PSContext()                       # address / worker id come from the launcher's environment
w_id, n = ps.worker_id, ps.num_workers
rng = np.random.RandomState(100 + w_id)
true_w = np.arange(1, 9, dtype=np.float32)
if w_id == 0:
    ps.init_dense("w", np.zeros(8, np.float32), opt="sgd", lr=0.1)
    ps.server.init_sparse(ps.key("emb"), 32, 4, np.zeros(32 * 4, np.float32).tolist(), __import__("hetu_b200")._C.PsOptimizer.SGD, 0.5)
ps.barrier()
for step in range(60):
    w = ps.pull("w")
    X = rng.randn(16, 8).astype(np.float32)
    g = X.T @ (X @ w - X @ true_w) / 16 / n
    ps.push("w", g)
    ps.barrier()
err = float(np.abs(ps.pull("w") - true_w).max())
tab = CacheSparseTable(ps, "emb", 32, 4, limit=8, policy="LRU", bound=2, lr=0.5)
ids = np.array([w_id, 10, 11])
for _ in range(4):
    tab.embedding_lookup(ids)
    tab.embedding_update(ids, np.ones((3, 4), np.float32))
ps.barrier()
out, partners = ps.preduce("pr", np.full(4, float(w_id), np.float32), min_workers=n, wait_ms=2000)
ps.ssp_init(1) if w_id == 0 else None
ps.barrier()
for clock in range(3):
    ps.ssp_sync(clock)
stats = ps.server.stats()
print(f"PSNET worker={w_id} err={err:.4f} preduce={out.tolist()} partners={sorted(partners)} pushes={stats.get('push', stats)}", flush=True)
