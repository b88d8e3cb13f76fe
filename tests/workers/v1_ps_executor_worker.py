"""worker of a v1 `Executor(comm_mode='PS')` job: softmax regression written with the v1 op API; the variables live on the
parameter server started by heturun, the server applies SGD, every worker must end with the same weights"""
import os

import numpy as np

import hetu_b200.v1 as ht

wid = int(os.environ.get("HETU_PS_WORKER_ID", "0"))
rng = np.random.RandomState(100 + wid)
w_true = np.random.RandomState(0).randn(8, 3).astype(np.float32)
x = ht.placeholder_op("x", [32, 8])
y = ht.placeholder_op("y", [32, 3])
w = ht.Variable("w", value=np.zeros((8, 3), np.float32))
b = ht.Variable("b", value=np.zeros((3,), np.float32))
logits = ht.add_op(ht.matmul_op(x, w), ht.broadcastto_op(b, ht.matmul_op(x, w)))
loss = ht.reduce_mean_op(ht.softmaxcrossentropy_op(logits, y), [0])
train = ht.optim.SGDOptimizer(learning_rate=0.5).minimize(loss)
ex = ht.Executor({"train": [loss, train], "eval": [loss]}, comm_mode="PS")
first = last = None
for step in range(60):
    xb = rng.randn(32, 8).astype(np.float32)
    yb = np.eye(3, dtype=np.float32)[(xb @ w_true).argmax(1)]
    out = ex.run("train", feed_dict={x: xb, y: yb})
    v = float(out[0].asnumpy())
    first = v if first is None else first
    last = v
wv = ex.graph.get_param(w).float().numpy()
print(f"V1PS worker={wid} first={first:.4f} last={last:.4f} wsum={float(np.abs(wv).sum()):.5f} w00={float(wv[0, 0]):.6f}", flush=True)
