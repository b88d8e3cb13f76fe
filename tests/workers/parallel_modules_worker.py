"""single-strategy tensor-parallel modules (hetu.nn.ColumnParallelLinear / RowParallelLinear / VocabParallelEmbedding /
ParallelLayerNorm with a device group + dp): a toy embedding -> MLP -> norm network, one SGD step.  argv: tp degree.
Prints the outputs before / after the step; they must not depend on tp (the reference's tests/test_parallel.py scenario)."""
import json
import os
import sys

import numpy as np
import torch

import hetu_b200 as ht

tp = int(sys.argv[1])
ht.init_comm_group(tp)
rank = int(os.environ.get("RANK", "0"))
ht.set_seed(21)
grp = ht.DeviceGroup([f"cpu:{i}" for i in range(tp)])
with ht.graph("define_and_run", create_new=True) as g:
    emb = ht.nn.VocabParallelEmbedding(32, 16, grp, dp=1, name="pm_emb")
    col = ht.nn.ColumnParallelLinear(16, 24, grp, dp=1, gather_output=False, name="pm_col")
    row = ht.nn.RowParallelLinear(24, 16, grp, dp=1, name="pm_row")
    ln = ht.nn.ParallelLayerNorm(16, grp, dp=1, name="pm_ln")
    ds_in = ht.DistributedStates(tp, {-1: tp}, [-1])
    ids = ht.parallel_placeholder("int64", [6], [ds_in], device_group_hierarchy=[grp], name="ids")
    tgt = ht.parallel_placeholder("float32", [6, 16], [ds_in], device_group_hierarchy=[grp], name="tgt")
    y = ln(row(ht.relu(col(emb(ids)))))
    loss = ht.mean(ht.mse_loss(y, tgt, reduction="none"))
    train = ht.SGDOptimizer(lr=0.5).minimize(loss)
rng = np.random.RandomState(0)
feed = {ids: torch.as_tensor(rng.randint(0, 32, 6)), tgt: torch.as_tensor(rng.randn(6, 16).astype(np.float32))}
o1 = g.run(loss, [y, loss, train], feed)
o2 = g.run(loss, [y, loss], feed)
if rank == 0:
    print("PM " + json.dumps({"y0": o1[0].flatten().tolist(), "l0": float(o1[1]), "y1": o2[0].flatten().tolist(), "l1": float(o2[1])}))
