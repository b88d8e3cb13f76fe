"""slice-wise collectives over heterogeneous holder groups: 3 ranks = a tp2 pipeline (ranks 0, 1) + a tp1 pipeline (rank 2)
holding a [8, 4] parameter split along dim 0: rank 0 rows 0-3, rank 1 rows 4-7, rank 2 all rows."""
import os

import numpy as np
import torch

import hetu_b200 as ht

ht.init_comm_group(3)
rank = int(os.environ["RANK"])
for grp in ([0, 2], [1, 2]):
    ht._C.comm_create_group(grp)            # collective over all ranks, same order everywhere
full = {r: np.arange(32, dtype=np.float32).reshape(8, 4) * (r + 1) for r in range(3)}
if rank == 2:
    local, offs, lens, groups = full[2], [0, 4], [4, 4], [[0, 2], [1, 2]]
else:
    local, offs, lens, groups = full[rank][rank * 4:(rank + 1) * 4], [0], [4], [[rank, 2]]
x = ht.from_numpy(torch.as_tensor(local.copy()))
ar = torch.as_tensor(ht.grouped_all_reduce(x, 0, offs, lens, groups).numpy())
rs = ht.grouped_reduce_scatter(x, 0, offs, lens, groups)
ag = torch.as_tensor(ht.grouped_all_gather(rs, 0, offs, lens, groups).numpy())
rs = torch.as_tensor(rs.numpy())
# expected: rows 0-3 summed over ranks {0, 2}, rows 4-7 over {1, 2}
tot = np.concatenate([full[0][:4] + full[2][:4], full[1][4:] + full[2][4:]])
want_ar = tot if rank == 2 else tot[rank * 4:(rank + 1) * 4]
if rank == 2:
    want_rs = np.concatenate([tot[2:4], tot[6:8]])          # second half of each slice (rank 2 is the 2nd member of both groups)
else:
    want_rs = tot[rank * 4:rank * 4 + 2]                    # first half of its slice
ok = np.allclose(ar.numpy(), want_ar) and np.allclose(rs.numpy(), want_rs) and np.allclose(ag.numpy(), want_ar)
print("GROUPED", rank, bool(ok), tuple(rs.shape), flush=True)
assert ok
