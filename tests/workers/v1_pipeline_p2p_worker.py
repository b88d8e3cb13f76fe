"""two ranks, a hand-placed v1 pipeline: rank 0 computes h = relu(x @ w0) and hands it over with pipeline_send_op; rank 1 takes it
with pipeline_receive_op and finishes y = h @ w1.  Also exercises the raw comm_send / comm_recv bindings."""
import json

import numpy as np
import torch
import torch.distributed as dist

import hetu_b200 as ht
import hetu_b200.v1 as v1

ht.init_comm_group()
rank = dist.get_rank()
rng = np.random.RandomState(0)
x, w0, w1 = rng.randn(4, 6).astype(np.float32), rng.randn(6, 5).astype(np.float32), rng.randn(5, 3).astype(np.float32)
if rank == 0:
    h = v1.relu_op(v1.matmul_op(v1.Variable("x", value=x, trainable=False), v1.Variable("w0", value=w0, trainable=False)))
    token = v1.pipeline_send_op(h, 1)
    out = v1.Executor([token]).run(feed_dict={}, convert_to_numpy_ret_vals=True)[0]
    assert out.shape == (1,)
    ht._C.comm_send(torch.arange(6, dtype=torch.float32).reshape(2, 3), 1, 1)
else:
    h = v1.pipeline_receive_op(0, shape=[4, 5])
    y = v1.matmul_op(h, v1.Variable("w1", value=w1, trainable=False))
    got = v1.Executor([y]).run(feed_dict={}, convert_to_numpy_ret_vals=True)[0]
    raw = ht._C.comm_recv([2, 3], "float32", 0, 1)
    want = np.maximum(x @ w0, 0) @ w1
    print("P2P " + json.dumps({"err": float(np.abs(got - want).max()), "raw": raw.reshape(-1).tolist()}), flush=True)
# the v1 communicator handle: world + a collectively created sub-group, numpy in / numpy out, v1 spellings
comm = v1.wrapped_mpi_nccl_init()
assert comm.nrank == 2 and comm.rank == rank and v1.get_mpi_communicate() is comm
total = comm.all_reduce(np.full(3, rank + 1.0, np.float32))
gathered = comm.all_gather(np.full((1, 2), float(rank), np.float32))
out = np.zeros(3, np.float32)
comm.dlarrayNcclAllReduce(np.full(3, 2.0, np.float32), out)
bc = comm.broadcast(np.array([7.0 + rank], np.float32), root=1)
sub = v1.new_group_comm([0, 1])
assert sub.nrank == 2
nprof = v1.NCCLProfiler()
t_ar = nprof.profile_allreduce(4096, [0, 1], num_iterations=3)
t_ag = nprof.profile_allreduce(4096, [0, 1], num_iterations=3, primitive=v1.NCCLOP.AllGather)
t_p2p = nprof.profile_sendrecv(4096, [0, 1], num_iterations=3)
assert t_ar > 0 and t_ag > 0 and t_p2p > 0, (t_ar, t_ag, t_p2p)
if rank == 1:
    print("COMM " + json.dumps({"sum": total.tolist(), "gather": gathered.tolist(), "out": out.tolist(), "bc": bc.tolist()}), flush=True)
dist.barrier()
