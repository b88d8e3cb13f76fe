"""DistributedStates algebra, comm classification and re-sharding planner (pure host logic)."""
import numpy as np
import pytest

import hetu_b200 as ht
from hetu_b200 import _C

DS = ht.DistributedStates
CT = _C.CommType


def g(n, kind="cpu"):
    return ht.DeviceGroup([f"{kind}:{i}" for i in range(n)])


def test_basic_states_and_order():
    ds = DS(8, {0: 2, -1: 4}, [0, -1])
    assert ds.device_num == 8 and ds.get_dim(0) == 2 and ds.get_dim(-1) == 4 and ds.get_dim(1) == 1
    assert ds.order == [0, -1]
    assert DS(4, {-1: 4}).check_pure_duplicate()
    with pytest.raises(_C.HetuError):
        DS(8, {0: 3})
    # default order: sorted keys
    assert DS(8, {1: 2, 0: 4}).order == [0, 1]
    assert ds.local_shape([16, 32]) == [8, 32] and ds.global_shape([8, 32]) == [16, 32]


def test_device_index_mapping():
    ds = DS(8, {0: 2, -1: 4}, [0, -1])        # dp outer, tp(dup) inner
    assert ds.map_device_to_state_index(5) == {0: 1, -1: 1}
    assert ds.get_dup_group_index(5) == 1
    assert ds.get_device_indices_by_dim(-1, 5) == [4, 5, 6, 7]
    assert ds.get_device_indices_by_dim(0, 5) == [1, 5]
    w = DS(8, {-1: 2, 0: 4}, [-1, 0])         # weight: dup over dp (outer), split over tp (inner)
    assert w.get_device_indices_by_dim(-1, 6) == [2, 6]
    b, s = w.local_slice([64, 16], 6)
    assert b == [32, 0] and s == [16, 16]


@pytest.mark.parametrize("src,dst,expect", [
    (DS(4, {-2: 4}, [-2]), DS(4, {-1: 4}, [-1]), CT.ALL_REDUCE),
    (DS(4, {0: 4}, [0]), DS(4, {-1: 4}, [-1]), CT.ALL_GATHER),
    (DS(4, {-2: 4}, [-2]), DS(4, {0: 4}, [0]), CT.REDUCE_SCATTER),
    (DS(4, {-1: 4}, [-1]), DS(4, {0: 4}, [0]), CT.SCATTER),
    (DS(8, {0: 2, -2: 4}, [0, -2]), DS(8, {0: 2, -1: 4}, [0, -1]), CT.ALL_REDUCE),       # row-parallel output
    (DS(8, {0: 2, -2: 4}, [0, -2]), DS(8, {0: 8}, [0]), CT.REDUCE_SCATTER),              # ... with sequence parallel
    (DS(8, {0: 8}, [0]), DS(8, {0: 2, -1: 4}, [0, -1]), CT.ALL_GATHER),                  # SP -> column-parallel input
    (DS(8, {-2: 2, 0: 4}, [-2, 0]), DS(8, {-1: 2, 0: 4}, [-1, 0]), CT.ALL_REDUCE),       # DP gradient
    (DS(8, {-2: 2, 0: 4}, [-2, 0]), DS(8, {0: 8}, [0]), CT.REDUCE_SCATTER),              # ZeRO gradient
    (DS(4, {0: 2, 1: 2}, [0, 1]), DS(4, {0: 4}, [0]), CT.BATCHED_ISEND_IRECV),           # irregular re-tiling
])
def test_comm_classification(src, dst, expect):
    grp = g(src.device_num)
    assert _C.classify_comm(src, grp, dst, grp) == expect


def test_p2p_and_unused():
    ds = DS(2, {0: 2}, [0])
    a, b = ht.DeviceGroup(["cpu:0", "cpu:1"]), ht.DeviceGroup(["cpu:2", "cpu:3"])
    assert _C.classify_comm(ds, a, ds, a) == CT.UNUSED
    assert _C.classify_comm(ds, a, ds, b) == CT.P2P
    assert _C.classify_comm(DS(2, {0: 2}), a, DS(2, {-1: 2}), b) == CT.BATCHED_ISEND_IRECV


def test_plan_comm_groups():
    src, dst = DS(8, {0: 2, -2: 4}, [0, -2]), DS(8, {0: 2, -1: 4}, [0, -1])
    t, dim, peers = _C.plan_comm(src, dst, g(8), 6)
    assert t == CT.ALL_REDUCE and peers == [4, 5, 6, 7]
    t, dim, peers = _C.plan_comm(DS(8, {-2: 2, 0: 4}, [-2, 0]), DS(8, {0: 8}, [0]), g(8), 6)
    assert t == CT.REDUCE_SCATTER and dim == 0 and peers == [2, 6]


def test_combine_and_reduce():
    ds = DS(8, {-2: 2, 0: 4}, [-2, 0])
    assert {k: v for k, v in ds.combine_states([-2], -1).items() if v > 1} == {-1: 2, 0: 4}
    assert ds.combine_order([-2], -1) == [-1, 0]
    assert ds.check_allreduce(DS(8, {-1: 2, 0: 4}, [-1, 0]))
    assert ds.check_reducescatter(DS(8, {0: 8}, [0]))
    assert not ds.check_allgather(DS(8, {0: 8}, [0]))


@pytest.mark.parametrize("algo", ["FCFS", "ROUND_ROBIN", "MULTI_NODE_ROUND_ROBIN", "GREEDY", "NEW_GREEDY"])
def test_resharding_plan_covers_every_destination_exactly_once(algo):
    """hot switch tp4 -> dp2 x tp2 of a [64, 32] weight: every destination element arrives exactly once"""
    shape = [64, 32]
    src = DS(4, {0: 4}, [0])
    dst = DS(4, {-1: 2, 0: 2}, [-1, 0])
    items, load = _C.plan_resharding(shape, src, [0, 1, 2, 3], dst, [0, 1, 2, 3], algo)
    cover = {r: np.zeros(shape, dtype=int) for r in range(4)}
    for s, d, begin, size in items:
        sl = tuple(slice(b, b + n) for b, n in zip(begin, size))
        cover[d][sl] += 1
        sb, ss = src.local_slice(shape, s)                      # the sender really owns the piece
        assert all(b >= x and b + n <= x + m for b, n, x, m in zip(begin, size, sb, ss))
    for r in range(4):
        db, dsz = dst.local_slice(shape, r)
        want = np.zeros(shape, dtype=int)
        want[tuple(slice(b, b + n) for b, n in zip(db, dsz))] = 1
        assert (cover[r] == want).all()


def test_resharding_balances_replicated_sources():
    shape = [64]
    src = DS(4, {-1: 4}, [-1])           # four replicas
    dst = DS(4, {0: 4}, [0])
    items, load = _C.plan_resharding(shape, src, [0, 1, 2, 3], dst, [4, 5, 6, 7], "GREEDY")
    senders = sorted(s for s, d, b, n in items)
    assert senders == [0, 1, 2, 3]        # greedy spreads the four pieces over the four replicas
    items, _ = _C.plan_resharding(shape, src, [0, 1, 2, 3], dst, [4, 5, 6, 7], "FCFS")
    assert all(s == 0 for s, d, b, n in items)


def test_union_and_hetero():
    u = ht.DistributedStatesUnion([DS(4, {0: 4}, [0])])
    h = u.to_hetero(0, 2)
    assert h.size() == 2 and h.hetero_dim == 0 and h.get(0).get_dim(0) == 2
    a = ht.DistributedStatesUnion([DS(2, {-2: 2}), DS(2, {-2: 2})], -2)
    b = ht.DistributedStatesUnion([DS(2, {-1: 2}), DS(2, {-1: 2})], -1)
    gu = ht.DeviceGroupUnion([ht.DeviceGroup(["cpu:0", "cpu:1"]), ht.DeviceGroup(["cpu:2", "cpu:3"])])
    assert _C.classify_comm_union(a, gu, b, gu) == CT.SPLIT_ALL_REDUCE


def test_int_symbol_and_schedules():
    s = ht.IntSymbol(12)
    e = (s + 4) * 2 // 8
    assert e.get_data() == 4
    s.set_data(28)
    assert e.get_data() == 8 and (s % 5).get_data() == 3
    sched = _C.generate_1f1b_schedule(4, 8)
    for st, tasks in enumerate(sched):
        f = [m for k, m in tasks if k == 0]
        b = [m for k, m in tasks if k == 1]
        assert f == list(range(8)) and b == list(range(8))
        # warm-up depth = number of later stages
        first_b = next(i for i, (k, m) in enumerate(tasks) if k == 1)
        assert first_b == min(4 - st - 1, 8) + 1
        # a micro-batch is never back-propagated before its forward
        seen = set()
        for k, m in tasks:
            if k == 0:
                seen.add(m)
            elif k == 1:
                assert m in seen
    gp = _C.generate_gpipe_schedule(2, 3)
    assert [k for k, m in gp[0]] == [0, 0, 0, 1, 1, 1]


def test_galvatron_dp_core():
    # 3 layers, 2 strategies: strategy 1 is faster but needs more memory; switching costs 5
    mem = [1, 3] * 3
    intra = [10.0, 4.0] * 3
    inter = [0.0, 5.0, 5.0, 0.0] * 3
    cost, strat, rem = _C.galvatron_dp(3, 10, 2, mem, intra, inter)
    assert strat == [1, 1, 1] and cost == 12.0
    cost, strat, rem = _C.galvatron_dp(3, 6, 2, mem, intra, inter)     # only 5 units usable -> one layer upgraded
    assert sum(strat) == 1 and cost == 10 + 10 + 4 + 5
    cost, strat, rem = _C.galvatron_dp(3, 3, 2, mem, intra, inter)
    assert strat == [] and cost == float("inf")


def test_embedding_cache_policies():
    c = _C.EmbeddingCache(2, 4, _C.CachePolicy.LRU, 0, 2)
    import torch
    out, miss = c.lookup([1, 2], [0, 0])
    assert miss == [0, 1]
    c.insert([1, 2], torch.ones(2, 4), [0, 0])
    out, miss = c.lookup([1], [0])
    assert miss == [] and out.sum() == 4
    c.insert([3], torch.full((1, 4), 3.0), [0])          # evicts key 2 (least recently used)
    assert c.contains(1) and c.contains(3) and not c.contains(2)
    pk, pg = c.update([1], torch.ones(1, 4), 0.1)
    assert pk == []                                       # below the push bound: kept locally
    for _ in range(2):
        pk, pg = c.update([1], torch.ones(1, 4), 0.1)
    assert pk == [1] and float(pg.sum()) == 12.0          # three accumulated gradients pushed at once
    out, miss = c.lookup([1], [5])                        # stale beyond pull bound -> must be refetched
    assert miss == [0]
    lfu = _C.EmbeddingCache(2, 1, _C.CachePolicy.LFU, 0, 0)
    lfu.insert([1, 2], torch.zeros(2, 1), [0, 0])
    lfu.lookup([1, 1, 1], [0, 0, 0])
    lfu.insert([3], torch.zeros(1, 1), [0])
    assert lfu.contains(1) and not lfu.contains(2)
