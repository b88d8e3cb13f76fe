"""Native runtime pieces (C++): caching memory pool, random-state bookkeeping, prefetching data loader, stream roles."""
import torch

import pytest

import hetu_b200 as ht
from hetu_b200 import _C


def test_memory_pool_split_merge_reuse_and_limits():
    pool = _C.MemoryPool("host", limit_mb=64)
    a = pool.alloc(300 << 10)          # small pool: carved from a 2 MiB segment
    b = pool.alloc(300 << 10)
    st = pool.stats()
    assert st["num_segment_alloc"] == 1 and st["num_split"] >= 2 and st["allocated"] >= 600 << 10
    pool.free(a)
    c = pool.alloc(200 << 10)          # fits in the hole left by `a`
    assert c == a and pool.stats()["cache_hits"] >= 1
    pool.free(b)
    pool.free(c)
    st = pool.stats()
    assert st["allocated"] == 0 and st["num_merge"] >= 2
    assert pool.empty_cache() == 2 << 20 and pool.stats()["reserved"] == 0
    big = pool.alloc(30 << 20)
    assert pool.stats()["reserved"] >= 30 << 20
    try:
        pool.alloc(60 << 20)
        raised = False
    except Exception as e:     # noqa: BLE001
        raised = "out of memory" in str(e)
    assert raised
    pool.free(big)
    # per-stream free lists: a block freed on stream 1 is not handed to stream 2
    p1 = pool.alloc(4 << 20, stream=1)
    pool.free(p1)
    p2 = pool.alloc(4 << 20, stream=2)
    assert p2 != p1
    p3 = pool.alloc(4 << 20, stream=1)
    assert p3 == p1
    assert "host pool" in pool.summary()


def test_random_state_offsets_are_reserved_monotonically():
    _C.random_set_seed(123)
    assert _C.random_seed() == 123 and _C.random_offset() == 0
    a = _C.random_next_offset(1000)
    b = _C.random_next_offset(10)
    assert (a, b) == (0, 1000) and _C.random_offset() == 1010
    _C.random_set_offset(a)          # replay (activation recompute) starts from the recorded offset
    assert _C.random_next_offset(1000) == 0


def test_native_dataloader_dp_slices_prefetch_and_reset():
    data = torch.arange(40).reshape(20, 2)
    d0 = _C.Dataloader(data, 8, dp_rank=0, dp_size=2)
    d1 = _C.Dataloader(data, 8, dp_rank=1, dp_size=2)
    assert d0.num_batches == 2
    a, b = d0.next(), d1.next()
    assert a[:, 0].tolist() == [0, 4, 8, 12] and b[:, 0].tolist() == [2, 6, 10, 14]
    d0.next()
    wrap = d0.next()                    # second epoch starts over
    assert wrap[:, 0].tolist() == [0, 4, 8, 12]
    d0.reset(1)
    assert d0.next()[:, 0].tolist() == [16, 20, 24, 28]
    sh = _C.Dataloader(data, 10, shuffle=True, seed=5)
    x, y = sh.next(), sh.next()
    assert sorted(torch.cat([x, y])[:, 0].tolist()) == list(range(0, 40, 2))


def test_stream_role_names():
    assert _C.stream_role_name(1) == "computing" and _C.stream_role_name(6) == "collective" and _C.stream_role_name(3) == "h2d"


def test_shared_memory_pool_and_registry(tmp_path):
    """POSIX shared-memory backend (AllocShareMemory): a block written through a zero-copy tensor view is readable from
    another process that maps the segment by name; the per-device registry hands out one pool per device string"""
    import subprocess
    import sys
    import torch
    pool = _C.get_memory_pool("shm")
    assert _C.get_memory_pool("shm") is pool and "shm" in _C.memory_pool_devices()
    assert _C.get_memory_pool("cpu") is not pool
    p = pool.alloc(4096 * 4)
    t = pool.as_tensor(p, [64, 64], "float32")
    t.copy_(torch.arange(4096, dtype=torch.float32).reshape(64, 64))
    name, off = pool.shm_locate(p)
    assert name.startswith("/hetu_b200_") and off >= 0
    code = ("import mmap, os, sys, numpy as np\n"
            f"fd = os.open('/dev/shm{name}', os.O_RDONLY)\n"
            "m = mmap.mmap(fd, 0, prot=mmap.PROT_READ)\n"
            f"a = np.frombuffer(m, dtype=np.float32, count=4096, offset={off})\n"
            "print(float(a.sum()), float(a[4095]))\n")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert out.stdout.split() == [str(float(sum(range(4096)))), "4095.0"]
    q = pool.alloc(1024)
    assert pool.shm_locate(q)[0] == name            # small blocks are carved out of the same segment
    pool.free(p); pool.free(q)
    assert _C.get_memory_pool("cpu").shm_locate(0) is None
    assert _C.empty_all_memory_caches() >= 0


def test_cuda_profiler_degrades_gracefully_without_a_gpu(tmp_path):
    """memory / NVLink / clock profiler (G17): without a GPU or NVML every probe returns zeros instead of failing, and the
    per-micro-batch memory log is still written"""
    import json
    from hetu_b200.utils.profiler import CUDAProfiler, get_cuda_profiler
    p = CUDAProfiler(log_file=str(tmp_path / "mem.jsonl"))
    info = p.get_current_memory_info()
    assert info.mempool_allocated >= 0 and info.all_reserved >= 0
    p.record_micro_batch(True, 0, 1, info, info)
    p.record_micro_batch(False, 0, 1, info, info)
    rows = [json.loads(l) for l in open(tmp_path / "mem.jsonl")]
    assert [r["is_forward"] for r in rows] == [True, False] and rows[0]["micro_batch_id"] == 1 and "begin" in rows[0]
    p.profile_nvlink_start()
    nv = p.profile_nvlink_end()
    assert nv["tx_bytes"] >= 0 and nv["seconds"] >= 0 and isinstance(p.clocks(), dict)
    assert get_cuda_profiler() is get_cuda_profiler()


def test_context_store_typed_entries_pop_and_migrate():
    """ref: hetu/utils/context_store.h -- typed put / get / pop, `contains`, migrate_from with a new key"""
    C = ht._C
    a, b = C.ContextStore(), C.ContextStore()
    a.put("flag", True); a.put("n", 7); a.put("eps", 1e-5); a.put("name", "ln"); a.put("shape", [2, 3]); a.put("scales", [0.5, 2.0])
    t = torch.arange(6.0).reshape(2, 3)
    a.put("saved", t)
    assert a.get("flag") is True and a.get("n") == 7 and a.get("eps") == 1e-5 and a.get("name") == "ln"
    assert a.get("shape") == [2, 3] and a.get("scales") == [0.5, 2.0] and torch.equal(a.get("saved"), t)
    assert a.get("missing") is None and a.get("missing", 3) == 3 and "n" in a and len(a) == 7
    assert torch.equal(a.pop("saved"), t) and "saved" not in a
    b.migrate_from(a, "n", "count")
    assert b.get("count") == 7 and "n" not in a
    b.migrate_from(a, "absent")              # nothing to move: no entry appears
    assert b.keys() == ["count"]
    with pytest.raises(ht.HetuError):
        a.pop("saved")


def test_task_queue_runs_tasks_on_workers_with_back_pressure_and_error_report():
    """ref: hetu/utils/task_queue.h -- N workers drain a bounded queue; `wait` returns when everything submitted has run and
    surfaces a task's exception; shutdown drains what is queued"""
    import threading
    import time
    C = ht._C
    q = C.TaskQueue("io", num_workers=3, max_pending=4)
    seen, lock, tids = [], threading.Lock(), set()

    def work(i):
        def run():
            time.sleep(0.005)
            with lock:
                seen.append(i); tids.add(threading.get_ident())
        return run
    for i in range(40):
        q.add(work(i))                       # blocks whenever 4 tasks are pending
    q.wait()
    assert sorted(seen) == list(range(40)) and q.completed == 40 and q.pending == 0 and q.num_workers == 3
    assert threading.get_ident() not in tids and len(tids) >= 2

    def boom():
        raise ValueError("disk full")
    q.add(boom)
    with pytest.raises(ht.HetuError, match="disk full"):
        q.wait()
    q.add(work(99)); q.wait()                # the queue stays usable after a failed task
    assert seen[-1] == 99
    for i in range(5):
        q.add(work(100 + i))
    q.shutdown()                             # drains
    assert not q.running and sorted(seen)[-5:] == [100, 101, 102, 103, 104]
    with pytest.raises(ht.HetuError):
        q.add(work(0))
    q2 = C.TaskQueue("gc", 1)
    q2.add(work(7))
    del q2                                   # destructor joins the worker without dead-locking on the GIL


def test_bfc_pool_best_fit_split_coalesce_and_region_growth():
    """ref: hetu/impl/memory/CUDABFCMemoryPool -- bins by power of two, best fit, split, coalescing with both neighbours,
    doubling regions, limit, empty_cache returns whole free regions"""
    MB = 1 << 20
    pool = _C.BFCMemoryPool("host", initial_region_mb=4, limit_mb=64)
    assert pool.kind == "bfc" and pool.num_regions == 0
    a = pool.alloc(1 * MB); b = pool.alloc(1 * MB); c = pool.alloc(1 * MB)
    assert pool.num_regions == 1 and b == a + MB and c == b + MB            # carved contiguously out of the first 4 MiB region
    st = pool.stats()
    assert st["reserved"] == 4 * MB and st["allocated"] == 3 * MB and st["num_split"] == 3
    pool.free(b)
    assert pool.largest_free_chunk == MB and pool.fragmentation == 0.5      # two free 1 MiB chunks, not adjacent
    # best fit: a 512 KiB request takes the 1 MiB hole (smallest chunk that fits), not the tail
    d = pool.alloc(512 << 10)
    assert d == b
    pool.free(d); pool.free(a)                                              # a + (d + its remainder) coalesce into 2 MiB
    assert pool.largest_free_chunk == 2 * MB
    pool.free(c)                                                            # everything merges back into one 4 MiB chunk
    assert pool.largest_free_chunk == 4 * MB and pool.fragmentation == 0.0 and sum(pool.bin_occupancy()) == 1
    assert pool.stats()["num_merge"] >= 4
    # growth: a request beyond the first region opens a second, larger one (doubling schedule)
    big = pool.alloc(6 * MB)
    assert pool.num_regions == 2 and pool.stats()["reserved"] == 4 * MB + 8 * MB
    # odd sizes are rounded to the 256-byte granularity and requests reuse freed chunks
    x = pool.alloc(1000); pool.free(x)
    assert pool.alloc(1024) == x
    hits = pool.stats()["cache_hits"]
    assert hits >= 3
    # limit: 64 MiB cap refuses a request that cannot fit
    assert pool.alloc(80 * MB) == 0
    pool.free(big); pool.free(x)
    released = pool.empty_cache()
    assert released == 12 * MB and pool.num_regions == 0 and pool.stats()["reserved"] == 0
    assert "bfc pool" in pool.summary()


def test_bfc_pool_random_workload_never_overlaps_and_returns_to_one_chunk_per_region():
    import random
    rng = random.Random(0)
    pool = _C.BFCMemoryPool("host", initial_region_mb=2, limit_mb=256)
    live = {}
    for step in range(3000):
        if live and (rng.random() < 0.45 or len(live) > 200):
            p = rng.choice(list(live)); pool.free(p); del live[p]
        else:
            n = rng.choice([300, 4096, 70_000, 1 << 20, 3 << 20]) + rng.randrange(0, 257)
            p = pool.alloc(n, stream=rng.choice([0, 0, 0, 7]))
            assert p != 0 and p % 256 == 0
            live[p] = n
        if step % 500 == 0:
            spans = sorted((p, p + n) for p, n in live.items())
            assert all(spans[i][1] <= spans[i + 1][0] for i in range(len(spans) - 1))      # no two live blocks overlap
    for p in list(live):
        pool.free(p)
    assert sum(pool.bin_occupancy()) == pool.num_regions and pool.stats()["allocated"] == 0
    st = pool.stats()
    assert st["num_alloc"] == st["num_free"] and st["peak_allocated"] <= st["peak_reserved"]


def test_stream_ordered_pool_bookkeeping_without_a_device_and_allocator_selection(monkeypatch):
    """ref: hetu/impl/memory/CUDAStreamOrderedMemoryPool -- on a GPU this is cudaMallocAsync on the default mempool; without one the
    same interface runs on host memory.  HETU_MEMORY_POOL picks the tensor allocator kind."""
    pool = _C.StreamOrderedMemoryPool(0)
    assert pool.kind == "stream_ordered" and (pool.on_device or not torch.cuda.is_available())
    p = pool.alloc(1 << 20, stream=0); q = pool.alloc(4096, stream=0)
    pool.mark_used_by_stream(p, 5)
    assert pool.stats()["allocated"] == (1 << 20) + 4096 and pool.stats()["num_alloc"] == 2
    pool.wait(p); pool.free(p); pool.free(q)
    assert pool.stats()["allocated"] == 0 and pool.stats()["num_free"] == 2 and "stream-ordered" in pool.summary()
    monkeypatch.setenv("HETU_MEMORY_POOL", "bfc")
    assert _C.tensor_allocator("cpu:91").kind == "bfc"
    monkeypatch.setenv("HETU_MEMORY_POOL", "caching")
    assert _C.tensor_allocator("cpu:92").kind == "caching"


@pytest.mark.parametrize("kind", ["caching", "bfc"])
def test_pools_are_thread_safe_under_concurrent_alloc_free(kind):
    """four threads hammer one pool: every live block keeps its own byte pattern (no two live blocks overlap), all memory returns"""
    import ctypes
    import random
    import threading
    pool = _C.MemoryPool("host", limit_mb=512) if kind == "caching" else _C.BFCMemoryPool("host", initial_region_mb=4, limit_mb=512)
    errors = []

    def worker(tid):
        rng = random.Random(tid)
        live = []
        try:
            for step in range(1500):
                if live and (rng.random() < 0.5 or len(live) > 40):
                    ptr, n, tag = live.pop(rng.randrange(len(live)))
                    buf = (ctypes.c_ubyte * n).from_address(ptr)
                    if buf[0] != tag or buf[n - 1] != tag or buf[n // 2] != tag:
                        errors.append(f"thread {tid}: block at {ptr:#x} was overwritten")
                        return
                    pool.free(ptr)
                else:
                    n = rng.choice([64, 700, 4096, 33_000, 260_000])
                    ptr = pool.alloc(n, stream=tid)
                    if ptr == 0:
                        errors.append("allocation failed")
                        return
                    tag = (tid * 37 + step) % 251 + 1
                    ctypes.memset(ptr, tag, n)
                    live.append((ptr, n, tag))
            for ptr, n, tag in live:
                pool.free(ptr)
        except Exception as e:      # noqa: BLE001
            errors.append(repr(e))
    ts = [threading.Thread(target=worker, args=(t,)) for t in range(4)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(120)
    assert not errors, errors[:3]
    st = pool.stats()
    assert st["allocated"] == 0 and st["num_alloc"] == st["num_free"] > 1000
