"""Periodic / rotating / asynchronous checkpointing (ref: python/hetu/utils/checkpoint/model_saver.py:889-1893).

* keeps the newest `save_copies` step directories, deletes older ones
* `async_save=True` snapshots the tensors on the training thread (a private host copy, so later optimizer steps cannot
  tear the checkpoint) and writes the files in the background, so the step loop only pays for the device->host copy.
  `async_mode="thread"` writes on a Python thread; `async_mode="process"` copies the snapshot into the POSIX shared-memory
  pool (`_C.get_memory_pool("shm")`) and forks a writer process, as the reference does, so file serialisation never
  competes with the training process for the GIL
* step-info CSV (`step_info.csv`: step, consumed_samples, loss, path) for resume
"""
from __future__ import annotations

import csv
import os
import shutil
import threading
import time
from typing import Optional

from .ht_safetensors import collect_split_state, temp_load_split, write_split_state


class ModelSaver:
    def __init__(self, save_dir: str, save_copies: int = 2, save_interval: int = 0, async_save: bool = False, only_lora=False,
                 save_dtype=None, async_mode: str = "thread", remote: Optional[str] = None):
        """remote: URI of a mirror (`hdfs://namenode/ckpt`, `hdfs-cli://ckpt`, `file:///shared/ckpt`, ...): every published step is
        uploaded there (each rank its own files, the bookkeeping rank the markers), rotation is applied remotely too, and
        `load_latest` on a node whose local directory is empty restores from the mirror"""
        assert async_mode in ("thread", "process")
        self.remote_uri = remote
        self._remote = None
        self.save_dir, self.save_copies, self.save_interval = save_dir, save_copies, save_interval
        self.async_save, self.only_lora, self.save_dtype, self.async_mode = async_save, only_lora, save_dtype, async_mode
        self._queue = None        # native TaskQueue (one writer thread) of the `thread` mode, created on first use
        self._child: Optional[int] = None
        self._shm_blocks = []
        self.last_write_error: Optional[str] = None
        self._pending = None      # (step, consumed, loss, path, writer_rank) of a checkpoint whose files are still being written
        os.makedirs(save_dir, exist_ok=True)

    def step_dir(self, step: int) -> str:
        return os.path.join(self.save_dir, f"step{step}")

    def if_need_save(self, step: int) -> bool:
        return self.save_interval > 0 and step > 0 and step % self.save_interval == 0

    def wait(self):
        if self._queue is not None:
            self._queue.wait()
        if self._child is not None:
            _, status = os.waitpid(self._child, 0)
            self._child = None
            if status != 0:
                self.last_write_error = f"checkpoint writer process exited with status {status}"
        if self._shm_blocks:
            from ... import _C
            pool = _C.get_memory_pool("shm")
            for ptr in self._shm_blocks:
                pool.free(ptr)
            self._shm_blocks = []
        err, self.last_write_error = self.last_write_error, None
        if self._pending is not None:
            # every rank's writer has finished here: only now does the checkpoint become visible (COMPLETE marker, CSV row)
            # and only now may older copies be rotated out -- a crash in the middle of a write leaves the previous
            # checkpoint as the newest published one
            self._publish(*self._pending, ok=err is None)
            self._pending = None
        if err is not None:
            raise RuntimeError(err)

    def _publish(self, step, consumed_samples, loss, path, writer_rank, ok=True):
        """collective: all ranks agree that their files are on disk, then `writer_rank` marks the step complete"""
        from ...distributed import all_ranks_ok, global_comm_barrier_rpc, rank
        ok = all_ranks_ok(ok)
        global_comm_barrier_rpc()
        if ok and self.remote_uri:
            ok = all_ranks_ok(self._mirror_own_files(step, path))        # every rank uploads what it wrote
        if ok and rank() == writer_rank:
            with open(os.path.join(path, "COMPLETE"), "w") as f:
                f.write(f"{step}\n")
            with open(os.path.join(self.save_dir, "step_info.csv"), "a", newline="") as f:
                csv.writer(f).writerow([step, consumed_samples, loss, path, time.time()])
            self._cleanup(step)
            if self.remote_uri:
                self._mirror_markers(step, path)
        global_comm_barrier_rpc()

    # ---- remote mirror ------------------------------------------------------------------------------------------------
    def remote(self):
        if self._remote is None and self.remote_uri:
            from .remote_fs import open_remote
            self._remote = open_remote(self.remote_uri)
        return self._remote

    def _mirror_own_files(self, step: int, path: str) -> bool:
        """upload the files of this step directory that are not on the mirror yet (ranks sharing a filesystem skip each other's)"""
        try:
            fs = self.remote()
            have = set(fs.listdir(f"step{step}"))
            for f in sorted(os.listdir(path)):
                if f != "COMPLETE" and f not in have:
                    fs.put_file(os.path.join(path, f), f"step{step}/{f}")
            return True
        except Exception as e:      # noqa: BLE001 -- surfaced as "checkpoint not published"
            self.last_write_error = f"remote mirror: {type(e).__name__}: {e}"
            return False

    def _mirror_markers(self, step: int, path: str):
        fs = self.remote()
        fs.put_file(os.path.join(path, "COMPLETE"), f"step{step}/COMPLETE")
        fs.put_file(os.path.join(self.save_dir, "step_info.csv"), "step_info.csv")
        steps = sorted(int(d[4:]) for d in fs.listdir("") if d.startswith("step") and d[4:].isdigit())
        done = [s for s in steps if fs.exists(f"step{s}/COMPLETE")]
        keep = set(done[-self.save_copies:]) if self.save_copies > 0 else set(done)
        for s in steps:
            if s not in keep and s < step:
                fs.remove_dir(f"step{s}")

    def restore_from_remote(self) -> Optional[int]:
        """download the newest complete step (and the step-info file) from the mirror into save_dir -> its step number"""
        fs = self.remote()
        if fs is None:
            return None
        steps = sorted(int(d[4:]) for d in fs.listdir("") if d.startswith("step") and d[4:].isdigit())
        done = [s for s in steps if fs.exists(f"step{s}/COMPLETE")]
        if not done:
            return None
        s = done[-1]
        fs.get_dir(f"step{s}", self.step_dir(s))
        info = os.path.join(self.save_dir, "step_info.csv")
        if fs.exists("step_info.csv"):
            fs.get_file("step_info.csv", info)
            # rows carry the path of the node that wrote them: point them at this node's directory
            rows = [r for r in csv.reader(open(info)) if r]
            with open(info, "w", newline="") as f:
                for r in rows:
                    r[3] = self.step_dir(int(r[0]))
                    csv.writer(f).writerow(r)
        return s

    def _snapshot_fn(self):
        """host copy owned by the saver: plain clone for the thread writer, shared-memory pool blocks for the process writer"""
        if self.async_mode == "thread":
            return lambda t: t.detach().to("cpu", copy=True)
        from ... import _C
        pool = _C.get_memory_pool("shm")
        names = {"torch.float32": "float32", "torch.bfloat16": "bfloat16", "torch.float16": "float16", "torch.int64": "int64",
                 "torch.int32": "int32", "torch.float64": "float64", "torch.bool": "bool", "torch.uint8": "uint8", "torch.int8": "int8"}

        def snap(t):
            t = t.detach()
            ptr = pool.alloc(max(t.numel() * t.element_size(), 1))
            self._shm_blocks.append(ptr)
            view = pool.as_tensor(ptr, list(t.shape), names[str(t.dtype)])
            view.copy_(t)
            return view
        return snap

    def save(self, model, optimizer, step: int, consumed_samples: int = 0, loss: float = float("nan"), writer_rank: int = 0):
        from ...distributed import global_comm_barrier_rpc, rank
        self.wait()
        path = self.step_dir(step)

        state = collect_split_state(model, optimizer, save_dtype=self.save_dtype, only_lora=self.only_lora, step=step,
                                    snapshot=self._snapshot_fn() if self.async_save else None)

        def work():
            try:
                write_split_state(path, state)
            except Exception as e:   # noqa: BLE001 -- surfaced by the next wait()
                self.last_write_error = f"{type(e).__name__}: {e}"

        if not self.async_save:
            write_split_state(path, state)
        elif self.async_mode == "thread":
            if self._queue is None:
                from ... import _C
                self._queue = _C.TaskQueue("checkpoint-writer", 1, 2)
            self._queue.add(work)
        else:
            pid = os.fork()
            if pid == 0:                 # writer process: only reads the shared snapshot and writes files
                code = 0
                try:
                    write_split_state(path, state)
                except BaseException:    # noqa: BLE001
                    code = 1
                os._exit(code)
            self._child = pid
        # `writer_rank` keeps the step-info CSV (an active rank when some ranks are idle)
        if not self.async_save:
            self._publish(step, consumed_samples, loss, path, writer_rank)
        else:
            self._pending = (step, consumed_samples, loss, path, writer_rank)     # published by the next wait() / save()

    def _cleanup(self, newest: int):
        """keep the newest `save_copies` COMPLETE checkpoints; torn directories (no marker) older than the newest go too"""
        steps = sorted(int(d[4:]) for d in os.listdir(self.save_dir) if d.startswith("step") and d[4:].isdigit())
        done = [s for s in steps if self._complete(self.step_dir(s))]
        keep = set(done[-self.save_copies:]) if self.save_copies > 0 else set(done)
        for s in steps:
            if s not in keep and s < newest:
                shutil.rmtree(self.step_dir(s), ignore_errors=True)

    @staticmethod
    def _complete(path: str) -> bool:
        return os.path.isdir(path) and os.path.exists(os.path.join(path, "COMPLETE"))

    def latest_step(self) -> Optional[int]:
        info = os.path.join(self.save_dir, "step_info.csv")
        if not os.path.exists(info):
            return None
        last = None
        with open(info) as f:
            for row in csv.reader(f):
                if row and self._complete(row[3]):
                    last = row
        return None if last is None else int(last[0])

    def load_latest(self, model, optimizer):
        """-> (step, consumed_samples) or None"""
        info = os.path.join(self.save_dir, "step_info.csv")
        if self.remote_uri and (not os.path.exists(info) or not any(self._complete(r[3]) for r in csv.reader(open(info)) if r)):
            self.restore_from_remote()                  # a fresh node: pull the newest complete step from the mirror
        if not os.path.exists(info):
            return None
        rows = [r for r in csv.reader(open(info)) if r and self._complete(r[3])]
        # newest complete checkpoint first; a copy that cannot be read (torn file) falls back to the one before it
        for row in reversed(rows):
            step, consumed, path = int(row[0]), int(float(row[1])), row[3]
            try:
                temp_load_split(model, optimizer, path, only_lora=self.only_lora, strict=False)
            except Exception as e:     # noqa: BLE001
                import warnings
                warnings.warn(f"checkpoint {path} is unreadable ({type(e).__name__}: {e}); trying the previous one")
                continue
            if hasattr(optimizer, "set_step"):
                optimizer.set_step(step)
            return step, consumed
        return None
