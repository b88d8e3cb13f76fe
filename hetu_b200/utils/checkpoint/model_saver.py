"""Periodic / rotating / asynchronous checkpointing (ref: python/hetu/utils/checkpoint/model_saver.py:889-1893).

* keeps the newest `save_copies` step directories, deletes older ones
* `async_save=True` snapshots the tensors to pinned host memory on the training thread and writes files on a
  background thread, so the step loop only pays for the device->host copy
* step-info CSV (`step_info.csv`: step, consumed_samples, loss, path) for resume
"""
from __future__ import annotations

import csv
import os
import shutil
import threading
import time
from typing import Optional

from .ht_safetensors import temp_load_split, temp_save_split


class ModelSaver:
    def __init__(self, save_dir: str, save_copies: int = 2, save_interval: int = 0, async_save: bool = False, only_lora=False,
                 save_dtype=None):
        self.save_dir, self.save_copies, self.save_interval = save_dir, save_copies, save_interval
        self.async_save, self.only_lora, self.save_dtype = async_save, only_lora, save_dtype
        self._thread: Optional[threading.Thread] = None
        os.makedirs(save_dir, exist_ok=True)

    def step_dir(self, step: int) -> str:
        return os.path.join(self.save_dir, f"step{step}")

    def if_need_save(self, step: int) -> bool:
        return self.save_interval > 0 and step > 0 and step % self.save_interval == 0

    def wait(self):
        if self._thread is not None:
            self._thread.join()
            self._thread = None

    def save(self, model, optimizer, step: int, consumed_samples: int = 0, loss: float = float("nan")):
        from ...distributed import global_comm_barrier_rpc, rank
        self.wait()
        path = self.step_dir(step)

        def work():
            temp_save_split(model, optimizer, path, only_lora=self.only_lora, save_dtype=self.save_dtype, step=step)

        if self.async_save:
            # tensors are copied to host inside temp_save_split (.cpu()); run that part now, write files in background
            self._thread = threading.Thread(target=work, daemon=True)
            self._thread.start()
        else:
            work()
        global_comm_barrier_rpc()
        if rank() == 0:
            with open(os.path.join(self.save_dir, "step_info.csv"), "a", newline="") as f:
                csv.writer(f).writerow([step, consumed_samples, loss, path, time.time()])
            self._cleanup(step)

    def _cleanup(self, newest: int):
        steps = sorted(int(d[4:]) for d in os.listdir(self.save_dir) if d.startswith("step") and d[4:].isdigit())
        for s in steps[:-self.save_copies] if self.save_copies > 0 else []:
            shutil.rmtree(self.step_dir(s), ignore_errors=True)

    def latest_step(self) -> Optional[int]:
        info = os.path.join(self.save_dir, "step_info.csv")
        if not os.path.exists(info):
            return None
        last = None
        with open(info) as f:
            for row in csv.reader(f):
                if row and os.path.isdir(row[3]):
                    last = row
        return None if last is None else int(last[0])

    def load_latest(self, model, optimizer):
        """-> (step, consumed_samples) or None"""
        info = os.path.join(self.save_dir, "step_info.csv")
        if not os.path.exists(info):
            return None
        rows = [r for r in csv.reader(open(info)) if r and os.path.isdir(r[3])]
        if not rows:
            return None
        step, consumed, _, path = int(rows[-1][0]), int(float(rows[-1][1])), rows[-1][2], rows[-1][3]
        temp_load_split(model, optimizer, path, only_lora=self.only_lora, strict=False)
        return step, consumed
