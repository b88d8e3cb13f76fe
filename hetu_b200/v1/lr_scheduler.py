"""v1 learning-rate schedulers (ref: hetu/v1/python/hetu/lr_scheduler.py: Fixed / Step / MultiStep / Exponential /
ReduceOnPlateau).  `get()` returns the current rate, `step()` advances; optimizers accept a scheduler as learning_rate."""
from __future__ import annotations

import bisect


class FixedScheduler:
    def __init__(self, learning_rate):
        assert learning_rate >= 0, "learning rate must be non-negative"
        self.lr = float(learning_rate)
        self.cnt = 0

    def get(self):
        return self.lr

    def step(self, *a, **k):
        self.cnt += 1
        return self.get()


class StepScheduler(FixedScheduler):
    """lr * gamma^(floor(step / step_size))"""

    def __init__(self, learning_rate, step_size, gamma=0.1):
        super().__init__(learning_rate)
        assert step_size > 0 and 0 < gamma <= 1
        self.step_size, self.gamma = int(step_size), float(gamma)

    def get(self):
        return self.lr * self.gamma ** (self.cnt // self.step_size)


class MultiStepScheduler(FixedScheduler):
    """lr * gamma^(number of milestones passed)"""

    def __init__(self, learning_rate, milestones, gamma=0.1):
        super().__init__(learning_rate)
        self.milestones, self.gamma = sorted(int(m) for m in milestones), float(gamma)

    def get(self):
        return self.lr * self.gamma ** bisect.bisect_right(self.milestones, self.cnt)


class ExponentialScheduler(FixedScheduler):
    def __init__(self, learning_rate, gamma=0.99):
        super().__init__(learning_rate)
        self.gamma = float(gamma)

    def get(self):
        return self.lr * self.gamma ** self.cnt


class ReduceOnPlateauScheduler(FixedScheduler):
    """multiply by `factor` after `patience` steps without improvement of the monitored value"""

    def __init__(self, learning_rate, mode="min", factor=0.1, patience=10, threshold=1e-4, threshold_mode="rel", cooldown=0, min_lr=0.0, eps=1e-8):
        super().__init__(learning_rate)
        assert mode in ("min", "max") and threshold_mode in ("rel", "abs") and 0 < factor < 1
        self.mode, self.factor, self.patience, self.threshold, self.threshold_mode = mode, factor, patience, threshold, threshold_mode
        self.cooldown, self.min_lr, self.eps = cooldown, min_lr, eps
        self.best = float("inf") if mode == "min" else -float("inf")
        self.bad, self.cool = 0, 0

    def _better(self, v):
        if self.mode == "min":
            return v < (self.best * (1 - self.threshold) if self.threshold_mode == "rel" else self.best - self.threshold)
        return v > (self.best * (1 + self.threshold) if self.threshold_mode == "rel" else self.best + self.threshold)

    def step(self, value=None):
        self.cnt += 1
        if value is None:
            return self.lr
        if self._better(value):
            self.best, self.bad = value, 0
        else:
            self.bad += 1
        if self.cool > 0:
            self.cool -= 1
            self.bad = 0
        if self.bad > self.patience:
            new = max(self.lr * self.factor, self.min_lr)
            if self.lr - new > self.eps:
                self.lr = new
            self.cool, self.bad = self.cooldown, 0
        return self.lr
