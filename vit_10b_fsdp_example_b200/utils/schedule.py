"""Linear-warmup + cosine learning-rate schedule (reference: utils.py:11-21)."""
from __future__ import annotations

import math


def warmup_cosine_ratio(step: int, warmup_iteration: int, max_iteration: int) -> float:
    if step < warmup_iteration:
        return step / float(warmup_iteration)
    where = (step - warmup_iteration) / float(max_iteration - warmup_iteration)
    return 0.5 * (1.0 + math.cos(math.pi * where))


class WarmupCosineSchedule:
    """LambdaLR-like scheduler over an optimizer that exposes ``param_groups`` with an ``lr`` entry.

    Like ``torch.optim.lr_scheduler.LambdaLR`` the schedule is applied once at construction
    (step 0 -> lr = 0 during warmup) and then after every ``step()``.
    """

    def __init__(self, optimizer, warmup_iteration: int, max_iteration: int):
        self.optimizer = optimizer
        self.warmup_iteration = warmup_iteration
        self.max_iteration = max_iteration
        self.base_lrs = [g.setdefault("initial_lr", g["lr"]) for g in optimizer.param_groups]
        self.last_epoch = 0
        self._apply()

    def _apply(self) -> None:
        ratio = warmup_cosine_ratio(self.last_epoch, self.warmup_iteration, self.max_iteration)
        for g, base in zip(self.optimizer.param_groups, self.base_lrs):
            g["lr"] = base * ratio

    def step(self) -> None:
        self.last_epoch += 1
        self._apply()

    def get_last_lr(self):
        return [g["lr"] for g in self.optimizer.param_groups]

    def state_dict(self):
        return {"last_epoch": self.last_epoch, "base_lrs": list(self.base_lrs),
                "warmup_iteration": self.warmup_iteration, "max_iteration": self.max_iteration}

    def load_state_dict(self, state) -> None:
        self.last_epoch = int(state["last_epoch"])
        self.base_lrs = list(state.get("base_lrs", self.base_lrs))
        self._apply()


def get_warmup_cosine_scheduler(optimizer, warmup_iteration: int, max_iteration: int) -> WarmupCosineSchedule:
    return WarmupCosineSchedule(optimizer, warmup_iteration, max_iteration)
