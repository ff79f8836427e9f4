"""Learning-rate schedules of the reference (utils/schedulers.py:11-65) for torch ≥ 2.x: the reference passes
`verbose=True` to LambdaLR (:50-52), which current torch rejects, so `dropin.install()` puts these in its place.
The same rule evaluated by the C ABI: `gb_warmup_cosine_lr` (tests/test_cabi.py pins the two together)."""
from __future__ import annotations

import math

from torch.optim import lr_scheduler
from torch.optim.lr_scheduler import LambdaLR


class WarmupCosineSchedule(LambdaLR):
    """Linear warm-up over `warmup_steps`, then cosine decay to 0 at `t_total` (:36-65)."""

    def __init__(self, optimizer, warmup_steps, t_total, cycles=0.5, last_epoch=-1):
        self.warmup_steps, self.t_total, self.cycles = warmup_steps, t_total, cycles
        super().__init__(optimizer, self.lr_lambda, last_epoch=last_epoch)

    def lr_lambda(self, step):
        if step < self.warmup_steps:
            return float(step) / float(max(1.0, self.warmup_steps))
        progress = float(step - self.warmup_steps) / float(max(1, self.t_total - self.warmup_steps))
        return max(0.0, 0.5 * (1.0 + math.cos(math.pi * float(self.cycles) * 2.0 * progress)))


def make_scheduler(optimizer, config, double=False, teacher=False):
    """:11-33 — SCHEDULER: cosine | one_warmup_epoch | step."""
    total = (config.t_EPOCHS if teacher else config.s_EPOCHS) if double else config.EPOCHS
    if config.SCHEDULER == "cosine":
        return WarmupCosineSchedule(optimizer, warmup_steps=config.WARMUP_EPOCHS, t_total=total)
    if config.SCHEDULER == "one_warmup_epoch":
        return LambdaLR(optimizer, lr_lambda=lambda epoch: config.WARMUP_LR / config.LR if epoch == 0 else 1)
    return lr_scheduler.StepLR(optimizer, step_size=config.STEP_SIZE, gamma=0.1)
