"""Warm-up + polynomial-decay learning-rate schedule of the DLRM example
(reference examples/dlrm/utils.py:45-88)."""
from __future__ import annotations


class LearningRateScheduler:

  def __init__(self, base_lr: float, warmup_steps: int, decay_start_step: int, decay_steps: int,
               poly_power: int = 2):
    self.base_lr = float(base_lr)
    self.warmup_steps = int(warmup_steps)
    self.decay_start_step = int(decay_start_step)
    self.decay_steps = int(decay_steps)
    self.decay_end_step = self.decay_start_step + self.decay_steps
    self.poly_power = poly_power
    self.step_count = 0

  def lr_at(self, step: int) -> float:
    if step < self.warmup_steps:
      factor = 1.0 - (self.warmup_steps - step) / self.warmup_steps
    elif step < self.decay_start_step:
      factor = 1.0
    else:
      factor = max(0.0, (self.decay_end_step - step) / self.decay_steps)**self.poly_power
    return self.base_lr * factor

  def step(self) -> float:
    lr = self.lr_at(self.step_count)
    self.step_count += 1
    return lr
