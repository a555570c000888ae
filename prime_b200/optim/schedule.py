"""Inner learning-rate schedules (cosine / linear / WSD-sqrt / constant), all with linear warm-up."""

from __future__ import annotations

import math


def lr_at(step: int, *, base_lr: float, sched_type: str, warmup_steps: int, total_steps: int, stable_steps: int = 0,
          min_ratio: float = 0.1) -> float:  # fmt: skip
    if warmup_steps > 0 and step < warmup_steps:
        return base_lr * (step + 1) / warmup_steps
    if sched_type == "constant":
        return base_lr
    if sched_type == "wsd-sqrt":
        if step < warmup_steps + stable_steps:
            return base_lr
        decay = max(1, total_steps - warmup_steps - stable_steps)
        p = min(1.0, (step - warmup_steps - stable_steps) / decay)
        return base_lr * max(min_ratio, 1.0 - math.sqrt(p))
    span = max(1, total_steps - warmup_steps)
    p = min(1.0, (step - warmup_steps) / span)
    if sched_type == "linear":
        return base_lr * max(min_ratio, 1.0 - p)
    return base_lr * (min_ratio + (1 - min_ratio) * 0.5 * (1.0 + math.cos(math.pi * p)))
