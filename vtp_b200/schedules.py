"""Training schedules: the reference ships one helper, `CosineScheduler`
(vtp/models/utils/text_utils.py:160-207) — a precomputed table  [freeze zeros | linear warm-up | cosine decay]  indexed by
the iteration, `final_value` past its end.  Restated here on numpy with the same arithmetic (float64 `linspace` / `cos`),
and uploaded to the device as fp32 tables: the fused optimiser looks lr / weight decay / teacher momentum up by its own
device-side step counter (`vtp_hyper_tick`), so a captured CUDA graph of the training step follows the schedule without
any host scalar."""
from __future__ import annotations

from typing import Sequence

import numpy as np


class CosineSchedule:
    """value(it) for it in [0, total_iters): 0 for `freeze_iters`, then linspace(start_warmup_value, base_value,
    warmup_iters), then final + 0.5 (base - final)(1 + cos(pi i / n)) for i in range(n), n = the remaining iterations;
    `final_value` for it >= total_iters  (text_utils.py:168-207)."""

    def __init__(self, base_value: float, final_value: float, total_iters: int, warmup_iters: int = 0,
                 start_warmup_value: float = 0.0, freeze_iters: int = 0):
        if warmup_iters + freeze_iters > total_iters:
            raise ValueError("warmup_iters + freeze_iters exceed total_iters")
        self.final_value, self.total_iters = float(final_value), int(total_iters)
        n = total_iters - warmup_iters - freeze_iters
        i = np.arange(n)
        cos = final_value + 0.5 * (base_value - final_value) * (1 + np.cos(np.pi * i / max(n, 1)))
        self.schedule = np.concatenate((np.zeros(freeze_iters), np.linspace(start_warmup_value, base_value, warmup_iters), cos))
        assert len(self.schedule) == self.total_iters

    def __getitem__(self, it: int) -> float:
        return self.final_value if it >= self.total_iters else float(self.schedule[it])

    def __len__(self) -> int:
        return self.total_iters

    def table(self) -> np.ndarray:
        """fp32 device table: every iteration's value followed by final_value (the clamp target of the kernel)."""
        return np.concatenate((self.schedule, [self.final_value])).astype(np.float32)


def as_table(s) -> np.ndarray:
    if isinstance(s, CosineSchedule):
        return s.table()
    a = np.asarray(list(s) if not isinstance(s, np.ndarray) else s, dtype=np.float32)
    if a.ndim != 1 or a.size == 0:
        raise ValueError("a schedule must be a non-empty 1-D sequence of values")
    return a


def pad_tables(tabs: Sequence[np.ndarray]) -> Sequence[np.ndarray]:
    """Tables of different lengths are extended with their last value to a common length (one n_tab in the kernel)."""
    n = max(t.size for t in tabs)
    return [np.concatenate((t, np.full(n - t.size, t[-1], dtype=np.float32))) for t in tabs]
