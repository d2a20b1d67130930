"""Device-side round timing: CUDA events recorded on the launching stream, elapsed
time reduced with MAX over ranks — never wall clock.  Usable as the ``profiler`` argument of
``optimizer.train(profiler=...)`` (it implements ``step()``), like the reference's torch.profiler hook
(optimizers/dinno.py:127-128)."""
from __future__ import annotations

from typing import List, Optional

import torch


class RoundTimer:
    def __init__(self, ctx=None, warmup: int = 3):
        self.ctx, self.warmup = ctx, int(warmup)
        self.events: List[torch.cuda.Event] = []
        self.cpu = not torch.cuda.is_available()
        self._t = []

    def step(self):
        if self.cpu:
            import time
            self._t.append(time.perf_counter())
            return
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        self.events.append(ev)

    def ms_per_round(self) -> Optional[float]:
        """Mean device time per round after ``warmup`` rounds, max over ranks."""
        if self.cpu:
            if len(self._t) <= self.warmup + 1:
                return None
            return (self._t[-1] - self._t[self.warmup]) / (len(self._t) - 1 - self.warmup) * 1e3
        if len(self.events) <= self.warmup + 1:
            return None
        torch.cuda.synchronize()
        ms = self.events[self.warmup].elapsed_time(self.events[-1]) / (len(self.events) - 1 - self.warmup)
        if self.ctx is not None and self.ctx.is_distributed:
            t = torch.tensor([ms], dtype=torch.float64, device=self.ctx.device)
            ms = float(self.ctx.all_reduce_max(t).item())
        return ms
