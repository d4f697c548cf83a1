"""CUDA-event timers on the launching stream (torch's current stream, which every C-ABI call of
this package uses).  Disabled timers cost nothing."""
from __future__ import annotations

import contextlib

import torch


class EventTimers(object):
    def __init__(self, enabled=False):
        self.enabled = enabled
        self._pairs = {}

    @contextlib.contextmanager
    def __call__(self, name):
        if not self.enabled:
            yield
            return
        a = torch.cuda.Event(enable_timing=True)
        b = torch.cuda.Event(enable_timing=True)
        a.record()
        try:
            yield
        finally:
            b.record()
            self._pairs.setdefault(name, []).append((a, b))

    def reset(self):
        self._pairs = {}

    def summary(self):
        """{name: (count, total_ms)}; synchronises."""
        torch.cuda.synchronize()
        return {k: (len(v), sum(a.elapsed_time(b) for a, b in v)) for k, v in self._pairs.items()}
