"""Running averages and device-side timing helpers."""
from __future__ import annotations

import contextlib
import time
from typing import Dict, List, Optional

import torch


class AverageMeter:
    """Running mean with the reference's printing format (``utils.py:120-141``)."""

    def __init__(self, name: str, fmt: str = ":f"):
        self.name, self.fmt = name, fmt
        self.reset()

    def reset(self):
        self.val = self.avg = self.sum = 0.0
        self.count = 0

    def update(self, val, n: int = 1):
        val = float(val)
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / max(self.count, 1)

    def __str__(self):
        return ("{name} {val" + self.fmt + "} ({avg" + self.fmt + "})").format(**self.__dict__)


class DeviceTimer:
    """CUDA-event timer (falls back to perf_counter on CPU).  Usage::

        t = DeviceTimer(); t.start(); ...; ms = t.stop()
    """

    def __init__(self, device: Optional[torch.device] = None):
        self.cuda = torch.cuda.is_available() and (device is None or torch.device(device).type == "cuda")
        if self.cuda:
            self._a = torch.cuda.Event(enable_timing=True)
            self._b = torch.cuda.Event(enable_timing=True)

    def start(self):
        if self.cuda:
            self._a.record()
        else:
            self._t0 = time.perf_counter()

    def stop(self) -> float:
        if self.cuda:
            self._b.record()
            self._b.synchronize()
            return self._a.elapsed_time(self._b)
        return (time.perf_counter() - self._t0) * 1e3


@contextlib.contextmanager
def nvtx_range(name: str):
    """NVTX range when CUDA is present (tracing subsystem the reference lacks)."""
    on = torch.cuda.is_available()
    if on:
        torch.cuda.nvtx.range_push(name)
    try:
        yield
    finally:
        if on:
            torch.cuda.nvtx.range_pop()


class PhaseProfiler:
    """Per-phase device time accumulation for one training step (enabled on demand)."""

    def __init__(self, enabled: bool = False):
        self.enabled = enabled
        self.records: Dict[str, List[float]] = {}
        self._open = []

    @contextlib.contextmanager
    def phase(self, name: str):
        if not self.enabled:
            with nvtx_range(name):
                yield
            return
        t = DeviceTimer()
        t.start()
        with nvtx_range(name):
            yield
        self.records.setdefault(name, []).append(t.stop())

    def summary(self) -> Dict[str, float]:
        return {k: sum(v) / len(v) for k, v in self.records.items() if v}
