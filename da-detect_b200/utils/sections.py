"""Opt-in wall-clock section timer (synchronising) used by tools/step_profile.py to see where a training
step spends its time; a no-op unless `enable()` was called."""
import contextlib
import time

import torch

_enabled = False
totals = {}


def enable(flag=True):
    global _enabled
    _enabled = flag
    totals.clear()


@contextlib.contextmanager
def section(name):
    if not _enabled:
        yield
        return
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    yield
    torch.cuda.synchronize()
    totals[name] = totals.get(name, 0.0) + (time.perf_counter() - t0)
