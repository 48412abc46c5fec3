"""Stage timing with CUDA events on the launching stream.

Disabled by default (zero overhead, no synchronisation).  bench.py enables it around its timed
region: every `stage(...)` block records a start/end event pair on torch's current stream --
the stream every kernel of this engine is launched on (native.stream()) -- and `collect()`
synchronises once at the end and returns the per-stage durations.
"""
from __future__ import annotations

from contextlib import contextmanager

import torch

_enabled = False
_records: list = []


def enable(flag: bool = True) -> None:
    global _enabled
    _enabled = flag
    _records.clear()


def enabled() -> bool:
    return _enabled


@contextmanager
def stage(name: str, **meta):
    if not _enabled:
        yield
        return
    start = torch.cuda.Event(enable_timing=True)
    end = torch.cuda.Event(enable_timing=True)
    start.record()
    try:
        yield
    finally:
        end.record()
        _records.append((name, start, end, meta))


def collect(reset: bool = True):
    """[(name, milliseconds, meta)] in program order."""
    torch.cuda.synchronize()
    out = [(n, s.elapsed_time(e), m) for (n, s, e, m) in _records]
    if reset:
        _records.clear()
    return out


def summarize(rows):
    """name -> {"ms": total, "calls": n, **summed numeric meta}"""
    acc: dict = {}
    for name, ms, meta in rows:
        d = acc.setdefault(name, {"ms": 0.0, "calls": 0})
        d["ms"] += ms
        d["calls"] += 1
        for k, v in meta.items():
            if isinstance(v, (int, float)):
                d[k] = d.get(k, 0) + v
    return acc
