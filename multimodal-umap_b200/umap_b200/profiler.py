"""Stage timing with CUDA events on the launching stream.

Disabled by default (zero overhead, no synchronisation).  bench.py enables it around its timed
region: every `stage(...)` block records a start/end event pair on torch's current stream --
the stream every kernel of this engine is launched on (native.stream()) -- and `collect()`
synchronises once at the end and returns the per-stage durations.
"""
from __future__ import annotations

from contextlib import contextmanager

import torch

_level = 0
_records: list = []


def enable(level: int | bool = 1) -> None:
    """0/False: off.  1: coarse stages + the dominant kernel (cheap enough for a timed region).
    2: additionally every small per-epoch kernel (a few microseconds of event overhead per launch)."""
    global _level
    _level = int(level)
    _records.clear()


def enabled(level: int = 1) -> bool:
    return _level >= level


@contextmanager
def stage(name: str, level: int = 1, **meta):
    if _level < level:
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
