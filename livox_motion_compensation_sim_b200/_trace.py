"""NVTX ranges for the host side of the path (tracing row).  Off by default; LMC_NVTX=1 in the environment (or
enable(True)) turns them on: every device operator of ops.py and the phases of run_simulation / save_results /
StreamingAligner.run then show up as named ranges under nsys / ncu --nvtx.  The reference's only instrumentation is a
wall-clock PerformanceMonitor around its scan loop (CS:1538-1601); kernel timing here is CUDA events + ncu (profiles/)."""
from __future__ import annotations

import contextlib
import os

import torch

_enabled = os.environ.get("LMC_NVTX", "0") not in ("", "0")
_depth = 0


def enable(on: bool = True) -> None:
    global _enabled
    _enabled = bool(on)


def enabled() -> bool:
    return _enabled


def depth() -> int:
    """Open ranges (tests)."""
    return _depth


@contextlib.contextmanager
def span(name: str):
    global _depth
    if not _enabled:
        yield
        return
    torch.cuda.nvtx.range_push(name)
    _depth += 1
    try:
        yield
    finally:
        _depth -= 1
        torch.cuda.nvtx.range_pop()


def traced(name: str):
    """Decorator: the whole call is one range."""
    import functools

    def deco(fn):
        @functools.wraps(fn)
        def wrapper(*a, **k):
            if not _enabled:
                return fn(*a, **k)
            with span(name):
                return fn(*a, **k)
        return wrapper
    return deco
