"""Optional per-kernel CUDA-event timing (used by bench.py for the roofline numbers).

Disabled by default: `region()` is then a no-op.  When enabled, every instrumented launch is
bracketed by two CUDA events recorded on the launching stream, and `summary()` reports, per
kernel family, launches, total device time and the algorithmic bytes the caller declared.
"""
import contextlib

import torch

_enabled = False
_records = []


def enable(flag=True):
    global _enabled
    _enabled = bool(flag)
    del _records[:]


def enabled():
    return _enabled


@contextlib.contextmanager
def region(name, nbytes=0, launches=1):
    """launches=None: counted from the library's own launch counter (regions whose number of kernels
    depends on the path taken, e.g. one-pass or per-axis transfers)."""
    if not _enabled:
        yield
        return
    a = torch.cuda.Event(enable_timing=True)
    b = torch.cuda.Event(enable_timing=True)
    n0 = 0
    if launches is None:
        from . import _lib
        n0 = _lib.lib().poms_launch_count()
    a.record()
    yield
    b.record()
    if launches is None:
        launches = _lib.lib().poms_launch_count() - n0
    _records.append((name, a, b, int(nbytes), int(launches)))


def summary():
    """{name: {"launches", "ms", "bytes", "gbs"}} -- call after torch.cuda.synchronize()."""
    out = {}
    for name, a, b, nbytes, launches in _records:
        d = out.setdefault(name, {"launches": 0, "ms": 0.0, "bytes": 0})
        d["launches"] += launches
        d["ms"] += a.elapsed_time(b)
        d["bytes"] += nbytes
    for d in out.values():
        d["gbs"] = d["bytes"] / (d["ms"] * 1e-3) / 1e9 if d["ms"] > 0 else 0.0
    return out


def summary_largest(frac=0.5):
    """Like summary(), restricted per kernel family to its LARGEST launches (declared bytes >= frac x
    the family's maximum): the fine-level launches of a multigrid run, whose rate is the kernel's
    roofline figure, without the latency-bound coarse-level launches of the same family."""
    top = {}
    for name, a, b, nbytes, launches in _records:
        top[name] = max(top.get(name, 0), nbytes)
    out = {}
    for name, a, b, nbytes, launches in _records:
        if nbytes <= 0 or nbytes < frac * top[name]:
            continue
        d = out.setdefault(name, {"launches": 0, "ms": 0.0, "bytes": 0})
        d["launches"] += launches
        d["ms"] += a.elapsed_time(b)
        d["bytes"] += nbytes
    for d in out.values():
        d["gbs"] = d["bytes"] / (d["ms"] * 1e-3) / 1e9 if d["ms"] > 0 else 0.0
    return out


def reset():
    del _records[:]
