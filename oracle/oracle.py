"""Python access to the CPU oracle (oracle/mpopis_oracle.c). TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
PARITY UNPINNED (no Julia in this image; the reference has no tests or golden vectors)."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

from mpopis_b200 import _abi
from mpopis_b200.engine import Engine

HERE = Path(__file__).resolve().parent
LIB = HERE / "libmpopis_oracle.so"
_bound = None


def build(force: bool = False) -> Path:
    src = [HERE / "mpopis_oracle.c", HERE / "mpopis_oracle.h", HERE.parent / "include" / "mpopis_b200.h"]
    if force or not LIB.exists() or any(s.stat().st_mtime > LIB.stat().st_mtime for s in src):
        subprocess.run(["make", "-C", str(HERE), "-B" if force else "-s"], check=True, capture_output=True)
    return LIB


def bound() -> _abi.Bound:
    global _bound
    if _bound is None:
        build()
        _bound = _abi.bind(C.CDLL(str(LIB)), "orc_")
    return _bound


def engine(**kwargs) -> Engine:
    """An Engine whose every call lands in the CPU oracle."""
    nthreads = kwargs.pop("nthreads", 1)
    e = Engine(bound(), **kwargs)
    e.b.set_threads(e.h, int(nthreads))
    return e


def sortperm(x):
    import numpy as np
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.zeros(x.size, dtype=np.int64)
    fn = bound().lib.orc_sortperm
    fn.restype, fn.argtypes = C.c_int, [C.POINTER(C.c_double), C.c_int64, C.POINTER(C.c_int64)]
    fn(x.ctypes.data_as(C.POINTER(C.c_double)), x.size, out.ctypes.data_as(C.POINTER(C.c_int64)))
    return out


def philox4x32_10(ctr, key):
    fn = bound().lib.orc_philox4x32_10
    U4, U2 = C.c_uint32 * 4, C.c_uint32 * 2
    fn.restype, fn.argtypes = None, [U4, U2, U4]
    out = U4()
    fn(U4(*ctr), U2(*key), out)
    return list(out)
