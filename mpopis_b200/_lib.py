"""Loader of the product shared library. There is no fallback of any kind: if the CUDA library has
not been built, cannot be loaded, or no sm_100 GPU is present, using the engine raises."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

from . import _abi

LIB_PATH = Path(__file__).resolve().parent / "libmpopis_b200.so"
_bound = None


class EngineUnavailable(RuntimeError):
    pass


def load_library() -> C.CDLL:
    """dlopen only (no device needed) — used by the CPU test-suite to check exported symbols."""
    if not LIB_PATH.exists():
        raise EngineUnavailable(
            f"{LIB_PATH} is missing: build it with `python -m mpopis_b200.build` "
            "(nvcc, sm_100a). The engine has no CPU fallback.")
    try:
        return C.CDLL(str(LIB_PATH))
    except OSError as e:  # pragma: no cover
        raise EngineUnavailable(f"cannot load {LIB_PATH}: {e}") from e


def product() -> _abi.Bound:
    """The bound C-ABI of libmpopis_b200.so."""
    global _bound
    if _bound is None:
        lib = load_library()
        b = _abi.bind(lib, "mpopis_b200_")
        if b.abi_version() != _abi.ABI_VERSION:
            raise EngineUnavailable("libmpopis_b200.so ABI version mismatch; rebuild")
        _bound = b
    return _bound


def comm_id() -> bytes:
    """ncclUniqueId (128 bytes) for a sharded policy; call on rank 0 and broadcast."""
    buf = C.create_string_buffer(128)
    b = product()
    rc = b.comm_id(buf)
    if rc != 0:
        raise EngineUnavailable(f"mpopis_b200_comm_id failed ({rc}): {b.error()}")
    return buf.raw


class LoopbackGroup:
    """`world` virtual ranks on one device in one process (include/mpopis_b200.h: loopback_create). Every
    handle of the group must be driven from its own host thread — see `sharding.run_virtual_ranks`."""

    def __init__(self, world: int):
        self.b = product()
        self.ptr = C.c_void_p()
        rc = self.b.loopback_create(int(world), C.byref(self.ptr))
        if rc != 0:
            raise EngineUnavailable(f"mpopis_b200_loopback_create failed ({rc}): {self.b.error()}")
        self.world = int(world)

    def close(self):
        if self.ptr:
            self.b.loopback_destroy(self.ptr)
            self.ptr = C.c_void_p()
