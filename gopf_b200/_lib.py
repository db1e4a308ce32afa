"""ctypes loader for libgopfcuda.so (the C ABI in include/gopf_cuda.h).

There is no fallback: if the shared library is missing or a call fails, an
exception is raised.  Nothing in this package imports ``oracle``.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libgopfcuda.so")


class GopfError(RuntimeError):
    """Non-zero status from the C ABI (the reference panics on these paths)."""


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GopfError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C gopf_b200/csrc`).  gopf_b200 has no CPU fallback.")
        _lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
        _lib.gopf_last_error.restype = ctypes.c_char_p
        _lib.gopf_abi_version.restype = ctypes.c_int
    return _lib


def check(status: int):
    if status != 0:
        msg = lib().gopf_last_error()
        raise GopfError(msg.decode("utf-8", "replace") if msg else f"libgopfcuda status {status}")


def int_array(values):
    arr = (ctypes.c_int * len(values))(*[int(v) for v in values])
    return arr
