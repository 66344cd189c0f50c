"""ctypes access to tests/emu/_build/libpaig_emu.so (the kernels compiled for the host through the
SIMT shim).  Test-only: checks kernel logic without a GPU.  numpy arrays stand in for device buffers."""
from __future__ import annotations

import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))

_lib = None


def lib():
    global _lib
    if _lib is None:
        import build_emu
        _lib = ctypes.CDLL(build_emu.build())
        _lib.paig_last_error.restype = ctypes.c_char_p
        if hasattr(_lib, "paig_workspace_bytes"):
            _lib.paig_workspace_bytes.restype = ctypes.c_size_t
    return _lib


def ptr(a):
    """Device-pointer stand-in for a numpy array (None -> NULL)."""
    if a is None:
        return ctypes.c_void_p(0)
    assert a.flags["C_CONTIGUOUS"], "emu buffers must be contiguous"
    return ctypes.c_void_p(a.ctypes.data)


def check(rc):
    if rc != 0:
        raise RuntimeError("paig error %d: %s" % (rc, lib().paig_last_error().decode()))


def f32(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float32))


def f64(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float64))
