"""Loader for libpaig_b200.so (the hand-written sm_100a kernels behind include/paig_b200.h).

There is no fallback of any kind: if the library is missing, has the wrong ABI or lacks a symbol, importing
code gets an exception.  The library is built in-tree by ``python -m paig_reproduction_b200.build``
(``__graft_entry__.build()`` does that)."""
from __future__ import annotations

import ctypes
import os

from . import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpaig_b200.so")

_lib = None


class PaigError(RuntimeError):
    pass


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PaigError("%s not found: build it with `python -m paig_reproduction_b200.build` "
                        "(there is no CPU or PyTorch fallback for this path)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    missing = _abi.declare(lib)
    if missing:
        raise PaigError("libpaig_b200.so lacks symbols declared in include/paig_b200.h: %s" % ", ".join(missing))
    if lib.paig_abi_version() != _abi.ABI_VERSION:
        raise PaigError("libpaig_b200.so ABI %d != expected %d" % (lib.paig_abi_version(), _abi.ABI_VERSION))
    if os.environ.get("PAIG_TRACE_LIB"):
        import sys
        sys.stderr.write("paig: loaded %s (ABI %d)\n" % (LIB_PATH, lib.paig_abi_version()))
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        raise PaigError("%s failed (%d): %s" % (what or "paig call", rc, load().paig_last_error().decode()))
