"""ctypes access to tests/emu/_build/libpaig_emu.so (the kernels compiled for the host through the
SIMT shim).  Test-only: checks kernel logic without a GPU.  numpy arrays stand in for device buffers."""
from __future__ import annotations

import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from paig_reproduction_b200 import _abi  # noqa: E402

_lib = None


def lib():
    global _lib
    if _lib is None:
        import build_emu
        _lib = ctypes.CDLL(build_emu.build())
        _abi.declare(_lib)
    return _lib


def ptr(a):
    """Device-pointer stand-in for a numpy array (None -> NULL)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"], "emu buffers must be contiguous"
    return a.ctypes.data


def check(rc):
    if rc != 0:
        raise RuntimeError("paig error %d: %s" % (rc, lib().paig_last_error().decode()))


def f32(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float32))


def f64(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float64))


def make_task(spec, seq_len=None, alpha=3.0, alt_vel=False, batch_global=0):
    return _abi.Task(_abi.CELL_IDS[spec.cell], spec.n_objs, spec.H, seq_len or spec.seq_len, spec.input_steps,
                     spec.pred_steps, int(alt_vel), int(spec.H >= 40), alpha, batch_global)


def make_params(spec, arrays, alt_vel=False):
    """arrays: dict state_dict-name -> numpy array (kept alive by the caller)."""
    unet = "unet" if spec.H >= 40 else "shallow_unet"
    n_convs = 18 if spec.H >= 40 else 13
    p = _abi.Params()
    _abi.fill_params(p, lambda k: arrays[k].ctypes.data, arrays.keys(), unet, n_convs, alt_vel, spec.cell)
    return p


def sd_to_numpy(sd):
    return {k: np.ascontiguousarray(v.detach().numpy()) for k, v in sd.items()}


def workspace(task, B):
    n = lib().paig_workspace_bytes(ctypes.byref(task), B)
    assert n > 0
    return np.zeros(n // 4 + 64, np.float32)
