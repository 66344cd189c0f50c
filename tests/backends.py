"""Two ways to run the C-ABI stage functions in tests:
   EmuBackend  -- kernels compiled for the host through tests/emu (numpy buffers; CPU suite, logic only)
   CudaBackend -- the real libpaig_b200.so on cuda:0 (torch buffers; -m gpu suite, the parity tests proper)
Both expose the same tiny interface so tests/stage_checks.py is written once."""
from __future__ import annotations

import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from paig_reproduction_b200 import _abi  # noqa: E402


class Buf:
    def __init__(self, be, obj):
        self.be, self.obj = be, obj

    @property
    def ptr(self):
        return self.be._ptr(self.obj)

    def np(self):
        return self.be._np(self.obj)


class _Base:
    def make_task(self, spec, seq_len=None, alpha=3.0, alt_vel=False, batch_global=0):
        return _abi.Task(_abi.CELL_IDS[spec.cell], spec.n_objs, spec.H, seq_len or spec.seq_len, spec.input_steps,
                         spec.pred_steps, int(alt_vel), int(spec.H >= 40), alpha, batch_global)

    def make_params(self, spec, bufs, alt_vel=False):
        """bufs: dict state_dict-name -> Buf (kept alive by the caller)."""
        unet = "unet" if spec.H >= 40 else "shallow_unet"
        n_convs = 18 if spec.H >= 40 else 13
        p = _abi.Params()
        _abi.fill_params(p, lambda k: bufs[k].ptr, bufs.keys(), unet, n_convs, alt_vel, spec.cell)
        return p

    def sd(self, sd):
        return {k: self.dev(v.detach().numpy()) for k, v in sd.items()}

    def workspace(self, task, B):
        n = self.lib.paig_workspace_bytes(ctypes.byref(task), B)
        assert n > 0, self.lib.paig_last_error()
        return self.zeros(n // 4 + 64)

    def check(self, rc):
        if rc != 0:
            raise RuntimeError("paig error %d: %s" % (rc, self.lib.paig_last_error().decode()))
        self.sync()

    @staticmethod
    def p(buf):
        return None if buf is None else buf.ptr


class EmuBackend(_Base):
    name = "emu"
    stream = None

    def __init__(self):
        sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
        import build_emu
        self.lib = ctypes.CDLL(build_emu.build())
        _abi.declare(self.lib)

    def dev(self, a):
        return Buf(self, np.ascontiguousarray(np.array(a)))

    def zeros(self, shape, dtype=np.float32):
        return Buf(self, np.zeros(shape, dtype))

    def full(self, shape, v, dtype=np.float32):
        return Buf(self, np.full(shape, v, dtype))

    def _ptr(self, a):
        return a.ctypes.data

    def _np(self, a):
        return a

    def sync(self):
        pass


class CudaBackend(_Base):
    name = "cuda"

    def __init__(self):
        import torch
        from paig_reproduction_b200 import _lib
        self.torch = torch
        self.lib = _lib.load()
        self.device = torch.device("cuda:0")
        self.stream = torch.cuda.current_stream(self.device).cuda_stream

    def dev(self, a):
        return Buf(self, self.torch.from_numpy(np.ascontiguousarray(np.array(a))).to(self.device))

    def zeros(self, shape, dtype=np.float32):
        return self.dev(np.zeros(shape, dtype))

    def full(self, shape, v, dtype=np.float32):
        return self.dev(np.full(shape, v, dtype))

    def _ptr(self, t):
        return t.data_ptr()

    def _np(self, t):
        return t.cpu().numpy()

    def sync(self):
        self.torch.cuda.synchronize()


_cache = {}


def get(name):
    if name not in _cache:
        _cache[name] = EmuBackend() if name == "emu" else CudaBackend()
    return _cache[name]
