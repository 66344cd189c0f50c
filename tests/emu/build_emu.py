"""Build tests/emu/_build/libpaig_emu.so: the package's .cu sources compiled by g++ against the
SIMT-on-CPU shim (emu_cuda.h).  TEST TOOL ONLY -- see emu_cuda.h.  The product never loads it."""
from __future__ import annotations

import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "paig_reproduction_b200", "csrc")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libpaig_emu.so")

FLAGS = ["-O2", "-g", "-std=c++17", "-fPIC", "-DPAIG_EMU", "-ffp-contract=off", "-fno-fast-math", "-I", HERE,
         "-Wno-unknown-pragmas", "-Wno-attributes"]


def build() -> str:
    os.makedirs(OUT, exist_ok=True)
    srcs = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs += [os.path.join(HERE, "emu_cuda.h"), os.path.join(ROOT, "include", "paig_b200.h")]
    newest_hdr = max(os.path.getmtime(h) for h in hdrs)
    jobs, objs = [], []
    for src in srcs + [os.path.join(HERE, "emu_runtime.cpp")]:
        obj = os.path.join(OUT, os.path.basename(src).rsplit(".", 1)[0] + ".o")
        objs.append(obj)
        if not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), newest_hdr):
            jobs.append((src, obj))

    def cc(job):
        src, obj = job
        r = subprocess.run(["g++"] + FLAGS + ["-x", "c++", "-c", src, "-o", obj], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("emu build failed for %s:\n%s" % (src, r.stderr[-6000:]))

    if jobs:
        with ThreadPoolExecutor(max_workers=8) as ex:
            list(ex.map(cc, jobs))
    if jobs or not os.path.exists(LIB):
        r = subprocess.run(["g++", "-shared", "-o", LIB] + objs, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("emu link failed:\n%s" % r.stderr[-4000:])
    return LIB


if __name__ == "__main__":
    print(build())
