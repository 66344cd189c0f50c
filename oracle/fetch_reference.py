"""Ship the UNMODIFIED reference to the GPU box: copy /root/reference into oracle/_ref/reference/.

TEST / MEASUREMENT INFRASTRUCTURE.  oracle/_ref/ is git-ignored (no reference source enters the history) but not
gpurun-ignored, so the copy travels with the snapshot like a built .so.  It gives the GPU box
  * the CPU arm of bench.py (``--impl reference`` and ``cpu_baseline``, kind "reference"): the reference's own
    PhysicsNet LIVE step on the box's host cores, through the reference's public API;
  * tests/test_runner_flow.py: the unmodified runners/torch_run_physics.py driving the drop-in;
  * tests comparing the CUDA path with the reference itself (not only with the oracle restatement).
The reference is pure Python: nothing is compiled.  Run in the build container (``__graft_entry__.build()`` does):

    python oracle/fetch_reference.py
"""
from __future__ import annotations

import importlib.machinery
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference"
DST = os.path.join(HERE, "_ref", "reference")


def fetch(force: bool = False) -> str | None:
    """Copy the checkout (python sources, README, LICENSE, requirements) -> oracle/_ref/reference.  Returns the path,
    or None when neither the checkout nor an earlier copy exists (e.g. on the GPU box before a fetch ever ran)."""
    if os.path.isdir(SRC) and (force or not os.path.isdir(DST) or _stale()):
        if os.path.isdir(DST):
            shutil.rmtree(DST)
        shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns(".git", "__pycache__", "*.pyc", "data"))
    return DST if os.path.isdir(DST) else None


def _stale() -> bool:
    for root, _d, files in os.walk(SRC):
        if ".git" in root:
            continue
        for f in files:
            if f.endswith(".py"):
                dst = os.path.join(DST, os.path.relpath(os.path.join(root, f), SRC))
                if not os.path.exists(dst) or open(dst, "rb").read() != open(os.path.join(root, f), "rb").read():
                    return True
    return False


def reference_path() -> str | None:
    """oracle/_ref/reference if it was fetched, else the live checkout, else None."""
    if os.path.isdir(DST):
        return DST
    return SRC if os.path.isdir(SRC) else None


def import_reference():
    """``nn.network.physics_models`` of the unmodified reference (SURVEY 8c recipe: inert stubs for the unused
    tensorflow / matplotlib imports, torchvision imported for real first)."""
    path = reference_path()
    if path is None:
        raise ImportError("reference not available: run oracle/fetch_reference.py in the build container")
    import torch  # noqa: F401
    import torchvision  # noqa: F401

    def stub(name):
        if name in sys.modules:
            return sys.modules[name]
        m = types.ModuleType(name)
        m.__spec__ = importlib.machinery.ModuleSpec(name, None)
        sys.modules[name] = m
        return m

    stub("tensorflow")
    try:
        import matplotlib  # noqa: F401
    except ImportError:
        mpl, cm, plt = stub("matplotlib"), stub("matplotlib.cm"), stub("matplotlib.pyplot")
        mpl.cm, mpl.pyplot = cm, plt
        plt.switch_backend = lambda *a, **k: None
    if path not in sys.path:
        sys.path.insert(0, path)
    sys.dont_write_bytecode = True
    from nn.network import physics_models
    return physics_models


if __name__ == "__main__":
    print(fetch(force="--force" in sys.argv))
