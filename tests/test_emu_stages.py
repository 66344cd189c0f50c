"""Kernel logic through the SIMT-on-CPU shim (tests/emu) vs the oracle.  CPU only, tiny sizes.
The same checks run against the real library on a B200 in tests/test_gpu_stages.py."""
import pytest

import backends
import stage_checks as sc

TASKS = ["spring_color", "3bp_color", "mnist_spring_color"]


@pytest.fixture(scope="module")
def be():
    return backends.get("emu")


@pytest.mark.parametrize("cell", list(sc.CELLS))
def test_rollout_forward_vs_reference_golden(be, golden_dir, cell):
    sc.check_rollout_forward_golden(be, golden_dir, cell)


@pytest.mark.parametrize("cell", list(sc.CELLS))
def test_rollout_backward_vs_autograd(be, cell):
    sc.check_rollout_backward(be, cell)


@pytest.mark.parametrize("task", TASKS)
def test_templates(be, task):
    sc.check_templates(be, task)


@pytest.mark.parametrize("task", TASKS)
@pytest.mark.parametrize("mode", ["dframes", "fused_loss"])
def test_decode(be, task, mode):
    sc.check_decode(be, task, mode)
