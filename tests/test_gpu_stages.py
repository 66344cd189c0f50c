"""Parity of each C-ABI stage of libpaig_b200.so on a B200 against the oracle / reference goldens."""
import pytest

import stage_checks as sc

pytestmark = pytest.mark.gpu
TASKS = ["spring_color", "3bp_color", "mnist_spring_color"]


@pytest.fixture(scope="module")
def be():
    import backends
    return backends.get("cuda")


@pytest.mark.parametrize("cell", list(sc.CELLS))
def test_rollout_forward_vs_reference_golden(be, golden_dir, cell):
    sc.check_rollout_forward_golden(be, golden_dir, cell)


@pytest.mark.parametrize("cell", list(sc.CELLS))
def test_rollout_backward_vs_autograd(be, cell):
    sc.check_rollout_backward(be, cell, B=300, steps=9)


@pytest.mark.parametrize("task", TASKS)
def test_templates(be, task):
    sc.check_templates(be, task)


@pytest.mark.parametrize("task", TASKS)
@pytest.mark.parametrize("mode", ["dframes", "fused_loss"])
def test_decode(be, task, mode):
    sc.check_decode(be, task, mode, F_=603)       # more frames than CTAs: exercises the persistent frame loop
