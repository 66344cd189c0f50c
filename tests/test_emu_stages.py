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


@pytest.mark.parametrize("task,B,kw", [
    ("spring_color", 2, {}),
    ("bouncing_balls", 1, {"alpha": 2.0}),
    ("spring_color", 1, {"alt_vel": True, "seed": 2}),
    # 3-body gravity amplifies rounding differences (the oracle's own fp32-vs-fp64 twin differs by 3.3e-5 in the
    # velocity-MLP gradients on this case, and ATen's fp32 sqrt is not correctly rounded): looser bound
    ("3bp_color", 1, {"alpha": 5.0, "tol": 1e-3, "traj_tol": 1e-3}),
])
def test_whole_step_vs_oracle(be, task, B, kw):
    sc.check_step(be, task, B, **kw)


@pytest.mark.parametrize("N,Cin,Cout,S,relu", [
    (3, 3, 8, 32, True), (2, 24, 8, 32, True), (5, 16, 16, 16, False), (19, 32, 32, 8, True),
    (2, 8, 8, 36, True), (3, 16, 16, 18, True), (11, 32, 32, 9, True), (1, 48, 16, 64, True), (2, 128, 32, 8, False),
    (2, 64, 32, 32, True), (2, 128, 128, 8, False),      # channel blocks of the TMA weight-gradient kernel
])
def test_conv3x3_primitive(be, N, Cin, Cout, S, relu):
    sc.check_conv3x3(be, N, Cin, Cout, S, relu)
