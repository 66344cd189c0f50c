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


@pytest.mark.parametrize("N,Cin,Cout,S,relu", [
    (300, 3, 8, 32, True), (300, 24, 8, 32, True), (257, 16, 16, 16, False), (999, 32, 32, 8, True),
    (64, 8, 8, 36, True), (65, 16, 16, 18, True), (130, 32, 32, 9, True), (20, 48, 16, 64, True), (40, 128, 128, 8, False),
    # wide layers: channel blocks of the TMA weight-gradient kernel (input x output channel ranges over grid.y), the
    # cp.async conv3x3; 18-px rows: the cp.async weight-gradient path with 8 channels per thread
    (37, 64, 32, 32, True), (21, 96, 64, 16, False), (11, 32, 32, 64, True), (150, 32, 16, 18, True), (70, 8, 16, 18, False),
])
def test_conv3x3_primitive(be, N, Cin, Cout, S, relu):
    sc.check_conv3x3(be, N, Cin, Cout, S, relu)


@pytest.mark.parametrize("task,B,kw", [
    ("spring_color", 3, {}),
    ("spring_color", 100, {}),                                  # BASELINE config 1/2 size
    ("bouncing_balls", 100, {"alpha": 2.0}),
    ("spring_color", 7, {"alt_vel": True, "seed": 2}),
    ("spring_color", 13, {"batch_global": 100}),                # a data-parallel shard: global-batch normalisers
    # 3-body rollouts amplify rounding differences; with the golden fixtures' g = log 8 and B = 100 some sequences
    # have close encounters and even the reference's fp32 and fp64 twins disagree by 7% in every gradient, so the
    # full-batch case uses a weaker coupling (g = -1) where parity is measurable
    ("3bp_color", 2, {"alpha": 5.0, "tol": 1e-3, "traj_tol": 1e-3}),
    ("3bp_color", 100, {"alpha": 5.0, "tol": 1e-3, "traj_tol": 1e-3, "phys": {"g": -1.0}, "seed": 1}),
    ("mnist_spring_color", 2, {}),
    ("mnist_spring_color", 16, {}),
])
def test_whole_step_vs_oracle(be, task, B, kw):
    report = {}
    try:
        sc.check_step(be, task, B, report=report, **kw)
    finally:
        _dump_report("%s_B%d%s" % (task, B, "".join("_%s%s" % (k, v) for k, v in kw.items() if k in ("alt_vel", "batch_global", "seed"))), report)


def _dump_report(name, report):
    """Parity numbers of this run -> gpurun_out/parity_report.json (copied to profiles/ by hand when judged)."""
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    path = os.path.join(root, "gpurun_out", "parity_report.json")
    try:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        data = json.load(open(path)) if os.path.exists(path) else {}
        worst = {}
        for k, v in report.items():
            if isinstance(v, dict) and not k.startswith("fwd/"):
                grp = k.split("/")[0]
                if v["err"] > worst.get(grp, ("", -1.0))[1]:
                    worst[grp] = (k, v["err"], v["err_vs_f64"], v["ref_noise"])
        data[name] = {"forward": {k[4:]: v for k, v in report.items() if k.startswith("fwd/")},
                      "relu_flips": report.get("relu_flips"), "worst_grad": worst}
        json.dump(data, open(path, "w"), indent=1)
    except OSError:
        pass
