"""Pin the CPU oracle against outputs of the unmodified reference (tests/golden, written by
oracle/make_golden.py in the build container).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import physicsnet_oracle as po
from oracle.make_golden import CASES, GRAD_SAMPLE, grad_digest

RTOL = 2e-5      # same ATen kernels on the same image: differences are summation-order only


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"))


def _close(a, b, rtol=RTOL, atol_scale=1.0):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = max(np.abs(b).max(), 1e-30)
    err = np.abs(a - b).max() / scale
    assert err <= rtol * atol_scale, "max-abs-diff / max-abs = %.3e" % err


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_oracle_matches_reference(golden_dir, case):
    name, task, batch, seq_len, seed, alpha, alt_vel, mode = case
    gold = _load(golden_dir, name)
    spec = po.TASKS[task]
    T = seq_len or spec.seq_len
    sd = po.init_state_dict(spec, seed, alt_vel)
    x = po.synthetic_frames(spec, batch, T, seed)
    if mode == "train":
        ff, ls, grads = po.live_step(sd, x, spec, alpha, alt_vel)
    else:
        with torch.no_grad():
            ff = po.feedforward(sd, x, spec, alt_vel)
            ls = po.losses(x, ff, spec, alpha)
        grads = {}
    got = np.array([ls["train"].item(), ls["pred"].item(), ls["extrap"].item(), ls["recons"].item()])
    _close(got, gold["losses"])
    _close(ff["enc_pos"].detach().numpy(), gold["enc_pos"])
    _close(ff["pos_vel_seq"].detach().numpy(), gold["pos_vel_seq"], atol_scale=5.0)
    _close(ff["output"].detach()[:, :, :, ::3, ::3].numpy(), gold["output_sub"])
    _close(ff["recons_out"].detach()[:, :, :, ::3, ::3].numpy(), gold["recons_sub"])
    _close(ff["output"].detach().double().sum((2, 3, 4)).numpy(), gold["output_sum"])
    _close(ff["enc_masks"].detach()[:, :, ::4, ::4].numpy(), gold["enc_masks_sub"])
    _close(ff["template"].detach().numpy(), gold["template"])
    gold_grads = [k[5:] for k in gold.files if k.startswith("grad/")]
    if mode == "train":
        assert sorted(gold_grads) == sorted(grads.keys())       # same set of live parameters (Q1/Q6)
        for k in gold_grads:
            _close(grad_digest(grads[k]), gold["grad/" + k], rtol=2e-4)
    else:
        assert not gold_grads


@pytest.mark.parametrize("cell,n", [("spring", 2), ("bouncing", 2), ("gravity", 3)])
def test_cells_bit_exact(golden_dir, cell, n):
    """The ODE cells restate cells.py op for op, so the fp32 trajectories are bit-identical."""
    gold = _load(golden_dir, "cells")
    pos = torch.from_numpy(gold[cell + "/pos0"])
    vel = torch.from_numpy(gold[cell + "/vel0"])
    spec = {"spring": po.TASKS["spring_color"], "bouncing": po.TASKS["bouncing_balls"],
            "gravity": po.TASKS["3bp_color"]}[cell]
    sd = {"rollout_cell.dt": torch.tensor(0.5 if cell == "gravity" else 0.3),
          "rollout_cell.k": torch.tensor(np.log(1.7)), "rollout_cell.equil": torch.tensor(np.log(2.5)),
          "rollout_cell.g": torch.tensor(np.log(30.0)), "rollout_cell.m": torch.tensor(np.log(1.0))}
    seq = []
    for _ in range(gold[cell + "/seq"].shape[1]):
        pos, vel = po.rollout_cell(sd, spec, pos, vel)
        seq.append(torch.cat([pos, vel], 1))
    got = torch.stack(seq, 1).numpy()
    assert np.array_equal(got, gold[cell + "/seq"])


def test_param_table_matches_reference_key_count():
    # 93 keys for spring / gravity tasks, 91 for bouncing
    assert len(po.param_shapes(po.TASKS["spring_color"])) == 93
    assert len(po.param_shapes(po.TASKS["3bp_color"])) == 93
    assert len(po.param_shapes(po.TASKS["bouncing_balls"])) == 91
    assert GRAD_SAMPLE == 48
