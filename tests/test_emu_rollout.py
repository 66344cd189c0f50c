"""Rollout kernels (rollout.cu) run through the SIMT-on-CPU shim vs the reference goldens / oracle.
CPU only; the same checks run on the real library under -m gpu (tests/test_gpu_stages.py)."""
import ctypes
import os

import numpy as np
import pytest
import torch

from oracle import physicsnet_oracle as po
import emu_util as eu

CELLS = {"spring": (0, 2, 0.3), "bouncing": (1, 2, 0.3), "gravity": (2, 3, 0.5)}
PHYS = {"spring": (np.log(1.7), np.log(2.5)), "bouncing": (0.0, 0.0), "gravity": (np.log(30.0), np.log(1.0))}


def _run_fwd(cell, pos0, vel0, steps):
    cid, n, dt = CELLS[cell]
    B = pos0.shape[0]
    seq = np.zeros((B, steps + 1, 4 * n), np.float32)
    seq[:, 0, :2 * n] = pos0
    seq[:, 0, 2 * n:] = vel0
    dt_a, p0, p1 = eu.f32([dt]), eu.f64([PHYS[cell][0]]), eu.f64([PHYS[cell][1]])
    eu.check(eu.lib().paig_rollout_forward(cid, n, B, steps, eu.ptr(dt_a), eu.ptr(p0), eu.ptr(p1), eu.ptr(seq), None))
    return seq


def _same_trajectory(cell, got, ref):
    if cell == "gravity":
        # ATen's AVX512 fp32 sqrt on this image is not correctly rounded (sqrt(9.049524307250977f) comes back one
        # ulp low), while sqrtf / __fsqrt_rn are; a few gravity trajectories therefore differ in the last bits.
        assert np.mean(np.any(got != ref, axis=(1, 2))) < 0.1
        np.testing.assert_allclose(got, ref, rtol=1e-5, atol=2e-5)
    else:
        assert np.array_equal(got, ref)          # spring / bouncing: bit-identical


@pytest.mark.parametrize("cell", list(CELLS))
def test_rollout_forward_bit_exact_vs_reference(golden_dir, cell):
    gold = np.load(os.path.join(golden_dir, "cells.npz"))
    ref = gold[cell + "/seq"]
    seq = _run_fwd(cell, gold[cell + "/pos0"], gold[cell + "/vel0"], ref.shape[1])
    _same_trajectory(cell, seq[:, 1:], ref)


@pytest.mark.parametrize("cell", list(CELLS))
def test_rollout_backward_vs_autograd(cell):
    cid, n, dt = CELLS[cell]
    g = torch.Generator().manual_seed(11)
    B, steps = 37, 7
    pos0 = torch.rand(B, 2 * n, generator=g) * 30 + 1
    vel0 = (torch.rand(B, 2 * n, generator=g) - 0.5) * 20
    w = torch.randn(B, steps + 1, 4 * n, generator=g)
    spec = {"spring": po.TASKS["spring_color"], "bouncing": po.TASKS["bouncing_balls"],
            "gravity": po.TASKS["3bp_color"]}[cell]
    sd = {"rollout_cell.dt": torch.tensor(dt),
          "rollout_cell.k": torch.tensor(PHYS["spring"][0], dtype=torch.float64, requires_grad=True),
          "rollout_cell.equil": torch.tensor(PHYS["spring"][1], dtype=torch.float64, requires_grad=True),
          "rollout_cell.g": torch.tensor(PHYS["gravity"][0], dtype=torch.float64, requires_grad=True),
          "rollout_cell.m": torch.tensor(PHYS["gravity"][1], dtype=torch.float64)}
    p, v = pos0.clone().requires_grad_(True), vel0.clone().requires_grad_(True)
    rows = [torch.cat([p, v], 1)]
    pp, vv = p, v
    for _ in range(steps):
        pp, vv = po.rollout_cell(sd, spec, pp, vv)
        rows.append(torch.cat([pp, vv], 1))
    seq_ref = torch.stack(rows, 1)
    (seq_ref * w).sum().backward()

    seq = _run_fwd(cell, pos0.numpy(), vel0.numpy(), steps)
    _same_trajectory(cell, seq, seq_ref.detach().numpy())
    d0 = np.zeros((B, 4 * n), np.float32)
    dphys = np.full(2, 123.0)
    dt_a, p0, p1 = eu.f32([dt]), eu.f64([PHYS[cell][0]]), eu.f64([PHYS[cell][1]])
    eu.check(eu.lib().paig_rollout_backward(cid, n, B, steps, eu.ptr(dt_a), eu.ptr(p0), eu.ptr(p1), eu.ptr(seq),
                                            eu.ptr(eu.f32(w.numpy())), eu.ptr(d0), eu.ptr(dphys), None))
    ref0 = torch.cat([p.grad, v.grad], 1).numpy()
    scale = np.abs(ref0).max()
    assert np.abs(d0 - ref0).max() / scale < 2e-5
    if cell == "spring":
        ref_phys = np.array([sd["rollout_cell.k"].grad.item(), sd["rollout_cell.equil"].grad.item()])
        assert np.allclose(dphys, ref_phys, rtol=2e-5, atol=1e-5 * np.abs(ref_phys).max())
    elif cell == "gravity":
        assert np.isclose(dphys[0], sd["rollout_cell.g"].grad.item(), rtol=2e-5)
