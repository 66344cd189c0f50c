"""Generate tests/golden/*.npz by executing the UNMODIFIED reference (/root/reference).

TEST INFRASTRUCTURE.  Runs only in the build container (the reference checkout does not
exist on the GPU box).  Recipe for importing the reference follows SURVEY.md section 8(c):
inert stubs for tensorflow / matplotlib, then ``from nn.network import physics_models``.

Weights and inputs are regenerated deterministically at test time from
``oracle.physicsnet_oracle.init_state_dict`` / ``synthetic_frames`` (seeded CPU generator), so
the fixtures hold only the reference's OUTPUTS: losses, positions, rollouts, sub-sampled
frames and per-parameter gradient digests (sum, L2 norm, a strided sample).

    python oracle/make_golden.py            # rewrites tests/golden/
"""
from __future__ import annotations

import importlib.machinery
import os
import sys
import types

import numpy as np
import torch
import torchvision  # noqa: F401  (real dep of the reference; import before stubbing)

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import physicsnet_oracle as po  # noqa: E402

REFERENCE = "/root/reference"

# (fixture name, task, batch, seq_len or None for train length, seed, alpha, alt_vel, mode)
CASES = [
    ("spring_color_b3", "spring_color", 3, None, 0, 3.0, False, "train"),
    ("spring_color_b2_s1", "spring_color", 2, None, 1, 3.0, False, "train"),
    ("spring_color_altvel_b2", "spring_color", 2, None, 2, 3.0, True, "train"),
    ("bouncing_balls_b3", "bouncing_balls", 3, None, 0, 2.0, False, "train"),
    ("3bp_color_b2", "3bp_color", 2, None, 0, 5.0, False, "train"),
    ("mnist_spring_color_b2", "mnist_spring_color", 2, None, 0, 3.0, False, "train"),
    ("spring_color_half_test_b2", "spring_color_half", 2, 30, 0, 3.0, False, "eval"),
    ("3bp_color_test_b2", "3bp_color", 2, 40, 1, 5.0, False, "eval"),
]

GRAD_SAMPLE = 48


def import_reference():
    def stub(name):
        m = types.ModuleType(name)
        m.__spec__ = importlib.machinery.ModuleSpec(name, None)
        sys.modules[name] = m
        return m

    stub("tensorflow")
    mpl, cm, plt = stub("matplotlib"), stub("matplotlib.cm"), stub("matplotlib.pyplot")
    mpl.cm, mpl.pyplot = cm, plt
    plt.switch_backend = lambda *a, **k: None
    sys.path.insert(0, REFERENCE)
    sys.dont_write_bytecode = True
    from nn.network import physics_models  # type: ignore
    return physics_models


def build_reference_net(pm, spec: po.TaskSpec, seq_len: int, alpha: float, alt_vel: bool):
    # positional order of runners/torch_run_physics.py:81-84
    return pm.PhysicsNet(spec.task, 100, 1, po.CELL_TYPE_NAMES[spec.cell], seq_len, spec.input_steps,
                         spec.pred_steps, alpha, alt_vel, True, spec.H * spec.H, "conv_encoder",
                         "conv_st_decoder")


def grad_digest(g: torch.Tensor) -> np.ndarray:
    """[sum, l2, then GRAD_SAMPLE strided samples] as float64."""
    f = g.detach().double().reshape(-1)
    idx = torch.linspace(0, f.numel() - 1, GRAD_SAMPLE).long()
    return torch.cat([f.sum()[None], f.norm()[None], f[idx]]).numpy()


def run_case(pm, name, task, batch, seq_len, seed, alpha, alt_vel, mode):
    spec = po.TASKS[task]
    T = seq_len or spec.seq_len
    sd = po.init_state_dict(spec, seed, alt_vel)
    x = po.synthetic_frames(spec, batch, T, seed)
    torch.manual_seed(0)
    net = build_reference_net(pm, spec, T, alpha, alt_vel)
    missing = net.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    if spec.cell == "gravity":                                      # refresh the cached A
        net.rollout_cell.A = torch.exp(net.rollout_cell.g) * torch.exp(2 * net.rollout_cell.m)
    out = {}
    if mode == "train":
        net.train()
        inp = x.clone().requires_grad_(True)                        # base.py:141
        net.output = net(inp)                                       # LIVE mode, SURVEY Q1 (base.py:195)
        train, (pred_alias, extrap, recons) = net.compute_loss()
        train.backward()
        for k, p in net.named_parameters():
            if p.grad is not None:
                out["grad/" + k] = grad_digest(p.grad)
    else:
        net.eval()
        with torch.no_grad():
            net.output = net.conv_feedforward(x)
            train, (pred_alias, extrap, recons) = net.compute_loss()
    pred = (train - alpha * recons) if alpha > 0 else train          # un-alias (Q4)
    out["losses"] = np.array([train.item(), pred.item(), extrap.item(), recons.item()], dtype=np.float64)
    out["enc_pos"] = net.enc_pos.detach().numpy()
    out["pos_vel_seq"] = net.pos_vel_seq.detach().numpy()
    out["output_sub"] = net.output.detach()[:, :, :, ::3, ::3].numpy()
    out["recons_sub"] = net.recons_out.detach()[:, :, :, ::3, ::3].numpy()
    out["output_sum"] = net.output.detach().double().sum((2, 3, 4)).numpy()
    out["recons_sum"] = net.recons_out.detach().double().sum((2, 3, 4)).numpy()
    out["enc_masks_sub"] = net.enc_masks.detach()[:, :, ::4, ::4].numpy()
    out["template"] = net.template.detach().numpy()
    out["contents_sub"] = net.contents.detach()[:, :, ::2, ::2].numpy()
    out["meta"] = np.array([batch, T, seed, alpha, float(alt_vel), float(mode == "train")], dtype=np.float64)
    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(path, **out)
    print("%-32s train=%.6f pred=%.6f extrap=%.6f recons=%.6f  %d arrays  %.0f KB" % (
        name, *out["losses"], len(out), os.path.getsize(path) / 1024))


def rollout_cases(pm):
    """Stage-level goldens for the three cells, with wall hits / near-coincident bodies."""
    import nn.network.cells as cells  # type: ignore
    g = torch.Generator().manual_seed(7)
    out = {}
    B, steps = 64, 12
    for cell, n in (("spring", 2), ("bouncing", 2), ("gravity", 3)):
        pos = torch.rand(B, 2 * n, generator=g) * 36 - 2
        vel = (torch.rand(B, 2 * n, generator=g) - 0.5) * 40
        if cell == "gravity":
            pos[:8, 2:4] = pos[:8, 0:2] + 0.05 * torch.rand(8, 2, generator=g)      # near-coincident bodies
            c = cells.gravity_ode_cell(2 * n, 2 * n)
            with torch.no_grad():
                c.g.fill_(np.log(30.0))
            c.A = torch.exp(c.g) * torch.exp(2 * c.m)
        elif cell == "spring":
            c = cells.spring_ode_cell(2 * n, 2 * n)
            with torch.no_grad():
                c.k.fill_(np.log(1.7))
                c.equil.fill_(np.log(2.5))
        else:
            c = cells.bouncing_ode_cell(2 * n, 2 * n)
        seq = []
        p, v = pos.clone(), vel.clone()
        with torch.no_grad():
            for _ in range(steps):
                p, v = c(p, v)
                seq.append(torch.cat([p, v], 1))
        out[cell + "/pos0"] = pos.numpy()
        out[cell + "/vel0"] = vel.numpy()
        out[cell + "/seq"] = torch.stack(seq, 1).numpy()
    path = os.path.join(ROOT, "tests", "golden", "cells.npz")
    np.savez_compressed(path, **out)
    print("cells.npz %.0f KB" % (os.path.getsize(path) / 1024))


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    pm = import_reference()
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    rollout_cases(pm)
    for case in CASES:
        run_case(pm, *case)


if __name__ == "__main__":
    main()
