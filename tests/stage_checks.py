"""Stage-level parity checks of the C-ABI entry points against the oracle / reference goldens, written
against the tests/backends.py interface so they run on the SIMT shim (CPU suite) and on the B200 (-m gpu)."""
from __future__ import annotations

import ctypes
import os

import numpy as np
import torch
import torch.nn.functional as F

from oracle import physicsnet_oracle as po

byref = ctypes.byref
CELLS = {"spring": (0, 2, 0.3), "bouncing": (1, 2, 0.3), "gravity": (2, 3, 0.5)}
PHYS = {"spring": (np.log(1.7), np.log(2.5)), "bouncing": (0.0, 0.0), "gravity": (np.log(30.0), np.log(1.0))}
CELL_SPEC = {"spring": "spring_color", "bouncing": "bouncing_balls", "gravity": "3bp_color"}


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


# ---------------------------------------------------------------------------------------- rollout
def _rollout_fwd(be, cell, pos0, vel0, steps):
    cid, n, dt = CELLS[cell]
    B = pos0.shape[0]
    seq0 = np.zeros((B, steps + 1, 4 * n), np.float32)
    seq0[:, 0, :2 * n] = pos0
    seq0[:, 0, 2 * n:] = vel0
    seq = be.dev(seq0)
    dt_a, p0, p1 = be.dev(np.float32([dt])), be.dev(np.float64([PHYS[cell][0]])), be.dev(np.float64([PHYS[cell][1]]))
    be.check(be.lib.paig_rollout_forward(cid, n, B, steps, dt_a.ptr, p0.ptr, p1.ptr, seq.ptr, be.stream))
    return seq, (dt_a, p0, p1)


def _same_trajectory(cell, got, ref, exact_sqrt_ref):
    if cell == "gravity" and not exact_sqrt_ref:
        # ATen's AVX512 fp32 sqrt on this image is not correctly rounded (sqrt(9.049524307250977f) comes back one
        # ulp low) while sqrtf / __fsqrt_rn are; a few gravity trajectories therefore differ in the last bits.
        assert np.mean(np.any(got != ref, axis=(1, 2))) < 0.1
        np.testing.assert_allclose(got, ref, rtol=1e-4, atol=2e-4)   # 3-body chaos amplifies the 1-ulp seeds
    else:
        assert np.array_equal(got, ref)          # spring / bouncing: bit-identical to the reference


def check_rollout_forward_golden(be, golden_dir, cell):
    gold = np.load(os.path.join(golden_dir, "cells.npz"))
    ref = gold[cell + "/seq"]
    seq, _ = _rollout_fwd(be, cell, gold[cell + "/pos0"], gold[cell + "/vel0"], ref.shape[1])
    _same_trajectory(cell, seq.np()[:, 1:], ref, False)


def check_rollout_backward(be, cell, B=37, steps=7):
    cid, n, dt = CELLS[cell]
    g = torch.Generator().manual_seed(11)
    pos0 = torch.rand(B, 2 * n, generator=g) * 30 + 1
    vel0 = (torch.rand(B, 2 * n, generator=g) - 0.5) * 20
    w = torch.randn(B, steps + 1, 4 * n, generator=g)
    spec = po.TASKS[CELL_SPEC[cell]]
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

    seq, (dt_a, p0, p1) = _rollout_fwd(be, cell, pos0.numpy(), vel0.numpy(), steps)
    _same_trajectory(cell, seq.np(), seq_ref.detach().numpy(), False)
    d0, dphys = be.zeros((B, 4 * n)), be.full(2, 123.0, np.float64)
    w_d = be.dev(w.numpy())           # keep every device buffer referenced until the call has finished
    be.check(be.lib.paig_rollout_backward(cid, n, B, steps, dt_a.ptr, p0.ptr, p1.ptr, seq.ptr,
                                          w_d.ptr, d0.ptr, dphys.ptr, be.stream))
    ref0 = torch.cat([p.grad, v.grad], 1).numpy()
    assert rel(d0.np(), ref0) < 2e-5
    if cell == "spring":
        ref_phys = np.array([sd["rollout_cell.k"].grad.item(), sd["rollout_cell.equil"].grad.item()])
        assert np.allclose(dphys.np(), ref_phys, rtol=2e-5, atol=1e-5 * np.abs(ref_phys).max())
    elif cell == "gravity":
        assert np.isclose(dphys.np()[0], sd["rollout_cell.g"].grad.item(), rtol=2e-5)


# ---------------------------------------------------------------------------------------- templates / decoder
def consts_from_oracle(sd, spec):
    tpl, con, bg = po.learned_tensors(sd, spec)
    return tpl, con, bg, torch.cat([(tpl + 5).reshape(-1), torch.sigmoid(con).reshape(-1),
                                    torch.sigmoid(bg).reshape(-1)])


def decode_from_consts(spec, T5, SC, SB, loc):
    """The oracle decoder (po.decoder) with the post-sigmoid constants as explicit leaves, so autograd
    yields the gradient w.r.t. the `consts` block.  Checked against po.decoder in the test."""
    n, t, H = spec.n_objs, spec.H // 2, spec.H
    N = loc.shape[0]
    joint = torch.cat([T5.reshape(n, 1, t, t).repeat(1, 3, 1, 1), SC.reshape(n, 3, t, t)], 1)
    one, zero = torch.ones(N, dtype=torch.float64), torch.zeros(N, dtype=torch.float64)
    sampled = []
    for o in range(n):
        lx, ly = loc[:, 2 * o], loc[:, 2 * o + 1]
        theta = torch.stack([one, zero, (H / 2 - lx) / t * 1.0, zero, one, (H / 2 - ly) / t * 1.0], 1)
        grid = F.affine_grid(theta.view(-1, 2, 3), torch.Size((N, 6, H, H)), align_corners=False)
        s = F.grid_sample(joint[o:o + 1].expand(N, -1, -1, -1).float(), grid.float(), mode="bilinear",
                          padding_mode="zeros", align_corners=False)
        sampled.append((s[:, :3], s[:, 3:]))
    bg = SB.reshape(1, 3, H, H).expand(N, -1, -1, -1)
    logits = torch.stack([m - 5 for m, _ in sampled] + [torch.ones_like(sampled[0][0])], 1)
    w = torch.softmax(logits, 1)
    layers = [c for _, c in sampled] + [bg]
    return sum(w[:, i] * layers[i] for i in range(n + 1))


def check_templates(be, task):
    spec = po.TASKS[task]
    sd = po.init_state_dict(spec, 3)
    bufs = be.sd(sd)
    tk = be.make_task(spec)
    P = be.make_params(spec, bufs)
    n, t, H = spec.n_objs, spec.H // 2, spec.H
    CN = n * t * t * 4 + 3 * H * H
    raw, consts, hidden = be.zeros(CN), be.zeros(CN), be.zeros(600)
    be.check(be.lib.paig_templates_forward(byref(tk), byref(P), raw.ptr, consts.ptr, hidden.ptr, be.stream))
    keys = [k for k in sd if k.startswith("var_net_")]
    leaves = {k: sd[k].clone().requires_grad_(True) for k in keys}
    tpl, con, bg, cref = consts_from_oracle({**sd, **leaves}, spec)
    assert rel(consts.np(), cref.detach().numpy()) < 1e-6
    assert rel(raw.np(), torch.cat([tpl.reshape(-1), con.reshape(-1), bg.reshape(-1)]).detach().numpy()) < 1e-6
    g = torch.Generator().manual_seed(5)
    dC = torch.randn(CN, generator=g)
    (cref * dC).sum().backward()
    grads = {k: be.full(tuple(sd[k].shape), 7.0) for k in keys}
    G = be.make_params(spec, grads)
    ws = be.workspace(tk, 1)
    dC_d = be.dev(dC.numpy())
    be.check(be.lib.paig_templates_backward(byref(tk), byref(P), byref(G), consts.ptr, hidden.ptr,
                                            dC_d.ptr, ws.ptr, be.stream))
    for k in keys:
        assert rel(grads[k].np(), leaves[k].grad.numpy()) < 2e-5, k


def check_decode(be, task, mode, F_=None):
    spec = po.TASKS[task]
    n, t, H = spec.n_objs, spec.H // 2, spec.H
    sd = po.init_state_dict(spec, 4)
    tk = be.make_task(spec)
    _, _, _, cref = consts_from_oracle(sd, spec)
    consts = be.dev(cref.numpy())
    g = torch.Generator().manual_seed(9)
    fps = 3
    Fn = F_ or (6 if H < 64 else 3)
    loc = (torch.rand(Fn, 2 * n, generator=g) * 1.4 - 0.2) * H            # spans [-0.2H, 1.2H]: window edges + outside
    loc[0] = H / 2                                                        # exactly centred (integer-aligned taps)
    CN = cref.numel()
    T5 = cref[:n * t * t].clone().requires_grad_(True)
    SC = cref[n * t * t:4 * n * t * t].clone().requires_grad_(True)
    SB = cref[4 * n * t * t:].clone().requires_grad_(True)
    locr = loc.clone().requires_grad_(True)
    out_ref = decode_from_consts(spec, T5, SC, SB, locr)
    assert rel(out_ref.detach().numpy(), po.decoder(sd, loc, spec).numpy()) < 1e-6

    Q = Fn // fps
    target = torch.rand(Q, fps + 2, 3, H, H, generator=g)                 # frames inside a longer sequence tensor
    tgt = be.dev(target.numpy())
    loc_d = be.dev(loc.numpy())
    frames, sse = be.zeros((Fn, 3, H, H)), be.zeros(Fn)
    be.check(be.lib.paig_decode_forward(byref(tk), consts.ptr, loc_d.ptr, Fn, frames.ptr, tgt.ptr,
                                        (fps + 2) * 3 * H * H, fps, sse.ptr, be.stream))
    assert rel(frames.np(), out_ref.detach().numpy()) < 2e-6
    tsel = target[:, :fps].reshape(Fn, 3, H, H)
    sse_ref = ((tsel - out_ref.detach()) ** 2).sum((1, 2, 3)).numpy()
    assert rel(sse.np(), sse_ref) < 1e-5

    scale = torch.tensor([0.7, 0.0, 1.3])                                  # a zero-weight (extrapolation) frame too
    dfr = None
    if mode == "dframes":
        dfr = torch.randn(Fn, 3, H, H, generator=g)
        (out_ref * dfr).sum().backward()
    else:
        per = ((out_ref - tsel) ** 2).sum((1, 2, 3)).reshape(Q, fps)
        (per * scale[None]).sum().backward()
    d_loc, d_consts, sse2 = be.full((Fn, 2 * n), 9.0), be.full(CN, 1.0), be.zeros(Fn)     # d_consts accumulates
    ws = be.workspace(tk, 1)
    dfr_d = be.dev(dfr.numpy()) if dfr is not None else None
    scale_d = be.dev(scale.numpy())
    be.check(be.lib.paig_decode_backward(
        byref(tk), consts.ptr, loc_d.ptr, Fn, be.p(dfr_d), tgt.ptr,
        (fps + 2) * 3 * H * H, fps, scale_d.ptr, d_loc.ptr, d_consts.ptr, sse2.ptr, ws.ptr, be.stream))
    assert rel(sse2.np(), sse_ref) < 1e-5
    assert rel(d_loc.np(), locr.grad.numpy()) < 2e-5
    dref = torch.cat([T5.grad, SC.grad, SB.grad]).numpy()
    dc = d_consts.np() - 1
    a, b = n * t * t, 4 * n * t * t
    assert rel(dc[:a], dref[:a]) < 2e-5
    assert rel(dc[a:b], dref[a:b]) < 2e-5
    assert rel(dc[b:], dref[b:]) < 2e-5
