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


# ---------------------------------------------------------------------------------------- whole step
def _grad_bufs(be, sd):
    return {k: be.full(tuple(v.shape), 7.0, np.float64 if v.dtype == torch.float64 else np.float32)
            for k, v in sd.items() if k != "rollout_cell.dt"}


def _compare_grads(grad_bufs, ref_grads, sd, tol, ref64=None, report=None, tag=""):
    """Per-tensor criterion (errors are max-abs-diff / max-abs-ref):
         err(ours, ref32) < tol                                                    -- the stated fp32 bound, or
         err(ours, ref64) < tol + 4 * err(ref32, ref64)                            -- the reference's own fp32 rounding
       noise (measured against its float64 twin under the same ReLU decisions) is as large as the difference: no
       independent fp32 implementation can sit closer to ref32 than ref32 sits to the exact gradient."""
    worst = ("", 0.0)
    fails = []
    for k, ref in ref_grads.items():
        got = grad_bufs[k].np()
        assert np.all(np.isfinite(got)), k
        err = rel(got, ref.numpy())
        ok = err < tol
        e64 = n64 = None
        if not ok and ref64 is not None:
            e64, n64 = rel(got, ref64[k].numpy()), rel(ref.numpy(), ref64[k].numpy())
            ok = e64 < tol + 4 * n64
        if report is not None:
            report[tag + k] = dict(err=err, err_vs_f64=e64, ref_noise=n64)
        if err > worst[1]:
            worst = (k, err)
        if not ok:
            fails.append("%s: err %.3e (vs f64 %s, reference fp32 noise %s)" % (k, err, e64, n64))
    assert not fails, "; ".join(fails)
    for k in sd:                       # parameters autograd leaves at grad=None must stay untouched
        if k not in ref_grads and k in grad_bufs:
            assert np.all(grad_bufs[k].np() == 7.0), k
    return worst


def f64_twin(sd, x, spec, alpha, alt_vel, force):
    """The oracle evaluated in float64 (same graph, same ReLU decisions): the exact gradient the fp32 reference
    approximates."""
    torch.set_default_dtype(torch.float64)
    try:
        sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
        f64 = None if force is None else {k: v.double() for k, v in force.items()}
        ff64, _, g64 = po.live_step(sd64, x.double(), spec, alpha, alt_vel, f64)
    finally:
        torch.set_default_dtype(torch.float32)
    return ff64, g64


def relu_decisions(be, tk, spec, B, ws):
    """Read OUR ReLU decisions (activation > 0) for every ReLU in the encoder out of the workspace, as the
    `force` dict of the oracle.  See oracle._relu: gradients are compared under identical kink decisions."""
    w = ws.np()
    N = B * spec.enc_steps
    force = {}
    n_convs = 18 if spec.H >= 40 else 13
    for layer in range(n_convs):
        view = (ctypes.c_long * 5)()
        be.check(be.lib.paig_debug_unet_conv_view(byref(tk), B, layer, byref(view)))
        off, bs, C, S, relu = [int(v) for v in view]
        if not relu:
            continue
        rows = np.stack([w[off + f * bs: off + f * bs + C * S * S] for f in range(N)]).reshape(N, C, S, S)
        force["c%d" % (layer + 1)] = torch.from_numpy((rows > 0).astype(np.float32))
    M = spec.n_objs * N
    for name in ("H1", "H2"):
        off = be.lib.paig_debug_workspace_offset(byref(tk), B, name.encode(), 0)
        force["l" + name[1]] = torch.from_numpy((w[off:off + M * 200].reshape(M, 200) > 0).astype(np.float32))
    return force


def kink_flips(force, sd, x, spec):
    """How many ReLU decisions differ from the oracle's own (unforced) forward: must be a vanishing fraction."""
    rec = {}
    orig = po._relu

    def spy(y, f, name):
        rec[name] = (y.detach() > 0).float()
        return orig(y, f, name)
    po._relu = spy
    try:
        with torch.no_grad():
            po.encoder(sd, x[:, :spec.enc_steps].reshape(-1, 3, spec.H, spec.H), spec)
    finally:
        po._relu = orig
    flips = sum(int((rec[k] != force[k]).sum()) for k in force)
    total = sum(force[k].numel() for k in force)
    return flips, total


def check_step(be, task, B, seed=0, alpha=3.0, alt_vel=False, seq_len=None, tol=1e-4, batch_global=0, fwd_tol=2e-5,
               traj_tol=5e-5, report=None, phys=None, horizon=None):
    """LIVE training step: paig_step_fused, then paig_step_forward + paig_step_backward, vs the oracle.
    horizon = (pred_steps, seq_len) shortens the rollout (the constructor accepts any in/pr/seq_len, physics_models.py:41-55):
    used where a long chaotic rollout would bury the comparison under the reference's own fp32 noise."""
    spec = po.TASKS[task]
    if horizon is not None:
        spec = po.TaskSpec(spec.task, spec.cell, horizon[1], spec.test_seq_len, spec.input_steps, horizon[0], spec.H, spec.n_objs)
    T = seq_len or spec.seq_len
    sd = po.init_state_dict(spec, seed, alt_vel, phys=phys)
    x = po.synthetic_frames(spec, B, T, seed)
    ff, ls, ref_grads = po.live_step(sd, x, spec, alpha, alt_vel)
    ff64_plain, _ = f64_twin(sd, x, spec, alpha, alt_vel, None)      # the reference's own fp32 rounding noise, per tensor
    Bg = batch_global or B
    n, H, e, steps = spec.n_objs, spec.H, spec.enc_steps, T - spec.input_steps
    tk = be.make_task(spec, T, alpha, alt_vel, batch_global)
    bufs = be.sd(sd)
    P = be.make_params(spec, bufs, alt_vel)
    gb = _grad_bufs(be, sd)
    G = be.make_params(spec, gb, alt_vel)
    ws = be.workspace(tk, B)
    xd = be.dev(x.numpy())
    from paig_reproduction_b200 import _abi
    ob = dict(output_seq=be.zeros((B, steps, 3, H, H)), recons_out=be.zeros((B, e, 3, H, H)),
              enc_pos=be.zeros((B, e, 2 * n)), pos_vel_seq=be.zeros((B, steps + 1, 4 * n)),
              enc_masks=be.zeros((B * e, n + 1, H, H)), masked_objs=be.zeros((n, B * e, 3, H, H)),
              templates=be.zeros(n * (H // 2) ** 2 * 4 + 3 * H * H), losses=be.zeros(4))
    scale = float(B) / Bg                                     # oracle means are over the local B
    ref_losses = np.array([ls["train"].item(), ls["pred"].item(), ls["extrap"].item(), ls["recons"].item()]) * scale

    # ---- fused LIVE step ----
    O = _abi.Outputs(None, None, ob["enc_pos"].ptr, ob["pos_vel_seq"].ptr, None, None, None, ob["losses"].ptr)
    be.check(be.lib.paig_step_fused(byref(tk), byref(P), byref(G), xd.ptr, B, byref(O), ws.ptr, be.stream))
    report = {} if report is None else report

    def fwd(name, got, ref, bound, ref64=None):
        """|ours - ref32| < bound, widened by 4x the reference's own fp32-vs-fp64 difference on the same tensor
        (3-body rollouts are chaotic: at B=100 the reference differs from its float64 twin by 3.6e-2)."""
        err = rel(got, ref)
        noise = rel(ref, ref64.detach().numpy()) if ref64 is not None else 0.0
        report["fwd/" + name] = dict(err=err, ref_noise=noise)
        assert err < bound + 4 * noise, "%s: rel err %.3e >= %.1e + 4 * %.1e" % (name, err, bound, noise)

    fwd("losses", ob["losses"].np(), ref_losses, fwd_tol)
    fwd("enc_pos", ob["enc_pos"].np(), ff["enc_pos"].detach().numpy(), fwd_tol, ff64_plain["enc_pos"])
    fwd("pos_vel_seq", ob["pos_vel_seq"].np(), ff["pos_vel_seq"].detach().numpy(), traj_tol, ff64_plain["pos_vel_seq"])
    # gradients are compared under identical ReLU decisions (oracle._relu); decisions may differ only on a
    # vanishing fraction of activations (those whose pre-activation is within rounding noise of zero)
    force = relu_decisions(be, tk, spec, B, ws)
    flips, total = kink_flips(force, sd, x, spec)
    assert flips <= max(2, total // 200000), "%d of %d ReLU decisions differ from the oracle" % (flips, total)
    report["relu_flips"] = (flips, total)
    if flips:
        _, _, ref_grads = po.live_step(sd, x, spec, alpha, alt_vel, force)
    ref64 = {k: v * scale for k, v in f64_twin(sd, x, spec, alpha, alt_vel, force)[1].items()}
    ref32 = {k: v * scale for k, v in ref_grads.items()}
    worst_fused = _compare_grads(gb, ref32, sd, tol, ref64, report, "fused/")

    # ---- forward materialising everything, then backward from explicit upstream gradients ----
    for b_ in gb.values():
        pass
    gb2 = _grad_bufs(be, sd)
    G2 = be.make_params(spec, gb2, alt_vel)
    O2 = _abi.Outputs(*[ob[k].ptr for k in ("output_seq", "recons_out", "enc_pos", "pos_vel_seq", "enc_masks",
                                            "masked_objs", "templates", "losses")])
    be.check(be.lib.paig_step_forward(byref(tk), byref(P), xd.ptr, B, byref(O2), ws.ptr, be.stream))
    fwd("losses2", ob["losses"].np(), ref_losses, fwd_tol)
    fwd("output_seq", ob["output_seq"].np(), ff["output"].detach().numpy(), max(fwd_tol, traj_tol), ff64_plain["output"])
    fwd("recons_out", ob["recons_out"].np(), ff["recons_out"].detach().numpy(), fwd_tol, ff64_plain["recons_out"])
    fwd("enc_masks", ob["enc_masks"].np(), ff["enc_masks"].detach().numpy(), fwd_tol, ff64_plain["enc_masks"])
    fwd("masked_objs", ob["masked_objs"].np(), torch.stack(ff["masked_objs"]).detach().numpy(), fwd_tol,
        torch.stack(ff64_plain["masked_objs"]))
    raw_ref = torch.cat([ff["template"].reshape(-1), ff["contents"].reshape(-1), ff["background"].reshape(-1)])
    fwd("templates", ob["templates"].np(), raw_ref.detach().numpy(), 1e-5)
    out_t, rec_t = torch.from_numpy(ob["output_seq"].np()), torch.from_numpy(ob["recons_out"].np())
    i, p_ = spec.input_steps, spec.pred_steps
    d_out = torch.zeros_like(out_t)
    d_out[:, :p_] = 2.0 * (out_t[:, :p_] - x[:, i:i + p_]) / (Bg * p_)
    d_rec = 2.0 * alpha * (rec_t - x[:, :e]) / (Bg * e)
    d_out_d, d_rec_d = be.dev(d_out.numpy()), be.dev(d_rec.numpy())
    be.check(be.lib.paig_step_backward(byref(tk), byref(P), byref(G2), xd.ptr, B, d_out_d.ptr, d_rec_d.ptr, None, None,
                                       ws.ptr, be.stream))
    worst_bwd = _compare_grads(gb2, ref32, sd, tol, ref64, report, "bwd/")
    return report


# ---------------------------------------------------------------------------------------- layer primitives
def check_conv3x3(be, N, Cin, Cout, S, relu):
    g = torch.Generator().manual_seed(S * 1000 + Cin * 10 + Cout)
    x = torch.randn(N, Cin, S, S, generator=g)
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) / (3.0 * Cin ** 0.5)).requires_grad_(True)
    b = (torch.randn(Cout, generator=g) * 0.1).requires_grad_(True)
    xr = x.clone().requires_grad_(True)
    xd, wd, bd = be.dev(x.numpy()), be.dev(w.detach().numpy()), be.dev(b.detach().numpy())
    yd = be.zeros((N, Cout, S, S))
    be.check(be.lib.paig_conv3x3_forward(xd.ptr, wd.ptr, bd.ptr, yd.ptr, N, Cin, Cout, S, int(relu), be.stream))
    y = F.conv2d(xr, w, b, padding="same")
    if relu:
        assert rel(yd.np(), F.relu(y).detach().numpy()) < 1e-5
        ours = torch.from_numpy((yd.np() > 0).astype(np.float32))          # compare gradients under OUR ReLU decisions
        assert (ours != (y.detach() > 0).float()).sum() <= max(2, y.numel() // 200000)
        y = y * ours
    dy = torch.randn(y.shape, generator=g)
    (y * dy).sum().backward()
    assert rel(yd.np(), y.detach().numpy()) < 1e-5
    dyd = be.dev(dy.numpy())
    dx, dw, db = be.full(tuple(x.shape), 3.0), be.full(tuple(w.shape), 3.0), be.full(tuple(b.shape), 3.0)
    ws = be.zeros(296 * (Cout * Cin * 9 + Cout) + 64)
    be.check(be.lib.paig_conv3x3_backward(xd.ptr, wd.ptr, yd.ptr, dyd.ptr, dx.ptr, dw.ptr, db.ptr, N, Cin, Cout, S,
                                          int(relu), ws.ptr, be.stream))
    for name, got, ref in (("dx", dx, xr.grad), ("dw", dw, w.grad), ("db", db, b.grad)):
        err = rel(got.np(), ref.numpy())
        assert err < 3e-5, "%s rel err %.3e" % (name, err)      # fp32 sums of up to N*S*S*9 terms in a different order
