"""Template generators (varnet.cu) and the fused STN decoder (decoder.cu) through the SIMT-on-CPU shim vs
the oracle (and its autograd).  CPU only; the same checks run on the real library under -m gpu."""
import ctypes

import numpy as np
import pytest
import torch

import emu_util as eu
from oracle import physicsnet_oracle as po

TASKS = ["spring_color", "3bp_color", "mnist_spring_color"]


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def _consts_from_oracle(sd, spec):
    tpl, con, bg = po.learned_tensors(sd, spec)
    return tpl, con, bg, torch.cat([(tpl + 5).reshape(-1), torch.sigmoid(con).reshape(-1),
                                    torch.sigmoid(bg).reshape(-1)])


def _decode_from_consts(spec, T5, SC, SB, loc):
    """oracle decoder with the constants as explicit leaves (so autograd gives d consts)."""
    n, t, H = spec.n_objs, spec.H // 2, spec.H
    tpl = (T5 - 5).reshape(n, 1, t, t)
    sd = {}
    learned = (tpl, torch.logit(SC.reshape(n, 3, t, t).double()).float(), torch.logit(SB.reshape(1, 3, H, H).double()).float())
    # re-derive through the oracle's own decoder; sigmoid(logit(x)) == x to 1 ulp, and gradients are taken w.r.t.
    # the post-sigmoid leaves below instead, so build the decoder inline from po pieces:
    import torch.nn.functional as F
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


@pytest.mark.parametrize("task", TASKS)
def test_templates_forward_backward(task):
    spec = po.TASKS[task]
    sd = po.init_state_dict(spec, 3)
    arrs = eu.sd_to_numpy(sd)
    tk = eu.make_task(spec)
    P = eu.make_params(spec, arrs)
    d = spec
    n, t, H = d.n_objs, d.H // 2, d.H
    CN = n * t * t * 4 + 3 * H * H
    raw, consts, hidden = np.zeros(CN, np.float32), np.zeros(CN, np.float32), np.zeros(600, np.float32)
    eu.check(eu.lib().paig_templates_forward(ctypes.byref(tk), ctypes.byref(P), eu.ptr(raw), eu.ptr(consts),
                                             eu.ptr(hidden), None))
    keys = [k for k in sd if k.startswith("var_net_")]
    leaves = {k: sd[k].clone().requires_grad_(True) for k in keys}
    tpl, con, bg, cref = _consts_from_oracle({**sd, **leaves}, spec)
    assert _rel(consts, cref.detach().numpy()) < 1e-6
    assert _rel(raw, torch.cat([tpl.reshape(-1), con.reshape(-1), bg.reshape(-1)]).detach().numpy()) < 1e-6
    g = torch.Generator().manual_seed(5)
    dC = torch.randn(CN, generator=g)
    (cref * dC).sum().backward()
    grads = {k: np.full(arrs[k].shape, 7.0, np.float32) for k in keys}
    G = eu.make_params(spec, grads)
    ws = eu.workspace(tk, 1)
    eu.check(eu.lib().paig_templates_backward(ctypes.byref(tk), ctypes.byref(P), ctypes.byref(G), eu.ptr(consts),
                                              eu.ptr(hidden), eu.ptr(eu.f32(dC.numpy())), eu.ptr(ws), None))
    for k in keys:
        assert _rel(grads[k], leaves[k].grad.numpy()) < 2e-5, k


@pytest.mark.parametrize("task", TASKS)
@pytest.mark.parametrize("mode", ["dframes", "fused_loss"])
def test_decode_forward_backward(task, mode):
    spec = po.TASKS[task]
    n, t, H = spec.n_objs, spec.H // 2, spec.H
    sd = po.init_state_dict(spec, 4)
    tk = eu.make_task(spec)
    _, _, _, cref = _consts_from_oracle(sd, spec)
    consts = eu.f32(cref.numpy())
    g = torch.Generator().manual_seed(9)
    fps = 3
    F = 6 if H < 64 else 3
    loc = (torch.rand(F, 2 * n, generator=g) * 1.4 - 0.2) * H            # spans [-0.2H, 1.2H]: window edges + outside
    loc[0] = H / 2                                                        # exactly centred (integer-aligned taps)
    CN = consts.size
    T5 = cref[:n * t * t].clone().requires_grad_(True)
    SC = cref[n * t * t:4 * n * t * t].clone().requires_grad_(True)
    SB = cref[4 * n * t * t:].clone().requires_grad_(True)
    locr = loc.clone().requires_grad_(True)
    out_ref = _decode_from_consts(spec, T5, SC, SB, locr)
    out_oracle = po.decoder(sd, loc, spec)
    assert _rel(out_ref.detach().numpy(), out_oracle.numpy()) < 1e-6       # the inline restatement == oracle decoder

    Q = F // fps
    target = torch.rand(Q, fps + 2, 3, H, H, generator=g)                 # frames of a longer "sequence" tensor
    tgt_np = eu.f32(target.numpy())
    frames = np.zeros((F, 3, H, H), np.float32)
    sse = np.zeros(F, np.float32)
    eu.check(eu.lib().paig_decode_forward(ctypes.byref(tk), eu.ptr(consts), eu.ptr(eu.f32(loc.numpy())), F,
                                          eu.ptr(frames), eu.ptr(tgt_np), (fps + 2) * 3 * H * H, fps, eu.ptr(sse), None))
    assert _rel(frames, out_ref.detach().numpy()) < 2e-6
    tsel = target[:, :fps].reshape(F, 3, H, H)
    sse_ref = ((tsel - out_ref.detach()) ** 2).sum((1, 2, 3)).numpy()
    assert _rel(sse, sse_ref) < 1e-5

    scale = torch.tensor([0.7, 0.0, 1.3])                                  # a zero-weight frame (extrapolation) too
    if mode == "dframes":
        dfr = torch.randn(F, 3, H, H, generator=g)
        (out_ref * dfr).sum().backward()
    else:
        per = ((out_ref - tsel) ** 2).sum((1, 2, 3)).reshape(Q, fps)
        (per * scale[None]).sum().backward()
    d_loc = np.full((F, 2 * n), 9.0, np.float32)
    d_consts = np.ones(CN, np.float32)                                     # accumulate semantics: starts at 1
    sse2 = np.zeros(F, np.float32)
    ws = eu.workspace(tk, 1)
    eu.check(eu.lib().paig_decode_backward(
        ctypes.byref(tk), eu.ptr(consts), eu.ptr(eu.f32(loc.numpy())), F,
        eu.ptr(eu.f32(dfr.numpy())) if mode == "dframes" else None, eu.ptr(tgt_np), (fps + 2) * 3 * H * H, fps,
        eu.ptr(eu.f32(scale.numpy())), eu.ptr(d_loc), eu.ptr(d_consts), eu.ptr(sse2), eu.ptr(ws), None))
    assert _rel(sse2, sse_ref) < 1e-5
    assert _rel(d_loc, locr.grad.numpy()) < 2e-5
    dref = torch.cat([T5.grad, SC.grad, SB.grad]).numpy()
    assert _rel(d_consts[:n * t * t] - 1, dref[:n * t * t]) < 2e-5
    assert _rel(d_consts[n * t * t:4 * n * t * t] - 1, dref[n * t * t:4 * n * t * t]) < 2e-5
    assert _rel(d_consts[4 * n * t * t:] - 1, dref[4 * n * t * t:]) < 2e-5
