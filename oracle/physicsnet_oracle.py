"""CPU oracle for the PhysicsNet per-sequence training step.

TEST INFRASTRUCTURE ONLY.  This module is a plain-PyTorch (CPU, fp32 + the
reference's fp64 scalars) restatement of the reference algorithm.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it; the product path
(``paig_reproduction_b200``) never does and fails loudly without its CUDA
library.

Parity status: PINNED.  The reference ships no tests or golden vectors
(SURVEY.md section 4), so the oracle is pinned against outputs of the
unmodified reference executed in the build container:
``oracle/make_golden.py`` imports ``/root/reference`` and writes
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` replays them.

It is written functionally over a ``state_dict`` (same 93 / 91 keys as the
reference) so the oracle, the reference and the CUDA path all share weights
by name.  Every function cites the reference lines it restates
(paths relative to the reference checkout).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ---------------------------------------------------------------------------
# Task table: runners/torch_run_physics.py:49-75, COORD_UNITS physics_models.py:31-37
# ---------------------------------------------------------------------------
@dataclass(frozen=True)
class TaskSpec:
    task: str
    cell: str            # "spring" | "bouncing" | "gravity"
    seq_len: int
    test_seq_len: int
    input_steps: int
    pred_steps: int
    H: int
    n_objs: int

    @property
    def enc_steps(self) -> int:
        return self.input_steps + self.pred_steps

    def with_seq_len(self, seq_len: int) -> "TaskSpec":
        return TaskSpec(self.task, self.cell, seq_len, self.test_seq_len, self.input_steps,
                        self.pred_steps, self.H, self.n_objs)


TASKS: Dict[str, TaskSpec] = {
    "bouncing_balls": TaskSpec("bouncing_balls", "bouncing", 12, 30, 4, 6, 32, 2),
    "spring_color": TaskSpec("spring_color", "spring", 12, 30, 4, 6, 32, 2),
    "spring_color_half": TaskSpec("spring_color_half", "spring", 12, 30, 4, 6, 32, 2),
    "3bp_color": TaskSpec("3bp_color", "gravity", 20, 40, 4, 12, 36, 3),
    "mnist_spring_color": TaskSpec("mnist_spring_color", "spring", 12, 30, 3, 7, 64, 2),
}

CELL_TYPE_NAMES = {"spring": "spring_ode_cell", "bouncing": "bouncing_ode_cell",
                   "gravity": "gravity_ode_cell"}


# ---------------------------------------------------------------------------
# Parameter table (names/shapes follow the reference's state_dict)
# ---------------------------------------------------------------------------
def param_shapes(spec: TaskSpec, alt_vel: bool = False) -> List[Tuple[str, Tuple[int, ...], torch.dtype]]:
    """Ordered (name, shape, dtype) list identical to ``PhysicsNet.state_dict()``.

    Order = construction order in physics_models.py:106-111, blocks.py:52-75,
    blocks.py:106-170 / 240-276, cells.py:6-8,24-29,55-58,87-94.
    """
    n, H, C = spec.n_objs, spec.H, 3
    t = H // 2
    out: List[Tuple[str, Tuple[int, ...], torch.dtype]] = []

    def lin(prefix, fin, fout):
        out.append((prefix + ".weight", (fout, fin), torch.float32))
        out.append((prefix + ".bias", (fout,), torch.float32))

    def conv(prefix, cin, cout, k=3):
        out.append((prefix + ".weight", (cout, cin, k, k), torch.float32))
        out.append((prefix + ".bias", (cout,), torch.float32))

    for name, numel in (("var_net_content", n * C * t * t), ("var_net_background", C * H * H),
                        ("var_net_template", n * t * t)):
        lin(name + ".l1", 10, 200)
        lin(name + ".l2", 200, numel)
    h = 8
    p = "encoder.shallow_unet."
    for i, (ci, co) in enumerate([(C, h), (h, h), (h, 2 * h), (2 * h, 2 * h), (2 * h, 4 * h), (4 * h, 4 * h),
                                  (4 * h, 2 * h), (4 * h, 2 * h), (2 * h, 2 * h), (2 * h, 2 * h),
                                  (3 * h, h), (h, h)], start=1):
        conv(p + "c%d" % i, ci, co)
    conv(p + "c13", h, n, 1)
    h = 16
    p = "encoder.unet."
    for i, (ci, co) in enumerate([(C, h), (h, h), (h, 2 * h), (2 * h, 2 * h), (2 * h, 4 * h), (4 * h, 4 * h),
                                  (4 * h, 8 * h), (8 * h, 8 * h), (8 * h, 2 * h), (6 * h, 4 * h), (4 * h, 4 * h),
                                  (4 * h, 2 * h), (4 * h, 2 * h), (2 * h, 2 * h), (2 * h, 2 * h), (3 * h, h),
                                  (h, h)], start=1):
        conv(p + "c%d" % i, ci, co)
    conv(p + "c18", h, n, 1)
    l1_in = H * H * C if H < 40 else (H // 2) * (H // 2) * C
    lin("encoder.l1", l1_in, 200)
    lin("encoder.l2", 200, 200)
    lin("encoder.l3", 200, 2)
    if alt_vel:
        lin("velocity_encoder.init_vel_linear", (spec.input_steps - 1) * 2, 2)
    else:
        lin("velocity_encoder.init_vel_mlp.0", spec.input_steps * 2, 100)
        lin("velocity_encoder.init_vel_mlp.2", 100, 100)
        lin("velocity_encoder.init_vel_mlp.4", 100, 2)
    hs = 2 * n
    out += [("rollout_cell.weight_ih", (hs, hs), torch.float32), ("rollout_cell.weight_hh", (hs, hs), torch.float32),
            ("rollout_cell.bias_ih", (hs,), torch.float32), ("rollout_cell.bias_hh", (hs,), torch.float32),
            ("rollout_cell.dt", (), torch.float32)]
    if spec.cell == "spring":
        out += [("rollout_cell.k", (), torch.float64), ("rollout_cell.equil", (), torch.float64)]
    elif spec.cell == "gravity":
        out += [("rollout_cell.g", (), torch.float64), ("rollout_cell.m", (), torch.float64)]
    return out


def init_state_dict(spec: TaskSpec, seed: int, alt_vel: bool = False, spread: float = 20.0,
                    phys: Dict[str, float] | None = None) -> Dict[str, Tensor]:
    """Deterministic synthetic weights (NOT the reference's init): every tensor is
    U(-1/sqrt(fan_in), 1/sqrt(fan_in)) from one CPU generator, walked in
    ``param_shapes`` order; ``encoder.l3.weight`` is scaled by ``spread`` so encoded
    positions cover the frame; physics scalars get non-trivial values.
    Both ``make_golden.py`` (which loads them into the real reference) and the tests
    call this, so no weights need to be committed."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {}
    for name, shape, dtype in param_shapes(spec, alt_vel):
        if len(shape) == 0:
            continue
        fan_in = 1
        if name.endswith(".weight") or name.startswith("rollout_cell.weight"):
            for s in shape[1:]:
                fan_in *= s
        else:
            fan_in = max(shape[0], 1)
        bound = 1.0 / math.sqrt(fan_in)
        if len(shape) == 4:                       # conv weights: He-uniform so ReLU stacks stay alive
            bound = math.sqrt(6.0 / fan_in)
        sd[name] = (torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) * bound
    sd["encoder.l3.weight"] = sd["encoder.l3.weight"] * spread
    for last in ("encoder.shallow_unet.c13", "encoder.unet.c18"):   # keep the mask logits off the ReLU floor
        sd[last + ".bias"] = sd[last + ".bias"] + 1.0
        sd[last + ".weight"] = sd[last + ".weight"] * 8.0
    # make templates / contents non-degenerate so masks and colours vary
    for vn, s in (("var_net_template", 6.0), ("var_net_content", 4.0), ("var_net_background", 2.0)):
        sd[vn + ".l2.weight"] = sd[vn + ".l2.weight"] * s
    phys = dict(phys or {})
    if spec.cell == "spring":
        sd["rollout_cell.dt"] = torch.tensor(0.3)
        sd["rollout_cell.k"] = torch.tensor(phys.get("k", math.log(1.7)), dtype=torch.float64)
        sd["rollout_cell.equil"] = torch.tensor(phys.get("equil", math.log(2.5)), dtype=torch.float64)
    elif spec.cell == "bouncing":
        sd["rollout_cell.dt"] = torch.tensor(0.3)
    else:
        sd["rollout_cell.dt"] = torch.tensor(0.5)
        sd["rollout_cell.g"] = torch.tensor(phys.get("g", math.log(8.0)), dtype=torch.float64)
        sd["rollout_cell.m"] = torch.tensor(phys.get("m", math.log(1.0)), dtype=torch.float64)
    return sd


def synthetic_frames(spec: TaskSpec, batch: int, seq_len: int, seed: int) -> Tensor:
    """Disc sequences + low-amplitude noise, float32 in [0,1], [B,T,3,H,H].

    Loosely follows the generators' physics (generators.py:311-329, 601-618: one disc of
    radius ~2-3 per object drawn into its own colour channel, constant-velocity motion);
    only used to give the encoder structured input, not to reproduce the datasets."""
    g = torch.Generator().manual_seed(1000 + seed)
    H, n = spec.H, spec.n_objs
    pos = torch.rand(batch, n, 2, generator=g) * (H - 10) + 5
    vel = (torch.rand(batch, n, 2, generator=g) - 0.5) * 3.0
    yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(H, dtype=torch.float32),
                            indexing="ij")
    x = torch.rand(batch, seq_len, 3, H, H, generator=g) * 0.08
    for tt in range(seq_len):
        p = pos + vel * tt
        p = (H - 1) - ((H - 1) - p.remainder(2 * (H - 1))).abs()          # reflect into [0,H-1]
        for o in range(n):
            d2 = (xx[None] - p[:, o, 0, None, None]) ** 2 + (yy[None] - p[:, o, 1, None, None]) ** 2
            disc = torch.clamp(1.5 - (d2.sqrt() - (H / 12.0)), 0.0, 1.0)
            ch = 2 - (o % 3)
            x[:, tt, ch] = torch.maximum(x[:, tt, ch], disc)
    return x.clamp_(0.0, 1.0).contiguous()


# ---------------------------------------------------------------------------
# Encoder
# ---------------------------------------------------------------------------
def _relu(y, force, name):
    """ReLU; with ``force[name]`` (a 0/1 tensor of y's shape) the kink decisions are taken from it instead of from
    sign(y).  Default (force=None) is the reference arithmetic.  Tests use forcing to compare gradients under
    IDENTICAL ReLU decisions: an independent fp32 implementation flips a decision wherever |pre-activation| is
    below rounding noise, which moves single-pixel gradient contributions between implementations."""
    if force is not None and name in force:
        return y * force[name]
    return F.relu(y)


def _conv(sd, key, x, relu, force=None):
    y = F.conv2d(x, sd[key + ".weight"], sd[key + ".bias"], padding="same")
    return _relu(y, force, key.rsplit(".", 1)[1]) if relu else y


def _up2(x):
    # tvtrans.Resize((2h,2w), BILINEAR) on a tensor dispatches to the antialias bilinear kernel
    # (blocks.py:260,269); for 2x UPsampling its weights equal plain bilinear
    # (out[2i] = .25 in[i-1] + .75 in[i], out[2i+1] = .75 in[i] + .25 in[i+1], edges clamped),
    # evaluated separably (W pass then H pass).  antialias=True keeps the oracle on the same ATen op.
    return F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False, antialias=True)


def shallow_unet(sd: Dict[str, Tensor], x: Tensor, prefix: str = "encoder.shallow_unet.", force=None) -> Tensor:
    """blocks.py:278-308.  ReLU after every conv except c7 and c10; c13 (1x1) keeps its ReLU (Q9)."""
    c = lambda i, v, r=True: _conv(sd, prefix + "c%d" % i, v, r, force)
    x1 = c(2, c(1, x))
    x2 = c(4, c(3, F.max_pool2d(x1, 2)))
    v = c(6, c(5, F.max_pool2d(x2, 2)))
    v = c(7, _up2(v), False)
    v = c(9, c(8, torch.cat([v, x2], 1)))
    v = c(10, _up2(v), False)
    v = c(12, c(11, torch.cat([v, x1], 1)))
    return c(13, v)


def deep_unet(sd: Dict[str, Tensor], x: Tensor, prefix: str = "encoder.unet.", force=None) -> Tensor:
    """blocks.py:172-237.  No ReLU after c9, c12, c15, c18."""
    c = lambda i, v, r=True: _conv(sd, prefix + "c%d" % i, v, r, force)
    x1 = c(2, c(1, x))
    x2 = c(4, c(3, F.max_pool2d(x1, 2)))
    x3 = c(6, c(5, F.max_pool2d(x2, 2)))
    v = c(8, c(7, F.max_pool2d(x3, 2)))
    v = c(9, _up2(v), False)
    v = c(11, c(10, torch.cat([v, x3], 1)))
    v = c(12, _up2(v), False)
    v = c(14, c(13, torch.cat([v, x2], 1)))
    v = c(15, _up2(v), False)
    v = c(17, c(16, torch.cat([v, x1], 1)))
    return c(18, v, False)


def encoder(sd: Dict[str, Tensor], frames: Tensor, spec: TaskSpec, force=None):
    """blocks.py:77-103.  frames [N,3,H,H] -> (enc_pos [N,2n], enc_masks [N,n+1,H,H], masked_objs list)."""
    n, H = spec.n_objs, spec.H
    logits = shallow_unet(sd, frames, force=force) if H < 40 else deep_unet(sd, frames, force=force)
    logits = torch.cat([logits, torch.ones_like(logits[:, :1])], 1)
    masks = torch.softmax(logits, 1)
    masked = [masks[:, o:o + 1] * frames for o in range(n)]
    a = torch.cat(masked, 0)                                   # object-major along batch
    if H >= 40:
        a = F.avg_pool2d(a, 2)
    a = a.reshape(a.shape[0], -1)
    a = _relu(F.linear(a, sd["encoder.l1.weight"], sd["encoder.l1.bias"]), force, "l1")
    a = _relu(F.linear(a, sd["encoder.l2.weight"], sd["encoder.l2.bias"]), force, "l2")
    a = F.linear(a, sd["encoder.l3.weight"], sd["encoder.l3.bias"])
    a = torch.cat(torch.split(a, a.shape[0] // n, 0), 1)       # [N, 2n] = x0,y0,x1,y1,...
    return torch.tanh(a) * (H / 2) + (H / 2), masks, masked


def velocity(sd: Dict[str, Tensor], enc_in: Tensor, spec: TaskSpec, alt_vel: bool = False) -> Tensor:
    """blocks.py:31-49.  enc_in [B,in,2n] -> vel [B,2n]."""
    n, B, steps = spec.n_objs, enc_in.shape[0], spec.input_steps
    if alt_vel:
        d = enc_in[:, 1:] - enc_in[:, :-1]                                           # [B,in-1,2n]
        d = torch.cat(torch.chunk(d, n, 2), 0).reshape(n * B, (steps - 1) * 2)
        v = F.linear(d, sd["velocity_encoder.init_vel_linear.weight"], sd["velocity_encoder.init_vel_linear.bias"])
    else:
        h = torch.cat(torch.chunk(enc_in, n, 2), 0).reshape(n * B, steps * 2)
        p = "velocity_encoder.init_vel_mlp."
        h = torch.tanh(F.linear(h, sd[p + "0.weight"], sd[p + "0.bias"]))
        h = torch.tanh(F.linear(h, sd[p + "2.weight"], sd[p + "2.bias"]))
        v = F.linear(h, sd[p + "4.weight"], sd[p + "4.bias"])
    return torch.cat(torch.chunk(v, n, 0), 1)


# ---------------------------------------------------------------------------
# ODE cells (cells.py).  Scalars keep the reference dtypes: dt fp32 0-dim, k/equil/g/m fp64
# 0-dim, so type promotion (0-dim fp64 with dimensioned fp32 -> fp32) is the reference's.
# ---------------------------------------------------------------------------
def cell_spring(pos: Tensor, vel: Tensor, dt: Tensor, k: Tensor, equil: Tensor):
    """cells.py:31-51.  split(...,1,dim=1) yields single COLUMNS: the spring acts between
    column 0 and column 1 only; columns 2.. pass through."""
    p = list(torch.split(pos, 1, 1))
    v = list(torch.split(vel, 1, 1))
    for _ in range(5):
        diff = p[0] - p[1]
        norm = torch.sqrt(torch.abs(torch.sum(diff ** 2, -1, keepdim=True)))
        direction = diff / (norm + 1e-4)
        force = torch.exp(k) * (norm - 2 * torch.exp(equil)) * direction
        v[0] = v[0] - dt / 5 * force
        v[1] = v[1] + dt / 5 * force
        p[0] = p[0] + dt / 5 * v[0]
        p[1] = p[1] + dt / 5 * v[1]
    return torch.cat(p, 1), torch.cat(v, 1)


def cell_bouncing(pos: Tensor, vel: Tensor, dt: Tensor):
    """cells.py:60-83.  Moves / reflects columns 0 and 1 only (Q2); walls at 0 and 32, radius 2."""
    p = list(torch.split(pos, 1, 1))
    v = list(torch.split(vel, 1, 1))
    for _ in range(5):
        p[0] = p[0] + dt / 5 * v[0]
        p[1] = p[1] + dt / 5 * v[1]
        for j in range(2):
            v[j] = torch.where(p[j] + 2 > 32, -v[j], v[j])
            v[j] = torch.where(0.0 > p[j] - 2, -v[j], v[j])
            p[j] = torch.where(p[j] + 2 > 32, 32 - (p[j] + 2 - 32) - 2, p[j])
            p[j] = torch.where(0.0 > p[j] - 2, -(p[j] - 2) + 2, p[j])
    return torch.cat(p, 1), torch.cat(v, 1)


def cell_gravity(pos: Tensor, vel: Tensor, dt: Tensor, g: Tensor, m: Tensor):
    """cells.py:96-106 with A = exp(g)*exp(2m) recomputed on every call."""
    A = torch.exp(g) * torch.exp(2 * m)
    for _ in range(5):
        vecs = [pos[:, 0:2] - pos[:, 2:4], pos[:, 2:4] - pos[:, 4:6], pos[:, 4:6] - pos[:, 0:2]]
        norms = [torch.sqrt(torch.clamp(torch.sum(v_ ** 2, -1, keepdim=True), min=1e-1, max=1e5)) for v_ in vecs]
        f = [v_ / torch.pow(torch.clamp(nr, min=1, max=170), 3) for v_, nr in zip(vecs, norms)]
        f = [f[0] - f[2], f[1] - f[0], f[2] - f[1]]
        f = torch.cat([-A * q for q in f], 1)
        vel = vel + dt / 5 * f
        pos = pos + dt / 5 * vel
    return pos, vel


def rollout_cell(sd: Dict[str, Tensor], spec: TaskSpec, pos: Tensor, vel: Tensor):
    dt = sd["rollout_cell.dt"]
    if spec.cell == "spring":
        return cell_spring(pos, vel, dt, sd["rollout_cell.k"], sd["rollout_cell.equil"])
    if spec.cell == "bouncing":
        return cell_bouncing(pos, vel, dt)
    return cell_gravity(pos, vel, dt, sd["rollout_cell.g"], sd["rollout_cell.m"])


# ---------------------------------------------------------------------------
# Decoder
# ---------------------------------------------------------------------------
def var_from_net(sd: Dict[str, Tensor], prefix: str, shape) -> Tensor:
    """blocks.py:318-322: l2(tanh(l1(ones[1,10]))) reshaped."""
    h = torch.tanh(F.linear(torch.ones(1, 10), sd[prefix + ".l1.weight"], sd[prefix + ".l1.bias"]))
    return F.linear(h, sd[prefix + ".l2.weight"], sd[prefix + ".l2.bias"]).reshape(shape)


def learned_tensors(sd: Dict[str, Tensor], spec: TaskSpec):
    n, H = spec.n_objs, spec.H
    t = H // 2
    template = var_from_net(sd, "var_net_template", (n, 1, t, t))
    contents = var_from_net(sd, "var_net_content", (n, 3, t, t))
    background = var_from_net(sd, "var_net_background", (1, 3, H, H))
    return template, contents, background


def decoder(sd: Dict[str, Tensor], loc: Tensor, spec: TaskSpec, learned=None, extras: dict | None = None) -> Tensor:
    """physics_models.py:151-199 + stn.py:5-16.  loc [N,2n] -> frames [N,3,H,H].

    theta is fp64 (sigma is a numpy float64, Q5) so affine_grid runs in fp64 and the grid is
    cast to fp32 before grid_sample (bilinear, zeros padding, align_corners=False)."""
    n, H = spec.n_objs, spec.H
    t = H // 2
    N = loc.shape[0]
    template, contents, background = learned if learned is not None else learned_tensors(sd, spec)
    joint = torch.cat([template.repeat(1, 3, 1, 1) + 5, torch.sigmoid(contents)], 1)        # [n,6,t,t]
    one = torch.ones(N, dtype=torch.float64)
    zero = torch.zeros(N, dtype=torch.float64)
    sampled = []
    for o in range(n):
        lx, ly = loc[:, 2 * o], loc[:, 2 * o + 1]
        theta = torch.stack([one, zero, (H / 2 - lx) / t * 1.0, zero, one, (H / 2 - ly) / t * 1.0], 1)
        grid = F.affine_grid(theta.view(-1, 2, 3), torch.Size((N, 6, H, H)), align_corners=False)
        s = F.grid_sample(joint[o:o + 1].expand(N, -1, -1, -1).float(), grid.float(), mode="bilinear",
                          padding_mode="zeros", align_corners=False)
        sampled.append((s[:, :3], s[:, 3:]))
    bg = torch.sigmoid(background).expand(N, -1, -1, -1)
    logits = torch.stack([m - 5 for m, _ in sampled] + [torch.ones_like(sampled[0][0])], 1)
    w = torch.softmax(logits, 1)
    layers = [c for _, c in sampled] + [bg]
    out = sum(w[:, i] * layers[i] for i in range(n + 1))
    if extras is not None:
        extras.update(template=template, contents=contents, background_content=torch.sigmoid(background),
                      transf_masks=list(torch.unbind(w, 1)), transf_contents=layers)
    return out


# ---------------------------------------------------------------------------
# Whole step
# ---------------------------------------------------------------------------
def feedforward(sd: Dict[str, Tensor], x: Tensor, spec: TaskSpec, alt_vel: bool = False, force=None) -> Dict[str, Tensor]:
    """physics_models.py:204-245.  x [B,T,3,H,H]."""
    B, T = x.shape[0], x.shape[1]
    n, H, e = spec.n_objs, spec.H, spec.enc_steps
    assert T > e
    frames = x[:, :e].reshape(B * e, 3, H, H)
    enc_pos, masks, masked = encoder(sd, frames, spec, force)
    learned = learned_tensors(sd, spec)
    recons = decoder(sd, enc_pos, spec, learned).reshape(B, e, 3, H, H)
    enc_pos = enc_pos.reshape(B, e, 2 * n)
    if spec.input_steps > 1:
        vel = velocity(sd, enc_pos[:, :spec.input_steps], spec, alt_vel)
    else:
        vel = torch.zeros(B, 2 * n)
    pos = enc_pos[:, spec.input_steps - 1]
    seq, outs = [torch.cat([pos, vel], 1)], []
    for _ in range(T - spec.input_steps):
        pos, vel = rollout_cell(sd, spec, pos, vel)
        outs.append(decoder(sd, pos, spec, learned))
        seq.append(torch.cat([pos, vel], 1))
    return dict(output=torch.stack(outs, 1), recons_out=recons, enc_pos=enc_pos,
                pos_vel_seq=torch.stack(seq, 1), enc_masks=masks, masked_objs=masked,
                template=learned[0], contents=learned[1], background=learned[2])


def losses(x: Tensor, ff: Dict[str, Tensor], spec: TaskSpec, alpha: float) -> Dict[str, Tensor]:
    """physics_models.py:119-142 (un-aliased: ``pred`` here is the pure prediction loss; the
    reference's returned eval_losses[0] is ``train`` because of the in-place add, Q4)."""
    e, i, p = spec.enc_steps, spec.input_steps, spec.pred_steps
    recons = ((x[:, :e] - ff["recons_out"]) ** 2).sum((2, 3, 4)).mean()
    per = ((x[:, i:] - ff["output"]) ** 2).sum((2, 3, 4))
    pred, extrap = per[:, :p].mean(), per[:, p:].mean()
    train = pred + alpha * recons if alpha > 0.0 else pred
    return dict(train=train, pred=pred, extrap=extrap, recons=recons, per_frame_pred=per)


def live_step(sd: Dict[str, Tensor], x: Tensor, spec: TaskSpec, alpha: float, alt_vel: bool = False, force=None):
    """One LIVE-mode training step: forward, loss, backward.  Returns
    (feedforward dict, losses dict, grads dict) -- grads only for tensors that receive one."""
    leaves = {k: v.detach().clone().requires_grad_(v.is_floating_point() and k != "rollout_cell.dt"
                                                   and k != "rollout_cell.m")
              for k, v in sd.items()}
    ff = feedforward(leaves, x, spec, alt_vel, force)
    ls = losses(x, ff, spec, alpha)
    ls["train"].backward()
    grads = {k: v.grad for k, v in leaves.items() if v.grad is not None}
    return ff, ls, grads
