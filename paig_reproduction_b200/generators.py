"""Synthetic video datasets of the PhysicsNet tasks (reference nn/datasets/generators.py) on numpy alone.

The reference draws with ``skimage.draw.circle`` (removed from current scikit-image) + ``skimage.transform.resize`` and
plots with matplotlib; neither is needed here.  What IS kept exactly is everything that decides which sequences a seed
produces: the order of the ``np.random`` draws, the initial-condition formulas, the sub-stepped explicit-Euler physics and
the rejection rules -- so a given ``np.random.seed`` yields the same trajectories as the reference functions
(tests/test_generators.py runs the unmodified reference functions with this module's rasteriser substituted for
skimage and compares the ``.npz`` byte for byte).  Rendering: discs are rasterised at ``scale`` x the frame size and
box-averaged down (the reference's anti-aliased resize is a Gaussian + linear interpolation; both are smooth
anti-aliased discs, the byte values differ slightly).

Output format of every generator (what nn/datasets/iterators.py:49-69 loads): ``np.savez_compressed(dest, train_x=...,
valid_x=..., test_x=...)`` with uint8 ``[N, T, H, W, C]`` frames (the bouncing-trajectory set stores ``[N, T, 2]``
positions)."""
from __future__ import annotations

from itertools import combinations

import numpy as np

from .viz import gallery, save_image


# ------------------------------------------------------------------------------------------------ rasteriser
def disc_indices(r, c, radius, shape):
    """Pixels (rr, cc) with (y - r)^2 + (x - c)^2 < radius^2 inside `shape` -- skimage.draw.circle's contract."""
    r0, r1 = max(int(np.ceil(r - radius)), 0), min(int(np.floor(r + radius)), shape[0] - 1)
    c0, c1 = max(int(np.ceil(c - radius)), 0), min(int(np.floor(c + radius)), shape[1] - 1)
    if r0 > r1 or c0 > c1:
        return np.zeros(0, np.intp), np.zeros(0, np.intp)
    yy, xx = np.mgrid[r0:r1 + 1, c0:c1 + 1]
    inside = (yy - r) ** 2 + (xx - c) ** 2 < radius ** 2
    return yy[inside], xx[inside]


def downscale(frame, out_size):
    """Box-average an [h*s, w*s(, C)] frame down to [h, w(, C)] (anti-aliased resize for integer factors)."""
    frame = np.asarray(frame, dtype=np.float32)
    squeeze = frame.ndim == 2
    if squeeze:
        frame = frame[:, :, None]
    h, w = int(out_size[0]), int(out_size[1])
    sy, sx = frame.shape[0] // h, frame.shape[1] // w
    out = frame[:h * sy, :w * sx].reshape(h, sy, w, sx, frame.shape[2]).mean(axis=(1, 3))
    return out[:, :, 0] if squeeze else out


# ------------------------------------------------------------------------------------------------ wall / object tests
def compute_wall_collision(pos, vel, radius, img_size):
    """generators.py:48-61: reflect position and velocity at the four walls (y first, then x), in place."""
    for axis in (1, 0):
        if pos[axis] - radius <= 0:
            vel[axis] = -vel[axis]
            pos[axis] = -(pos[axis] - radius) + radius
        if pos[axis] + radius >= img_size[axis]:
            vel[axis] = -vel[axis]
            pos[axis] = img_size[axis] - (pos[axis] + radius - img_size[axis]) - radius
    return pos, vel


def verify_wall_collision(pos, vel, radius, img_size):
    """generators.py:64-73."""
    return bool(pos[1] - radius <= 0 or pos[1] + radius >= img_size[1] or pos[0] - radius <= 0 or
                pos[0] + radius >= img_size[0])


def verify_object_collision(poss, radius):
    """generators.py:76-80."""
    return any(np.linalg.norm(a - b) <= radius for a, b in combinations(poss, 2))


# ------------------------------------------------------------------------------------------------ shared plumbing
def _save_splits(dest, data, n_train, n_valid):
    np.savez_compressed(dest, train_x=data[:n_train], valid_x=data[n_train:n_train + n_valid], test_x=data[n_train + n_valid:])
    print("Saved to file %s" % dest)


def _save_samples(dest, sequences):
    """The reference's '<dest>_samples.jpg': the first 10 sequences as a gallery, one row per sequence."""
    frames = np.concatenate(sequences[:10] / 255)
    save_image(dest.split(".")[0] + "_samples.jpg", gallery(frames, ncols=sequences.shape[1]))


def _render_discs(poss, radius, scale, scaled, img_size, color, background=None):
    """One frame: object j as a disc of `radius` in channel 2-j (colour) or channel 0, over `background` or black."""
    frame = np.zeros(list(scaled) + [3 if color else 1], dtype=np.float32) if background is None else background
    for j, pos in enumerate(poss):
        rr, cc = disc_indices(int(pos[1] * scale), int(pos[0] * scale), radius * scale, scaled)
        frame[rr, cc, (2 - j) if color else 0] = 1.0
    return (downscale(frame, img_size) * 255).astype(np.uint8)


def _collect(generate_sequence, total):
    sequences = []
    for i in range(total):
        if i % 100 == 0:
            print(i)
        sequences.append(generate_sequence())
    return np.array(sequences, dtype=np.uint8)


# ------------------------------------------------------------------------------------------------ datasets
def generate_bouncing_ball_dataset(dest, train_set_size, valid_set_size, test_set_size, seq_len, box_size):
    """generators.py:9-45: positions only, one point bouncing in a box; seeds numpy itself (seed 0)."""
    np.random.seed(0)

    def trajectory():
        x = np.random.rand(2) * box_size
        speed = np.random.rand() + 1
        angle = np.random.rand() * 2 * np.pi
        v = np.array([speed * np.cos(angle), speed * np.sin(angle)])
        traj = []
        for _ in range(seq_len):
            traj.append(x)
            for axis in (0, 1):
                if x[axis] + v[axis] > box_size or x[axis] + v[axis] < 0.0:
                    v[axis] = -v[axis]
            x = x + v
        return traj

    data = np.array([trajectory() for _ in range(train_set_size + valid_set_size + test_set_size)])
    _save_splits(dest, data, train_set_size, valid_set_size)


def generate_falling_bouncing_ball_dataset(dest, train_set_size, valid_set_size, test_set_size, seq_len, img_size=None,
                                           radius=3, dt=0.30, g=9.8, vx0_max=0.0, vy0_max=0.0, ode_steps=10):
    """generators.py:149-240 (without the CIFAR background, which needs a download): one grey ball under gravity `g`
    bouncing off the walls."""
    img_size = [32, 32] if img_size is None else img_size
    scale = 10
    scaled = [img_size[0] * scale, img_size[1] * scale]

    def generate_sequence():
        pos = np.random.rand(2)
        pos[0] = radius + (img_size[0] - 2 * radius) * pos[0]
        pos[1] = radius + (img_size[1] - 2 * radius) * pos[1] * (1.0 if g == 0.0 else 0.5)
        angle = np.random.rand() * 2 * np.pi
        vel = np.array([np.cos(angle) * vx0_max, np.sin(angle) * vy0_max])
        seq = []
        for _ in range(seq_len):
            frame = np.zeros(scaled, dtype=np.float32)
            rr, cc = disc_indices(int(pos[1] * scale), int(pos[0] * scale), radius * scale, scaled)
            frame[rr, cc] = 1.0
            seq.append((downscale(frame, img_size)[:, :, None] * 255).astype(np.uint8))
            for _ in range(ode_steps):
                vel[1] = vel[1] + dt / ode_steps * g
                pos[1] = pos[1] + dt / ode_steps * vel[1]
                pos[0] = pos[0] + dt / ode_steps * vel[0]
                pos, vel = compute_wall_collision(pos, vel, radius, img_size)
        return seq

    sequences = _collect(generate_sequence, train_set_size + valid_set_size + test_set_size)
    _save_splits(dest, sequences, train_set_size, valid_set_size)
    _save_samples(dest, sequences)


def _spring_initial_state(radius, equil, img_size, vx0_max, vy0_max):
    """generators.py:277-293: centre of mass inside the frame, the pair on a random axis at 0.5-1.5 x equil."""
    cm = np.random.rand(2)
    cm[0] = radius + equil + (img_size[0] - 2 * (radius + equil)) * cm[0]
    cm[1] = radius + equil + (img_size[1] - 2 * (radius + equil)) * cm[1]
    angle = np.random.rand() * 2 * np.pi
    r = np.random.rand() + 0.5
    poss = np.array([[np.cos(angle) * equil * r + cm[0], np.sin(angle) * equil * r + cm[1]],
                     [np.cos(angle + np.pi) * equil * r + cm[0], np.sin(angle + np.pi) * equil * r + cm[1]]])
    angles = np.random.rand(2) * 2 * np.pi
    vels = np.array([[np.cos(a) * vx0_max, np.sin(a) * vy0_max] for a in angles])
    return poss, vels


def _spring_substeps(poss, vels, k, equil, dt, ode_steps, radius, img_size):
    """generators.py:322-335: `ode_steps` explicit-Euler substeps; True as soon as a ball touches a wall."""
    for _ in range(ode_steps):
        norm = np.linalg.norm(poss[0] - poss[1])
        F = k * (norm - 2 * equil) * ((poss[0] - poss[1]) / norm)
        vels[0] = vels[0] - dt / ode_steps * F
        vels[1] = vels[1] + dt / ode_steps * F
        poss = poss + dt / ode_steps * vels
        if verify_wall_collision(poss[0], vels[0], radius, img_size) or verify_wall_collision(poss[1], vels[1], radius, img_size):
            return poss, True
    return poss, False


def generate_spring_balls_dataset(dest, train_set_size, valid_set_size, test_set_size, seq_len, img_size=None, radius=3,
                                  dt=0.3, k=3, equil=5, vx0_max=0.0, vy0_max=0.0, color=False, ode_steps=10):
    """generators.py:243-365: two balls joined by a spring (spring_color / spring_color_half: radius 2, k 4, equil 6,
    vx/vy 8 or 4 -- the shipped file names, runners/torch_run_physics.py:55-62).  Sequences in which a ball would touch
    a wall are discarded and redrawn."""
    img_size = [32, 32] if img_size is None else img_size
    scale = 10
    scaled = [img_size[0] * scale, img_size[1] * scale]

    def generate_sequence():
        while True:
            poss, vels = _spring_initial_state(radius, equil, img_size, vx0_max, vy0_max)
            seq, collision = [], False
            for _ in range(seq_len):
                seq.append(_render_discs(poss, radius, scale, scaled, img_size, color))
                poss, collision = _spring_substeps(poss, vels, k, equil, dt, ode_steps, radius, img_size)
                if collision:
                    break
            if not collision:
                return seq

    sequences = _collect(generate_sequence, train_set_size + valid_set_size + test_set_size)
    _save_splits(dest, sequences, train_set_size, valid_set_size)
    _save_samples(dest, sequences)


def generate_3_body_problem_dataset(dest, train_set_size, valid_set_size, test_set_size, seq_len, img_size=None, radius=3,
                                    dt=0.3, g=9.8, m=1.0, vx0_max=0.0, vy0_max=0.0, color=False, ode_steps=10):
    """generators.py:517-652: three equal masses near the vertices of a triangle about the frame centre, tangential
    initial velocities (3bp_color: radius 2, g 60, m 1, dt 0.5, vx/vy 2).  Redrawn on wall contact or when two bodies
    come within radius + 1."""
    img_size = [32, 32] if img_size is None else img_size
    scale = 10
    scaled = [img_size[0] * scale, img_size[1] * scale]

    def generate_sequence():
        while True:
            np.random.rand(2)                                             # the reference draws a centre and discards it
            cm = np.array(img_size) / 2
            a1 = np.random.rand() * 2 * np.pi
            a2 = a1 + 2 * np.pi / 3 + (np.random.rand() - 0.5) / 2
            a3 = a1 + 4 * np.pi / 3 + (np.random.rand() - 0.5) / 2
            r = (np.random.rand() / 2 + 0.75) * img_size[0] / 4
            poss = np.array([[np.cos(a) * r + cm[0], np.sin(a) * r + cm[1]] for a in (a1, a2, a3)])
            turn = np.random.randint(0, 2) * 2 - 1
            noise = np.random.rand(2) - 0.5
            vels = np.array([[np.cos(a + turn * np.pi / 2) * vx0_max + noise[0], np.sin(a + turn * np.pi / 2) * vy0_max + noise[1]]
                             for a in (a1, a2, a3)])
            seq, collision = [], False
            for _ in range(seq_len):
                seq.append(_render_discs(poss, radius, scale, scaled, img_size, color))
                for _ in range(ode_steps):
                    v01, v12, v20 = poss[0] - poss[1], poss[1] - poss[2], poss[2] - poss[0]
                    f01, f12, f20 = (v / np.linalg.norm(v) ** 3 for v in (v01, v12, v20))
                    F = -g * m * m * np.array([f01 - f20, f12 - f01, f20 - f12])
                    vels = vels + dt / ode_steps * F
                    poss = poss + dt / ode_steps * vels
                    collision = any(verify_wall_collision(p, v, radius, img_size) for p, v in zip(poss, vels)) or \
                        verify_object_collision(poss, radius + 1)
                    if collision:
                        break
                if collision:
                    break
            if not collision:
                return seq

    sequences = _collect(generate_sequence, train_set_size + valid_set_size + test_set_size)
    _save_splits(dest, sequences, train_set_size, valid_set_size)
    _save_samples(dest, sequences)


def generate_spring_mnist_dataset(dest, train_set_size, valid_set_size, test_set_size, seq_len, digits, background=None,
                                  img_size=None, dt=0.3, k=3, equil=5, vx0_max=0.0, vy0_max=0.0, ode_steps=10):
    """generators.py:367-514 with the data it downloads passed in: `digits` = two [22, 22] glyphs in [0, 1] (the
    reference crops the first two MNIST training digits to 22 x 22), `background` = an RGB [h, w, 3] image in [0, 1] or
    None (the reference uses one CIFAR image).  Two glyphs of 'radius' 11 joined by a spring (mnist_spring_color: 64 x 64,
    k 2, equil 12); glyph j is painted into channel 2-j over the background."""
    img_size = [32, 32] if img_size is None else img_size
    scale, radius = 5, 11
    scaled = [img_size[0] * scale, img_size[1] * scale]
    glyphs = [np.kron(np.asarray(d, dtype=np.float32), np.ones((scale, scale), np.float32)) for d in digits]
    bg = None
    if background is not None:
        b = np.asarray(background, dtype=np.float32)
        ry, rx = scaled[0] // b.shape[0] + 1, scaled[1] // b.shape[1] + 1
        bg = np.clip(np.kron(b, np.ones((ry, rx, 1), np.float32))[:scaled[0], :scaled[1]] - 0.2, 0.0, 1.0)

    def render(poss):
        frame = np.zeros(scaled + [3], dtype=np.float32) if bg is None else bg.copy()
        half = 11 * scale
        for j, pos in enumerate(poss):
            cy, cx = int(pos[1] * scale), int(pos[0] * scale)
            y0, x0 = cy - half, cx - half
            ys, xs = slice(max(y0, 0), min(y0 + 2 * half, scaled[0])), slice(max(x0, 0), min(x0 + 2 * half, scaled[1]))
            patch = glyphs[j][ys.start - y0:ys.stop - y0, xs.start - x0:xs.stop - x0]
            frame[ys, xs, 2 - j] = np.maximum(frame[ys, xs, 2 - j], patch)
        return (downscale(frame, img_size) * 255).astype(np.uint8)

    def generate_sequence():
        while True:
            poss, vels = _spring_initial_state(radius, equil, img_size, vx0_max, vy0_max)
            seq, collision = [], False
            for _ in range(seq_len):
                seq.append(render(poss))
                poss, collision = _spring_substeps(poss, vels, k, equil, dt, ode_steps, radius, img_size)
                if collision:
                    break
            if not collision:
                return seq

    sequences = _collect(generate_sequence, train_set_size + valid_set_size + test_set_size)
    _save_splits(dest, sequences, train_set_size, valid_set_size)
    _save_samples(dest, sequences)
