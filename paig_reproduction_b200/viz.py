"""Figures and side tensors the reference writes after every validation / test pass
(``PhysicsNet.visualize_sequence``, nn/network/physics_models.py:247-330; helpers nn/utils/viz.py:4-22 ``gallery`` and
:24-61 ``gif``).  The reference draws with matplotlib + moviepy; neither is a dependency here: images are laid out
with numpy exactly like ``gallery`` and encoded with Pillow when it is importable (JPEG / GIF), else skipped.  The data
artefact ``extra_outputs.npz`` is always written.  Host-side, off the hot path."""
from __future__ import annotations

import os

import numpy as np


def gallery(array: np.ndarray, ncols: int = 3) -> np.ndarray:
    """viz.py:4-22: [n, h, w, c] tiles with a 1-pixel 0.5-grey border, laid out row-major in `ncols` columns."""
    n, h, w, c = array.shape
    framed = np.full((n, h + 2, w + 2, c), 0.5, dtype=np.float64)
    framed[:, 1:-1, 1:-1, :] = array
    nrows = n // ncols
    assert n == nrows * ncols
    return framed.reshape(nrows, ncols, h + 2, w + 2, c).swapaxes(1, 2).reshape((h + 2) * nrows, (w + 2) * ncols, c)


def _to_u8(img: np.ndarray) -> np.ndarray:
    img = np.clip(np.asarray(img, dtype=np.float64), 0.0, 1.0)          # plt.Normalize(0, 1)
    if img.ndim == 3 and img.shape[-1] == 1:
        img = np.repeat(img, 3, axis=-1)                                  # Greys_r on a single channel
    return (img * 255.0 + 0.5).astype(np.uint8)


def save_image(path: str, img: np.ndarray, scale: int = 1) -> bool:
    try:
        from PIL import Image
    except ImportError:
        return False
    im = Image.fromarray(_to_u8(img))
    if scale != 1:
        im = im.resize((im.width * scale, im.height * scale), Image.NEAREST)
    im.save(path, quality=95)
    return True


def save_gif(path: str, frames: np.ndarray, fps: int = 7, scale: int = 3) -> bool:
    """viz.py:24-61: `frames` [T, h, w, 3] in 0..255."""
    try:
        from PIL import Image
    except ImportError:
        return False
    path = os.path.splitext(path)[0] + ".gif"
    ims = []
    for f in frames:
        im = Image.fromarray(np.clip(f, 0, 255).astype(np.uint8))
        ims.append(im.resize((im.width * scale, im.height * scale), Image.NEAREST))
    ims[0].save(path, save_all=True, append_images=ims[1:], duration=int(round(1000.0 / fps)), loop=0)
    return True


def _np(x):
    return x.detach().cpu().numpy() if hasattr(x, "detach") else np.asarray(x)


def visualize_sequence(net, logger=None):
    """physics_models.py:247-330, step for step.  Like the reference it pairs a FRESH test batch with the outputs of
    the model's LAST forward pass (``net.output`` / ``recons_out`` / ``pos_vel_seq``)."""
    batch_size = net.batch_size
    _feed, (batch_x, _) = net.get_batch(batch_size, net.test_iterator)
    batch_x = _np(batch_x)
    C, H = net.conv_ch, net.input_shape[1]
    output_seq, recons_seq = _np(net.output), _np(net.recons_out)
    pos_vel_seq = getattr(net, "pos_vel_seq", None)
    output_seq = np.concatenate([batch_x[:, :net.input_steps], output_seq], axis=1)
    recons_seq = np.concatenate([recons_seq, np.zeros((batch_size, net.extrap_steps) + recons_seq.shape[2:])], axis=1)
    i = 0
    for i in range(batch_x.shape[0]):
        if pos_vel_seq is not None and i in (0, 1) and logger is not None:
            logger.info(pos_vel_seq[i])
        total = np.concatenate([output_seq[i], batch_x[i], recons_seq[i]], axis=0)
        total = total.reshape([total.shape[0], H, H, C])              # a reshape, like the reference (not a permute)
        save_image(os.path.join(net.save_dir, "example%d.jpg" % i), gallery(total, ncols=batch_x.shape[1]))
    # animation of predictions over ground truth, sequences side by side (written once, for the last index)
    T = net.seq_len
    bordered_out = 0.5 * np.ones([batch_size, T, H + 2, H + 2, 3])
    bordered_x = 0.5 * np.ones([batch_size, T, H + 2, H + 2, 3])
    bordered_out[:, :, 1:-1, 1:-1] = output_seq.reshape([batch_size, T, H, H, C])
    bordered_x[:, :, 1:-1, 1:-1] = batch_x.reshape([batch_size, T, H, H, C])
    out_strip = np.concatenate(np.split(bordered_out, batch_size, 0), axis=-2).squeeze(0)
    x_strip = np.concatenate(np.split(bordered_x, batch_size, 0), axis=-2).squeeze(0)
    save_gif(os.path.join(net.save_dir, "animation%d.gif" % i), np.concatenate([out_strip, x_strip], axis=1) * 255,
             fps=7, scale=3)
    # extra tensors (physics_models.py:305-313)
    results = {"contents": _np(net.contents), "templates": _np(net.template),
               "background_content": _np(net.background_content),
               "transf_contents": np.stack([_np(c) for c in net.transf_contents]),
               "transf_masks": np.stack([_np(m) for m in net.transf_masks]),
               "enc_masks": _np(net.enc_masks), "masked_objs": np.stack([_np(m) for m in net.masked_objs])}
    np.savez_compressed(os.path.join(net.save_dir, "extra_outputs.npz"), **results)
    contents = np.swapaxes(results["contents"], 1, -1)
    templates = np.swapaxes(results["templates"], 1, -1)
    contents = 1 / (1 + np.exp(-contents))
    templates = 1 / (1 + np.exp(-(templates - 5)))
    if C == 1:
        contents = np.tile(contents, [1, 1, 1, 3])
    templates = np.tile(templates, [1, 1, 1, 3])
    save_image(os.path.join(net.save_dir, "templates.jpg"),
               gallery(np.concatenate([contents, templates], axis=0), ncols=net.n_objs), scale=4)
