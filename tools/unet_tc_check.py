"""A/B check of the tcgen05 ShallowUNet forward (csrc/unet_tc.cu, PAIG_UNET_TC=1) against the FMA kernel
(csrc/unet_fused.cu): every saved activation, the logits and the encoder outputs of paig_step_forward on the same
input, plus the float64 oracle's logits as the yardstick, plus timing of the whole forward.

    python tools/unet_tc_check.py [B]         (run under `timeout`: a pipeline bug would hang the kernel)
"""
import ctypes
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import physicsnet_oracle as po          # noqa: E402  (tools/ is test infrastructure)
import backends                                      # noqa: E402
from paig_reproduction_b200 import _abi             # noqa: E402

byref = ctypes.byref


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    task = sys.argv[2] if len(sys.argv) > 2 else "spring_color"
    be = backends.get("cuda")
    spec = po.TASKS[task]
    T = spec.seq_len
    sd = po.init_state_dict(spec, 0, False)
    x = po.synthetic_frames(spec, B, T, 0)
    n, H, e, steps = spec.n_objs, spec.H, spec.enc_steps, T - spec.input_steps
    tk = be.make_task(spec, T, 3.0, False, 0)
    bufs = be.sd(sd)
    P = be.make_params(spec, bufs, False)
    xd = be.dev(x.numpy())
    N = B * e

    def run(tc):
        os.environ["PAIG_UNET_TC"] = "1" if tc else "0"
        ws = be.workspace(tk, B)
        ob = dict(output_seq=be.zeros((B, steps, 3, H, H)), recons_out=be.zeros((B, e, 3, H, H)),
                  enc_pos=be.zeros((B, e, 2 * n)), pos_vel_seq=be.zeros((B, steps + 1, 4 * n)),
                  enc_masks=be.zeros((B * e, n + 1, H, H)), masked_objs=be.zeros((n, B * e, 3, H, H)),
                  templates=be.zeros(n * (H // 2) ** 2 * 4 + 3 * H * H), losses=be.zeros(4))
        O = _abi.Outputs(*[ob[k].ptr for k in ("output_seq", "recons_out", "enc_pos", "pos_vel_seq", "enc_masks",
                                               "masked_objs", "templates", "losses")])
        be.check(be.lib.paig_step_forward(byref(tk), byref(P), xd.ptr, B, byref(O), ws.ptr, be.stream))
        w = ws.np()
        acts = {}
        for layer in range(13):
            view = (ctypes.c_long * 5)()
            be.check(be.lib.paig_debug_unet_conv_view(byref(tk), B, layer, byref(view)))
            off, bs, C, S, relu = [int(v) for v in view]
            acts["c%d" % (layer + 1)] = np.stack([w[off + f * bs: off + f * bs + C * S * S] for f in range(min(N, 40))]).reshape(-1, C, S, S)
        out = {k: v.np() for k, v in ob.items()}
        # timing of the whole forward
        be.lib.paig_profile_begin()
        for _ in range(5):
            be.check(be.lib.paig_step_forward(byref(tk), byref(P), xd.ptr, B, byref(O), ws.ptr, be.stream))
        buf = ctypes.create_string_buffer(1 << 16)
        be.lib.paig_profile_end(buf, len(buf))
        prof = buf.value.decode()
        return acts, out, prof

    a0, o0, p0 = run(False)
    a1, o1, p1 = run(True)
    res = {"B": B, "task": task}
    for k in a0:
        res[k] = rel(a1[k], a0[k])
    for k in ("enc_pos", "enc_masks", "output_seq", "losses"):
        res[k] = rel(o1[k], o0[k])
    # float64 oracle logits as the yardstick (first frames only)
    nf = min(N, 40)
    with torch.no_grad():
        sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
        torch.set_default_dtype(torch.float64)
        try:
            frames = x[:, :spec.enc_steps].reshape(-1, 3, H, H)[:nf].double()
            rec = {}
            orig = po._relu

            def spy(y, f, name):
                out = orig(y, f, name)
                rec[name] = out.detach()
                return out
            po._relu = spy
            try:
                po.encoder(sd64, frames, spec)
            finally:
                po._relu = orig
        finally:
            torch.set_default_dtype(torch.float32)
    for name in sorted(rec):
        if name in a0 and rec[name].shape == a0[name][:nf].shape:
            ref = rec[name].numpy()
            res["f64/" + name] = {"fma": rel(a0[name][:nf], ref), "tc": rel(a1[name][:nf], ref),
                                  "tc_mean_signed": float(np.mean((a1[name][:nf] - ref)[ref > 1e-3] / ref[ref > 1e-3])),
                                  "fma_mean_signed": float(np.mean((a0[name][:nf] - ref)[ref > 1e-3] / ref[ref > 1e-3]))}
    if os.environ.get("COMPACT"):
        print("kappa", os.environ.get("PAIG_UNET_TC_KAPPA"), "max A/B diff", max(v for k, v in res.items() if isinstance(v, float)))
        for k in sorted((k for k in res if k.startswith("f64/")), key=lambda s: int(s[5:])):
            v = res[k]
            print("  %-8s err tc %.2e fma %.2e   mean signed tc %+.2e fma %+.2e" % (k, v["tc"], v["fma"], v["tc_mean_signed"], v["fma_mean_signed"]))
    else:
        print(json.dumps(res, indent=1))

    def pick(prof):
        return [ln for ln in prof.splitlines() if "unet" in ln or "pack" in ln]
    print("FMA :", pick(p0))
    print("TC  :", pick(p1))


if __name__ == "__main__":
    main()
