"""How far, and how COHERENTLY, do the encoder's forward results sit from the exact (float64) ones?  For one task / batch
this runs paig_step_forward and compares the UNet logits, encoder.l1 pre-activations' effect (H1) and the encoded
positions with the oracle evaluated in float64 -- and, for scale, does the same for the oracle in float32 (the
reference's own arithmetic).  A coherent component (the slope of error against value, the mean signed relative error)
is what a long network amplifies; incoherent rounding noise averages out.
    python tools/logit_bias.py mnist_spring_color 16      (run under PAIG_NO_CONV_TC=1, PAIG_CONV_TC_NOCOMP=1, ... to compare)"""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import backends
from oracle import physicsnet_oracle as po
from paig_reproduction_b200 import _abi

task, B = sys.argv[1], int(sys.argv[2])
spec = po.TASKS[task]
sd = po.init_state_dict(spec, 0)
x = po.synthetic_frames(spec, B, spec.seq_len, 0)
frames = x[:, :spec.enc_steps].reshape(-1, 3, spec.H, spec.H)
unet = po.deep_unet if spec.H >= 40 else po.shallow_unet
with torch.no_grad():
    z32 = unet(sd, frames)
    pos32 = po.encoder(sd, frames, spec)[0]
    torch.set_default_dtype(torch.float64)
    sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
    z64 = unet(sd64, frames.double())
    pos64 = po.encoder(sd64, frames.double(), spec)[0]
    torch.set_default_dtype(torch.float32)
be = backends.get("cuda")
tk = be.make_task(spec, spec.seq_len, 3.0)
bufs = be.sd(sd)
P = be.make_params(spec, bufs)
ws = be.workspace(tk, B)
xd = be.dev(x.numpy())
n, H, e, steps = spec.n_objs, spec.H, spec.enc_steps, spec.seq_len - spec.input_steps
ob = dict(enc_pos=be.zeros((B, e, 2 * n)), losses=be.zeros(4))
O = _abi.Outputs(None, None, ob["enc_pos"].ptr, None, None, None, None, ob["losses"].ptr)
be.check(be.lib.paig_step_forward(ctypes.byref(tk), ctypes.byref(P), xd.ptr, B, ctypes.byref(O), ws.ptr, be.stream))
off = be.lib.paig_debug_workspace_offset(ctypes.byref(tk), B, b"logits", 0)
N = B * e
z = torch.from_numpy(ws.np()[off:off + N * n * H * H].reshape(N, n, H, H).copy()).double()
pos = torch.from_numpy(ob["enc_pos"].np().reshape(N, 2 * n)).double()


def stats(a, ref):
    d = (a - ref).reshape(-1)
    r = ref.reshape(-1)
    big = r.abs() > 0.05 * r.abs().max()
    slope = float((d * r).sum() / (r * r).sum())                 # least-squares scale error: a ~ (1 + slope) ref
    return {"max_rel": float(d.abs().max() / r.abs().max()), "scale_error": slope,
            "mean_signed_rel": float((d[big] / r[big]).mean()), "rms_rel": float((d[big] / r[big]).pow(2).mean().sqrt())}


env = {k: v for k, v in os.environ.items() if k.startswith("PAIG_")}
print(json.dumps({"task": task, "B": B, "env": env,
                  "logits_cuda_vs_f64": stats(z, z64), "logits_ref32_vs_f64": stats(z32.double(), z64),
                  "pos_cuda_vs_f64": stats(pos, pos64), "pos_ref32_vs_f64": stats(pos32.double(), pos64)}), flush=True)
