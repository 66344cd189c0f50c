"""Accuracy / timing probe of the tcgen05 weight-gradient kernel (csrc/wgrad_tc.cu) against float64 and against the
TMA / mma.sync kernels it replaces (PAIG_NO_WGRAD_TC=1).  Each path runs in its own process under a time limit (a pipeline bug
would hang the kernel; the last shape printed tells which).

    python tools/wgrad_tc_probe.py            # all shapes
    python tools/wgrad_tc_probe.py small      # correctness shapes only"""
import json
import os
import subprocess
import sys

SMALL = [  # frames, Cin, Cout, S
    (3, 32, 32, 32), (5, 64, 32, 32), (2, 32, 32, 64), (7, 64, 64, 16), (6, 32, 64, 16), (5, 96, 64, 16), (9, 128, 128, 8),
    (8, 64, 128, 8), (6, 128, 32, 16), (40, 32, 32, 8), (37, 64, 64, 32),
]
BIG = [  # the 64-px UNet's layers at B = 100 (1000 frames), and ShallowUNet c6
    (1000, 32, 32, 32), (1000, 64, 32, 32), (1000, 32, 32, 64), (1000, 64, 64, 16), (1000, 32, 64, 16), (1000, 96, 64, 16),
    (1000, 128, 128, 8), (1000, 64, 128, 8), (1000, 128, 32, 16), (1000, 32, 32, 8),
]


def one(shape):
    import torch
    import torch.nn.functional as F
    sys.path.insert(0, ".")
    from paig_reproduction_b200 import _lib
    lib = _lib.load()
    N, Cin, Cout, S = shape
    g = torch.Generator(device="cuda").manual_seed(S * 1000 + Cin * 10 + Cout)
    x = torch.randn(N, Cin, S, S, device="cuda", generator=g)
    x = torch.relu(x) + 0.01 * x                                  # mostly positive, like post-ReLU activations
    dy = torch.randn(N, Cout, S, S, device="cuda", generator=g) + 0.3          # a coherent part: truncation bias would show
    w = torch.zeros(Cout, Cin, 3, 3, device="cuda")
    dw = torch.full((Cout, Cin, 3, 3), 3.0, device="cuda")
    db = torch.full((Cout,), 3.0, device="cuda")
    ws = torch.zeros(296 * (Cout * Cin * 9 + Cout) + 64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream

    def run():
        return lib.paig_conv3x3_backward(x.data_ptr(), w.data_ptr(), None, dy.data_ptr(), None, dw.data_ptr(), db.data_ptr(),
                                         N, Cin, Cout, S, 0, ws.data_ptr(), st)
    _lib.check(run(), "conv3x3_backward")
    torch.cuda.synchronize()
    # float64 reference in chunks of frames
    rw = torch.zeros(Cout, Cin, 3, 3, device="cuda", dtype=torch.float64)
    rb = torch.zeros(Cout, device="cuda", dtype=torch.float64)
    for i in range(0, N, 50):
        xs, ds = x[i:i + 50].double(), dy[i:i + 50].double()
        rw += torch.nn.grad.conv2d_weight(xs, (Cout, Cin, 3, 3), ds, padding=1)
        rb += ds.sum((0, 2, 3))
    res = {"shape": shape, "tc": os.environ.get("PAIG_NO_WGRAD_TC") is None,
           "dw_max_rel_err": ((dw.double() - rw).abs().max() / rw.abs().max()).item(),
           "db_max_rel_err": ((db.double() - rb).abs().max() / rb.abs().max()).item(),
           "dw_mean_signed_rel_err": ((dw.double() - rw) / rw.abs().clamp_min(1e-30) * rw.sign()).mean().item()}
    if N >= 100:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            run()
        e0.record()
        for _ in range(10):
            run()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 100.0
        res["us"] = us
        res["tflops"] = 2.0 * N * S * S * 9 * Cin * Cout / (us * 1e-6) / 1e12
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    # one process per path (a hang shows as the last shape printed before the time limit)
    if len(sys.argv) > 1 and sys.argv[1] == "--run":
        for sh in json.loads(sys.argv[2]):
            one(tuple(sh))
        sys.exit(0)
    small_only = len(sys.argv) > 1 and sys.argv[1] == "small"
    jobs = [(SMALL, False)] if small_only else [(SMALL + BIG, False), (BIG, True)]
    if len(sys.argv) > 1 and sys.argv[1] == "big":
        jobs = [(BIG, False)]
    for shapes, legacy in jobs:
        env = dict(os.environ)
        if legacy:
            env["PAIG_NO_WGRAD_TC"] = "1"
        try:
            r = subprocess.run([sys.executable, __file__, "--run", json.dumps(shapes)], env=env, capture_output=True, text=True,
                               timeout=240)
            print(r.stdout.strip(), flush=True)
            if r.returncode:
                print(json.dumps({"legacy": legacy, "rc": r.returncode, "err": r.stderr[-600:]}), flush=True)
        except subprocess.TimeoutExpired as e:
            print((e.stdout or b"").decode() if isinstance(e.stdout, bytes) else (e.stdout or ""), flush=True)
            print(json.dumps({"legacy": legacy, "hang": True}), flush=True)
