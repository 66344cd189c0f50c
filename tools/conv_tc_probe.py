"""Accuracy / timing probe of the tcgen05 implicit-GEMM convolution (csrc/conv_tc.cu) against float64 and against the
CUDA-core kernel.  Each shape runs in its own process under a time limit (a pipeline bug would hang the kernel)."""
import json
import subprocess
import sys

SHAPES = [  # frames, Cin, Cout, S, relu, transposed
    (4, 16, 16, 64, 1, 0), (3, 32, 32, 32, 1, 0), (7, 64, 64, 16, 1, 0), (5, 128, 128, 8, 0, 0), (6, 64, 128, 8, 1, 0),
    (4, 128, 32, 16, 0, 0), (3, 96, 64, 16, 1, 0), (2, 48, 16, 64, 1, 0), (3, 64, 96, 16, 0, 1), (3, 16, 48, 64, 0, 1),
    (1000, 64, 64, 16, 1, 0), (1000, 32, 32, 32, 1, 0), (1000, 16, 16, 64, 1, 0), (1000, 128, 128, 8, 1, 0),
    (1000, 16, 16, 32, 0, 0),       # ShallowUNet c10 (16 -> 16 at 32 px): the tensor-core experiment DESIGN.md records
]


def one(shape):
    import torch
    import torch.nn.functional as F
    sys.path.insert(0, ".")
    from paig_reproduction_b200 import _lib
    lib = _lib.load()
    N, Cin, Cout, S, relu, tr = shape
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.rand(N, Cin, S, S, device="cuda", generator=g) + 0.05            # all positive: a truncation bias would show
    wshape = (Cin, Cout, 3, 3) if tr else (Cout, Cin, 3, 3)
    w = (torch.rand(wshape, device="cuda", generator=g) + 0.05) / (9 * Cin)
    b = torch.rand(Cout, device="cuda", generator=g) * 0.1
    y = torch.full((N, Cout, S, S), 7.0, device="cuda")
    scratch = torch.empty(9 * Cin * Cout + 64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.paig_debug_conv3x3_tc(x.data_ptr(), w.data_ptr(), None if tr else b.data_ptr(), y.data_ptr(), N, Cin, Cout, S,
                                         relu, tr, scratch.data_ptr(), st), "conv3x3_tc")
    torch.cuda.synchronize()
    n_ref = min(N, 8)
    xr, yr = x[:n_ref].double(), y[:n_ref].double()
    if tr:
        ref = F.conv_transpose2d(xr, w.double(), padding=1)
    else:
        ref = F.conv2d(xr, w.double(), b.double(), padding=1)
    if relu:
        ref = F.relu(ref)
    err = (yr - ref).abs().max().item() / ref.abs().max().item()
    bias = ((yr - ref) / ref.clamp_min(1e-30)).mean().item()
    yl = torch.empty_like(y)
    res = {"shape": shape, "max_rel_err": err, "mean_signed_rel_err": bias}
    if not tr:
        _lib.check(lib.paig_conv3x3_forward(x.data_ptr(), w.data_ptr(), b.data_ptr(), yl.data_ptr(), N, Cin, Cout, S, relu, st))
        torch.cuda.synchronize()
        res["fma_max_rel_err"] = (yl[:n_ref].double() - ref).abs().max().item() / ref.abs().max().item()
        res["fma_mean_signed_rel_err"] = ((yl[:n_ref].double() - ref) / ref.clamp_min(1e-30)).mean().item()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for name, fn in (("tc_us", lambda: lib.paig_debug_conv3x3_tc(x.data_ptr(), w.data_ptr(), None if tr else b.data_ptr(), y.data_ptr(),
                                                                  N, Cin, Cout, S, relu, tr, scratch.data_ptr(), st)),
                     ("fma_us", None if tr else (lambda: lib.paig_conv3x3_forward(x.data_ptr(), w.data_ptr(), b.data_ptr(), yl.data_ptr(),
                                                                                  N, Cin, Cout, S, relu, st)))):
        if fn is None or N < 100:
            continue
        for _ in range(3):
            fn()
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 100
        res[name] = us
        res[name.replace("_us", "_tflops")] = 2.0 * 9 * Cin * Cout * S * S * N / us * 1e-6
    print(json.dumps(res), flush=True)


BIG = [s for s in SHAPES if s[0] >= 1000]

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "big":
        SHAPES, sys.argv = BIG, sys.argv[:1]
    if len(sys.argv) > 1:
        one(tuple(json.loads(sys.argv[1])))
    else:
        for s in SHAPES:
            try:
                r = subprocess.run([sys.executable, __file__, json.dumps(s)], capture_output=True, text=True, timeout=90)
                out = r.stdout.strip().splitlines()
                print(out[-1] if out and r.returncode == 0 else json.dumps({"shape": s, "rc": r.returncode, "err": (r.stderr or r.stdout)[-600:]}),
                      flush=True)
            except subprocess.TimeoutExpired:
                print(json.dumps({"shape": s, "hang": True}), flush=True)
                break
