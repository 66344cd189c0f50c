"""Every output element of the tcgen05 convolution at full batch (1000 frames: several super-tiles per persistent CTA)
against cuDNN's fp32 convolution (TF32 off) -- the probe's float64 check only covers the first frames."""
import json, sys
sys.path.insert(0, ".")
import torch
import torch.nn.functional as F
from paig_reproduction_b200 import _lib
lib = _lib.load()
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
SHAPES = [(1000, 64, 64, 16, 1, 0), (1000, 32, 32, 32, 1, 0), (1000, 32, 32, 64, 1, 0), (1000, 128, 128, 8, 1, 0), (999, 64, 128, 8, 1, 0),
          (1000, 64, 96, 16, 0, 1), (1000, 32, 64, 16, 0, 1), (1000, 16, 48, 64, 0, 1), (1000, 128, 64, 8, 0, 1), (1000, 32, 32, 64, 0, 1),
          (1000, 96, 64, 16, 1, 0), (1000, 64, 32, 32, 0, 0), (1000, 128, 32, 16, 0, 0)]
for N, Cin, Cout, S, relu, tr in SHAPES:
    g = torch.Generator(device="cuda").manual_seed(N + Cin + S)
    x = torch.randn(N, Cin, S, S, device="cuda", generator=g)
    w = torch.randn((Cin, Cout, 3, 3) if tr else (Cout, Cin, 3, 3), device="cuda", generator=g) / (3 * Cin ** 0.5)
    b = torch.randn(Cout, device="cuda", generator=g) * 0.1
    y = torch.full((N, Cout, S, S), 7.0, device="cuda")
    scratch = torch.empty(9 * Cin * Cout + 64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    worst = 0.0
    for rep in range(3):                                     # repeated launches: barrier phases carry no state across them
        y.fill_(7.0)
        _lib.check(lib.paig_debug_conv3x3_tc(x.data_ptr(), w.data_ptr(), None if tr else b.data_ptr(), y.data_ptr(), N, Cin, Cout, S,
                                             relu, tr, scratch.data_ptr(), st), "conv3x3_tc")
        torch.cuda.synchronize()
        ref = F.conv_transpose2d(x, w, padding=1) if tr else F.conv2d(x, w, b, padding=1)
        if relu:
            ref = F.relu(ref)
        err = (y - ref).abs().amax(dim=(1, 2, 3)) / ref.abs().max()
        worst = max(worst, float(err.max()))
    bad = int((err > 2e-5).sum())
    print(json.dumps({"shape": [N, Cin, Cout, S, relu, tr], "max_rel_err_all_frames": worst, "frames_off": bad,
                      "worst_frame": int(err.argmax())}), flush=True)
