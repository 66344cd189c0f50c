"""Accuracy / timing probe of the tcgen05 3xTF32 GEMM (csrc/gemm_tc.cu) against float64."""
import sys, time, torch
sys.path.insert(0, '.')
from paig_reproduction_b200 import _lib
lib = _lib.load()
torch.manual_seed(0)
for (M, N, K, fixed) in [(2000, 200, 3072, 1), (2000, 3072, 200, 1), (200, 3072, 2000, 0), (512, 256, 256, 1)]:
    A = torch.randn(M, K, device="cuda"); B = torch.randn(N, K, device="cuda") * 0.1
    C = torch.zeros(M, N, device="cuda"); ws = torch.zeros(16 * M * N, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.paig_debug_gemm_tc(A.data_ptr(), B.data_ptr(), C.data_ptr(), M, N, K, fixed, ws.data_ptr(), ws.numel(), st))
    torch.cuda.synchronize()
    ref = A.double() @ B.double().t()
    err = (C.double() - ref).abs().max().item() / ref.abs().max().item()
    f32 = ((A @ B.t()).double() - ref).abs().max().item() / ref.abs().max().item()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        lib.paig_debug_gemm_tc(A.data_ptr(), B.data_ptr(), C.data_ptr(), M, N, K, fixed, ws.data_ptr(), ws.numel(), st)
    e1.record(); torch.cuda.synchronize()
    print("M=%d N=%d K=%d: rel err %.2e (torch fp32 matmul %.2e)  %.1f us" % (M, N, K, err, f32, e0.elapsed_time(e1) * 50), flush=True)
