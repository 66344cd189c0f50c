// Per-frame squared-error sums of compute_loss (physics_models.py:122-131) and their gradient, for the
// drop-in compute_loss() path (the fused step forms both inside the decoder kernel instead).
#include "common.cuh"
#include "internal.h"

namespace paig {

// sse[b, f] = sum_chw (x[b, first+f] - pred[b, f])^2 ; one CTA per frame, float4 loads, shuffle reduction
__global__ void __launch_bounds__(256) frame_sse_kernel(const float* __restrict__ x, long x_seq_stride, int first,
                                                        const float* __restrict__ pred, int F, int chw,
                                                        float* __restrict__ sse) {
    __shared__ float scratch[33];
    const int fr = blockIdx.x, b = fr / F, f = fr % F;
    const float4* xp = reinterpret_cast<const float4*>(x + (long)b * x_seq_stride + (long)(first + f) * chw);
    const float4* pp = reinterpret_cast<const float4*>(pred + (long)fr * chw);
    float s = 0.f;
    for (int i = threadIdx.x; i < chw / 4; i += blockDim.x) {
        const float4 a = xp[i], c = pp[i];
        const float d0 = a.x - c.x, d1 = a.y - c.y, d2 = a.z - c.z, d3 = a.w - c.w;
        s += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
    }
    s = block_sum(s, scratch);
    if (threadIdx.x == 0) sse[fr] = s;
}

// d_pred[b, f] = 2 * d_sse[b, f] * (pred[b, f] - x[b, first+f])
__global__ void __launch_bounds__(256) frame_sse_bwd_kernel(const float* __restrict__ x, long x_seq_stride, int first,
                                                            const float* __restrict__ pred, int F, int chw,
                                                            const float* __restrict__ d_sse, float* __restrict__ d_pred) {
    const int fr = blockIdx.x, b = fr / F, f = fr % F;
    const float g = 2.f * d_sse[fr];
    const float4* xp = reinterpret_cast<const float4*>(x + (long)b * x_seq_stride + (long)(first + f) * chw);
    const float4* pp = reinterpret_cast<const float4*>(pred + (long)fr * chw);
    float4* dp = reinterpret_cast<float4*>(d_pred + (long)fr * chw);
    for (int i = threadIdx.x; i < chw / 4; i += blockDim.x) {
        const float4 a = xp[i], c = pp[i];
        dp[i] = make_float4(g * (c.x - a.x), g * (c.y - a.y), g * (c.z - a.z), g * (c.w - a.w));
    }
}

}  // namespace paig

using namespace paig;

extern "C" {

int paig_frame_sse_forward(const float* x, long x_seq_stride, int first, const float* pred, int B, int F, int chw,
                           float* sse, void* stream) {
    if (B * F <= 0) return 0;
    if (chw % 4) {
        set_error("frame_sse: C*H*W must be a multiple of 4");
        return 1;
    }
    launch(frame_sse_kernel, dim3(B * F), dim3(256), 0, (cudaStream_t)stream, x, x_seq_stride, first, pred, F, chw, sse);
    return check_launch("frame_sse");
}

int paig_frame_sse_backward(const float* x, long x_seq_stride, int first, const float* pred, int B, int F, int chw,
                            const float* d_sse, float* d_pred, void* stream) {
    if (B * F <= 0) return 0;
    if (chw % 4) {
        set_error("frame_sse: C*H*W must be a multiple of 4");
        return 1;
    }
    launch(frame_sse_bwd_kernel, dim3(B * F), dim3(256), 0, (cudaStream_t)stream, x, x_seq_stride, first, pred, F, chw,
           d_sse, d_pred);
    return check_launch("frame_sse_bwd");
}

}  // extern "C"
