// The two steps either side of the hot path in the reference's training loop:
//   * optimizer.step() over one flat parameter / gradient buffer (base.py:12-17 table: adam, rmsprop, momentum, sgd
//     with torch.optim defaults; base.py:152) -- one launch instead of ~920 ATen calls;
//   * get_batch (physics_models.py:113-117 -> iterators.py:26-40) with the uint8 dataset resident on the device:
//     gather the shuffled sequences and apply iterators.py:64's `astype(float32) / 255` in the same pass.  The
//     reference's [N,T,H,W,C] -> [N,T,C,H,W] step is a reshape (a reinterpretation), so the gather is a
//     flat per-sequence copy.
#include "common.cuh"
#include "internal.h"

#include <cstdint>

namespace paig {

enum { OPT_SGD = 0, OPT_MOMENTUM = 1, OPT_RMSPROP = 2, OPT_ADAM = 3 };

// torch.optim semantics (defaults): SGD; SGD(momentum=0.9): buf = g on the first step, then 0.9 buf + g;
// RMSprop(alpha=0.99, eps=1e-8): v = a v + (1-a) g^2, p -= lr g / (sqrt(v) + eps);
// Adam(betas=(0.9,0.999), eps=1e-8): m, v moments with bias corrections c1 = 1-b1^t, c2 = 1-b2^t,
//                                    p -= (lr / c1) * m / (sqrt(v) / sqrt(c2) + eps).
template <typename T>
__global__ void __launch_bounds__(256) optimizer_step_kernel(int kind, T* __restrict__ p, const T* __restrict__ g,
                                                             T* __restrict__ s0, T* __restrict__ s1, long n, T lr, int step,
                                                             T c1, T sqrt_c2) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const T gi = g[i];
    T pi = p[i];
    if (kind == OPT_SGD) {
        pi -= lr * gi;
    } else if (kind == OPT_MOMENTUM) {
        const T b = step <= 1 ? gi : (T)0.9 * s0[i] + gi;
        s0[i] = b;
        pi -= lr * b;
    } else if (kind == OPT_RMSPROP) {
        const T v = (T)0.99 * s0[i] + ((T)1 - (T)0.99) * gi * gi;
        s0[i] = v;
        pi -= lr * (gi / (sqrt(v) + (T)1e-8));
    } else {
        const T m = (T)0.9 * s0[i] + ((T)1 - (T)0.9) * gi;
        const T v = (T)0.999 * s1[i] + ((T)1 - (T)0.999) * gi * gi;
        s0[i] = m;
        s1[i] = v;
        pi -= (lr / c1) * (m / (sqrt(v) / sqrt_c2 + (T)1e-8));
    }
    p[i] = pi;
}

// out[b][j] = data[idx[b]][j] / 255   (float32 division, as numpy's astype(float32) / 255)
__global__ void __launch_bounds__(256) gather_batch_u8_kernel(const uint8_t* __restrict__ data, long seq_elems,
                                                              const long* __restrict__ idx, int B, float* __restrict__ out) {
    const long q4 = seq_elems / 4;                                      // 4 bytes -> 4 floats per thread
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long)B * q4) return;
    const int b = (int)(t / q4);
    const long j = (t % q4) * 4;
    const uchar4 u = *reinterpret_cast<const uchar4*>(data + idx[b] * seq_elems + j);
    float4 o;
    o.x = __fdiv_rn((float)u.x, 255.f); o.y = __fdiv_rn((float)u.y, 255.f);
    o.z = __fdiv_rn((float)u.z, 255.f); o.w = __fdiv_rn((float)u.w, 255.f);
    *reinterpret_cast<float4*>(out + (long)b * seq_elems + j) = o;
}

}  // namespace paig

using namespace paig;

extern "C" {

int paig_optimizer_step(int kind, float* params, const float* grads, float* state0, float* state1, long n, float lr,
                        int step, void* stream) {
    if (kind < OPT_SGD || kind > OPT_ADAM || n < 0 || step < 1) {
        set_error("optimizer_step: bad kind %d / n %ld / step %d", kind, n, step);
        return 1;
    }
    if (n == 0) return 0;
    const double c1 = 1.0 - pow(0.9, (double)step), c2 = 1.0 - pow(0.999, (double)step);
    launch(optimizer_step_kernel<float>, dim3(cdiv(n, 256)), dim3(256), 0, (cudaStream_t)stream, kind, params, grads, state0,
           state1, n, lr, step, (float)c1, (float)sqrt(c2));
    return check_launch("optimizer_step");
}

int paig_optimizer_step_f64(int kind, double* params, const double* grads, double* state0, double* state1, long n,
                            double lr, int step, void* stream) {
    if (kind < OPT_SGD || kind > OPT_ADAM || n < 0 || step < 1) {
        set_error("optimizer_step_f64: bad kind %d / n %ld / step %d", kind, n, step);
        return 1;
    }
    if (n == 0) return 0;
    const double c1 = 1.0 - pow(0.9, (double)step), c2 = 1.0 - pow(0.999, (double)step);
    launch(optimizer_step_kernel<double>, dim3(cdiv(n, 256)), dim3(256), 0, (cudaStream_t)stream, kind, params, grads,
           state0, state1, n, lr, step, c1, sqrt(c2));
    return check_launch("optimizer_step_f64");
}

int paig_gather_batch_u8(const uint8_t* data, long seq_elems, const long* idx, int B, float* out, void* stream) {
    if (seq_elems <= 0 || (seq_elems % 4) != 0 || ((uintptr_t)data % 4) != 0 || ((uintptr_t)out % 16) != 0) {
        set_error("gather_batch_u8: sequences of %ld bytes must be a multiple of 4 and 4-byte aligned", seq_elems);
        return 1;
    }
    if (B <= 0) return 0;
    launch(gather_batch_u8_kernel, dim3(cdiv((long)B * (seq_elems / 4), 256)), dim3(256), 0, (cudaStream_t)stream, data,
           seq_elems, idx, B, out);
    return check_launch("gather_batch_u8");
}

}  // extern "C"
