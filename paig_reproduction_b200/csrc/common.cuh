// Shared helpers for the sm_100a kernels of libpaig_b200.so.
#pragma once

#ifdef PAIG_EMU
#include "emu_cuda.h"   // tests/emu: SIMT-on-CPU shim (test tool, never shipped)
#else
#include <cuda_runtime.h>
#endif

#include <cstdarg>
#include <cstdio>

#include "../../include/paig_b200.h"

namespace paig {

constexpr int kHidden = 200;      // VariableFromNetwork / encoder MLP width (blocks.py:314, physics_models.py:109)
constexpr int kVarIn = 10;        // VariableFromNetwork input ones[1,10] (blocks.py:319)
constexpr int kVelHidden = 100;   // VelocityEncoder MLP width (blocks.py:25-29)
constexpr int kMaxObjs = 3;

void set_error(const char* fmt, ...);
int check_launch(const char* what);
const char* layer_name(const char* family, int Cin, int Cout, int S);   // per-layer profile names (PAIG_PROFILE_LAYERS)

struct Dims {
    int n, H, t, e, in, pr, T, steps, HW, CHW;
};
inline Dims dims_of(const paig_task* k) {
    Dims d;
    d.n = k->n_objs;
    d.H = k->H;
    d.t = k->H / 2;
    d.in = k->input_steps;
    d.pr = k->pred_steps;
    d.e = d.in + d.pr;
    d.T = k->seq_len;
    d.steps = d.T - d.in;
    d.HW = d.H * d.H;
    d.CHW = 3 * d.HW;
    return d;
}

// launch accounting / optional per-launch CUDA-event timing (api.cu); used by bench.py for "gpu_launches" and for
// the live per-kernel durations behind its roofline line
extern long g_launch_count;
extern bool g_profiling;
void prof_before(cudaStream_t st);
void prof_after(cudaStream_t st);

// Kernel launch through a function pointer so the same call compiles for the GPU and for tests/emu.
template <typename... KArgs, typename... Args>
inline void launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
#ifdef PAIG_EMU
    (void)stream;
    emu::launch(grid, block, smem, [&]() { kern(KArgs(args)...); });
#else
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    ++g_launch_count;
    if (g_profiling) prof_before(stream);
    kern<<<grid, block, smem, stream>>>(KArgs(args)...);
    if (g_profiling) prof_after(stream);
#endif
}

#ifdef PAIG_EMU
#define PAIG_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(::emu::g_dyn_smem)
#else
#define PAIG_DYN_SMEM(type, name)                                  \
    extern __shared__ __align__(128) unsigned char _paig_dyn_smem[]; \
    type* name = reinterpret_cast<type*>(_paig_dyn_smem)
#endif

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum; result valid in every thread.  `scratch` holds >= 33 floats.  blockDim.x*y*z must be a
// multiple of 32.  Fixed order => deterministic.
__device__ __forceinline__ float block_sum(float v, float* scratch) {
    const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
    const int nw = (blockDim.x * blockDim.y * blockDim.z) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if ((tid & 31) == 0) scratch[tid >> 5] = v;
    __syncthreads();
    if (tid < 32) {
        float s = tid < nw ? scratch[tid] : 0.f;
        s = warp_sum(s);
        if (tid == 0) scratch[32] = s;
    }
    __syncthreads();
    return scratch[32];
}

// Stage `bytes` (multiple of 16, both pointers 16-byte aligned) of global memory into shared memory with one
// TMA bulk copy (cp.async.bulk -> SASS UBLKCP) completing on an mbarrier; every thread of the block calls this
// and returns once the data is visible.  `bar` is a block-shared 8-byte slot, `phase` the parity of this use.
__device__ __forceinline__ void stage_bulk(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* bar,
                                           unsigned phase) {
    const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
#if defined(PAIG_EMU)
    (void)bar;
    (void)phase;
    const int nt = blockDim.x * blockDim.y * blockDim.z;
    for (unsigned i = tid; i < bytes / 4; i += nt) ((float*)smem_dst)[i] = ((const float*)gmem_src)[i];
    __syncthreads();
#else
    const unsigned bar_a = (unsigned)__cvta_generic_to_shared(bar);
    const unsigned dst_a = (unsigned)__cvta_generic_to_shared(smem_dst);
    if (tid == 0) {
        if (phase == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes) : "memory");
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_a),
            "l"(__cvta_generic_to_global(gmem_src)), "r"(bytes), "r"(bar_a)
            : "memory");
    }
    __syncthreads();   // barrier initialised (phase 0) before anyone polls it
    unsigned done = 0;
    while (!done) {
        asm volatile(
            "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
            : "=r"(done)
            : "r"(bar_a), "r"(phase & 1u)
            : "memory");
    }
#endif
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

inline int cdiv(long a, long b) { return (int)((a + b - 1) / b); }

}  // namespace paig
