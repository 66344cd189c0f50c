// SIMT-on-CPU shim: lets the package's .cu kernel sources compile with g++ and run their
// thread blocks as cooperative fibers on one host thread.
//
// TEST TOOL ONLY.  It exists because the build container has no GPU: it checks kernel
// indexing / reduction / barrier logic against the oracle at tiny sizes before a kernel is
// spent on a B200 box.  The product package never loads the library built from this shim
// (paig_reproduction_b200/_lib.py only opens libpaig_b200.so and raises if it is missing);
// only tests/test_emu_*.py build and open it.
//
// Supported subset: threadIdx/blockIdx/blockDim/gridDim, static and dynamic __shared__,
// __syncthreads, __syncwarp, full-mask warp shuffles, atomicAdd (float/double/int/unsigned),
// float2/float4, the usual math intrinsics, and the handful of runtime calls the library uses
// (treated as host memory operations).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))
#define __restrict__ __restrict
#define __constant__ static

struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct float2 { float x, y; };
struct __attribute__((aligned(16))) float4 { float x, y, z, w; };
struct int2 { int x, y; };
struct __attribute__((aligned(4))) uchar4 { unsigned char x, y, z, w; };
struct __attribute__((aligned(16))) int4 { int x, y, z, w; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline int2 make_int2(int x, int y) { return int2{x, y}; }

typedef void* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3,
                      cudaMemcpyDefault = 4 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };

namespace emu {
extern uint3 g_threadIdx, g_blockIdx;
extern dim3 g_blockDim, g_gridDim;
extern unsigned char* g_dyn_smem;
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);
void syncthreads();
void syncwarp();
uint64_t warp_exchange(uint64_t v, int src_lane);   // every lane of the warp must call
int lane_id();
}  // namespace emu

#define threadIdx (::emu::g_threadIdx)
#define blockIdx (::emu::g_blockIdx)
#define blockDim (::emu::g_blockDim)
#define gridDim (::emu::g_gridDim)
#define warpSize 32

static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline void __syncthreads() { emu::syncthreads(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::syncwarp(); }

template <typename T>
static inline T emu_shfl(T v, int src) {
    static_assert(sizeof(T) <= 8, "shuffle payload");
    uint64_t raw = 0;
    memcpy(&raw, &v, sizeof(T));
    raw = emu::warp_exchange(raw, src);
    T out;
    memcpy(&out, &raw, sizeof(T));
    return out;
}
template <typename T>
static inline T __shfl_sync(unsigned, T v, int src, int width = 32) {
    int l = emu::lane_id();
    return emu_shfl(v, (l / width) * width + (src % width));
}
template <typename T>
static inline T __shfl_xor_sync(unsigned, T v, int m, int width = 32) {
    (void)width;
    return emu_shfl(v, emu::lane_id() ^ m);
}
template <typename T>
static inline T __shfl_down_sync(unsigned, T v, unsigned d, int width = 32) {
    int l = emu::lane_id();
    int s = l + (int)d;
    if ((s / width) != (l / width)) s = l;
    return emu_shfl(v, s);
}
template <typename T>
static inline T __shfl_up_sync(unsigned, T v, unsigned d, int width = 32) {
    int l = emu::lane_id();
    int s = l - (int)d;
    if (s < 0 || (s / width) != (l / width)) s = l;
    return emu_shfl(v, s);
}

template <typename T>
static inline T atomicAdd(T* p, T v) { T old = *p; *p = old + v; return old; }   // fibers: one host thread
static inline float atomicExch(float* p, float v) { float o = *p; *p = v; return o; }
static inline void __threadfence() {}                                            // one host thread: nothing to order

static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
#define __expf(x) expf(x)
static inline float __fdividef(float a, float b) { return a / b; }
static inline float __frcp_rn(float a) { return 1.0f / a; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __saturatef(float a) { return a < 0.f ? 0.f : (a > 1.f ? 1.f : a); }
template <typename T>
static inline T __ldg(const T* p) { return *p; }
static inline float rsqrtf(float a) { return 1.0f / sqrtf(a); }

static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t = 0) { memset(p, v, n); return 0; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = 0) {
    memmove(d, s, n);
    return 0;
}
static inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dpitch, const void* s, size_t spitch, size_t width,
                                            size_t height, cudaMemcpyKind, cudaStream_t = 0) {
    for (size_t r = 0; r < height; ++r) memmove((char*)d + r * dpitch, (const char*)s + r * spitch, width);
    return 0;
}
template <typename F>
static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
