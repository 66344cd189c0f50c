// Microbenchmark: what limits HMMA.1688.F32.TF32 issue inside a realistic instruction mix on sm_100a?
// Variants (16 warps per SM, 1 CTA per SM): V0 same A/B registers, 8 accumulators (the rate test); V1 distinct A/B per MMA;
// V2 = V1 with every third MMA starting from a zero accumulator + 4 FADDs (the 3xTF32 pattern); V3 = V2 + the split ALU
// work (2 IADD + 2 LOP + 2 FADD per n-tile); V4 = V3 + 2 LDS per n-tile.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma_acc(float (&d)[4], const float (&a)[4], float b0, float b1) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(__float_as_uint(a[0])), "r"(__float_as_uint(a[1])), "r"(__float_as_uint(a[2])), "r"(__float_as_uint(a[3])), "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)));
}
__device__ __forceinline__ void mma_zero(float (&d)[4], const float (&a)[4], float b0, float b1) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
        : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
        : "r"(__float_as_uint(a[0])), "r"(__float_as_uint(a[1])), "r"(__float_as_uint(a[2])), "r"(__float_as_uint(a[3])), "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)), "f"(0.f));
}
__device__ __forceinline__ float hi(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }
template <int V>
__global__ void __launch_bounds__(512, 1) k(float* out, const float* in, int iters, long long* clk) {
    __shared__ float sm[9216];
    for (int i = threadIdx.x; i < 9216; i += blockDim.x) sm[i] = in[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    float ah[4], al[4], sum[4][4], ca[4][4], cb[4][4];
    for (int i = 0; i < 4; ++i) { ah[i] = hi(in[threadIdx.x + i]); al[i] = in[threadIdx.x + i] - ah[i]; }
    for (int n = 0; n < 4; ++n) for (int i = 0; i < 4; ++i) sum[n][i] = ca[n][i] = cb[n][i] = 0.f;
    const float* pb = sm + t * 1224 + g;
    float b0[4], b1[4];
    for (int n = 0; n < 4; ++n) { b0[n] = in[lane + n]; b1[n] = in[lane + 8 + n]; }
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
            float d[4][4], h0[4], h1[4], l0[4], l1[4];
#pragma unroll
            for (int n = 0; n < 4; ++n) {
                float x0 = b0[n], x1 = b1[n];
                if (V >= 4) { x0 = pb[(tap / 3) * 36 + n * 8 + tap % 3 + (it & 1) * 4]; x1 = pb[4 * 1224 + (tap / 3) * 36 + n * 8 + tap % 3 + (it & 1) * 4]; }
                if (V >= 3) { h0[n] = hi(x0); h1[n] = hi(x1); l0[n] = x0 - h0[n]; l1[n] = x1 - h1[n]; }
                else { h0[n] = x0; h1[n] = x1; l0[n] = x1; l1[n] = x0; }
                if (V == 0) { mma_acc(sum[n], ah, b0[0], b1[0]); }
                else if (V == 1) { mma_acc(sum[n], ah, h0[n], h1[n]); }
                else mma_zero(d[n], ah, h0[n], h1[n]);
            }
#pragma unroll
            for (int n = 0; n < 4; ++n) { if (V == 0) mma_acc(ca[n], ah, b0[0], b1[0]); else mma_acc(ca[n], ah, l0[n], l1[n]); }
#pragma unroll
            for (int n = 0; n < 4; ++n) { if (V == 0) mma_acc(cb[n], ah, b0[0], b1[0]); else mma_acc(cb[n], al, h0[n], h1[n]); }
            if (V >= 2) {
#pragma unroll
                for (int n = 0; n < 4; ++n) for (int i = 0; i < 4; ++i) sum[n][i] += d[n][i];
            }
        }
    }
    long long t1 = clock64();
    float s = 0;
    for (int n = 0; n < 4; ++n) for (int i = 0; i < 4; ++i) s += sum[n][i] + ca[n][i] + cb[n][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}
template <int V> void run(float* out, float* in, long long* clk, int warps) {
    const int iters = 200;
    k<V><<<148, warps * 32>>>(out, in, iters, clk); cudaDeviceSynchronize();
    k<V><<<148, warps * 32>>>(out, in, iters, clk); cudaError_t e = cudaDeviceSynchronize();
    double hmma_per_sched = (double)iters * 9 * 12 * warps / 4.0;
    printf("{\"variant\": %d, \"warps_per_sm\": %d, \"clk\": %lld, \"clk_per_hmma_per_scheduler\": %.2f, \"err\": %d}\n", V, warps, *clk, (double)*clk / hmma_per_sched, (int)e);
}
int main() {
    float *out, *in; long long* clk; cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&in, 16384 * 4); cudaMemset(in, 0, 16384 * 4); cudaMallocManaged(&clk, 8);
    for (int warps = 4; warps <= 16; warps *= 2) { run<0>(out, in, clk, warps); run<1>(out, in, clk, warps); run<2>(out, in, clk, warps); run<3>(out, in, clk, warps); run<4>(out, in, clk, warps); }
    return 0;
}
