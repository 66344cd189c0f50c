// The inner loop of the tensor-core fused UNet kernels (csrc/unet_fused.cu conv_accumulate_mma<32, 2>) in isolation, with parts
// switched off, to find what keeps the tensor pipe at 40 %:  MODE bits: 1 = B operands from shared memory (else registers),
// 2 = hi/lo split arithmetic, 4 = FADD of the hi.hi partials, 8 = weights from shared memory + split.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }
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
constexpr int P = 36, PLANE = 34 * 36, NT = 4;
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(float* out, const float* in, int nkc, int reps, long long* clk) {
    extern __shared__ float sm[];
    for (int i = threadIdx.x; i < 16 * PLANE + 4096; i += blockDim.x) sm[i] = in[i & 16383];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    float sum[NT][4], corr[NT][2][4];
    for (int n = 0; n < NT; ++n) for (int i = 0; i < 4; ++i) sum[n][i] = corr[n][0][i] = corr[n][1][i] = 0.f;
    float dp[NT][4];
    for (int n = 0; n < NT; ++n) for (int i = 0; i < 4; ++i) dp[n][i] = 0.f;
    float rb[8];
    for (int i = 0; i < 8; ++i) rb[i] = in[lane + i];
    float ra[4];
    for (int i = 0; i < 4; ++i) ra[i] = in[lane * 4 + i];
    long long t0 = clock64();
    for (int rep = 0; rep < reps; ++rep) {
        const float* pb = sm + t * PLANE + (warp + 16 * (rep & 1)) * P + g;
        const float4* wf = reinterpret_cast<const float4*>(sm + 16 * PLANE) + lane;
#pragma unroll 1
        for (int kc = 0; kc < nkc; ++kc) {
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
                const int ky = tap / 3, kx = tap % 3;
                float ah[4], al[4];
                if (MODE & 8) {
                    const float4 w4 = wf[tap * 32];
                    ah[0] = tf32_hi(w4.x); al[0] = w4.x - ah[0]; ah[1] = tf32_hi(w4.y); al[1] = w4.y - ah[1];
                    ah[2] = tf32_hi(w4.z); al[2] = w4.z - ah[2]; ah[3] = tf32_hi(w4.w); al[3] = w4.w - ah[3];
                } else {
                    for (int i = 0; i < 4; ++i) { ah[i] = ra[i]; al[i] = ra[3 - i]; }
                }
                float d[NT][4], h0[NT], h1[NT], l0[NT], l1[NT];
                if (MODE & 16) {       // FADDs of the PREVIOUS tap's partials: their MMAs finished long ago
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) { sum[nt][0] += dp[nt][0]; sum[nt][1] += dp[nt][1]; sum[nt][2] += dp[nt][2]; sum[nt][3] += dp[nt][3]; }
                }
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    float b0, b1;
                    if (MODE & 1) { const float* q = pb + ky * P + nt * 8 + kx; b0 = q[0]; b1 = q[4 * PLANE]; }
                    else { b0 = rb[(nt + tap) & 7]; b1 = rb[(nt + tap + 3) & 7]; }
                    if (MODE & 2) { h0[nt] = tf32_hi(b0); h1[nt] = tf32_hi(b1); l0[nt] = b0 - h0[nt]; l1[nt] = b1 - h1[nt]; }
                    else { h0[nt] = b0; h1[nt] = b1; l0[nt] = b1; l1[nt] = b0; }
                    mma_zero(d[nt], ah, h0[nt], h1[nt]);
                }
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) mma_acc(corr[nt][0], ah, l0[nt], l1[nt]);
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) mma_acc(corr[nt][1], al, h0[nt], h1[nt]);
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    if (MODE & 16) { dp[nt][0] = d[nt][0]; dp[nt][1] = d[nt][1]; dp[nt][2] = d[nt][2]; dp[nt][3] = d[nt][3]; }
                    else if (MODE & 4) { sum[nt][0] += d[nt][0]; sum[nt][1] += d[nt][1]; sum[nt][2] += d[nt][2]; sum[nt][3] += d[nt][3]; }
                    else { sum[nt][0] = d[nt][0]; sum[nt][1] = d[nt][1]; sum[nt][2] = d[nt][2]; sum[nt][3] = d[nt][3]; }
                }
            }
            pb += PLANE;           // (stays inside the 16 planes for nkc <= 4)
            wf += 9 * 32;
        }
    }
    long long t1 = clock64();
    float s = 0;
    for (int n = 0; n < NT; ++n) for (int i = 0; i < 4; ++i) s += sum[n][i] + corr[n][0][i] + corr[n][1][i] + dp[n][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}
template <int MODE> void run(float* out, float* in, long long* clk) {
    const int nkc = 2, reps = 200;
    const size_t smem = (16 * PLANE + 4096) * 4;
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<MODE><<<148, 512, smem>>>(out, in, nkc, reps, clk); cudaDeviceSynchronize();
    k<MODE><<<148, 512, smem>>>(out, in, nkc, reps, clk); cudaError_t e = cudaDeviceSynchronize();
    double hmma_per_sched = (double)reps * nkc * 108 * 4;
    printf("{\"mode\": %d, \"clk_per_hmma_per_scheduler\": %.2f, \"err\": %d}\n", MODE, (double)*clk / hmma_per_sched, (int)e);
}
int main() {
    float *out, *in; long long* clk; cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&in, 16384 * 4);
    float* h = (float*)malloc(16384 * 4); for (int i = 0; i < 16384; ++i) h[i] = (float)((i * 2654435761u) >> 8) / 16777216.f - 0.5f;
    cudaMemcpy(in, h, 16384 * 4, cudaMemcpyHostToDevice); cudaMallocManaged(&clk, 8);
    run<0>(out, in, clk); run<1>(out, in, clk); run<2>(out, in, clk); run<3>(out, in, clk); run<4>(out, in, clk); run<6>(out, in, clk);
    run<7>(out, in, clk); run<15>(out, in, clk); run<23>(out, in, clk); run<31>(out, in, clk);
    return 0;
}
