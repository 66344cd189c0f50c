// Does HMMA.1688.F32.TF32 issue overlap with other instructions of the same scheduler?  Per loop iteration: 8 independent
// accumulating MMAs (volatile), each followed by K filler instructions (volatile asm: FADD on private registers, or
// IADD/LOP3 pairs, or LDS).  16 warps per SM.  Output: cycles per MMA per scheduler for K = 0..8.
#include <cstdio>
#include <cuda_runtime.h>
template <int K, int KIND>
__global__ void __launch_bounds__(512, 1) k(float* out, int iters, long long* clk) {
    __shared__ float sm[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = i;
    __syncthreads();
    unsigned a[4] = {threadIdx.x, threadIdx.x * 3u, threadIdx.x * 5u, threadIdx.x * 7u}, b[2] = {threadIdx.x * 11u, threadIdx.x * 13u};
    float c[8][4], f[8];
    unsigned u[8];
    for (int i = 0; i < 8; ++i) { f[i] = i; u[i] = threadIdx.x + i; for (int j = 0; j < 4; ++j) c[i][j] = 0.f; }
    const float* p = sm + threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
#pragma unroll
            for (int j = 0; j < K; ++j) {
                if (KIND == 0) asm volatile("add.f32 %0, %0, %1;" : "+f"(f[(i + j) & 7]) : "f"(1.0f));
                else if (KIND == 1) asm volatile("add.u32 %0, %0, 4096;\n\tand.b32 %0, %0, 0xffffe001;" : "+r"(u[(i + j) & 7]));
                else asm volatile("ld.shared.f32 %0, [%1];" : "=f"(f[(i + j) & 7]) : "l"(p + ((i + j) & 7) * 32));
            }
        }
    }
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < 8; ++i) { s += f[i] + u[i]; for (int j = 0; j < 4; ++j) s += c[i][j]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}
template <int K, int KIND> void run(float* out, long long* clk) {
    const int iters = 500;
    k<K, KIND><<<148, 512>>>(out, iters, clk); cudaDeviceSynchronize();
    k<K, KIND><<<148, 512>>>(out, iters, clk); cudaError_t e = cudaDeviceSynchronize();
    printf("{\"filler\": \"%s\", \"per_mma\": %d, \"clk_per_mma_per_scheduler\": %.2f, \"err\": %d}\n",
           KIND == 0 ? "FADD" : (KIND == 1 ? "IADD+LOP3 pair" : "LDS"), K, (double)*clk / (iters * 8 * 4.0), (int)e);
}
int main() {
    float* out; long long* clk; cudaMalloc(&out, 148 * 512 * 4); cudaMallocManaged(&clk, 8);
    run<0, 0>(out, clk); run<2, 0>(out, clk); run<4, 0>(out, clk); run<6, 0>(out, clk); run<8, 0>(out, clk);
    run<1, 1>(out, clk); run<2, 1>(out, clk); run<3, 1>(out, clk); run<4, 1>(out, clk);
    run<1, 2>(out, clk); run<2, 2>(out, clk); run<3, 2>(out, clk);
    return 0;
}
