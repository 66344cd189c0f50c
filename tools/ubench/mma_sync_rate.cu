// Microbenchmark: legacy mma.sync (HMMA) issue rate on sm_100a.  MAC/clk/SM for m16n8k8 tf32 and m16n8k16 bf16, with
// 1..16 warps per SM, 8 independent accumulator chains per warp.  nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
template <int KIND>
__global__ void k(float* out, int iters, long long* clk) {
    unsigned a[4] = {threadIdx.x, threadIdx.x * 3u, threadIdx.x * 5u, threadIdx.x * 7u}, b[2] = {threadIdx.x * 11u, threadIdx.x * 13u};
    float c[8][4];
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (KIND == 0)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
            else
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
        }
    }
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}
int main() {
    float* out; long long* clk; cudaMalloc(&out, 148 * 1024 * 4); cudaMallocManaged(&clk, 8);
    const int iters = 2000;
    for (int kind = 0; kind < 2; ++kind)
        for (int warps = 1; warps <= 16; warps *= 2) {
            if (kind == 0) k<0><<<148, warps * 32>>>(out, iters, clk); else k<1><<<148, warps * 32>>>(out, iters, clk);
            cudaDeviceSynchronize();
            if (kind == 0) k<0><<<148, warps * 32>>>(out, iters, clk); else k<1><<<148, warps * 32>>>(out, iters, clk);
            cudaError_t e = cudaDeviceSynchronize();
            double macs = (double)iters * 8 * warps * 16 * 8 * (kind ? 16 : 8);
            printf("{\"kind\": \"%s\", \"warps_per_sm\": %d, \"clk\": %lld, \"mac_per_clk_per_sm\": %.1f, \"err\": %d}\n",
                   kind ? "m16n8k16.bf16" : "m16n8k8.tf32", warps, *clk, macs / (double)*clk, (int)e);
        }
    return 0;
}
