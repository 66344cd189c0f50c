// Standalone probe of cp.async.bulk.tensor.4d with out-of-bounds boxes (development aid for csrc/wgrad_tma.cu).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
struct Args { CUtensorMap tm; int x, y, c, n; int floats; float* out; int variant; };
__global__ void k(const __grid_constant__ Args a) {
    extern __shared__ __align__(128) unsigned char raw[];
    __shared__ unsigned long long bar;
    float* sm = (float*)raw;
    unsigned pad = (128u - ((unsigned)__cvta_generic_to_shared(raw) & 127u)) & 127u;
    sm += pad / 4;
    unsigned bar_a = (unsigned)__cvta_generic_to_shared(&bar), dst = (unsigned)__cvta_generic_to_shared(sm);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(a.floats * 4) : "memory");
        if (a.variant == 0)
            asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                         ::"r"(dst), "l"(reinterpret_cast<uint64_t>(&a.tm)), "r"(a.x), "r"(a.y), "r"(a.c), "r"(a.n), "r"(bar_a) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                         ::"r"(dst), "l"(reinterpret_cast<uint64_t>(&a.tm)), "r"(a.x), "r"(a.y), "r"(a.c), "r"(a.n), "r"(bar_a) : "memory");
    }
    unsigned done = 0;
    while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar_a), "r"(0u) : "memory");
    for (int i = threadIdx.x; i < a.floats; i += blockDim.x) a.out[i] = sm[i];
}
int main(int argc, char** argv) {
    int S = atoi(argv[1]), C = atoi(argv[2]), N = atoi(argv[3]), bx = atoi(argv[4]), by = atoi(argv[5]), x = atoi(argv[6]), y = atoi(argv[7]);
    int variant = argc > 8 ? atoi(argv[8]) : 0;
    std::vector<float> h((size_t)N * C * S * S);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 1000) + 1.f;
    float *d, *out;
    cudaMalloc(&d, h.size() * 4); cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    int floats = bx * by * C;
    cudaMalloc(&out, floats * 4);
    typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    Args a; memset(&a, 0, sizeof(a));
    cuuint64_t dims[4] = {(cuuint64_t)S, (cuuint64_t)S, (cuuint64_t)C, (cuuint64_t)N};
    cuuint64_t str[3] = {(cuuint64_t)S * 4, (cuuint64_t)S * S * 4, (cuuint64_t)C * S * S * 4};
    cuuint32_t box[4] = {(cuuint32_t)bx, (cuuint32_t)by, (cuuint32_t)C, 1}, es[4] = {1, 1, 1, 1};
    CUresult r = ((Fn)p)(&a.tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d q=%d\n", (int)r, (int)q);
    if (r) return 1;
    a.x = x; a.y = y; a.c = 0; a.n = N - 1; a.floats = floats; a.out = out; a.variant = variant;
    k<<<1, 128, floats * 4 + 256>>>(a);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    if (e) return 2;
    std::vector<float> o(floats);
    cudaMemcpy(o.data(), out, floats * 4, cudaMemcpyDeviceToHost);
    // check against the definition
    int bad = 0;
    for (int c = 0; c < C; ++c) for (int r2 = 0; r2 < by; ++r2) for (int kx = 0; kx < bx; ++kx) {
        int gx = x + kx, gy = y + r2; float want = 0.f;
        if (gx >= 0 && gx < S && gy >= 0 && gy < S) want = h[(((size_t)(N - 1) * C + c) * S + gy) * S + gx];
        if (o[(c * by + r2) * bx + kx] != want) ++bad;
    }
    printf("mismatches: %d of %d\n", bad, floats);
    return bad != 0;
}
