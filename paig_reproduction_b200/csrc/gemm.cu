// fp32 GEMM on the CUDA cores for the encoder / velocity MLPs (blocks.py:98-100, 43-48) and their
// backward passes.  C[M,N] = epilogue(sum_k A(m,k) * B(k,n) + bias[n]) with fully strided operands, so the
// same kernel serves   Y = X W^T,   dX = dY W   and   dW = dY^T X   without materialising transposes.
// fp32 accumulate in registers: the 1e-4 parity bound of the step rules out single-pass TF32.
#include "common.cuh"
#include "internal.h"

namespace paig {

constexpr int kBM = 64, kBN = 64, kBK = 16, kGemmThreads = 256;

// A_KFAST: A's k index is the contiguous one (stride_ak == 1) -> threads sweep k first when loading.
template <bool A_KFAST, bool B_KFAST>
__global__ void __launch_bounds__(kGemmThreads) sgemm_kernel(GemmArgs g) {
    __shared__ __align__(16) float As[kBK][kBM + 4];
    __shared__ __align__(16) float Bs[kBK][kBN + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * kBM, n0 = blockIdx.x * kBN;
    const int tx = tid & 15, ty = tid >> 4;            // 16 x 16 threads, 4 x 4 outputs each
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < g.K; k0 += kBK) {
        // ---- stage A tile (64 x 16) and B tile (16 x 64), zero-filled outside the matrix ----
#pragma unroll
        for (int it = 0; it < (kBM * kBK) / kGemmThreads; ++it) {
            const int e = tid + it * kGemmThreads;
            int mm, kk;
            if (A_KFAST) { kk = e % kBK; mm = e / kBK; } else { mm = e % kBM; kk = e / kBM; }
            const int m = m0 + mm, k = k0 + kk;
            As[kk][mm] = (m < g.M && k < g.K) ? g.A[(long)m * g.sam + (long)k * g.sak] : 0.f;
        }
#pragma unroll
        for (int it = 0; it < (kBN * kBK) / kGemmThreads; ++it) {
            const int e = tid + it * kGemmThreads;
            int nn, kk;
            if (B_KFAST) { kk = e % kBK; nn = e / kBK; } else { nn = e % kBN; kk = e / kBN; }
            const int n = n0 + nn, k = k0 + kk;
            Bs[kk][nn] = (n < g.N && k < g.K) ? g.B[(long)k * g.sbk + (long)n * g.sbn] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < kBK; ++kk) {
            const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] += av[i] * bv[j];
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= g.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= g.N) continue;
            float v = acc[i][j];
            if (g.bias) v += g.bias[n];
            const long ci = (long)m * g.ldc + n;
            switch (g.epi) {
                case EPI_RELU: v = fmaxf(v, 0.f); break;
                case EPI_TANH: v = tanhf(v); break;
                case EPI_MASK_RELU: v = g.aux[(long)m * g.ldaux + n] > 0.f ? v : 0.f; break;
                case EPI_MASK_TANH: { const float h = g.aux[(long)m * g.ldaux + n]; v *= 1.f - h * h; } break;
                default: break;
            }
            if (g.accumulate) v += g.C[ci];
            g.C[ci] = v;
        }
    }
}

int gemm(const GemmArgs& g, cudaStream_t st) {
    if (g.M <= 0 || g.N <= 0) return 0;
    dim3 grid(cdiv(g.N, kBN), cdiv(g.M, kBM));
    const bool ak = g.sak == 1, bk = g.sbk == 1;
    if (ak && bk) launch(sgemm_kernel<true, true>, grid, dim3(kGemmThreads), 0, st, g);
    else if (ak) launch(sgemm_kernel<true, false>, grid, dim3(kGemmThreads), 0, st, g);
    else if (bk) launch(sgemm_kernel<false, true>, grid, dim3(kGemmThreads), 0, st, g);
    else launch(sgemm_kernel<false, false>, grid, dim3(kGemmThreads), 0, st, g);
    return check_launch("sgemm");
}

// out[n] = sum_m X[m*ld + n]   (bias gradients).  One warp per 32 columns, fixed order.
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ X, int M, int N, int ld,
                                                     float* __restrict__ out) {
    __shared__ float part[8][33];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int n = blockIdx.x * 32 + lane;
    float s = 0.f;
    if (n < N)
        for (int m = w; m < M; m += 8) s += X[(long)m * ld + n];
    part[w][lane] = s;
    __syncthreads();
    if (w == 0 && n < N) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += part[i][lane];
        out[n] = t;
    }
}

int colsum(const float* X, int M, int N, int ld, float* out, cudaStream_t st) {
    if (N <= 0) return 0;
    launch(colsum_kernel, dim3(cdiv(N, 32)), dim3(256), 0, st, X, M, N, ld, out);
    return check_launch("colsum");
}

// Linear layer helpers over row-major X[M,K], W[N,K] (torch.nn.Linear layout), Y[M,N].
int linear_forward(const float* X, const float* W, const float* b, float* Y, int M, int K, int N, int epi,
                   cudaStream_t st) {
    GemmArgs g;
    g.A = X; g.sam = K; g.sak = 1;
    g.B = W; g.sbk = 1; g.sbn = K;
    g.C = Y; g.ldc = N; g.M = M; g.N = N; g.K = K;
    g.bias = b; g.epi = epi;
    return gemm(g, st);
}

// dX[M,K] = dY[M,N] W[N,K]  (optionally masked by the activation of X's producer: aux[M,K]).
int linear_dgrad(const float* dY, const float* W, float* dX, int M, int K, int N, int epi, const float* aux,
                 cudaStream_t st) {
    GemmArgs g;
    g.A = dY; g.sam = N; g.sak = 1;
    g.B = W; g.sbk = K; g.sbn = 1;
    g.C = dX; g.ldc = K; g.M = M; g.N = K; g.K = N;
    g.epi = epi; g.aux = aux; g.ldaux = K;
    return gemm(g, st);
}

// dW[N,K] = dY[M,N]^T X[M,K] ; db[N] = colsum(dY).
int linear_wgrad(const float* dY, const float* X, float* dW, float* db, int M, int K, int N, cudaStream_t st) {
    GemmArgs g;
    g.A = dY; g.sam = 1; g.sak = N;
    g.B = X; g.sbk = K; g.sbn = 1;
    g.C = dW; g.ldc = K; g.M = N; g.N = K; g.K = M;
    int rc = gemm(g, st);
    if (rc) return rc;
    return db ? colsum(dY, M, N, N, db, st) : 0;
}

}  // namespace paig
