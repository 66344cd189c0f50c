// fp32 GEMM on the CUDA cores for the encoder / velocity MLPs (blocks.py:98-100, 43-48) and their
// backward passes.  C[M,N] = epilogue(sum_k A(m,k) * B(k,n) + bias[n]) with fully strided operands, so the
// same kernel serves   Y = X W^T,   dX = dY W   and   dW = dY^T X   without materialising transposes.
// fp32 accumulate in registers: the 1e-4 parity bound of the step rules out single-pass TF32.
#include "common.cuh"
#include "internal.h"

#include <cstdint>
#include <cstdlib>

namespace paig {

// Tiling: a CTA of 256 threads owns a 128 x BN tile (BN = 64 or 128) of C and a contiguous K range (split-K over
// gridDim.z when the tile count alone cannot fill 148 SMs); a thread owns an 8 x (BN/16) register tile read from shared
// memory as 16-byte vectors along m / n.  The next K slab is fetched into registers while the current one is
// multiplied (software double buffering, one __syncthreads pair per slab).  Operands may be contiguous along k
// (16-byte global loads along k, transposed on the way into shared memory) or along m / n (straight vector copy).
constexpr int kBK = 16, kGemmThreads = 256;
constexpr int kSplitTargetCtas = 296;     // 2 resident CTAs per SM
static int split_fixed_k() {               // SPLIT_FIXED chunk length (multiple of kBK)
    static const int v = getenv("PAIG_GEMM_SPLITK") ? atoi(getenv("PAIG_GEMM_SPLITK")) : 512;
    return v;
}
#define kSplitFixedK split_fixed_k()

// four consecutive-k (KFAST) or consecutive-mn elements of an operand, zero outside the matrix
template <bool KFAST>
__device__ __forceinline__ float4 load4(const float* __restrict__ P, long s_mn, long s_k, int mn, int k, int MN, int kend,
                                        bool vec_ok) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (KFAST) {
        if (mn >= MN) return v;
        const float* p = P + (long)mn * s_mn + k;
        if (vec_ok && k + 3 < kend) return *reinterpret_cast<const float4*>(p);
        if (k < kend) v.x = p[0];
        if (k + 1 < kend) v.y = p[1];
        if (k + 2 < kend) v.z = p[2];
        if (k + 3 < kend) v.w = p[3];
    } else {
        if (k >= kend) return v;
        const float* p = P + (long)k * s_k + mn;
        if (vec_ok && mn + 3 < MN) return *reinterpret_cast<const float4*>(p);
        if (mn < MN) v.x = p[0];
        if (mn + 1 < MN) v.y = p[1];
        if (mn + 2 < MN) v.z = p[2];
        if (mn + 3 < MN) v.w = p[3];
    }
    return v;
}

// smem tile T[kBK][LD]: element (k, mn).  KFAST: the 4 values are k..k+3 of one mn; else mn..mn+3 of one k.
template <bool KFAST, int LD>
__device__ __forceinline__ void store4(float* T, int mn, int k, const float4& v) {
    if (KFAST) {
        T[(k + 0) * LD + mn] = v.x; T[(k + 1) * LD + mn] = v.y; T[(k + 2) * LD + mn] = v.z; T[(k + 3) * LD + mn] = v.w;
    } else {
        *reinterpret_cast<float4*>(T + k * LD + mn) = v;
    }
}

__device__ __forceinline__ float gemm_epilogue(const GemmArgs& g, float v, int m, int n) {
    if (g.bias) v += g.bias[n];
    switch (g.epi) {
        case EPI_RELU: v = fmaxf(v, 0.f); break;
        case EPI_TANH: v = tanhf(v); break;
        case EPI_MASK_RELU: v = g.aux[(long)m * g.ldaux + n] > 0.f ? v : 0.f; break;
        case EPI_MASK_TANH: { const float h = g.aux[(long)m * g.ldaux + n]; v *= 1.f - h * h; } break;
        default: break;
    }
    if (g.accumulate) v += g.C[(long)m * g.ldc + n];
    return v;
}

template <bool A_KFAST, bool B_KFAST, int BM, int BN>
__global__ void __launch_bounds__(kGemmThreads, (BM == 128 ? 1 : 2)) sgemm_kernel(GemmArgs g) {
    constexpr int TM = BM / 16, TN = BN / 16;          // rows / columns per thread: 4 or 8
    constexpr int LDA = BM + 4, LDB = BN + 4;
    constexpr int A_IT = BM * kBK / 4 / kGemmThreads;   // float4 per thread per slab: 1 or 2
    constexpr int B_IT = BN * kBK / 4 / kGemmThreads;   // 1 or 2
    __shared__ __align__(16) float As[kBK * LDA];
    __shared__ __align__(16) float Bs[kBK * LDB];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int tx = tid & 15, ty = tid >> 4;
    // K range of this split
    const int kper = g.kper;
    const int kbeg = blockIdx.z * kper, kend = min(g.K, kbeg + kper);

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    // per-thread staging coordinates
    int a_mn[A_IT], a_k[A_IT], b_mn[B_IT], b_k[B_IT];
#pragma unroll
    for (int it = 0; it < A_IT; ++it) {
        const int e = tid + it * kGemmThreads;
        if (A_KFAST) { a_k[it] = (e % (kBK / 4)) * 4; a_mn[it] = e / (kBK / 4); }
        else { a_mn[it] = (e % (BM / 4)) * 4; a_k[it] = e / (BM / 4); }
    }
#pragma unroll
    for (int it = 0; it < B_IT; ++it) {
        const int e = tid + it * kGemmThreads;
        if (B_KFAST) { b_k[it] = (e % (kBK / 4)) * 4; b_mn[it] = e / (kBK / 4); }
        else { b_mn[it] = (e % (BN / 4)) * 4; b_k[it] = e / (BN / 4); }
    }
    float4 ra[A_IT], rb[B_IT];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int it = 0; it < A_IT; ++it)
            ra[it] = load4<A_KFAST>(g.A, g.sam, g.sak, m0 + a_mn[it], k0 + a_k[it], g.M, kend, g.a_vec != 0);
#pragma unroll
        for (int it = 0; it < B_IT; ++it)
            rb[it] = load4<B_KFAST>(g.B, g.sbn, g.sbk, n0 + b_mn[it], k0 + b_k[it], g.N, kend, g.b_vec != 0);
    };
    if (kbeg < kend) fetch(kbeg);
    for (int k0 = kbeg; k0 < kend; k0 += kBK) {
#pragma unroll
        for (int it = 0; it < A_IT; ++it) store4<A_KFAST, LDA>(As, a_mn[it], a_k[it], ra[it]);
#pragma unroll
        for (int it = 0; it < B_IT; ++it) store4<B_KFAST, LDB>(Bs, b_mn[it], b_k[it], rb[it]);
        __syncthreads();
        if (k0 + kBK < kend) fetch(k0 + kBK);          // in flight while this slab is multiplied
#pragma unroll
        for (int kk = 0; kk < kBK; ++kk) {
            float av[TM], bv[TN];
#pragma unroll
            for (int i4 = 0; i4 < TM / 4; ++i4) {
                const float4 a = *reinterpret_cast<const float4*>(&As[kk * LDA + i4 * 64 + ty * 4]);
                av[i4 * 4 + 0] = a.x; av[i4 * 4 + 1] = a.y; av[i4 * 4 + 2] = a.z; av[i4 * 4 + 3] = a.w;
            }
#pragma unroll
            for (int j4 = 0; j4 < TN / 4; ++j4) {
                const float4 b = *reinterpret_cast<const float4*>(&Bs[kk * LDB + j4 * 64 + tx * 4]);
                bv[j4 * 4 + 0] = b.x; bv[j4 * 4 + 1] = b.y; bv[j4 * 4 + 2] = b.z; bv[j4 * 4 + 3] = b.w;
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] += av[i] * bv[j];
        }
        __syncthreads();
    }
    // rows i4*64 + ty*4 + i; columns j4*64 + tx*4 + j
    const bool partial = gridDim.z > 1;
    float* P = partial ? g.splitk_ws + (size_t)blockIdx.z * g.M * g.N : nullptr;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int m = m0 + (i / 4) * 64 + ty * 4 + (i % 4);
        if (m >= g.M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int n = n0 + (j / 4) * 64 + tx * 4 + (j % 4);
            if (n >= g.N) continue;
            if (partial) P[(size_t)m * g.N + n] = acc[i][j];
            else g.C[(long)m * g.ldc + n] = gemm_epilogue(g, acc[i][j], m, n);
        }
    }
}

// C = epilogue(sum over splits, fixed order => deterministic)
__global__ void __launch_bounds__(256) splitk_reduce_kernel(GemmArgs g, int splits) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long MN = (long)g.M * g.N;
    if (idx >= MN) return;
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += g.splitk_ws[(size_t)z * MN + idx];
    const int m = (int)(idx / g.N), n = (int)(idx % g.N);
    g.C[(long)m * g.ldc + n] = gemm_epilogue(g, s, m, n);
}

template <int BM, int BN>
static void launch_sgemm(const GemmArgs& g, dim3 grid, cudaStream_t st) {
    const bool ak = g.sak == 1, bk = g.sbk == 1;
    if (ak && bk) launch(sgemm_kernel<true, true, BM, BN>, grid, dim3(kGemmThreads), 0, st, g);
    else if (ak) launch(sgemm_kernel<true, false, BM, BN>, grid, dim3(kGemmThreads), 0, st, g);
    else if (bk) launch(sgemm_kernel<false, true, BM, BN>, grid, dim3(kGemmThreads), 0, st, g);
    else launch(sgemm_kernel<false, false, BM, BN>, grid, dim3(kGemmThreads), 0, st, g);
}

int gemm(const GemmArgs& in, cudaStream_t st) {
    GemmArgs g = in;
    if (g.M <= 0 || g.N <= 0) return 0;
    const bool ak = g.sak == 1, bk = g.sbk == 1;
    if (!ak && g.sam != 1) { set_error("gemm: A must be contiguous along m or k"); return 1; }
    if (!bk && g.sbn != 1) { set_error("gemm: B must be contiguous along n or k"); return 1; }
    // 16-byte loads need an aligned base and a leading stride that keeps every vector aligned
    g.a_vec = ((uintptr_t)g.A % 16 == 0) && ((ak ? g.sam : g.sak) % 4 == 0);
    g.b_vec = ((uintptr_t)g.B % 16 == 0) && ((bk ? g.sbn : g.sbk) % 4 == 0);
    // 64x64 tiles (4x4 per thread, two CTAs per SM) measured faster than 128x128 (8x8 per thread) on every GEMM of the
    // step (B200: l1 dgrad 0.089 vs 0.111 ms); the big tile stays selectable for experiments (PAIG_GEMM_BIG=1)
    static const int force_big = getenv("PAIG_GEMM_BIG") ? atoi(getenv("PAIG_GEMM_BIG")) : -1;
    bool big = false;
    if (force_big >= 0 && g.N > 64 && g.M > 64) big = force_big != 0;
    const int BM = big ? 128 : 64, BN = big ? 128 : 64;
    const int tiles = cdiv(g.M, BM) * cdiv(g.N, BN);
    int splits = 1;
    if (g.split_mode == SPLIT_FIXED && g.K >= 2 * kSplitFixedK) {
        // K chunks of a fixed length: the summation order depends on (K) only, never on M, so a row of C is
        // bit-identical whatever batch it is computed in (tests: per-sequence independence of the eval sweep)
        splits = cdiv(g.K, kSplitFixedK);
        if (!g.splitk_ws || g.splitk_floats < (size_t)splits * g.M * g.N) {
            set_error("gemm: split-K scratch too small (%zu floats for %d x %d x %d)", g.splitk_floats, splits, g.M, g.N);
            return 1;
        }
    } else if (g.split_mode == SPLIT_AUTO && g.splitk_ws && tiles < kSplitTargetCtas && g.K >= 8 * kBK) {
        splits = cdiv(kSplitTargetCtas, tiles);
        const int max_by_k = g.K / (4 * kBK);                       // at least 4 slabs per split
        if (splits > max_by_k) splits = max_by_k;
        const size_t cap = g.splitk_floats / ((size_t)g.M * g.N);
        if ((size_t)splits > cap) splits = (int)cap;
        if (splits < 1) splits = 1;
        const int kper = cdiv(cdiv(g.K, splits), kBK) * kBK;        // drop empty trailing splits
        splits = cdiv(g.K, kper);
    }
    g.kper = g.split_mode == SPLIT_FIXED && splits > 1 ? kSplitFixedK : cdiv(cdiv(g.K, splits), kBK) * kBK;
    dim3 grid(cdiv(g.N, BN), cdiv(g.M, BM), splits);
    if (big) launch_sgemm<128, 128>(g, grid, st);
    else launch_sgemm<64, 64>(g, grid, st);
    int rc = check_launch(g.tag ? g.tag : "sgemm");
    if (rc || splits == 1) return rc;
    launch(splitk_reduce_kernel, dim3(cdiv((long)g.M * g.N, 256)), dim3(256), 0, st, g, splits);
    return check_launch("splitk_reduce");
}

int gemm_fold_partials(const GemmArgs& g, int splits, cudaStream_t st) {
    launch(splitk_reduce_kernel, dim3(cdiv((long)g.M * g.N, 256)), dim3(256), 0, st, g, splits);
    return check_launch("splitk_reduce");
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
                   cudaStream_t st, float* skws, size_t skfl, const char* tag) {
    GemmArgs g;
    g.tag = tag;
    g.splitk_ws = skws; g.splitk_floats = skfl; g.split_mode = skws ? SPLIT_FIXED : SPLIT_NONE;
    g.A = X; g.sam = K; g.sak = 1;
    g.B = W; g.sbk = 1; g.sbn = K;
    g.C = Y; g.ldc = N; g.M = M; g.N = N; g.K = K;
    g.bias = b; g.epi = epi;
    return gemm(g, st);
}

// dX[M,K] = dY[M,N] W[N,K]  (optionally masked by the activation of X's producer: aux[M,K]).
int linear_dgrad(const float* dY, const float* W, float* dX, int M, int K, int N, int epi, const float* aux,
                 cudaStream_t st, float* skws, size_t skfl, const char* tag) {
    GemmArgs g;
    g.tag = tag;
    g.splitk_ws = skws; g.splitk_floats = skfl;
    g.A = dY; g.sam = N; g.sak = 1;
    g.B = W; g.sbk = K; g.sbn = 1;
    g.C = dX; g.ldc = K; g.M = M; g.N = K; g.K = N;
    g.epi = epi; g.aux = aux; g.ldaux = K;
    return gemm(g, st);
}

// dW[N,K] = dY[M,N]^T X[M,K] ; db[N] = colsum(dY).
int linear_wgrad(const float* dY, const float* X, float* dW, float* db, int M, int K, int N, cudaStream_t st,
                 float* skws, size_t skfl, const char* tag) {
    GemmArgs g;
    g.tag = tag;
    g.splitk_ws = skws; g.splitk_floats = skfl; g.split_mode = skws ? SPLIT_AUTO : SPLIT_NONE;
    g.A = dY; g.sam = 1; g.sak = N;
    g.B = X; g.sbk = K; g.sbn = 1;
    g.C = dW; g.ldc = K; g.M = N; g.N = K; g.K = M;
    int rc = gemm(g, st);
    if (rc) return rc;
    return db ? colsum(dY, M, N, N, db, st) : 0;
}

}  // namespace paig
