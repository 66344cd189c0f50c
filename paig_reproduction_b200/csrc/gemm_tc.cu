// encoder.l1 (blocks.py:98: 3072/3888 -> 200) forward, data gradient and weight gradient on the 5th-generation tensor
// cores: C[M,N] = A[M,K] . B[N,K]^T with both operands K-major, as a 3xTF32 product so the result keeps fp32 accuracy
// (the step's 1e-4 parity bound rules out a single TF32 pass):
//
//      x = hi + lo,  hi = rn_tf32(x),  lo = rn_tf32(x - hi)   (x - hi is exact in fp32; |x - hi - lo| <= 2^-23 |x|)
//      A.B ~= A_hi.B_hi + A_hi.B_lo + A_lo.B_hi            (the dropped lo.lo term is ~2^-20 relative)
//
// This is the one dense contraction of the step large enough for tcgen05 (M=2000, N=200, K=3072 and its two
// backward GEMMs).  One CTA per 128 x BN output tile and K range (split-K over gridDim.z):
//   warp 0      TMA producer: 2-D tensor maps, 128-byte swizzle, boxes of 32 floats (one swizzle span) x rows; two
//               stages on full/empty mbarriers
//   warps 2-9   split each landed stage in place into hi (over the raw tile) and lo (a second tile with the identical
//               swizzled byte layout, so no address arithmetic), then fence.proxy.async and arrive (the drained kernel:
//               warps 6-9 convert, warps 2-5 drain the accumulator, so the two overlap)
//   warp 1      allocates TMEM (256 columns), issues 4 K-steps x 3 tcgen05.mma.kind::tf32 per stage from one thread
//               (M=128, N=BN, K=8, fp32 accumulate in TMEM) and commits to the stage's empty barrier
//   warps 2-5   epilogue: tcgen05.ld 32x32b (a warp owns its 32 TMEM lanes = 32 rows), write the split's partial tile;
//               the shared split-K fold of gemm.cu applies bias / activation / ReLU-or-tanh gate
// The fixed split (by K only) keeps a row's summation order independent of the batch, like the CUDA-core path.
#include "common.cuh"
#include "internal.h"

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#ifndef PAIG_EMU
#include <cuda.h>

namespace paig {

constexpr int kTcThreads = 320;      // warp 0 TMA, warp 1 MMA issue, warps 2..9 converters (2..5 also own the TMEM lane quarters)
constexpr int kTcConv = 256;
constexpr int kTcBM = 128, kTcBK = 32, kTcStages = 2;

struct TcArgs {
    CUtensorMap tmA, tmB;
    float* partials;          // [splits][M][N]
    int M, N, K, BN;
    int kb_per;               // K blocks (of 32) per split
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity) {
    const unsigned a = smem_u32(b);
    unsigned done = 0;
    while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
// K-major operand tile, rows of 128 bytes, 128-byte swizzle: 8-row groups 1024 bytes apart (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t umma_desc(unsigned smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);      // start address, 16-byte units
    d |= (uint64_t)1 << 16;                            // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                  // stride byte offset: 8 rows x 128 B
    d |= (uint64_t)1 << 46;                            // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                            // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ void umma_tf32(unsigned tmem_d, uint64_t da, uint64_t db, unsigned idesc, unsigned accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
        "}\n"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
// round to nearest TF32 (10-bit mantissa); the result is an fp32 value with the low 13 mantissa bits clear
// x rounded to TF32 (nearest, ties away from zero: what cvt.rna.tf32.f32 computes for finite values).  That instruction is
// emulated on sm_100a (FSETP + SEL + LOP3 + IADD); the converter warps are what bounds these kernels, and on the bit
// pattern the rounding is one add and one mask.  (The lo part is rounded the same way: left to the tensor core's own
// truncation of the low 13 bits, the 64-px task lost its thin parity margin at B = 100, profiles/r2q_cheap_split.txt.)
__device__ __forceinline__ float tf32_rn(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }
__device__ __forceinline__ void umma_commit(unsigned long long* b) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(b)) : "memory");
}

__global__ void __launch_bounds__(kTcThreads, 1) gemm_tf32x3_kernel(const __grid_constant__ TcArgs a) {
    extern __shared__ __align__(1024) unsigned char tc_raw[];
    __shared__ unsigned long long full[kTcStages], conv[kTcStages], empty[kTcStages], done;
    __shared__ unsigned tmem_slot;
    // 1024-byte aligned tiles (swizzle atom = 8 rows x 128 B)
    unsigned char* base = tc_raw + ((1024u - (smem_u32(tc_raw) & 1023u)) & 1023u);
    const int BN = a.BN;
    const unsigned a_bytes = kTcBM * 128u, b_bytes = (unsigned)BN * 128u;
    const unsigned stage_bytes = 2 * a_bytes + 2 * b_bytes;              // A_hi | A_lo | B_hi | B_lo
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * kTcBM, n0 = blockIdx.x * BN;
    const int nkb = (a.K + kTcBK - 1) / kTcBK;
    const int kb0 = blockIdx.z * a.kb_per;
    const int count = min(a.kb_per, nkb - kb0);

    if (threadIdx.x == 0) {
        for (int s = 0; s < kTcStages; ++s) { mbar_init(&full[s], 1); mbar_init(&conv[s], kTcConv); mbar_init(&empty[s], 1); }
        mbar_init(&done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {                                                     // TMEM: 256 fp32 columns x 128 lanes
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(256u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int i = 0; i < count; ++i) {
                const int s = i % kTcStages;
                mbar_wait(&empty[s], (((unsigned)(i / kTcStages)) & 1u) ^ 1u);
                unsigned char* st = base + (size_t)s * stage_bytes;
                const unsigned bar = smem_u32(&full[s]);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(a_bytes + b_bytes) : "memory");
                const int k = (kb0 + i) * kTcBK;
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                             ::"r"(smem_u32(st)), "l"(reinterpret_cast<uint64_t>(&a.tmA)), "r"(k), "r"(m0), "r"(bar) : "memory");
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                             ::"r"(smem_u32(st + 2 * a_bytes)), "l"(reinterpret_cast<uint64_t>(&a.tmB)), "r"(k), "r"(n0), "r"(bar) : "memory");
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both K-major, N >> 3, M >> 4
            const unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(BN >> 3) << 17) | ((unsigned)(kTcBM >> 4) << 24);
            for (int i = 0; i < count; ++i) {
                const int s = i % kTcStages;
                mbar_wait(&conv[s], ((unsigned)(i / kTcStages)) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const unsigned sa = smem_u32(base + (size_t)s * stage_bytes);
                const unsigned sa_lo = sa + a_bytes, sb = sa + 2 * a_bytes, sb_lo = sb + b_bytes;
#pragma unroll
                for (int k = 0; k < kTcBK / 8; ++k) {                    // UMMA_K = 8 TF32 = 32 bytes inside the swizzle span
                    const uint64_t ah = umma_desc(sa + 32u * k), al = umma_desc(sa_lo + 32u * k);
                    const uint64_t bh = umma_desc(sb + 32u * k), bl = umma_desc(sb_lo + 32u * k);
                    umma_tf32(tmem, ah, bh, idesc, (i > 0 || k > 0) ? 1u : 0u);
                    umma_tf32(tmem, ah, bl, idesc, 1u);
                    umma_tf32(tmem, al, bh, idesc, 1u);
                }
                umma_commit(&empty[s]);                                  // frees the stage when these MMAs retire
            }
            umma_commit(&done);
        }
    } else {
        const int t = threadIdx.x - 64;                                  // 0..255
        for (int i = 0; i < count; ++i) {
            const int s = i % kTcStages;
            mbar_wait(&full[s], ((unsigned)(i / kTcStages)) & 1u);
            float4* st = reinterpret_cast<float4*>(base + (size_t)s * stage_bytes);
            const int a4 = a_bytes / 16, b4 = b_bytes / 16;
            // A_hi | A_lo occupy [0, a4) and [a4, 2 a4); B_hi | B_lo follow.  Four loads in flight per thread: one load per
            // trip left the converters -- which bound this kernel -- waiting out a shared-memory round trip per 16 bytes
            for (int e0 = t; e0 < a4 + b4; e0 += 4 * kTcConv) {
                float4 x[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int e = e0 + u * kTcConv;
                    if (e < a4 + b4) x[u] = *(e < a4 ? st + e : st + 2 * a4 + (e - a4));
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int e = e0 + u * kTcConv;
                    if (e < a4 + b4) {
                        float4* hi = e < a4 ? st + e : st + 2 * a4 + (e - a4);
                        float4* lo = e < a4 ? hi + a4 : hi + b4;
                        float4 h, l;
                        h.x = tf32_rn(x[u].x); l.x = tf32_rn(x[u].x - h.x);
                        h.y = tf32_rn(x[u].y); l.y = tf32_rn(x[u].y - h.y);
                        h.z = tf32_rn(x[u].z); l.z = tf32_rn(x[u].z - h.z);
                        h.w = tf32_rn(x[u].w); l.w = tf32_rn(x[u].w - h.w);
                        *hi = h;
                        *lo = l;
                    }
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> visible to the tensor core
            mbar_arrive(&conv[s]);
        }
        // ---- epilogue: this warp's 32 TMEM lanes = rows m0 + 32*(warp%4) + lane ----
        mbar_wait(&done, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int q = warp & 3;
        const int row = m0 + q * 32 + lane;
        float* out = a.partials + ((size_t)blockIdx.z * a.M + row) * a.N + n0;
        const int chalf = ((BN / 16 + 1) / 2) * 16;                      // warps 2..5 take columns [0, chalf), warps 6..9 the rest
        for (int c = warp < 6 ? 0 : chalf; c < (warp < 6 ? chalf : BN); c += 16) {
            unsigned v[16];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                           "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                         : "r"(tmem + ((unsigned)(q * 32) << 16) + (unsigned)c));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (row < a.M) {
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    const int n = n0 + c + j;
                    if (n + 3 < a.N) {
                        *reinterpret_cast<float4*>(out + c + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                              __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
                    } else {
                        for (int e = 0; e < 4; ++e)
                            if (n + e < a.N) out[c + j + e] = __uint_as_float(v[j + e]);
                    }
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u));
    }
}

// The same product for operands that feed a long ReLU / tanh chain (encoder.l1 FORWARD): the tensor core truncates its
// fp32 accumulator after every MMA, and 144 accumulations in one TMEM accumulator shrink every pre-activation by ~2e-6,
// which the position head amplifies past the gradient bar.  Here the hi.hi products of ONE stage
// (4 accumulations) go to a double-buffered accumulator that the converter warps drain after every stage into register
// sums with round-to-nearest FMAs (with the expected truncation loss of a 4-chain added back in the same FMA, see
// csrc/conv_tc.cu), and the small hi.lo + lo.hi corrections keep their own accumulator.  BN <= 112 (3 x BN TMEM columns,
// BN register sums per thread); the caller tiles N.
template <int BN>
__global__ void __launch_bounds__(kTcThreads, 1) gemm_tf32x3_drained_kernel(const __grid_constant__ TcArgs a) {
    extern __shared__ __align__(1024) unsigned char tc_raw[];
    __shared__ unsigned long long full[kTcStages], conv[kTcStages], empty[kTcStages], accfull[2], accfree[2];
    __shared__ unsigned tmem_slot;
    unsigned char* base = tc_raw + ((1024u - (smem_u32(tc_raw) & 1023u)) & 1023u);
    constexpr unsigned a_bytes = kTcBM * 128u, b_bytes = (unsigned)BN * 128u;
    constexpr unsigned stage_bytes = 2 * a_bytes + 2 * b_bytes;          // A_hi | A_lo | B_hi | B_lo
    constexpr unsigned kCols = 3 * BN <= 256 ? 256u : 512u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * kTcBM, n0 = blockIdx.x * BN;
    const int nkb = (a.K + kTcBK - 1) / kTcBK;
    const int kb0 = blockIdx.z * a.kb_per;
    const int count = min(a.kb_per, nkb - kb0);

    if (threadIdx.x == 0) {
        for (int s = 0; s < kTcStages; ++s) { mbar_init(&full[s], 1); mbar_init(&conv[s], 128); mbar_init(&empty[s], 1); }   // 128 converter threads
        for (int b = 0; b < 2; ++b) { mbar_init(&accfull[b], 1); mbar_init(&accfree[b], 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(kCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int i = 0; i < count; ++i) {
                const int s = i % kTcStages;
                mbar_wait(&empty[s], (((unsigned)(i / kTcStages)) & 1u) ^ 1u);
                unsigned char* st = base + (size_t)s * stage_bytes;
                const unsigned bar = smem_u32(&full[s]);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(a_bytes + b_bytes) : "memory");
                const int k = (kb0 + i) * kTcBK;
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                             ::"r"(smem_u32(st)), "l"(reinterpret_cast<uint64_t>(&a.tmA)), "r"(k), "r"(m0), "r"(bar) : "memory");
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                             ::"r"(smem_u32(st + 2 * a_bytes)), "l"(reinterpret_cast<uint64_t>(&a.tmB)), "r"(k), "r"(n0), "r"(bar) : "memory");
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(BN >> 3) << 17) | ((unsigned)(kTcBM >> 4) << 24);
            for (int i = 0; i < count; ++i) {
                const int s = i % kTcStages;
                const unsigned b = i & 1u, use = (unsigned)i >> 1;
                mbar_wait(&conv[s], ((unsigned)(i / kTcStages)) & 1u);
                if (use > 0) mbar_wait(&accfree[b], (use - 1) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const unsigned sa = smem_u32(base + (size_t)s * stage_bytes);
                const unsigned sa_lo = sa + a_bytes, sb = sa + 2 * a_bytes, sb_lo = sb + b_bytes;
                const unsigned d_main = tmem + b * BN, d_corr = tmem + 2 * BN;
#pragma unroll
                for (int k = 0; k < kTcBK / 8; ++k) {
                    const uint64_t ah = umma_desc(sa + 32u * k), al = umma_desc(sa_lo + 32u * k);
                    const uint64_t bh = umma_desc(sb + 32u * k), bl = umma_desc(sb_lo + 32u * k);
                    umma_tf32(d_main, ah, bh, idesc, k > 0 ? 1u : 0u);
                    umma_tf32(d_corr, ah, bl, idesc, (i > 0 || k > 0) ? 1u : 0u);
                    umma_tf32(d_corr, al, bh, idesc, 1u);
                }
                umma_commit(&empty[s]);
                umma_commit(&accfull[b]);
            }
        }
    } else if (warp >= 6) {
        // ===== converters (warps 6..9): split each landed stage in place into hi | lo =====
        const int t = threadIdx.x - 192;                                 // 0..127
        for (int i = 0; i < count; ++i) {
            const int s = i % kTcStages;
            mbar_wait(&full[s], ((unsigned)(i / kTcStages)) & 1u);
            float4* st = reinterpret_cast<float4*>(base + (size_t)s * stage_bytes);
            constexpr int a4 = a_bytes / 16, b4 = b_bytes / 16;
            for (int e0 = t; e0 < a4 + b4; e0 += 5 * 128) {              // 15 chunks per thread: 3 trips of 5 loads in flight
                float4 x[5];
#pragma unroll
                for (int u = 0; u < 5; ++u) {
                    const int e = e0 + u * 128;
                    if (e < a4 + b4) x[u] = *(e < a4 ? st + e : st + 2 * a4 + (e - a4));
                }
#pragma unroll
                for (int u = 0; u < 5; ++u) {
                    const int e = e0 + u * 128;
                    if (e < a4 + b4) {
                        float4* hi = e < a4 ? st + e : st + 2 * a4 + (e - a4);
                        float4* lo = e < a4 ? hi + a4 : hi + b4;
                        float4 h, l;
                        h.x = tf32_rn(x[u].x); l.x = tf32_rn(x[u].x - h.x);
                        h.y = tf32_rn(x[u].y); l.y = tf32_rn(x[u].y - h.y);
                        h.z = tf32_rn(x[u].z); l.z = tf32_rn(x[u].z - h.z);
                        h.w = tf32_rn(x[u].w); l.w = tf32_rn(x[u].w - h.w);
                        *hi = h;
                        *lo = l;
                    }
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(&conv[s]);
        }
    } else {
        // ===== drainers (warps 2..5, TMEM lane quarter = warp % 4): after every stage the hi.hi accumulator goes into
        // register sums; in their own warps the drain of stage i overlaps the conversion of stage i+1 and the MMAs of i+1 =====
        const int q = warp & 3;
        const unsigned lane_base = tmem + ((unsigned)(q * 32) << 16);
        float sum[BN];
#pragma unroll
        for (int c = 0; c < BN; ++c) sum[c] = 0.f;
        for (int i = 0; i < count; ++i) {
            const unsigned b = (unsigned)i & 1u, use = (unsigned)i >> 1;
            mbar_wait(&accfull[b], use & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int c = 0; c < BN; c += 16) {
                unsigned v[16];
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                             : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                               "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                             : "r"(lane_base + b * BN + (unsigned)c));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 16; ++j) sum[c + j] = fmaf(__uint_as_float(v[j]), 1.00000012f, sum[c + j]);   // chain of 4: see header
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(&accfree[b]);
        }
        // the last commit also covered the correction accumulator
#pragma unroll
        for (int c = 0; c < BN; c += 16) {
            unsigned v[16];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                           "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                         : "r"(lane_base + 2 * BN + (unsigned)c));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 16; ++j) sum[c + j] += __uint_as_float(v[j]);
        }
        const int row = m0 + q * 32 + lane;
        if (row < a.M) {
            float* out = a.partials + ((size_t)blockIdx.z * a.M + row) * a.N + n0;
            // 16-byte stores where the row allows (ncu: lg_throttle was this kernel's third stall -- 112 scalar stores per
            // thread, every one touching 32 sectors per warp)
            const bool vec = (a.N % 4) == 0 && (reinterpret_cast<uintptr_t>(a.partials) % 16) == 0;
#pragma unroll
            for (int c = 0; c < BN; c += 4) {
                if (vec && n0 + c + 3 < a.N) {
                    *reinterpret_cast<float4*>(out + c) = make_float4(sum[c], sum[c + 1], sum[c + 2], sum[c + 3]);
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (n0 + c + e < a.N) out[c + e] = sum[c + e];
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kCols));
    }
}

// dst[c][r] = src[r][c]   (32x32 tiles through shared memory)
__global__ void __launch_bounds__(256) transpose_kernel(const float* __restrict__ src, float* __restrict__ dst, int R, int C) {
    __shared__ float tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int j = ty; j < 32; j += 8)
        if (r0 + j < R && c0 + tx < C) tile[j][tx] = src[(long)(r0 + j) * C + c0 + tx];
    __syncthreads();
    for (int j = ty; j < 32; j += 8)
        if (c0 + j < C && r0 + tx < R) dst[(long)(c0 + j) * R + r0 + tx] = tile[tx][j];
}

namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn tc_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}
// row-major [rows][K] fp32, box = 32 floats (128 B, one swizzle span) x box_rows
bool tc_map(CUtensorMap* tm, const float* p, int rows, int K, int box_rows) {
    EncodeTiledFn fn = tc_encode_fn();
    if (!fn) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)K * 4};
    const cuuint32_t box[2] = {(cuuint32_t)kTcBK, (cuuint32_t)box_rows};
    const cuuint32_t es[2] = {1, 1};
    return fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)p, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
}  // namespace

int transpose(const float* src, float* dst, int R, int C, cudaStream_t st) {
    launch(transpose_kernel, dim3(cdiv(C, 32), cdiv(R, 32)), dim3(256), 0, st, src, dst, R, C);
    return check_launch("transpose");
}

// partials[z][M][N] = A[M, kz] . B[N, kz]^T for every K split z.  Returns the number of splits (>= 1) on success, 0 on
// a launch error (paig_last_error), -1 when the shape does not qualify (caller uses the CUDA-core GEMM).
// Drained variant (see gemm_tf32x3_drained_kernel): N tiled by 112 or 96 columns, fixed K split of 12 blocks.
int gemm_tc_partials_drained(const float* A, const float* B, int M, int N, int K, float* partials, size_t partial_floats,
                             const char* tag, cudaStream_t st) {
    static const bool off = getenv("PAIG_NO_TCGEN05") != nullptr;
    // (any M: rows past the end of A arrive as zeros and are not written -- a sequence must get bit-identical results
    // alone or inside a large batch, tests/test_gpu_module.py::test_eval_batch_sweep_properties, so small batches must
    // not switch to another GEMM)
    if (off || M < 1 || N < 96 || K < 128) return -1;
    if ((K % 4) != 0 || ((uintptr_t)A % 16) || ((uintptr_t)B % 16)) return -1;
    TcArgs a;
    memset(&a, 0, sizeof(a));
    a.M = M; a.N = N; a.K = K;
    a.BN = (N % 112 == 0 || N > 96) ? 112 : 96;
    const int nkb = cdiv(K, kTcBK);
    int splits = nkb >= 24 ? nkb / 12 : 1;                               // by K only: batch-invariant summation order
    while (splits > 1 && (size_t)splits * M * N > partial_floats) --splits;
    if ((size_t)splits * M * N > partial_floats) return -1;
    a.kb_per = cdiv(nkb, splits);
    splits = cdiv(nkb, a.kb_per);
    a.partials = partials;
    if (!tc_map(&a.tmA, A, M, K, kTcBM) || !tc_map(&a.tmB, B, N, K, a.BN)) return -1;
    const size_t smem = (size_t)kTcStages * (2 * kTcBM * 128 + 2 * (size_t)a.BN * 128) + 1024;
    const dim3 grid(cdiv(N, a.BN), cdiv(M, kTcBM), splits);
    if (a.BN == 112) launch(gemm_tf32x3_drained_kernel<112>, grid, dim3(kTcThreads), smem, st, a);
    else launch(gemm_tf32x3_drained_kernel<96>, grid, dim3(kTcThreads), smem, st, a);
    if (check_launch(tag ? tag : "gemm_tf32x3_drained")) return 0;
    return splits;
}

int gemm_tc_partials(const float* A, const float* B, int M, int N, int K, bool fixed_split, float* partials,
                     size_t partial_floats, const char* tag, cudaStream_t st) {
    static const bool off = getenv("PAIG_NO_TCGEN05") != nullptr;
    if (off || M < 128 || N < 128 || K < 128) return -1;
    if ((K % 4) != 0 || ((uintptr_t)A % 16) || ((uintptr_t)B % 16) || (N % 4) != 0) return -1;
    TcArgs a;
    memset(&a, 0, sizeof(a));
    a.M = M; a.N = N; a.K = K;
    a.BN = N <= 256 ? ((N + 15) / 16) * 16 : 256;
    const int nkb = cdiv(K, kTcBK);
    const int tiles = cdiv(M, kTcBM) * cdiv(N, a.BN);
    int splits;
    if (fixed_split) splits = nkb >= 24 ? nkb / 12 : 1;                 // by K only: batch-invariant summation order
    else splits = tiles >= 120 ? 1 : (148 / tiles > 0 ? 148 / tiles : 1);     // one resident wave: 24 tiles x 7 splits = 168 CTAs ran as 148 + 20
    if (splits > nkb / 2) splits = nkb / 2 > 0 ? nkb / 2 : 1;
    while (splits > 1 && (size_t)splits * M * N > partial_floats) --splits;
    if ((size_t)splits * M * N > partial_floats) return -1;
    a.kb_per = cdiv(nkb, splits);
    splits = cdiv(nkb, a.kb_per);
    a.partials = partials;
    if (!tc_map(&a.tmA, A, M, K, kTcBM) || !tc_map(&a.tmB, B, N, K, a.BN)) return -1;
    const size_t smem = (size_t)kTcStages * (2 * kTcBM * 128 + 2 * (size_t)a.BN * 128) + 1024;
    launch(gemm_tf32x3_kernel, dim3(cdiv(N, a.BN), cdiv(M, kTcBM), splits), dim3(kTcThreads), smem, st, a);
    if (check_launch(tag ? tag : "gemm_tf32x3")) return 0;
    return splits;
}

}  // namespace paig

#else   // PAIG_EMU: the tensor-core path needs the device; the SIMT-on-CPU shim runs the CUDA-core GEMM instead

namespace paig {
int transpose(const float*, float*, int, int, cudaStream_t) { return 1; }
int gemm_tc_partials(const float*, const float*, int, int, int, bool, float*, size_t, const char*, cudaStream_t) { return -1; }
int gemm_tc_partials_drained(const float*, const float*, int, int, int, float*, size_t, const char*, cudaStream_t) { return -1; }
}  // namespace paig

#endif
