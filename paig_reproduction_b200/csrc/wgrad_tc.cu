// Weight (and bias) gradient of a 3x3 "same" convolution on the 5th-generation tensor cores, for the layers of the 64-px
// UNet whose channel counts make it a dense contraction (blocks.py:113-170: c4 .. c15, 32 .. 128 channels):
//
//      dW[co][ci][ky][kx] = sum_{f,y,x} g[f][co][y][x] . in[f][ci][y+ky-1][x+kx-1]              (zero padding)
//
// The contraction runs over pixels, so pixels are the K dimension of the MMA and BOTH operands are K-major with one
// row per channel.  A tap is a relative shift of the two operands; splitting it as "ky on the gradient, kx on the
// input" turns the nine taps into ONE product of two stacked operands,
//
//      D[(ky, co), (kx, ci)] = sum_K  G_ky[co][K] . X_kx[ci][K],   G_ky = g shifted by 1-ky rows, X_kx = in shifted by kx-1 columns
//
// with M = 3 Cout rows (tiles of 128) and N = 3 x 32 input-channel columns (+ 16: the bias column) per CTA: MMAs of
// 128 x 224 x 8 and 128 x 112 x 8 instead of nine products 128 x Cin x 8, which is what keeps the tensor pipe fed (an
// M = 128 tf32 MMA costs ~100 cycles of operand-A fetch whatever its N).
//   * K block = 32 pixels = one 128-byte swizzle span: one image row at 32 px, half a row at 64 px, 2 / 4 rows at 16 / 8 px.
//   * Raw boxes arrive through two 4-D tensor maps (no swizzle) into a small ring of landing stages.  The row shift of G_ky
//     is the box's y coordinate; the TMA unit's out-of-bounds zero fill is the padding.
//   * The innermost TMA coordinate must stay 16-byte aligned (tools/tma_probe.cu), so the column shift cannot be a
//     coordinate: the input arrives once per K block as a raw box with a 4-pixel halo on both sides.  Converter warps
//     split every raw element into hi | lo and write the MMA tiles in the 128-byte-swizzled K-major layout themselves,
//     16 bytes at a time: a chunk of a gradient row goes to the chunk index XOR (row & 7); a thread splits the six raw
//     neighbours of four input pixels once and stores the three column-shifted chunks from registers (conflict free).
//   * 3xTF32 (hi.hi + hi.lo + lo.hi, split as in gemm_tc.cu) in two MMAs per K step: [main | corr] += G_hi x [X_hi | X_lo]
//     (N doubled, one fetch of G_hi -- the kernel is bound by the MMAs' shared-memory operand fetch) and corr += G_lo x X_hi.
//     The tensor core accumulates with truncation, so the accumulators run in chains of `chain` K blocks that the
//     converter warps drain into fp32 registers (round to nearest).
//   * The bias gradient rides along as one extra column: a row of ones appended to the X tile; its products with the
//     unshifted gradient rows (ky = 1) are sum g.
//   * grid = (K splits, M tiles, N tiles); every CTA writes its share of partial `split` in the final
//     [co][ci][3][3] | [co] layout, folded in fixed order by the shared reduce_partials pass (bit-reproducible).
//   * All loop bookkeeping of the single-thread roles is incremental: with `kb / bpf`, `%` by run-time values in the TMA
//     producer's loop that one thread's trip time (~1900 cycles) bounded the kernel.
// Measured bounds and the experiments behind the constants below: profiles/r2w_wgrad_tc_probe.txt.
#include "common.cuh"
#include "internal.h"

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#ifndef PAIG_EMU
#include <cuda.h>

namespace paig {
namespace {

constexpr int kWcConvWarps = 12;       // converter warps (8 and 12 measure the same: the kernel is bound by shared-memory bandwidth)
constexpr int kWcThreads = 64 + 32 * kWcConvWarps;   // warp 0 TMA producer, warp 1 MMA issuer / TMEM owner, then the converters
constexpr int kWcConv = 32 * kWcConvWarps;
constexpr int kWcDrain = 256;          // the first eight converter warps also drain the main accumulator
constexpr int kWcStages = 3;           // converted (hi | lo) operand stages the MMAs read (2 with a deeper raw ring measures the same)
constexpr int kWcMaxRaw = 6;           // raw TMA landing stages
constexpr int kWcMaxCt = 32;           // input channels per N tile
constexpr unsigned kWcABytes = 128 * 128;

struct WgTcArgs {
    CUtensorMap tmG, tmIn;
    float* partials;
    int Cin, Cout, S;
    int bx, by, lbx;          // K block = bx x by pixels (= 32); lbx = log2(bx)
    int bpr, bpf;             // K blocks per row group (S / bx) and per frame (S*S / 32)
    long total_blocks;        // frames * bpf
    int splits, chain, nraw, nconv, dbg;
    int Ct;                   // input channels per N tile
    int stride;               // floats per partial: 9 Cin Cout + Cout
};

__device__ __forceinline__ unsigned wc_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void wc_init(unsigned long long* b, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(wc_u32(b)), "r"(count));
}
__device__ __forceinline__ void wc_wait(unsigned long long* b, unsigned parity) {
    const unsigned a = wc_u32(b);
    unsigned done = 0;
    while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ void wc_arrive(unsigned long long* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(wc_u32(b)) : "memory");
}
// K-major tile, rows of 128 bytes, 128-byte swizzle, 8-row groups 1024 bytes apart (as gemm_tc.cu)
__device__ __forceinline__ uint64_t wc_desc(unsigned smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ void wc_mma(unsigned tmem_d, uint64_t da, uint64_t db, unsigned idesc, unsigned accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
        "}\n"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void wc_commit(unsigned long long* b) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(wc_u32(b)) : "memory");
}
__device__ __forceinline__ float wc_rn(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }
__device__ __forceinline__ void wc_tma4(void* dst, const CUtensorMap* tm, int x, int y, int c, int n, unsigned long long* bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
        ::"r"(wc_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(x), "r"(y), "r"(c), "r"(n), "r"(wc_u32(bar))
        : "memory");
}
__device__ __forceinline__ void wc_ld8(unsigned taddr, float (&v)[8]) {
    unsigned r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(r[e]);
}

constexpr int kWcAccMax = (3 * kWcMaxCt + 16) / 2;    // accumulator columns per drainer thread

__global__ void __launch_bounds__(kWcThreads, 1) conv3x3_wgrad_tc_kernel(const __grid_constant__ WgTcArgs a) {
    extern __shared__ __align__(1024) unsigned char wc_raw[];
    __shared__ unsigned long long rfull[kWcMaxRaw], rempty[kWcMaxRaw], conv[kWcStages], empty[kWcStages], chain_done, drained;
    __shared__ unsigned tmem_slot;
    unsigned char* base = wc_raw + ((1024u - (wc_u32(wc_raw) & 1023u)) & 1023u);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Ct = a.Ct, rowsB = 3 * Ct, NT = rowsB + 16;
    const int m0 = blockIdx.y * 128, c0 = blockIdx.z * Ct;
    const int rowsA = min(128, 3 * a.Cout - m0), nA = rowsA >> 5;            // gradient boxes of 32 channels
    const unsigned b_bytes = (unsigned)NT * 128u;
    const int rawx = a.bx + 8;                                               // raw input row: 4-pixel halo either side
    const unsigned rawx_bytes = ((unsigned)(Ct * a.by * rawx * 4) + 1023u) & ~1023u;
    const unsigned raw_bytes = kWcABytes + rawx_bytes;                       // raw stage: gradient boxes | input box
    const unsigned stage_bytes = 2 * kWcABytes + 2 * b_bytes;                // MMA stage: G_hi | G_lo | X_hi | X_lo
    const int nconv = a.nconv;
    unsigned char* rbase = base + nconv * (size_t)stage_bytes;
    const int nraw = a.nraw;
    const long kb0 = a.total_blocks * blockIdx.x / a.splits, kb1 = a.total_blocks * (blockIdx.x + 1) / a.splits;
    const int count = (int)(kb1 - kb0);

    if (threadIdx.x == 0) {
        for (int s = 0; s < kWcStages; ++s) { wc_init(&conv[s], kWcConv); wc_init(&empty[s], 1); }
        for (int s = 0; s < kWcMaxRaw; ++s) { wc_init(&rfull[s], 1); wc_init(&rempty[s], kWcConv); }
        wc_init(&chain_done, 1);
        wc_init(&drained, kWcDrain);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(wc_u32(&tmem_slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    // constant rows: gradient rows past 3 Cout (M padding) are zero; the 16 rows after the X tile are [ones | 15 x zero]
    for (int s = 0; s < nconv; ++s) {
        float4* st = reinterpret_cast<float4*>(base + (size_t)s * stage_bytes);
        for (int e = threadIdx.x; e < (128 - rowsA) * 8; e += kWcThreads) {
            st[rowsA * 8 + e] = make_float4(0.f, 0.f, 0.f, 0.f);
            st[kWcABytes / 16 + rowsA * 8 + e] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float4* xh = st + 2 * (kWcABytes / 16) + rowsB * 8;
        float4* xl = xh + b_bytes / 16;
        for (int e = threadIdx.x; e < 16 * 8; e += kWcThreads) {
            const float v = e < 8 ? 1.f : 0.f;
            xh[e] = make_float4(v, v, v, v);
            xl[e] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = tmem_slot;
    const int chain = a.chain;

    if (warp == 0) {
        if (lane == 0) {
            // no divisions inside the loop: this one thread's trip time bounds the whole pipeline (with kb / bpf etc. per
            // block it took ~1900 cycles per K block, more than the MMAs)
            int box_ky[4], box_co[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) { box_ky[j] = (m0 + 32 * j) / a.Cout; box_co[j] = (m0 + 32 * j) % a.Cout; }
            int f = (int)(kb0 / a.bpf), r = (int)(kb0 % a.bpf);
            int xb = r % a.bpr, yb = r / a.bpr;
            const unsigned tx = (unsigned)nA * 4096u + (unsigned)(Ct * a.by * rawx * 4);
            int s = 0;
            unsigned ph = 1;
            for (int i = 0; i < count; ++i) {
                wc_wait(&rempty[s], ph);
                unsigned char* st = rbase + (size_t)s * raw_bytes;
                const int x0 = xb * a.bx, y0 = yb * a.by;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(wc_u32(&rfull[s])), "r"(tx) : "memory");
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (j < nA)                                              // anchor row y pairs with g row y - (ky - 1)
                        wc_tma4(st + (size_t)j * 4096, &a.tmG, x0, y0 + 1 - box_ky[j], box_co[j], f, &rfull[s]);
                wc_tma4(st + kWcABytes, &a.tmIn, x0 - 4, y0, c0, f, &rfull[s]);
                if (++s == nraw) { s = 0; ph ^= 1u; }
                if (++xb == a.bpr) {
                    xb = 0;
                    if (++yb * a.by == a.S) { yb = 0; ++f; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // per K step two MMAs instead of three: [main | corr] += G_hi x [X_hi | X_lo] (the two X tiles are adjacent rows:
            // one descriptor, N = 2 NT, one fetch of G_hi) and corr += G_lo x X_hi
            const unsigned idesc1 = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)((2 * NT) >> 3) << 17) | ((unsigned)(128 >> 4) << 24);
            const unsigned idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(NT >> 3) << 17) | ((unsigned)(128 >> 4) << 24);
            int in_chain = 0;
            unsigned drained_ph = 0;
            int s = 0;
            unsigned cph = 0;
            for (int i = 0; i < count; ++i) {
                wc_wait(&conv[s], cph);
                const bool first = in_chain == 0;
                if (first && i > 0) { wc_wait(&drained, drained_ph); drained_ph ^= 1u; }     // the main accumulator was read out
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const unsigned sa = wc_u32(base + (size_t)s * stage_bytes);
                const unsigned sa_lo = sa + kWcABytes, sb = sa + 2 * kWcABytes;
                if (!(a.dbg & 1))
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint64_t ah = wc_desc(sa + 32u * k), al = wc_desc(sa_lo + 32u * k), bh = wc_desc(sb + 32u * k);
                    wc_mma(tmem, ah, bh, idesc1, (first && k == 0) ? 0u : 1u);
                    wc_mma(tmem + (unsigned)NT, al, bh, idesc2, 1u);
                }
                wc_commit(&empty[s]);
                if (++in_chain == chain || i == count - 1) { wc_commit(&chain_done); in_chain = 0; }
                if (++s == nconv) { s = 0; cph ^= 1u; }
            }
        }
    } else {
        const int t = threadIdx.x - 64;                                      // 0 .. kWcConv-1
        const bool drainer = warp < 10;                                      // warps 2..9: (TMEM lane quarter, column half)
        const int q = warp & 3, h = ((warp - 2) >> 2) & 1;
        const int colsPer = NT >> 1;
        float acc[kWcAccMax];
#pragma unroll
        for (int e = 0; e < kWcAccMax; ++e) acc[e] = 0.f;
        const unsigned tbase = tmem + ((unsigned)(q * 32) << 16) + (unsigned)(h * colsPer);
        auto drain = [&](int c) {                                            // acc += main + corr of chain c (both restart per chain)
            wc_wait(&chain_done, (unsigned)c & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int j = 0; j < kWcAccMax / 8; ++j) {
                if (j * 8 < colsPer) {
                    float v[8], u[8];
                    wc_ld8(tbase + (unsigned)(j * 8), v);
                    wc_ld8(tbase + (unsigned)(NT + j * 8), u);
#pragma unroll
                    for (int e = 0; e < 8; ++e) acc[j * 8 + e] += v[e] + u[e];
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        };
        int pending = -1;
        const int lbx = a.lbx, bxm = a.bx - 1;
        int rs = 0, in_chain = 0, chain_no = 0, s = 0;
        unsigned rph = 0, eph = 1;
        for (int i = 0; i < count; ++i) {
            wc_wait(&rfull[rs], rph);
            wc_wait(&empty[s], eph);                                        // the MMAs that last read this stage retired
            unsigned char* st = base + (size_t)s * stage_bytes;
            const unsigned char* rst = rbase + (size_t)rs * raw_bytes;
            if (!(a.dbg & 2))
            {   // gradient tiles: dense [row][32 px] boxes -> hi | lo with the 16-byte chunks swizzled by the row; one
                // 16-byte chunk of every box per thread, the loads of all boxes in flight together
                const float4* gr = reinterpret_cast<const float4*>(rst);
                float4* gh = reinterpret_cast<float4*>(st);
                constexpr int kTrips = (4 * 256 + kWcConv - 1) / kWcConv;
                float4 x[kTrips];
#pragma unroll
                for (int j = 0; j < kTrips; ++j)
                    if (t + kWcConv * j < nA * 256) x[j] = gr[t + kWcConv * j];
#pragma unroll
                for (int j = 0; j < kTrips; ++j) {
                    if (t + kWcConv * j < nA * 256) {
                        const int e = t + kWcConv * j, ch = e & 7;
                        float4 hh, ll;
                        hh.x = wc_rn(x[j].x); ll.x = wc_rn(x[j].x - hh.x);
                        hh.y = wc_rn(x[j].y); ll.y = wc_rn(x[j].y - hh.y);
                        hh.z = wc_rn(x[j].z); ll.z = wc_rn(x[j].z - hh.z);
                        hh.w = wc_rn(x[j].w); ll.w = wc_rn(x[j].w - hh.w);
                        const int row = e >> 3, d = (row << 3) + (ch ^ (row & 7));
                        gh[d] = hh;
                        gh[d + kWcABytes / 16] = ll;
                    }
                }
            }
            if (!(a.dbg & 4))
            {   // input tiles: one 16-byte chunk (4 pixels of one channel) per thread and trip; its six raw neighbours are
                // split once and leave as the three column-shifted copies, 16-byte stores in the swizzled K-major layout
                const float* raw = reinterpret_cast<const float*>(rst + kWcABytes);
                unsigned char* xh = st + 2 * kWcABytes;
                unsigned char* xl = xh + b_bytes;
                for (int e = t; e < Ct * 8; e += kWcConv) {
                    const int c = e >> 3, ch = e & 7, p0 = ch * 4;
                    const float* src = raw + (c * a.by + (p0 >> lbx)) * rawx + (p0 & bxm) + 3;
                    const float4 mid = *reinterpret_cast<const float4*>(src + 1);
                    const float v[6] = {src[0], mid.x, mid.y, mid.z, mid.w, src[5]};
                    float hh[6], ll[6];
#pragma unroll
                    for (int k = 0; k < 6; ++k) { hh[k] = wc_rn(v[k]); ll[k] = wc_rn(v[k] - hh[k]); }
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        const unsigned row = (unsigned)(kx * Ct + c);
                        const unsigned off = row * 128u + (((unsigned)ch ^ (row & 7u)) << 4);
                        *reinterpret_cast<float4*>(xh + off) = make_float4(hh[kx], hh[kx + 1], hh[kx + 2], hh[kx + 3]);
                        *reinterpret_cast<float4*>(xl + off) = make_float4(ll[kx], ll[kx + 1], ll[kx + 2], ll[kx + 3]);
                    }
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            wc_arrive(&conv[s]);
            wc_arrive(&rempty[rs]);
            if (++rs == nraw) { rs = 0; rph ^= 1u; }
            if (++s == nconv) { s = 0; eph ^= 1u; }
            if (pending >= 0 && drainer) {                                   // the previous chain, drained behind this block's conversion
                drain(pending);
                wc_arrive(&drained);
            }
            pending = -1;
            if (++in_chain == chain || i == count - 1) { pending = chain_no++; in_chain = 0; }
        }
        if (drainer) {
            drain(pending);                                                  // the last block always closes a chain
        }
        // ---- write this CTA's share of partial blockIdx.x: rows = (ky, co), columns = (kx, ci) | ones ----
        const int R = m0 + q * 32 + lane;
        if (drainer && R < 3 * a.Cout) {
            const int ky = R / a.Cout, co = R % a.Cout;
            float* out = a.partials + (size_t)blockIdx.x * a.stride;
            float* ow = out + ((size_t)co * a.Cin + c0) * 9 + ky * 3;
#pragma unroll
            for (int e = 0; e < kWcAccMax; ++e) {
                const int col = h * colsPer + e;
                if (e < colsPer) {
                    const float v = acc[e];
                    if (col < rowsB) {
                        const int kx = col / Ct, ci = col - kx * Ct;
                        ow[ci * 9 + kx] = v;
                    } else if (col == rowsB && ky == 1 && blockIdx.z == 0) {
                        out[(size_t)a.Cout * a.Cin * 9 + co] = v;
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
    }
}

typedef CUresult (*WcEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
WcEncodeFn wc_encode_fn() {
    static WcEncodeFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (WcEncodeFn)p;
    }
    return fn;
}
int wc_sm_count() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

}  // namespace

// 0 ok, > 0 error, -1: the layer does not qualify (the caller's TMA / CUDA-core kernels run)
int conv3x3_wgrad_tc(const WgradArgs& w, float* dW, float* db, cudaStream_t st) {
    static const bool off = getenv("PAIG_NO_WGRAD_TC") != nullptr || getenv("PAIG_NO_TCGEN05") != nullptr;
    static const int min_s = getenv("PAIG_WGRAD_TC_MIN_S") ? atoi(getenv("PAIG_WGRAD_TC_MIN_S")) : 8;
    static const int chain_env = getenv("PAIG_WGRAD_TC_CHAIN") ? atoi(getenv("PAIG_WGRAD_TC_CHAIN")) : 16;
    if (off || w.N <= 0 || w.in_mask || w.act) return -1;
    const int S = w.S;
    if (S != 8 && S != 16 && S != 32 && S != 64) return -1;
    if (S < min_s) return -1;
    if (w.Cin % 32 || w.Cout % 32 || w.Cin > 128 || w.Cout > 128) return -1;
    if (((uintptr_t)w.in % 16) || ((uintptr_t)w.g % 16) || (w.in_bs % 4) || (w.g_bs % 4)) return -1;
    WcEncodeFn enc = wc_encode_fn();
    if (!enc) return -1;
    WgTcArgs a;
    memset(&a, 0, sizeof(a));
    a.Cin = w.Cin; a.Cout = w.Cout; a.S = S; a.partials = w.partials;
    a.bx = S >= 32 ? 32 : S;
    a.by = 32 / a.bx;
    a.lbx = a.bx == 32 ? 5 : (a.bx == 16 ? 4 : 3);
    a.bpr = S / a.bx;
    a.bpf = S * S / 32;
    a.total_blocks = (long)w.N * a.bpf;
    a.chain = chain_env > 0 ? chain_env : 16;
    static const int dbg = getenv("PAIG_WGRAD_TC_DBG") ? atoi(getenv("PAIG_WGRAD_TC_DBG")) : 0;   // timing experiments: 1 no MMAs, 2 / 4 no gradient / input conversion
    a.dbg = dbg;
    // input channels per N tile: the largest divisor of Cin that is a multiple of 16 and at most 64 (N = 3 Ct + 16 <= 208)
    a.Ct = 0;
    for (int ct = kWcMaxCt; ct >= 16; ct -= 16)
        if (w.Cin % ct == 0) { a.Ct = ct; break; }
    if (!a.Ct) return -1;
    const int n_tiles = w.Cin / a.Ct, m_tiles = cdiv(3 * w.Cout, 128);
    int splits = wc_sm_count() / (n_tiles * m_tiles);
    if (splits < 1) splits = 1;
    if (splits > kWgradMaxCtas) splits = kWgradMaxCtas;
    if ((long)splits > a.total_blocks) splits = (int)a.total_blocks;
    a.splits = splits;
    a.stride = w.Cout * w.Cin * 9 + w.Cout;
    const cuuint64_t gdims[4] = {(cuuint64_t)S, (cuuint64_t)S, (cuuint64_t)w.Cout, (cuuint64_t)w.N};
    const cuuint64_t gstr[3] = {(cuuint64_t)S * 4, (cuuint64_t)S * S * 4, (cuuint64_t)w.g_bs * 4};
    const cuuint32_t gbox[4] = {(cuuint32_t)a.bx, (cuuint32_t)a.by, 32, 1};        // dense [32 channels][32 pixels]
    const cuuint64_t idims[4] = {(cuuint64_t)S, (cuuint64_t)S, (cuuint64_t)w.Cin, (cuuint64_t)w.N};
    const cuuint64_t istr[3] = {(cuuint64_t)S * 4, (cuuint64_t)S * S * 4, (cuuint64_t)w.in_bs * 4};
    const cuuint32_t ibox[4] = {(cuuint32_t)(a.bx + 8), (cuuint32_t)a.by, (cuuint32_t)a.Ct, 1};
    const cuuint32_t es[4] = {1, 1, 1, 1};
    if (enc(&a.tmG, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)w.g, gdims, gstr, gbox, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return -1;
    if (enc(&a.tmIn, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)w.in, idims, istr, ibox, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return -1;
    const int NT = 3 * a.Ct + 16;
    const size_t raw_bytes = kWcABytes + (((size_t)a.Ct * a.by * (a.bx + 8) * 4 + 1023) & ~(size_t)1023);
    static const int raw_env = getenv("PAIG_WGRAD_TC_RAW") ? atoi(getenv("PAIG_WGRAD_TC_RAW")) : kWcMaxRaw;
    static const int conv_env = getenv("PAIG_WGRAD_TC_STAGES") ? atoi(getenv("PAIG_WGRAD_TC_STAGES")) : kWcStages;
    const size_t conv_stage = 2 * (size_t)kWcABytes + 2 * (size_t)NT * 128, budget = 226 * 1024 - 1024;
    a.nconv = (conv_env >= 3 && 3 * conv_stage + 2 * raw_bytes <= budget) ? 3 : 2;     // 8-px layers: the raw box is 8 KB, two stages
    const size_t conv_bytes = a.nconv * conv_stage;
    a.nraw = (int)((budget - conv_bytes) / raw_bytes);
    if (a.nraw > raw_env) a.nraw = raw_env;
    if (a.nraw > kWcMaxRaw) a.nraw = kWcMaxRaw;
    if (a.nraw < 2) return -1;
    const size_t smem = conv_bytes + a.nraw * raw_bytes + 1024;
    static const bool debug = getenv("PAIG_DEBUG") != nullptr;
    if (debug)
        fprintf(stderr, "[paig] wgrad_tc %d->%d S=%d N=%d tiles %dx%d (Ct=%d, MMA N=%d) splits=%d blocks/split=%.1f chain=%d stages=%d raw stages=%d smem=%zu\n",
                w.Cin, w.Cout, S, w.N, m_tiles, n_tiles, a.Ct, NT, splits, (double)a.total_blocks / splits, a.chain, a.nconv, a.nraw, smem);
    launch(conv3x3_wgrad_tc_kernel, dim3(splits, m_tiles, n_tiles), dim3(kWcThreads), smem, st, a);
    int rc = check_launch(layer_name("conv3x3_wgrad_tc", w.Cin, w.Cout, S));
    if (rc) return rc;
    const int nW = w.Cout * w.Cin * 9;
    if (w.defer) return w.defer->add(a.partials, splits, nW + w.Cout, nW, dW, w.Cout, db) ? 0 : 1;
    return reduce_partials(a.partials, splits, nW + w.Cout, nW, dW, w.Cout, db, st);
}

}  // namespace paig

#else   // PAIG_EMU

namespace paig {
int conv3x3_wgrad_tc(const WgradArgs&, float*, float*, cudaStream_t) { return -1; }
}  // namespace paig

#endif
