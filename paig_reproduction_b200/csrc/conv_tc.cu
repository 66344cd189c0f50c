// 3x3 "same" convolution (blocks.py:113-170: the 64-px UNet's layers, forward and data gradient) as an implicit GEMM on
// the 5th-generation tensor cores, with fp32 accuracy:
//
//      out[pixel, co] = sum_{tap, ci} in[pixel + tap, ci] * w[co, ci, tap]          M = pixels, N = Cout, K = 9 * Cin
//
// as a 3xTF32 product (x = hi + lo, hi = rn_tf32(x), lo = rn_tf32(x - hi);  a.b ~ hi.hi + hi.lo + lo.hi), the same
// split as csrc/gemm_tc.cu.  What makes it a convolution and not a GEMM with an im2col matrix:
//
//   * A super-tile of T tiles of 16 rows x 8 columns of output pixels (8-px images: 2 frames x 8 rows, rows interleaved)
//     is owned by one CTA.  Per chunk of 8 input channels the producer warp fetches each tile's HALOED input patch
//     (18 x 16 pixels around the tile, NCHW) with ONE 4-D TMA tensor-map box whose out-of-bounds zero fill is the
//     convolution's padding, and the chunk's 9 x 8 x N weights with one bulk copy from a pre-packed array.
//   * Four converter warps re-lay the patch out as [channel quad][row][10 pixels][4 channels] -- pixel pitch 16 bytes,
//     the K-major "no swizzle" core-matrix layout of tcgen05 (8 pixels x 16 bytes contiguous) -- writing the hi and the
//     lo tile in the same pass, and split the weights in place.  In this layout the 9 taps are 9 START ADDRESSES into
//     the same patch (row stride 160 bytes = the descriptor's stride-byte-offset, channel-quad planes = its
//     leading-byte-offset): no im2col copy, every input element is fetched and split once and multiplied 9 N times.
//   * One thread issues, per chunk and tile, 9 taps x 3 products of tcgen05.mma.kind::tf32 (M = 128, N = Cout, K = 8)
//     into TMEM.  The tensor core accumulates fp32 with truncation; a long chain of accumulations shrinks the result
//     coherently (measured on encoder.l1).  So the hi.hi products of ONE chunk (9 accumulations)
//     go to a double-buffered TMEM accumulator that four epilogue warps drain after every chunk (tcgen05.ld) and sum
//     in registers with round-to-nearest adds; the small correction products (hi.lo + lo.hi, 2^-11 of the result)
//     keep their own accumulator for the whole K range.  Bias / ReLU are applied in registers, outputs go straight to
//     NCHW global memory (32-byte row segments).
//
// Warp roles (320 threads, 1 CTA per SM, persistent over super-tiles): 0 producer, 1 MMA issuer (+ TMEM allocation),
// 2-5 converters, 6-9 epilogue (TMEM lane quarter = warp % 4).
#include "common.cuh"
#include "internal.h"

#include <cstdint>
#include <cstdlib>
#include <cstring>

#ifndef PAIG_EMU
#include <cuda.h>

namespace paig {

constexpr int kCtThreads = 352;
constexpr int kCtStages = 2;
constexpr int kCtRawTile = 10240;        // bytes: [<=2 frames][8 ch][<=18 rows][16 px] fp32 as TMA delivers it
constexpr int kCtRawSlots = 3;           // raw patches in flight (their own ring: a slot is free again once converted)
constexpr int kCtPlane = 3200;           // bytes: one channel quad of a converted tile, 20 virtual rows x 10 px x 16 B
constexpr int kCtCvTile = 4 * kCtPlane;  // hi quad0, hi quad1, lo quad0, lo quad1
constexpr int kCtRow = 160;              // bytes between virtual rows of a converted tile (10 pixels x 16 B)

struct ConvTcArgs {
    CUtensorMap tm;           // input view: dims (x, y, c, n), box (16, R+2, 8, fpt)
    const float* wpack;       // [Cin/8][9 taps][2 quads][N][4]
    const float* bias;        // nullable
    float* out; long out_bs;
    int S, Nframes, Cin, relu;
    int R;                    // image rows per tile (16, or 8 at S = 8)
    int fpt;                  // frames per tile (1, or 2 at S = 8)
    int tiles_x, tpg;         // tiles per row, tiles per frame group
    int ntiles, nsuper, drain;
    float comp;               // expected relative truncation loss of a drained partial, added back in the epilogue (0: off)
    int dbg;                  // timing experiments (PAIG_CONV_TC_DBG): 1 skip patch conversion, 2 skip weight split, 4 skip MMAs, 8 skip drains
};

__device__ __forceinline__ unsigned ct_smem(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ct_bar_init(unsigned long long* b, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ct_smem(b)), "r"(count));
}
__device__ __forceinline__ void ct_wait(unsigned long long* b, unsigned parity) {
    const unsigned a = ct_smem(b);
    unsigned done = 0;
    while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ void ct_arrive(unsigned long long* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ct_smem(b)) : "memory");
}
__device__ __forceinline__ void ct_commit(unsigned long long* b) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(ct_smem(b)) : "memory");
}
// K-major operand without swizzle: core matrix = 8 rows x 16 bytes, contiguous (128 B); lbo = bytes between the two
// 16-byte K chunks of one MMA (K = 8 tf32), sbo = bytes between 8-row groups (cute::UMMA::SmemDescriptor, INTERLEAVE)
__device__ __forceinline__ uint64_t ct_desc(unsigned addr, unsigned lbo, unsigned sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(lbo >> 4) << 16;
    d |= (uint64_t)(sbo >> 4) << 32;
    d |= (uint64_t)1 << 46;                       // descriptor version (Blackwell)
    return d;                                     // layout type 0: no swizzle
}
__device__ __forceinline__ void ct_mma(unsigned tmem_d, uint64_t da, uint64_t db, unsigned idesc, unsigned accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
        "}\n"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
// x rounded to TF32 (nearest, ties away from zero: what cvt.rna.tf32.f32 computes for finite values).  That instruction is
// emulated on sm_100a (FSETP + SEL + LOP3 + IADD); the converter warps are what bounds these kernels, and on the bit
// pattern the rounding is one add and one mask.  (The lo part is rounded the same way: left to the tensor core's own
// truncation of the low 13 bits, the 64-px task lost its thin parity margin at B = 100, profiles/r2q_cheap_split.txt.)
__device__ __forceinline__ float ct_tf32(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }
__device__ __forceinline__ void ct_split4(const float4 x, float4& h, float4& l) {
    h.x = ct_tf32(x.x); l.x = ct_tf32(x.x - h.x);
    h.y = ct_tf32(x.y); l.y = ct_tf32(x.y - h.y);
    h.z = ct_tf32(x.z); l.z = ct_tf32(x.z - h.z);
    h.w = ct_tf32(x.w); l.w = ct_tf32(x.w - h.w);
}
__device__ __forceinline__ void ct_ld16(unsigned taddr, unsigned (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void ct_ld16_nowait(unsigned taddr, unsigned (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
}

// tile index -> frame group, first image row / column of the tile
struct TilePos { int grp, y0, x0; };
__device__ __forceinline__ TilePos ct_tile(const ConvTcArgs& a, int tile) {
    TilePos p;
    p.grp = tile / a.tpg;
    const int r = tile - p.grp * a.tpg;
    const int ty = r / a.tiles_x;
    p.y0 = ty * a.R;
    p.x0 = (r - ty * a.tiles_x) * 8;
    return p;
}

// F2: 8-px images, a tile is two frames of 8 rows with their rows interleaved (compile-time so that the converters'
// index arithmetic is divisions by constants)
// DRAIN: taps whose hi.hi products share one TMEM accumulation chain before the epilogue warps drain it (9: once per
// chunk, 3: once per tap row, 1: every tap).  NB main accumulators rotate so that draining overlaps the next MMAs.
template <int N, int T, bool F2, int DRAIN>
__global__ void __launch_bounds__(kCtThreads, 1) conv3x3_tc_kernel(const __grid_constant__ ConvTcArgs a) {
    extern __shared__ __align__(1024) unsigned char ct_raw[];
    __shared__ unsigned long long rawfull[kCtRawSlots], rawfree[kCtRawSlots];
    __shared__ unsigned long long wfull[kCtStages], ready[kCtStages], empty[kCtStages];
    constexpr int NB = (512 / (T * N) - 1) >= 3 ? 3 : 2;                  // main accumulators next to the correction one
    constexpr int UNITS = 9 / DRAIN;                                      // drains per chunk
    __shared__ unsigned long long accfull[NB], accfree[NB], corrfree;
    __shared__ unsigned tmem_slot;
    // A pair of tiles that are horizontal neighbours (every tile pair at S >= 16) arrives as ONE box of 24 columns
    // (x0-4 .. x0+19): half the rows for the TMA unit to walk, which is what bounds the load rate of 64-byte rows
    constexpr bool PAIR = (T == 2) && !F2;
    constexpr int RAWW = PAIR ? 24 : 16;                                  // floats per row of a raw patch
    constexpr unsigned kRawSlot = PAIR ? 24u * 18u * 8u * 4u + 512u : (unsigned)T * kCtRawTile;   // 14336 | T * 10240
    constexpr unsigned kRawBytes = PAIR ? 24u * 18u * 8u * 4u : (F2 ? (unsigned)T * 10240u : (unsigned)T * 9216u);
    constexpr unsigned kWBytes = 9u * 2u * N * 16u;                       // one chunk of weights (hi or lo)
    constexpr unsigned kStage = T * kCtCvTile + 2 * kWBytes;              // converted tiles | W hi | W lo
    constexpr unsigned kCols = ((NB + 1) * T * N <= 128) ? 128u : ((NB + 1) * T * N <= 256 ? 256u : 512u);
    static_assert((NB + 1) * T * N <= 512, "TMEM: NB hi.hi buffers and one correction accumulator per tile");
    unsigned char* rawbase = ct_raw + ((1024u - (ct_smem(ct_raw) & 1023u)) & 1023u);
    unsigned char* base = rawbase + kCtRawSlots * kRawSlot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nchunks = a.Cin / 8;

    if (threadIdx.x == 0) {
        for (int r = 0; r < kCtRawSlots; ++r) { ct_bar_init(&rawfull[r], 1); ct_bar_init(&rawfree[r], 128); }
        for (int s = 0; s < kCtStages; ++s) { ct_bar_init(&wfull[s], 1); ct_bar_init(&ready[s], 128); ct_bar_init(&empty[s], 1); }
        for (int b = 0; b < NB; ++b) { ct_bar_init(&accfull[b], 1); ct_bar_init(&accfree[b], 128); }
        ct_bar_init(&corrfree, 128);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ct_smem(&tmem_slot)), "r"(kCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = tmem_slot;

    if (warp == 0) {
        // ===== patch producer: TMA tensor-map boxes into the raw ring, up to kCtRawSlots chunks ahead =====
        if (lane == 0) {
            unsigned it = 0;
            for (int st = blockIdx.x; st < a.nsuper; st += gridDim.x) {
                for (int kc = 0; kc < nchunks; ++kc, ++it) {
                    const int r = it % kCtRawSlots;
                    ct_wait(&rawfree[r], ((it / kCtRawSlots) & 1u) ^ 1u);
                    unsigned char* rb = rawbase + (size_t)r * kRawSlot;
                    const unsigned bar = ct_smem(&rawfull[r]);
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kRawBytes) : "memory");
#pragma unroll
                    for (int t = 0; t < (PAIR ? 1 : T); ++t) {
                        const TilePos p = ct_tile(a, min(st * T + t, a.ntiles - 1));
                        asm volatile(
                            "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                            ::"r"(ct_smem(rb + t * kCtRawTile)), "l"(reinterpret_cast<uint64_t>(&a.tm)), "r"(p.x0 - 4), "r"(p.y0 - 1),
                              "r"(kc * 8), "r"(p.grp * a.fpt), "r"(bar) : "memory");
                    }
                }
            }
        }
    } else if (warp == 10) {
        // ===== weight producer: one bulk copy of the chunk's 9 x 8 x N weights into the stage the MMAs released =====
        if (lane == 0) {
            unsigned it = 0;
            for (int st = blockIdx.x; st < a.nsuper; st += gridDim.x) {
                for (int kc = 0; kc < nchunks; ++kc, ++it) {
                    const int s = it % kCtStages;
                    ct_wait(&empty[s], ((it / kCtStages) & 1u) ^ 1u);
                    const unsigned bar = ct_smem(&wfull[s]);
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kWBytes) : "memory");
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(ct_smem(base + (size_t)s * kStage + T * kCtCvTile)),
                                   "l"(__cvta_generic_to_global(a.wpack + (size_t)kc * (kWBytes / 4))), "r"(kWBytes), "r"(bar) : "memory");
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            const unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(N >> 3) << 17) | ((unsigned)(128 >> 4) << 24);
            constexpr unsigned rowmul = F2 ? 2u : 1u;         // an image row is `fpt` virtual rows of the converted tile
            unsigned it = 0, sti = 0, un = 0;                 // chunk, super-tile and drain-unit counters of this CTA
            for (int st = blockIdx.x; st < a.nsuper; st += gridDim.x, ++sti) {
                for (int kc = 0; kc < nchunks; ++kc, ++it) {
                    const int s = it % kCtStages;
                    ct_wait(&ready[s], (it / kCtStages) & 1u);
                    if (kc == 0 && sti > 0) ct_wait(&corrfree, (sti - 1) & 1u);
                    const unsigned sb = ct_smem(base + (size_t)s * kStage);
                    const unsigned w_hi = sb + T * kCtCvTile, w_lo = w_hi + kWBytes;
#pragma unroll
                    for (int g = 0; g < UNITS; ++g, ++un) {
                        const unsigned b = un % NB, use = un / NB;
                        if (use > 0) ct_wait(&accfree[b], (use - 1) & 1u);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        if (!(a.dbg & 4)) {
#pragma unroll
                            for (int t = 0; t < T; ++t) {
                                const unsigned cv = sb + t * kCtCvTile;
                                const unsigned d_main = tmem + (b * T + t) * N, d_corr = tmem + (NB * T + t) * N;
#pragma unroll
                                for (int j = 0; j < DRAIN; ++j) {
                                    const unsigned tap = g * DRAIN + j;
                                    const unsigned dy = tap / 3, dx = tap - 3 * (tap / 3);
                                    const unsigned aoff = (dy * rowmul * 10u + dx) * 16u;
                                    const uint64_t ah = ct_desc(cv + aoff, kCtPlane, kCtRow);
                                    const uint64_t al = ct_desc(cv + 2 * kCtPlane + aoff, kCtPlane, kCtRow);
                                    const uint64_t bh = ct_desc(w_hi + tap * (2u * N * 16u), N * 16u, 128u);
                                    const uint64_t bl = ct_desc(w_lo + tap * (2u * N * 16u), N * 16u, 128u);
                                    ct_mma(d_main, ah, bh, idesc, j > 0 ? 1u : 0u);
                                    ct_mma(d_corr, ah, bl, idesc, (kc > 0 || tap > 0) ? 1u : 0u);
                                    ct_mma(d_corr, al, bh, idesc, 1u);
                                }
                            }
                        }
                        if (g == UNITS - 1) ct_commit(&empty[s]);   // the stage's smem is free once these MMAs have read it
                        ct_commit(&accfull[b]);                     // this unit's hi.hi sums (on the very last unit also corr) are final
                    }
                }
            }
        }
    } else if (warp < 6) {
        // ===== converters: NCHW patch -> [quad][row][10 px][4 ch] hi / lo; weights -> hi (in place) / lo =====
        const int ct = threadIdx.x - 64;                                  // 0..127
        constexpr int vrows = F2 ? 20 : 18;                               // virtual rows of a converted tile
        constexpr int raw_rows = F2 ? 10 : 18;
        unsigned it = 0;
        for (int st = blockIdx.x; st < a.nsuper; st += gridDim.x) {
            for (int kc = 0; kc < nchunks; ++kc, ++it) {
                const int s = it % kCtStages, r = it % kCtRawSlots;
                ct_wait(&rawfull[r], (it / kCtRawSlots) & 1u);                       // the patch has landed ...
                ct_wait(&empty[s], ((it / kCtStages) & 1u) ^ 1u);                    // ... and the MMAs are done with this stage
                unsigned char* sb = base + (size_t)s * kStage;
                const unsigned char* rb = rawbase + (size_t)r * kRawSlot;
                constexpr int per_tile = 2 * vrows * 10;
                for (int e = ct; e < ((a.dbg & 1) ? 0 : T * per_tile); e += 128) {
                    const int t = e / per_tile, rr = e - t * per_tile;
                    const int q = rr / (vrows * 10), pr = rr - q * (vrows * 10);
                    const int v = pr / 10, px = pr - v * 10;
                    const int f = F2 ? (v & 1) : 0, ry = F2 ? (v >> 1) : v;
                    const float* raw = reinterpret_cast<const float*>(rb + (PAIR ? 0 : t * kCtRawTile)) +
                                       ((f * 8 + q * 4) * raw_rows + ry) * RAWW + px + 3 + (PAIR ? 8 * t : 0);
                    constexpr int cs = raw_rows * RAWW;                   // floats between channels of the raw patch
                    float4 x = make_float4(raw[0], raw[cs], raw[2 * cs], raw[3 * cs]), h, l;
                    ct_split4(x, h, l);
                    unsigned char* dst = sb + t * kCtCvTile + q * kCtPlane + (v * 10 + px) * 16;
                    *reinterpret_cast<float4*>(dst) = h;
                    *reinterpret_cast<float4*>(dst + 2 * kCtPlane) = l;
                }
                ct_arrive(&rawfree[r]);                                   // (generic-proxy reads done: the slot may be refilled)
                ct_wait(&wfull[s], (it / kCtStages) & 1u);
                float4* wh = reinterpret_cast<float4*>(sb + T * kCtCvTile);
                float4* wl = reinterpret_cast<float4*>(sb + T * kCtCvTile + kWBytes);
                for (int e = ct; e < ((a.dbg & 2) ? 0 : (int)(kWBytes / 16)); e += 128) {
                    float4 h, l;
                    ct_split4(wh[e], h, l);
                    wh[e] = h;
                    wl[e] = l;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic writes -> visible to the tensor core
                ct_arrive(&ready[s]);
            }
        }
    } else if (warp < 10) {
        // ===== epilogue: drain the hi.hi accumulator after every chunk, sum in registers (round to nearest) =====
        const int q = warp & 3;                                           // TMEM lanes 32q .. 32q+31
        const int m = q * 32 + lane;                                      // pixel of the tile: virtual row m / 8, column m % 8
        const unsigned lane_base = tmem + ((unsigned)(q * 32) << 16);
        unsigned un = 0;
        for (int st = blockIdx.x; st < a.nsuper; st += gridDim.x) {
            float sum[T][N];
#pragma unroll
            for (int t = 0; t < T; ++t)
#pragma unroll
                for (int c = 0; c < N; ++c) sum[t][c] = 0.f;
            for (int kc = 0; kc < nchunks; ++kc) {
#pragma unroll
                for (int g = 0; g < UNITS; ++g, ++un) {
                    const unsigned b = un % NB, use = un / NB;
                    ct_wait(&accfull[b], use & 1u);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (!(a.dbg & 8)) {
#pragma unroll
                        for (int t = 0; t < T; ++t)
#pragma unroll
                            for (int c = 0; c < N; c += 32) {
                                if (c + 32 <= N) {                  // two loads in flight per wait
                                    unsigned v[16], u[16];
                                    ct_ld16_nowait(lane_base + (b * T + t) * N + c, v);
                                    ct_ld16_nowait(lane_base + (b * T + t) * N + c + 16, u);
                                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                                    for (int j = 0; j < 16; ++j) sum[t][c + j] += __uint_as_float(v[j]);
#pragma unroll
                                    for (int j = 0; j < 16; ++j) sum[t][c + 16 + j] += __uint_as_float(u[j]);
                                } else {
                                    unsigned v[16];
                                    ct_ld16(lane_base + (b * T + t) * N + c, v);
#pragma unroll
                                    for (int j = 0; j < 16; ++j) sum[t][c + j] += __uint_as_float(v[j]);
                                }
                            }
                    }
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    ct_arrive(&accfree[b]);
                }
            }
            // the last commit also covered the correction accumulator
#pragma unroll
            for (int t = 0; t < T; ++t)
#pragma unroll
                for (int c = 0; c < N; c += 16) {
                    unsigned v[16];
                    ct_ld16(lane_base + (NB * T + t) * N + c, v);
                    // The tensor core truncates (round toward zero) the sum of an MMA's 8 products and its fp32 accumulator:
                    // every drained partial is short by an EXPECTED kappa = 0.28 / 0.62 / 2.1 x 2^-23 of itself for chains
                    // of 1 / 3 / 9 accumulations (measured, profiles/r2d_conv_tc_sweep.txt).  With one MMA per drained
                    // partial (DRAIN = 1) the loss is proportional to each partial whatever its sign, so the sum of the
                    // partials is short by kappa x itself -- a coherent shrink of every pre-activation that a deep ReLU
                    // network amplifies ~1000x into the gradients (mnist_spring_color B = 100 fails the parity bar
                    // without it).  It is added back once, next to the small hi.lo + lo.hi correction, BEFORE the one
                    // round-to-nearest addition to the sum: kappa x sum is below half an ulp of sum, on its own it would
                    // always round away.  (Scaling single partials by 1 + 2^-23, the finest factor fp32 has, on a subset
                    // of the taps was measured too: it distorts the taps against each other and made parity worse.)
#pragma unroll
                    for (int j = 0; j < 16; ++j) sum[t][c + j] += fmaf(sum[t][c + j], a.comp, __uint_as_float(v[j]));
                }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            ct_arrive(&corrfree);
#pragma unroll
            for (int t = 0; t < T; ++t) {
                const int tile = st * T + t;
                if (tile >= a.ntiles) continue;
                const TilePos p = ct_tile(a, tile);
                const int v = m >> 3, x = p.x0 + (m & 7);
                const int f = F2 ? (v & 1) : 0, y = p.y0 + (F2 ? (v >> 1) : v);
                const int frame = p.grp * a.fpt + f;
                if (frame >= a.Nframes) continue;
                float* o = a.out + (size_t)frame * a.out_bs + (size_t)y * a.S + x;
                const size_t cs = (size_t)a.S * a.S;
#pragma unroll
                for (int c = 0; c < N; ++c) {
                    float r = sum[t][c] + (a.bias ? __ldg(a.bias + c) : 0.f);
                    if (a.relu) r = fmaxf(r, 0.f);
                    o[c * cs] = r;
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kCols));
    }
}

// wpack[kc][tap][quad][n][4] = B[k = 8 kc + 4 quad + j][n] at tap (dy, dx):
//   forward     w[n][k][dy][dx]          (w: [Cout = N][Cin = K][3][3])
//   transposed  w[k][n][2 - dy][2 - dx]  (w: the layer's [Cout_l = K][Cin_l = N][3][3]: data gradient)
__global__ void __launch_bounds__(256) conv_tc_pack_kernel(const float* __restrict__ w, float* __restrict__ dst, int K, int N,
                                                           int transposed) {
    const int total = K * N * 9;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int j = e & 3;
        int r = e >> 2;
        const int n = r % N; r /= N;
        const int quad = r & 1; r >>= 1;
        const int tap = r % 9;
        const int kc = r / 9;
        const int k = kc * 8 + quad * 4 + j;
        const int dy = tap / 3, dx = tap % 3;
        dst[e] = transposed ? w[((size_t)k * N + n) * 9 + (2 - dy) * 3 + (2 - dx)] : w[((size_t)n * K + k) * 9 + dy * 3 + dx];
    }
}

namespace {
typedef CUresult (*CtEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
CtEncodeFn ct_encode_fn() {
    static CtEncodeFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (CtEncodeFn)p;
    }
    return fn;
}

int ct_sm_count() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

// Taps whose hi.hi products share one TMEM accumulation chain.  Measured (profiles/r2d_conv_tc_sweep.txt, all-positive
// operands) mean signed relative error without compensation: -3e-8 / -7e-8 / -2.5e-7 for 1 / 3 / 9 (FMA kernel: 1e-9).
// Forward convolutions feed ReLU stacks whose coherent shrink the position head amplifies ~1000x into the gradients:
// uncompensated, or compensated tap by tap, the 64-px task misses the gradient bar at B = 100; with the expected loss of
// the whole sum added back once in the epilogue, chains of 1 and of 3 both pass it below 1e-4 without the noise allowance
// (profiles/r2m_parity_sweep*.txt).  They drain per tap row (3): 25.8 ms per mnist step against 27.6 with chains of 1.
// Data gradients only get scaled by 1 - 2.5e-7 (compensated to ~0): they drain per chunk (9).
int ct_drain(bool backward) {
    static const int f = getenv("PAIG_CONV_TC_DRAIN") ? atoi(getenv("PAIG_CONV_TC_DRAIN")) : 3;
    static const int b = getenv("PAIG_CONV_TC_DRAIN_BWD") ? atoi(getenv("PAIG_CONV_TC_DRAIN_BWD")) : 9;
    const int v = backward ? b : f;
    return v == 9 ? 9 : (v == 1 ? 1 : 3);
}
float ct_comp(int drain) {
    static const bool off = getenv("PAIG_CONV_TC_NOCOMP") != nullptr;
    static const float k1 = getenv("PAIG_CONV_TC_KAPPA") ? (float)atof(getenv("PAIG_CONV_TC_KAPPA")) : 0.28f;
    const float ulp = 1.1920929e-7f;
    return off ? 0.f : (drain == 1 ? k1 : (drain == 3 ? 0.62f : 2.1f)) * ulp;
}

template <int N, int T, bool F2>
void ct_launch2(const ConvTcArgs& a, int grid, cudaStream_t st) {
    const size_t raw_slot = (T == 2 && !F2) ? 24 * 18 * 8 * 4 + 512 : (size_t)T * kCtRawTile;
    const size_t smem = kCtRawSlots * raw_slot + kCtStages * ((size_t)T * kCtCvTile + 2 * (size_t)(9 * 2 * N * 16)) + 1024;
    switch (a.drain) {
        case 9: launch(conv3x3_tc_kernel<N, T, F2, 9>, dim3(grid), dim3(kCtThreads), smem, st, a); break;
        case 3: launch(conv3x3_tc_kernel<N, T, F2, 3>, dim3(grid), dim3(kCtThreads), smem, st, a); break;
        default: launch(conv3x3_tc_kernel<N, T, F2, 1>, dim3(grid), dim3(kCtThreads), smem, st, a);
    }
}
template <int N, int T>
void ct_launch(const ConvTcArgs& a, int grid, cudaStream_t st) {
    if (a.fpt == 2) ct_launch2<N, T, true>(a, grid, st);
    else ct_launch2<N, T, false>(a, grid, st);
}
}  // namespace

size_t conv_tc_scratch_floats(int Cin, int Cout) { return (size_t)9 * Cin * Cout; }

bool conv_tc_enabled() {
    static const bool off = getenv("PAIG_NO_CONV_TC") != nullptr || getenv("PAIG_NO_TCGEN05") != nullptr;
    return !off;
}

// -1: the layer does not qualify (caller runs the CUDA-core kernel); 0 ok; > 0 error
int conv3x3_tc(const ConvArgs& c, float* scratch, cudaStream_t st) {
    const int N = c.Cout, K = c.Cin, S = c.S;
    if (!conv_tc_enabled() || !scratch || c.mask) return -1;
    if (K % 8 || N % 16 || N > 128 || N < 16) return -1;
    // 16 output channels: an M=128 x N=16 MMA still reads a full 4 KB pixel slab from shared memory, the tensor pipe
    // idles and the kernel only ties with the FMA one (measured, r2d sweep) -- those layers keep the CUDA-core path
    static const int min_n = getenv("PAIG_CONV_TC_MIN_N") ? atoi(getenv("PAIG_CONV_TC_MIN_N")) : 32;
    if (N < min_n && !c.force_tc) return -1;
    if (S != 8 && S != 16 && S != 32 && S != 64) return -1;
    static const int dirs = getenv("PAIG_CONV_TC_DIR") ? atoi(getenv("PAIG_CONV_TC_DIR")) : 3;   // bit 0 forward, bit 1 data gradient
    if (!(dirs & (c.transposed ? 2 : 1)) && !c.force_tc) return -1;
    if (((uintptr_t)c.in % 16) || (c.in_bs % 4) || c.N <= 0) return -1;
    CtEncodeFn enc = ct_encode_fn();
    if (!enc) return -1;
    ConvTcArgs a;
    memset(&a, 0, sizeof(a));
    a.S = S; a.Nframes = c.N; a.Cin = K; a.relu = c.relu;
    a.R = S >= 16 ? 16 : 8;
    a.fpt = S >= 16 ? 1 : 2;
    a.tiles_x = S / 8;
    a.tpg = a.tiles_x * (S / a.R);
    a.ntiles = cdiv(c.N, a.fpt) * a.tpg;
    const int T = N > 64 ? 1 : 2;
    a.nsuper = cdiv(a.ntiles, T);
    a.wpack = scratch; a.bias = c.b; a.out = c.out; a.out_bs = c.out_bs;
    static const int dbg = getenv("PAIG_CONV_TC_DBG") ? atoi(getenv("PAIG_CONV_TC_DBG")) : 0;
    a.dbg = dbg;
    a.drain = ct_drain(c.transposed != 0);
    a.comp = ct_comp(a.drain);
    const cuuint64_t dims[4] = {(cuuint64_t)S, (cuuint64_t)S, (cuuint64_t)K, (cuuint64_t)c.N};
    const cuuint64_t strides[3] = {(cuuint64_t)S * 4, (cuuint64_t)S * S * 4, (cuuint64_t)c.in_bs * 4};
    const cuuint32_t box[4] = {(cuuint32_t)((T == 2 && a.fpt == 1) ? 24 : 16), (cuuint32_t)(a.R + 2), 8, (cuuint32_t)a.fpt};
    const cuuint32_t es[4] = {1, 1, 1, 1};
    if (enc(&a.tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)c.in, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return -1;
    launch(conv_tc_pack_kernel, dim3(cdiv((long)K * N * 9, 256 * 4)), dim3(256), 0, st, c.w, scratch, K, N, c.transposed);
    int rc = check_launch("conv_tc_pack");
    if (rc) return rc;
    const int grid = a.nsuper < ct_sm_count() ? a.nsuper : ct_sm_count();
    switch (N) {
        case 16: ct_launch<16, 2>(a, grid, st); break;
        case 32: ct_launch<32, 2>(a, grid, st); break;
        case 48: ct_launch<48, 2>(a, grid, st); break;
        case 64: ct_launch<64, 2>(a, grid, st); break;
        case 96: ct_launch<96, 1>(a, grid, st); break;
        case 128: ct_launch<128, 1>(a, grid, st); break;
        default: return -1;
    }
    return check_launch(layer_name("conv3x3_tc", K, N, S));
}

}  // namespace paig

#else   // PAIG_EMU: tensor-core path needs the device

namespace paig {
size_t conv_tc_scratch_floats(int Cin, int Cout) { return (size_t)9 * Cin * Cout; }
bool conv_tc_enabled() { return false; }
int conv3x3_tc(const ConvArgs&, float*, cudaStream_t) { return -1; }
}  // namespace paig

#endif
