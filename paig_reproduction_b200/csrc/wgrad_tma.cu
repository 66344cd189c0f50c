// conv3x3 weight + bias gradient with TMA-staged, double-buffered strips.
//
//   dW[co,ci,ky,kx] = sum_{n,y,x} g[n,co,y,x] * in[n,ci,y+ky-1,x+kx-1]        db[co] = sum g[n,co,y,x]
//
// Same scheme as conv.cu's kernel (a thread keeps the 3x3 taps of a block of output channels x 1 input channel in
// registers across every strip its CTA walks; persistent CTAs; one partial per CTA; fixed-order reduce), but the
// strips arrive through two 4-D tensor maps (cp.async.bulk.tensor -> UTMALDG): the box of the layer input starts at
// (x,y) = (-4, y0-1) -- the innermost start coordinate must stay 16-byte aligned (x = -1 raises an illegal
// instruction, tools/tma_probe.cu) -- so the TMA unit's out-of-bounds zero fill *is* the conv's zero padding: no
// halo code, no per-element index arithmetic, no thread touches global memory, and the next strip lands in the
// other buffer while this one is multiplied (ncu on the old kernel: long_scoreboard, i.e. exposed staging, was the
// top stall).  A thread keeps 8 output channels x 1 input channel x 9 taps (72 accumulators): per 4-pixel step it
// reads 3 x (4 + 16 + 4) bytes of input and 8 x 16 bytes of g for 288 FMAs, which keeps the shared-memory pipe
// below the FMA pipe.
// Box widths/heights are chosen so that channel planes sit 4 banks apart: the 32 lanes of a warp (= 32 different
// input channels / channel groups at the same pixel) read conflict-free, and all lanes sharing an output-channel
// block read g as a broadcast.
#include "common.cuh"
#include "internal.h"

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#ifndef PAIG_EMU
#include <cuda.h>
#else
struct alignas(64) CUtensorMap { unsigned char opaque[128]; };
#endif

namespace paig {

constexpr int kWtThreads = 256;
constexpr int kWtMaxStages = 4;

struct TmaView {                 // what the tensor map describes (also drives the emulation path)
    const float* base;           // element (x=0, y=0, c=0, n=0)
    int S, C, N;
    long nstride;                // floats between frames
    int boxx, boxy, boxc;        // box = boxx x boxy x boxc x 1
};
struct WgradTmaArgs {
    CUtensorMap tm_in, tm_g;
    TmaView vin, vg;
    int Cin, Cout, S, N, R, strips;
    int Cc, Oc, csets;           // channels per CTA: blockIdx.y = oset * csets + cset owns ci in [cset*Cc, +Cc) x co in [oset*Oc, +Oc)
    int in_plane, g_plane;       // boxx * boxy of each view
    int stage_floats, g_off;     // per stage; offset of g inside a stage
    int stages;                  // 2..4 strips in flight (as many as fit next to the fold buffer)
    float* partials;
};

__device__ __forceinline__ void tma_box_load(float* dst, const CUtensorMap* tm, const TmaView& v, int x, int y, int c0,
                                             int n, unsigned long long* bar) {
#ifdef PAIG_EMU
    (void)tm; (void)bar;
    for (int c = 0; c < v.boxc; ++c)
        for (int r = 0; r < v.boxy; ++r)
            for (int k = 0; k < v.boxx; ++k) {
                const int gx = x + k, gy = y + r;
                float val = 0.f;
                if (gx >= 0 && gx < v.S && gy >= 0 && gy < v.S && n < v.N && c0 + c < v.C)
                    val = v.base[(long)n * v.nstride + ((long)(c0 + c) * v.S + gy) * v.S + gx];
                dst[(c * v.boxy + r) * v.boxx + k] = val;
            }
#else
    const unsigned dst_a = (unsigned)__cvta_generic_to_shared(dst);
    const unsigned bar_a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
        ::"r"(dst_a), "l"(reinterpret_cast<uint64_t>(tm)), "r"(x), "r"(y), "r"(c0), "r"(n), "r"(bar_a)
        : "memory");
#endif
}

#ifdef PAIG_EMU
template <int COB>
__global__ void conv3x3_wgrad_tma_kernel(const WgradTmaArgs a) {
#else
template <int COB>
__global__ void __launch_bounds__(kWtThreads, 2) conv3x3_wgrad_tma_kernel(const __grid_constant__ WgradTmaArgs a) {
#endif
    constexpr int kRed = COB * 10;                                    // floats a thread leaves for the final fold
    PAIG_DYN_SMEM(float, smem_raw);
    __shared__ unsigned long long full[kWtMaxStages];
    // TMA tensor destinations must be 128-byte aligned in the shared window; the dynamic segment follows the static
    // barriers, so align explicitly (the launch asks for 128 spare bytes)
#ifdef PAIG_EMU
    float* smem = smem_raw;
#else
    float* smem = smem_raw + (((128u - ((unsigned)__cvta_generic_to_shared(smem_raw) & 127u)) & 127u) >> 2);
#endif
    const int S = a.S, QX = (S + 3) / 4, R = a.R, kWtStages = a.stages;
    const int tid = threadIdx.x;
    // this CTA's channel block (layers too wide for one CTA's shared memory / 256 owner threads are split over grid.y)
    const int Cc = a.Cc, Oc = a.Oc;
    const int c_base = ((int)blockIdx.y % a.csets) * Cc, o_base = ((int)blockIdx.y / a.csets) * Oc;
    const int G_per = (Oc / COB) * Cc;                                // owner groups (<= kWtThreads, host)
    const int P = max(1, kWtThreads / G_per);                         // pixel partitions
    const int grp_local = tid % G_per, part = tid / G_per;
    const bool owner = part < P;
    const int ci = owner ? grp_local % Cc : 0, cob = owner ? grp_local / Cc : 0;   // local to the channel block

    float acc[COB][9], bacc[COB];
#pragma unroll
    for (int c = 0; c < COB; ++c) {
        bacc[c] = 0.f;
#pragma unroll
        for (int t = 0; t < 9; ++t) acc[c][t] = 0.f;
    }
    const int items = a.N * a.strips;
    const int n_my = blockIdx.x < items ? (items - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const unsigned stage_bytes = (unsigned)((Cc * a.in_plane + Oc * a.g_plane) * sizeof(float));
#ifndef PAIG_EMU
    if (tid == 0) {
        for (int s = 0; s < kWtMaxStages; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&full[s])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
#endif
    __syncthreads();
    auto issue = [&](int k) {                                          // thread 0 only
        const int item = blockIdx.x + k * gridDim.x;
        const int f = item / a.strips, y0 = (item % a.strips) * R;
        float* st = smem + (size_t)(k % kWtStages) * a.stage_floats;
#ifndef PAIG_EMU
        const unsigned bar_a = (unsigned)__cvta_generic_to_shared(&full[k % kWtStages]);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(stage_bytes) : "memory");
#endif
        tma_box_load(st, &a.tm_in, a.vin, -4, y0 - 1, c_base, f, &full[k % kWtStages]);
        tma_box_load(st + a.g_off, &a.tm_g, a.vg, 0, y0, o_base, f, &full[k % kWtStages]);
    };
    if (tid == 0)
        for (int k = 0; k < kWtStages && k < n_my; ++k) issue(k);

    for (int k = 0; k < n_my; ++k) {
        const int item = blockIdx.x + k * gridDim.x;
        const int y0 = (item % a.strips) * R;
        const int rows = min(R, S - y0);
        const float* sIn = smem + (size_t)(k % kWtStages) * a.stage_floats;
        const float* sG = sIn + a.g_off;
#ifdef PAIG_EMU
        __syncthreads();
#else
        {
            const unsigned bar_a = (unsigned)__cvta_generic_to_shared(&full[k % kWtStages]);
            const unsigned phase = (unsigned)(k / kWtStages) & 1u;
            unsigned done = 0;
            while (!done)
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                             : "=r"(done) : "r"(bar_a), "r"(phase) : "memory");
        }
#endif
        if (owner) {
            const int nq = rows * QX;
            const float* ipc = sIn + ci * a.in_plane;
            const float* gpc = sG + cob * COB * a.g_plane;
            const int bxi = a.vin.boxx, bxg = a.vg.boxx, gpl = a.g_plane;
            // quads q = part, part + P, ...: (row, quad-in-row) advance without a division per step
            const int dr = P / QX, dq = P % QX;
            int r = part / QX, qx = part % QX;
            for (int q = part; q < nq; q += P) {
                const float* ip = ipc + r * bxi + 4 * qx;              // tile column = image column + 4
                const float* gp = gpc + r * bxg + 4 * qx;
                float v[3][6];
#pragma unroll
                for (int kk = 0; kk < 3; ++kk) {
                    const float* row = ip + kk * bxi;
                    const float4 p4 = *reinterpret_cast<const float4*>(row + 4);
                    v[kk][0] = row[3]; v[kk][1] = p4.x; v[kk][2] = p4.y; v[kk][3] = p4.z; v[kk][4] = p4.w; v[kk][5] = row[8];
                }
#pragma unroll
                for (int c = 0; c < COB; ++c) {
                    const float4 g4 = *reinterpret_cast<const float4*>(gp + c * gpl);
                    const float gv[4] = {g4.x, g4.y, g4.z, g4.w};
                    bacc[c] += (gv[0] + gv[1]) + (gv[2] + gv[3]);
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                            for (int p = 0; p < 4; ++p) acc[c][ky * 3 + kx] += gv[p] * v[ky][kx + p];
                }
                r += dr; qx += dq;
                if (qx >= QX) { qx -= QX; ++r; }
            }
        }
        __syncthreads();                                               // everyone is done with this stage
        if (tid == 0 && k + kWtStages < n_my) issue(k + kWtStages);
    }
    // ---- fold pixel partitions (fixed order) and write this CTA's partial ----
    __syncthreads();
    float* sRed = smem;                                                // [P][G_per][kRed]
    if (owner) {
        float* dst = sRed + ((size_t)part * G_per + grp_local) * kRed;
#pragma unroll
        for (int c = 0; c < COB; ++c) {
#pragma unroll
            for (int t = 0; t < 9; ++t) dst[c * 9 + t] = acc[c][t];
            dst[COB * 9 + c] = bacc[c];
        }
    }
    __syncthreads();
    const int nW = a.Cout * a.Cin * 9;
    float* out = a.partials + (size_t)blockIdx.x * (nW + a.Cout);
    for (int e = tid; e < G_per * kRed; e += kWtThreads) {
        const int k = e % kRed, gl = e / kRed;
        float s = 0.f;
        for (int p = 0; p < P; ++p) s += sRed[((size_t)p * G_per + gl) * kRed + k];
        const int gci = c_base + gl % Cc, gcob = gl / Cc;
        if (k < COB * 9) {
            const int co = o_base + gcob * COB + k / 9;
            out[((size_t)co * a.Cin + gci) * 9 + (k % 9)] = s;
        } else if (gci == 0) {
            const int co = o_base + gcob * COB + (k - COB * 9);
            out[nW + co] = s;
        }
    }
}

#ifndef PAIG_EMU
// ---- tensor-core variant ---------------------------------------------------------------------------------------------
// The same strips, the same barriers, the products on the warp-level tensor-core path (mma.sync.m16n8k8 TF32, SASS
// HMMA.1688.F32.TF32) as a 3xTF32 split (a.b ~ hi.hi + hi.lo + lo.hi, see unet_fused.cu):
//
//     dW[co, ci](tap) += sum_px g[co, px] * in[ci, px + tap]         M = 16 output channels, N = 8 input channels, K = 8 pixels of a row
//
// A fragments come straight out of the g strip (lane (q, t): channel q / q+8, pixel t / t+4), B fragments out of the
// input strip shifted by the tap (lane (q, t): channel q, pixel t + kx / +4) -- both conflict-free because pick_boxx puts
// the channel planes 4 banks apart.  A warp owns one (16 co x 8 ci) block x 9 taps = 36 accumulators that live INSIDE the
// MMA for a whole strip and are drained to a register sum once per strip (chains of <= ~50 truncating accumulations:
// a relative loss of ~1e-6 of the partial, harmless for an end product compared at 1e-4; the forward/backward-data
// convolutions, whose error is amplified by the network behind them, drain every MMA).  The g fragment of an 8-pixel
// segment is split once and reused by 27 MMAs; per MMA the loop issues ~3 other instructions, against ~8 tensor-pipe
// cycles: the kernel is bound by the tensor pipe (512 MAC/clk/SM, 170 after the split) where the FMA loop reached 47.
// Pairs (m-tile, n-tile) x pixel partitions (units = (row, 8-pixel segment), dealt round-robin) go to the 8 warps; the host
// picks channel blocks with at most 8 pairs.  Measured and not adopted (profiles/r2o_*): one accumulator set per product type
// without the drain (108 accumulators: spills, 2x slower) and a (pair, tap row) item per warp with two units in flight
// (12 + 24 accumulators, same-register MMAs six slots apart, but the g fragment split three times and 6 of 8 warps busy:
// 16->16@32 0.155 ms against 0.123).
__device__ __forceinline__ float wg_tf32_hi(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }
__device__ __forceinline__ void wg_mma(float (&d)[4], const float (&a)[4], float b0, float b1) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(__float_as_uint(a[0])), "r"(__float_as_uint(a[1])), "r"(__float_as_uint(a[2])), "r"(__float_as_uint(a[3])),
          "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)));
}

__global__ void __launch_bounds__(kWtThreads, 2) conv3x3_wgrad_mma_kernel(const __grid_constant__ WgradTmaArgs a) {
    PAIG_DYN_SMEM(float, smem_raw);
    __shared__ unsigned long long full[kWtMaxStages];
    float* smem = smem_raw + (((128u - ((unsigned)__cvta_generic_to_shared(smem_raw) & 127u)) & 127u) >> 2);
    const int S = a.S, R = a.R, kWtStages = a.stages;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, q = lane >> 2, t = lane & 3;
    const int Cc = a.Cc, Oc = a.Oc;
    const int c_base = ((int)blockIdx.y % a.csets) * Cc, o_base = ((int)blockIdx.y / a.csets) * Oc;
    const int MT = (Oc + 15) >> 4, NTL = (Cc + 7) >> 3, G = MT * NTL;            // G <= 8 (host)
    const int P = (kWtThreads / 32) / G;                                          // pixel partitions (rows)
    const int pair = warp % G, part = warp / G;
    const bool active = part < P;
    const int mt = pair / NTL, nt = pair % NTL;
    const int co0 = mt * 16 + q, co1 = co0 + 8, cil = nt * 8 + q;                 // local to the channel block
    const bool v0 = co0 < Oc, v1 = co1 < Oc, vc = cil < Cc;

    float acc[9][4], sum[9][4], bs0 = 0.f, bs1 = 0.f;
#pragma unroll
    for (int tp = 0; tp < 9; ++tp)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[tp][i] = sum[tp][i] = 0.f;
    const int items = a.N * a.strips;
    const int n_my = blockIdx.x < items ? (items - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const unsigned stage_bytes = (unsigned)((Cc * a.in_plane + Oc * a.g_plane) * sizeof(float));
    if (tid == 0) {
        for (int s = 0; s < kWtMaxStages; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&full[s])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int k) {                                          // thread 0 only
        const int item = blockIdx.x + k * gridDim.x;
        const int f = item / a.strips, y0 = (item % a.strips) * R;
        float* st = smem + (size_t)(k % kWtStages) * a.stage_floats;
        const unsigned bar_a = (unsigned)__cvta_generic_to_shared(&full[k % kWtStages]);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(stage_bytes) : "memory");
        tma_box_load(st, &a.tm_in, a.vin, -4, y0 - 1, c_base, f, &full[k % kWtStages]);
        tma_box_load(st + a.g_off, &a.tm_g, a.vg, 0, y0, o_base, f, &full[k % kWtStages]);
    };
    if (tid == 0)
        for (int k = 0; k < kWtStages && k < n_my; ++k) issue(k);

    const int bxi = a.vin.boxx, bxg = a.vg.boxx;
    const int nseg = S >> 3;
    for (int k = 0; k < n_my; ++k) {
        const int item = blockIdx.x + k * gridDim.x;
        const int y0 = (item % a.strips) * R;
        const int rows = min(R, S - y0);
        const float* sIn = smem + (size_t)(k % kWtStages) * a.stage_floats;
        const float* sG = sIn + a.g_off;
        {
            const unsigned bar_a = (unsigned)__cvta_generic_to_shared(&full[k % kWtStages]);
            const unsigned phase = (unsigned)(k / kWtStages) & 1u;
            unsigned done = 0;
            while (!done)
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                             : "=r"(done) : "r"(bar_a), "r"(phase) : "memory");
        }
        if (active) {
            const float* g0 = sG + (v0 ? co0 : 0) * a.g_plane + t;
            const float* g1 = sG + (v1 ? co1 : 0) * a.g_plane + t;
            const float* xin = sIn + (vc ? cil : 0) * a.in_plane + t + 3;          // tile column = image column + 4, tap kx - 1
            // units = (row, 8-pixel segment) dealt round-robin to the pixel partitions: 44 units over 8 partitions waste 8 %,
            // whole rows (11 over 8) wasted 31 %
            const int dr = P / nseg, ds = P % nseg;
            int r = part / nseg, seg = part % nseg;
            for (int u = part; u < rows * nseg; u += P) {
                {
                    const int x0 = seg * 8;
                    float av[4];
                    av[0] = v0 ? g0[r * bxg + x0] : 0.f; av[2] = v0 ? g0[r * bxg + x0 + 4] : 0.f;
                    av[1] = v1 ? g1[r * bxg + x0] : 0.f; av[3] = v1 ? g1[r * bxg + x0 + 4] : 0.f;
                    bs0 += av[0] + av[2];
                    bs1 += av[1] + av[3];
                    float ah[4], al[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) { ah[i] = wg_tf32_hi(av[i]); al[i] = av[i] - ah[i]; }
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky) {
                        const float* xr = xin + (r + ky) * bxi + x0;
                        float h0[3], h1[3], l0[3], l1[3];
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx) {
                            const float b0 = vc ? xr[kx] : 0.f, b1 = vc ? xr[kx + 4] : 0.f;
                            h0[kx] = wg_tf32_hi(b0); h1[kx] = wg_tf32_hi(b1);
                            l0[kx] = b0 - h0[kx]; l1[kx] = b1 - h1[kx];
                        }
                        // (issue order: MMAs on the same accumulator are three slots apart)
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx) wg_mma(acc[ky * 3 + kx], ah, h0[kx], h1[kx]);
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx) wg_mma(acc[ky * 3 + kx], ah, l0[kx], l1[kx]);
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx) wg_mma(acc[ky * 3 + kx], al, h0[kx], h1[kx]);
                    }
                }
                r += dr; seg += ds;
                if (seg >= nseg) { seg -= nseg; ++r; }
            }
            // drain the strip's chains into the register sums (round to nearest)
#pragma unroll
            for (int tp = 0; tp < 9; ++tp)
#pragma unroll
                for (int i = 0; i < 4; ++i) { sum[tp][i] += acc[tp][i]; acc[tp][i] = 0.f; }
        }
        __syncthreads();                                               // everyone is done with this stage
        if (tid == 0 && k + kWtStages < n_my) issue(k + kWtStages);
    }
    // ---- fold pixel partitions (fixed order) and write this CTA's partial ----
    __syncthreads();
    float* sRed = smem;                                                // [P][Oc][Cc][9] | [P][Oc]
    float* sBias = smem + (size_t)P * Oc * Cc * 9;
    if (active) {
#pragma unroll
        for (int tp = 0; tp < 9; ++tp)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int co = mt * 16 + q + ((i & 2) ? 8 : 0), ci = nt * 8 + 2 * t + (i & 1);
                if (co < Oc && ci < Cc) sRed[(((size_t)part * Oc + co) * Cc + ci) * 9 + tp] = sum[tp][i];
            }
        bs0 += __shfl_xor_sync(0xffffffffu, bs0, 1); bs0 += __shfl_xor_sync(0xffffffffu, bs0, 2);
        bs1 += __shfl_xor_sync(0xffffffffu, bs1, 1); bs1 += __shfl_xor_sync(0xffffffffu, bs1, 2);
        if (nt == 0 && t == 0) {
            if (v0) sBias[part * Oc + co0] = bs0;
            if (v1) sBias[part * Oc + co1] = bs1;
        }
    }
    __syncthreads();
    const int nW = a.Cout * a.Cin * 9;
    float* out = a.partials + (size_t)blockIdx.x * (nW + a.Cout);
    const int blockW = Oc * Cc * 9;
    for (int e = tid; e < blockW + (c_base == 0 ? Oc : 0); e += kWtThreads) {
        float s = 0.f;
        if (e < blockW) {
            for (int p = 0; p < P; ++p) s += sRed[(size_t)p * blockW + e];
            const int co = e / (Cc * 9), ci = (e / 9) % Cc, tp = e % 9;
            out[((size_t)(o_base + co) * a.Cin + c_base + ci) * 9 + tp] = s;
        } else {
            const int co = e - blockW;
            for (int p = 0; p < P; ++p) s += sBias[p * Oc + co];
            out[nW + o_base + co] = s;
        }
    }
}
#endif   // !PAIG_EMU

// ---- host ---------------------------------------------------------------------------------------------------
namespace {

// Smallest box width >= need that keeps 16-byte loads of 8 consecutive lanes (= 8 channel planes) conflict-free:
// plane = rows * boxx floats must be 4 * odd (mod 32), i.e. rows odd and boxx = 4 * odd -- the planes' first banks
// are then the 8 distinct multiples of 4.
int pick_boxx(int need, int rows) {
    if ((rows & 1) == 0) return -1;
    int k = (need + 3) / 4;
    if ((k & 1) == 0) ++k;
    return 4 * k;
}

#ifndef PAIG_EMU
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
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
#endif

bool make_map(CUtensorMap* tm, const TmaView& v) {
#ifdef PAIG_EMU
    (void)tm; (void)v;
    return true;
#else
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    const cuuint64_t dims[4] = {(cuuint64_t)v.S, (cuuint64_t)v.S, (cuuint64_t)v.C, (cuuint64_t)v.N};
    const cuuint64_t strides[3] = {(cuuint64_t)v.S * 4, (cuuint64_t)v.S * v.S * 4, (cuuint64_t)v.nstride * 4};
    const cuuint32_t box[4] = {(cuuint32_t)v.boxx, (cuuint32_t)v.boxy, (cuuint32_t)v.boxc, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    return fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)v.base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
#endif
}

}  // namespace

// Returns 0 on success, 1 on error, -1 when this layer's geometry cannot be described to the TMA unit (row pitch not
// a multiple of 16 bytes, masked input, too many channels for one box) -- the caller then uses conv.cu's kernel.
int conv3x3_wgrad_tma(const WgradArgs& w, float* dW, float* db, cudaStream_t st) {
    static const bool off = getenv("PAIG_WGRAD_LEGACY") != nullptr;
    if (off || w.N <= 0) return -1;
    if (w.in_mask || w.act) return -1;                                 // gating must already be applied to g
    const int S = w.S;
    if ((S * 4) % 16 != 0 || w.Cin > 256 || w.Cout > 256) return -1;
    if (((uintptr_t)w.in % 16) || ((uintptr_t)w.g % 16) || (w.in_bs % 4) || (w.g_bs % 4)) return -1;
    WgradTmaArgs a;
    memset(&a, 0, sizeof(a));
    a.Cin = w.Cin; a.Cout = w.Cout; a.S = S; a.N = w.N; a.partials = w.partials;
    // Strip height: odd (so R and R+2 are odd, pick_boxx); the tallest of {whole image, 13, 11, 9, 7, 5, 3} that
    // still leaves two strips in flight per CTA and wastes the fewest rows -- fewer, fuller strips per frame mean
    // fewer barrier round trips and less halo re-read.
    static const int forced_R = getenv("PAIG_WGRAD_R") ? atoi(getenv("PAIG_WGRAD_R")) : 0;
    const size_t budget = (110 * 1024 - 128) / sizeof(float);
    const int COB = (a.Cout % 8) == 0 ? 8 : 4;
    if (a.Cout % COB) return -1;
#ifdef PAIG_EMU
    const bool mma = false;
#else
    static const bool mma_off = getenv("PAIG_NO_WGRAD_MMA") != nullptr || getenv("PAIG_NO_MMA") != nullptr;
    // tensor-core variant: rows are whole 8-pixel segments; m-tiles of 16 output channels (with 8, half of every MMA is
    // padding and the FMA kernel wins: 24->8@32 0.187 vs 0.109 ms, 8->8@32 0.128 vs 0.094, profiles/r2o_wgrad_layers.txt)
    const bool mma = !mma_off && (S % 8) == 0 && (a.Cout % 16) == 0;
#endif
    const int cands[7] = {(S + 1) | 1, 13, 11, 9, 7, 5, 3};
    bool found = false;
    WgradTmaArgs best = a;
    // Channel blocks: the fewest CTAs-per-strip (csets x osets) whose stage fits twice next to a second CTA and whose
    // owner groups fit the 256 threads; ties prefer splitting the input channels (their planes carry the halo).
    for (int sets = 1; sets <= 64 && !found; ++sets) {
        for (int osets = 1; osets <= sets && !found; ++osets) {
            if (sets % osets) continue;
            const int csets = sets / osets;
            if (a.Cin % csets || a.Cout % osets) continue;
            const int Cc = a.Cin / csets, Oc = a.Cout / osets;
            if (Oc % COB || (Oc / COB) * Cc > kWtThreads || Cc > 256 || Oc > 256) continue;
            if (mma && ((Oc + 15) / 16) * ((Cc + 7) / 8) > kWtThreads / 32) continue;    // one (m-tile, n-tile) pair per warp
            double best_cost = 0;
            for (int ci = 0; ci < 7; ++ci) {
                const int R = forced_R > 0 ? forced_R : cands[ci];
                if (R > ((S + 1) | 1) || (R & 1) == 0) continue;
                WgradTmaArgs c = a;
                c.R = R;
                c.strips = cdiv(S, R);
                c.Cc = Cc; c.Oc = Oc; c.csets = csets;
                c.vin = TmaView{w.in, S, w.Cin, w.N, w.in_bs, pick_boxx(S + 5, R + 2), R + 2, Cc};    // image columns -4 .. S
                c.vg = TmaView{w.g, S, w.Cout, w.N, w.g_bs, pick_boxx(S, R), R, Oc};
                if (c.vin.boxx < 0 || c.vg.boxx < 0 || c.vin.boxx > 256 || c.vg.boxx > 256 || R + 2 > 256) continue;
                c.in_plane = c.vin.boxx * c.vin.boxy;
                c.g_plane = c.vg.boxx * c.vg.boxy;
                c.g_off = (Cc * c.in_plane + 31) & ~31;                    // 128-byte aligned TMA destinations
                c.stage_floats = (c.g_off + Oc * c.g_plane + 31) & ~31;
                c.stages = (int)(budget / c.stage_floats);
                if (c.stages > kWtMaxStages) c.stages = kWtMaxStages;
                if (c.stages < 2) continue;                                // two CTAs per SM, two strips in flight
                // rows fetched per frame (halo re-reads + padding of the last strip) plus a fixed cost per strip
                const double cost = (double)c.strips * (R + 2) + 3.0 * c.strips;
                if (!found || cost < best_cost) { found = true; best_cost = cost; best = c; }
                if (forced_R > 0) break;
            }
        }
    }
    if (!found) return -1;
    a = best;
    const int sets = a.csets * (a.Cout / a.Oc);
    const int G_per = (a.Oc / COB) * a.Cc;
    const int P = kWtThreads / G_per > 0 ? kWtThreads / G_per : 1;
    size_t red = (size_t)P * G_per * COB * 10;
    if (mma) {
        const int Pm = (kWtThreads / 32) / (((a.Oc + 15) / 16) * ((a.Cc + 7) / 8));
        red = (size_t)Pm * a.Oc * (a.Cc * 9 + 1);
    }
    const size_t tile = (size_t)a.stages * a.stage_floats;
    const size_t smem = (tile > red ? tile : red) * sizeof(float) + 128;
    if (smem > 110 * 1024) return -1;
    if (!make_map(&a.tm_in, a.vin) || !make_map(&a.tm_g, a.vg)) return -1;
    // one resident wave: 296 CTAs shared between the channel blocks; every CTA column writes one partial
    int ctas = a.N * a.strips;
    const int cap = kWgradMaxCtas / sets > 0 ? kWgradMaxCtas / sets : 1;
    if (ctas > cap) ctas = cap;
    static const bool debug = getenv("PAIG_DEBUG") != nullptr;
    if (debug)
        fprintf(stderr, "[paig] wgrad_tma %d->%d S=%d N=%d R=%d strips=%d stages=%d in box %dx%d g box %dx%d stage=%d floats smem=%zu ctas=%d blocks=%dx%d (Cc=%d Oc=%d)\n",
                a.Cin, a.Cout, S, a.N, a.R, a.strips, a.stages, a.vin.boxx, a.vin.boxy, a.vg.boxx, a.vg.boxy, a.stage_floats, smem, ctas,
                a.csets, sets / a.csets, a.Cc, a.Oc);
#ifndef PAIG_EMU
    if (mma) launch(conv3x3_wgrad_mma_kernel, dim3(ctas, sets), dim3(kWtThreads), smem, st, a);
    else
#endif
    if (COB == 8) launch(conv3x3_wgrad_tma_kernel<8>, dim3(ctas, sets), dim3(kWtThreads), smem, st, a);
    else launch(conv3x3_wgrad_tma_kernel<4>, dim3(ctas, sets), dim3(kWtThreads), smem, st, a);
    int rc = check_launch(layer_name("conv3x3_wgrad", -a.Cin, a.Cout, a.S));   // (negative Cin marks the TMA kernel)
    if (rc) return rc;
    const int nW = a.Cout * a.Cin * 9;
    if (w.defer) return w.defer->add(a.partials, ctas, nW + a.Cout, nW, dW, a.Cout, db) ? 0 : 1;
    return reduce_partials(a.partials, ctas, nW + a.Cout, nW, dW, a.Cout, db, st);
}

}  // namespace paig
