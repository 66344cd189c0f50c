// ShallowUNet forward (blocks.py:278-308) as ONE persistent kernel: a CTA keeps a whole frame's activations in
// shared memory (<= 227 KB) and walks the layer list -- 3x3 convs (+ReLU), 2x2 max-pools, 2x bilinear
// upsamples fused into the conv that consumes them, the 1x1 head -- without a round trip to HBM between layers.
// Activations the backward pass needs leave the SM as fire-and-forget stores in the layout the layer kernels
// of conv.cu use, so encoder_backward is unchanged.
//
// Why this shape: per layer the step sits at ~18 FLOP/B, next to B200's balance point, so per-layer kernels
// alternate between waiting on HBM/L2 and on the FMA pipe (profiles/r1b_ncu_summary.md: 45 % FMA-pipe
// utilisation, long_scoreboard stalls while tiles are staged).  On chip the only traffic is 12 KB of frame in and
// the saved activations out; the convs run from shared memory at the FMA pipe's pace.
//
//   * 512 threads / CTA, one CTA per SM, grid = min(frames, SMs); CTA b handles frames b, b+grid, ...
//   * planes are stored [C][S+2][4*ceil(S/4)+4] with a zero halo (image (y,x) at tile (y+1,x+1)), so a thread's
//     4-pixel x CO-output tile reads rows as one 16-byte + one 8-byte shared load, no bounds checks
//   * weights are pre-packed once per step to [ci][tap][co] (+bias) and brought in per layer by a TMA bulk copy
//     (cp.async.bulk -> UBLKCP) issued one layer ahead, completing on an mbarrier
//   * the host plans shared-memory offsets from buffer lifetimes (first-fit), and refuses (caller falls back to
//     the per-layer kernels) if a network does not fit -- the 64x64 UNet of the mnist task does not
#include "common.cuh"
#include "internal.h"
#include "layout.h"
#include "unet_plan.h"

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace paig {

constexpr int kFusedThreads = 512;
// (FusedOp / FusedPlan / PackPlan: unet_plan.h)
constexpr int kIssueThread = kFusedThreads - 32;
constexpr size_t kFusedSmemLimit = 227 * 1024 - 4096;   // dynamic part; the op table and barriers are static

struct Geo {
    int S, nqx, P, plane;
};
__host__ __device__ inline Geo geo_of(int S) {
    Geo g;
    g.S = S;
    g.nqx = (S + 3) / 4;
    g.P = 4 * g.nqx + 4;
    g.plane = (S + 2) * g.P;
    return g;
}

// ---- TMA bulk copy split into issue (one thread) and wait (everyone) ------------------------------------------
__device__ __forceinline__ void bulk_issue(float* smem_dst, const float* gmem_src, unsigned bytes, unsigned long long* bar) {
#ifdef PAIG_EMU
    (void)bar;
    memcpy(smem_dst, gmem_src, bytes);
#else
    const unsigned bar_a = (unsigned)__cvta_generic_to_shared(bar);
    const unsigned dst_a = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic accesses to the region come first
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_a),
                 "l"(__cvta_generic_to_global(gmem_src)), "r"(bytes), "r"(bar_a)
                 : "memory");
#endif
}
__device__ __forceinline__ void bulk_wait(unsigned long long* bar, unsigned phase) {
#ifdef PAIG_EMU
    (void)bar;
    (void)phase;
    __syncthreads();
#else
    const unsigned bar_a = (unsigned)__cvta_generic_to_shared(bar);
    unsigned done = 0;
    while (!done) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done)
                     : "r"(bar_a), "r"(phase & 1u)
                     : "memory");
    }
#endif
}

#ifdef PAIG_EMU
#define PAIG_STAMP(tm, i)
#else
#define PAIG_STAMP(tm, i) do { if (tm) (tm)[i] = clock64(); } while (0)
#endif

// zero the halo (row 0, row S+1, column 0, columns S+1..P-1) of C consecutive planes: a thread owns one halo cell
// (<= 2P + S(P-S) = 224 cells for S <= 36) and walks the planes, so the index arithmetic is done once
__device__ __forceinline__ void zero_halo_planes(float* base, int C, const Geo g, int tid, int nthr) {
    const int side = g.P - g.S;                        // halo cells per interior row
    const int per = 2 * g.P + g.S * side;
    for (int k = tid; k < per; k += nthr) {
        int row, col;
        if (k < 2 * g.P) {
            row = k < g.P ? 0 : g.S + 1;
            col = k < g.P ? k : k - g.P;
        } else {
            const int k2 = k - 2 * g.P, j = k2 % side;
            row = 1 + k2 / side;
            col = j == 0 ? 0 : g.S + j;
        }
        float* p = base + row * g.P + col;
        for (int c = 0; c < C; ++c) p[c * g.plane] = 0.f;
    }
}

// acc[r][c][p] += sum_{ci < nci} sum_taps w[ci][tap][c] * in[ci][y + r + ky - 1][4 qx + p + kx - 1]
// A thread owns PY rows x 4 pixels x CO output channels.  COUT is a template parameter so that every weight load is
// [pointer + immediate]; the row pointer advances by one plane per input channel.  Shared-memory traffic per input
// channel is (PY+2) x (16 B + 8 B) of pixels and 9 x CO x 4 B of (broadcast) weights for 36 x CO x PY FMAs: the
// taller tile halves the wavefronts per FMA, which is what bounds the 4 x 1 tile (profiles/r1d_fusedfwd_*).
template <int CO, int COUT, int PY>
__device__ __forceinline__ void conv_accumulate(float (&acc)[PY][CO][4], const float* __restrict__ planes, int nci,
                                                const Geo g, const float* __restrict__ w, int y, int qx) {
    const float* r0 = planes + y * g.P + 4 * qx;
    const int plane = g.plane, P = g.P;
#pragma unroll 1
    for (int ci = 0; ci < nci; ++ci) {
        float v[PY + 2][6];
#pragma unroll
        for (int r = 0; r < PY + 2; ++r) {
            const float4 a4 = *reinterpret_cast<const float4*>(r0 + r * P);
            const float2 a2 = *reinterpret_cast<const float2*>(r0 + r * P + 4);
            v[r][0] = a4.x; v[r][1] = a4.y; v[r][2] = a4.z; v[r][3] = a4.w; v[r][4] = a2.x; v[r][5] = a2.y;
        }
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
                for (int c4 = 0; c4 < CO; c4 += 4) {
                    const float4 w4 = *reinterpret_cast<const float4*>(w + (ky * 3 + kx) * COUT + c4);
                    const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                    for (int r = 0; r < PY; ++r)
#pragma unroll
                        for (int c = 0; c < 4; ++c)
#pragma unroll
                            for (int p = 0; p < 4; ++p) acc[r][c4 + c][p] += wv[c] * v[r + ky][kx + p];
                }
            }
        r0 += plane;
        w += 9 * COUT;
    }
}

__device__ __forceinline__ void up_taps_f(int o, int Si, int& i0, int& i1, float& w0, float& w1) {
    const int k = o >> 1;
    if (o & 1) { i0 = k; i1 = min(k + 1, Si - 1); w0 = 0.75f; w1 = 0.25f; }
    else { i0 = max(k - 1, 0); i1 = k; w0 = 0.25f; w1 = 0.75f; }
}

template <int CO, int PY>
__device__ __forceinline__ void conv_epilogue(const float (&acc)[PY][CO][4], const FusedOp& op, const Geo g, float* sm,
                                              const float* bias, int cg, int y0, int qx, int f) {
    const int S = g.S, x0 = 4 * qx;
    const bool vec = (S & 3) == 0;
    // ReLU adjoint gates: four channel rows' gate loads go out before the first of their stores (the stores may alias
    // them as far as the compiler knows, which would serialise one L2 round trip per output channel)
    constexpr int KB = PY * CO > 32 ? 1 : 4;      // (64-accumulator tiles have no registers to spare)
#pragma unroll
    for (int r = 0; r < PY; ++r)
#pragma unroll
    for (int c0 = 0; c0 < CO; c0 += KB) {
    float4 gate[KB];
#pragma unroll
    for (int k = 0; k < KB; ++k) {
        gate[k] = make_float4(1.f, 1.f, 1.f, 1.f);
        if (op.gmask) {
            const float* m = op.gmask + (long)f * op.gmask_bs + ((long)(cg * CO + c0 + k) * S + y0 + r) * S + x0;
            if (vec) {
                gate[k] = *reinterpret_cast<const float4*>(m);
            } else {
                gate[k].x = m[0];
                if (x0 + 1 < S) gate[k].y = m[1];
                if (x0 + 2 < S) gate[k].z = m[2];
                if (x0 + 3 < S) gate[k].w = m[3];
            }
        }
    }
#pragma unroll
    for (int k = 0; k < KB; ++k) {
        const int c = c0 + k;
        const int co = cg * CO + c, y = y0 + r;
        const float b = op.bias ? bias[co] : 0.f;
        float o[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            o[p] = acc[r][c][p] + b;
            if (op.relu) o[p] = fmaxf(o[p], 0.f);
        }
        {                                     // pass the gradient where the forward activation was positive
            const float4 m4 = gate[k];
            o[0] = m4.x > 0.f ? o[0] : 0.f; o[1] = m4.y > 0.f ? o[1] : 0.f;
            o[2] = m4.z > 0.f ? o[2] : 0.f; o[3] = m4.w > 0.f ? o[3] : 0.f;
        }
        if (op.out >= 0) {
            float* d = sm + op.out + co * g.plane + (y + 1) * g.P + x0 + 1;
#pragma unroll
            for (int p = 0; p < 4; ++p)
                if (vec || x0 + p < S) d[p] = o[p];
        }
        if (op.gout) {
            float* d = op.gout + (long)f * op.gout_bs + ((long)co * S + y) * S + x0;
            if (vec) {
                *reinterpret_cast<float4*>(d) = make_float4(o[0], o[1], o[2], o[3]);
            } else {
#pragma unroll
                for (int p = 0; p < 4; ++p)
                    if (x0 + p < S) d[p] = o[p];
            }
        }
    }
    }
}

// 2x bilinear upsample (align_corners=False) of channels [c0, c0+nch) of a half-resolution buffer into `chunk`
// (and to global memory for the backward pass).  A thread owns one 4-pixel output quad position and strides over
// the channels; per channel it reads 2 source rows x 4 source columns (two 8-byte loads each) and blends W first,
// then H -- the arithmetic of conv.cu's upsample2_kernel.  No per-element index arithmetic.
__device__ __forceinline__ void upsample_chunk(const float* __restrict__ src, const Geo gl, float* __restrict__ chunk,
                                               const Geo g, int nch, float* __restrict__ gdst, int tid) {
    const int S = g.S, Si = gl.S;
    const int nq = S * g.nqx;                       // output quads per channel
    if (tid >= (kFusedThreads / nq) * nq && nq <= kFusedThreads) return;
    const int q = tid % nq, sub = tid / nq, nsub = kFusedThreads / nq > 0 ? kFusedThreads / nq : 1;
    if (tid >= nq && nq > kFusedThreads) return;    // (never: nq <= 324)
    const int yy = q / g.nqx, qx = q % g.nqx, x0 = 4 * qx;
    // rows: out row yy reads source rows ya, yb with weights wya, wyb
    int ya, yb;
    float wya, wyb;
    up_taps_f(yy, Si, ya, yb, wya, wyb);
    // columns: outputs x0..x0+3 read source columns k-1..k+2 with k = x0/2 (even x: .25 in[k-1] + .75 in[k]; odd x:
    // .75 in[k] + .25 in[k+1]); clamped at the borders
    const int k = x0 >> 1;
    const bool left = k == 0;                        // column k-1 is outside: clamp to column 0
    const int kmax = Si - 1;
    const float* pa = src + (ya + 1) * gl.P + k;     // tile column of source column k-1 is k
    const float* pb = src + (yb + 1) * gl.P + k;
    float* cd = chunk + (yy + 1) * g.P + x0 + 1;
    const bool vec = (S & 3) == 0;
    for (int c = sub; c < nch; c += nsub) {
        const float2 a01 = *reinterpret_cast<const float2*>(pa + c * gl.plane);
        const float2 a23 = *reinterpret_cast<const float2*>(pa + c * gl.plane + 2);
        const float2 b01 = *reinterpret_cast<const float2*>(pb + c * gl.plane);
        const float2 b23 = *reinterpret_cast<const float2*>(pb + c * gl.plane + 2);
        float a[4] = {a01.x, a01.y, a23.x, a23.y}, b[4] = {b01.x, b01.y, b23.x, b23.y};
        if (left) { a[0] = a[1]; b[0] = b[1]; }
        if (k + 1 > kmax) { a[2] = a[1]; b[2] = b[1]; }           // (only when the quad is partly outside the image)
        if (k + 2 > kmax) { a[3] = k + 1 > kmax ? a[1] : a[2]; b[3] = k + 1 > kmax ? b[1] : b[2]; }
        // W pass: x0 (even): .25 s[k-1] + .75 s[k]; x0+1: .75 s[k] + .25 s[k+1]; x0+2: .25 s[k] + .75 s[k+1]; x0+3: .75 s[k+1] + .25 s[k+2]
        float top[4], bot[4], o[4];
        top[0] = 0.25f * a[0] + 0.75f * a[1]; bot[0] = 0.25f * b[0] + 0.75f * b[1];
        top[1] = 0.75f * a[1] + 0.25f * a[2]; bot[1] = 0.75f * b[1] + 0.25f * b[2];
        top[2] = 0.25f * a[1] + 0.75f * a[2]; bot[2] = 0.25f * b[1] + 0.75f * b[2];
        top[3] = 0.75f * a[2] + 0.25f * a[3]; bot[3] = 0.75f * b[2] + 0.25f * b[3];
#pragma unroll
        for (int p = 0; p < 4; ++p) o[p] = wya * top[p] + wyb * bot[p];
        float* d = cd + c * g.plane;
#pragma unroll
        for (int p = 0; p < 4; ++p)
            if (vec || x0 + p < S) d[p] = o[p];
        if (gdst) {
            float* gd = gdst + ((long)c * S + yy) * S + x0;
            if (vec) *reinterpret_cast<float4*>(gd) = make_float4(o[0], o[1], o[2], o[3]);
            else {
#pragma unroll
                for (int p = 0; p < 4; ++p)
                    if (x0 + p < S) gd[p] = o[p];
            }
        }
    }
}

template <int CO, int COUT, int PY>
__device__ __forceinline__ void run_conv(const FusedOp& op, float* sm, int f, int tid, long long* tm) {
    const Geo g = geo_of(op.S);
    const int nq = (op.S / PY) * g.nqx, nitems = nq * (COUT / CO);       // PY divides S (planner)
    const float* w = sm + op.wsm;
    const float* bias = w + (op.Cin0 + op.Cin1) * 9 * COUT;
    if (op.out >= 0) zero_halo_planes(sm + op.out, COUT, g, tid, kFusedThreads);
    PAIG_STAMP(tm, 1);
    float acc[PY][CO][4];
    if (!op.up) {
        const bool pow2 = (nq & (nq - 1)) == 0 && (g.nqx & (g.nqx - 1)) == 0;
        const int lq = 31 - __clz(nq), lx = 31 - __clz(g.nqx);
        for (int item = tid; item < nitems; item += kFusedThreads) {
            int cg, q, yr, qx;
            if (pow2) { cg = item >> lq; q = item & (nq - 1); yr = q >> lx; qx = q & (g.nqx - 1); }
            else { cg = item / nq; q = item % nq; yr = q / g.nqx; qx = q % g.nqx; }
            const int y = yr * PY;
#pragma unroll
            for (int r = 0; r < PY; ++r)
#pragma unroll
                for (int c = 0; c < CO; ++c)
#pragma unroll
                    for (int p = 0; p < 4; ++p) acc[r][c][p] = 0.f;
#ifndef PAIG_EMU
            if (op.gmask) {                  // the epilogue's ReLU gates live in HBM: pull their lines into L2 while we compute
#pragma unroll
                for (int r = 0; r < PY; ++r)
#pragma unroll
                    for (int c = 0; c < CO; ++c)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(op.gmask + (long)f * op.gmask_bs +
                                                                       ((long)(cg * CO + c) * op.S + y + r) * op.S + 4 * qx));
            }
#endif
            // the two input segments of a concat share one copy of the accumulate loop (code size: the kernel's
            // instruction footprint per frame is what the instruction cache has to hold)
#pragma unroll 1
            for (int seg = 0; seg < 2; ++seg) {
                const int nci = seg ? op.Cin1 : op.Cin0;
                if (nci)
                    conv_accumulate<CO, COUT, PY>(acc, sm + (seg ? op.in1 : op.in0), nci, g,
                                                  w + (seg ? op.Cin0 * 9 * COUT : 0) + cg * CO, y, qx);
            }
            PAIG_STAMP(tm, 2);
            conv_epilogue<CO, PY>(acc, op, g, sm, bias, cg, y, qx, f);
        }
        PAIG_STAMP(tm, 3);
    } else {
        // the planner guarantees nitems <= kFusedThreads here: accumulators persist across the channel chunks
        const bool active = tid < nitems;
        const int cg = active ? tid / nq : 0, q = active ? tid % nq : 0, y = (q / g.nqx) * PY, qx = q % g.nqx;
#pragma unroll
        for (int r = 0; r < PY; ++r)
#pragma unroll
            for (int c = 0; c < CO; ++c)
#pragma unroll
                for (int p = 0; p < 4; ++p) acc[r][c][p] = 0.f;
        const int S = op.S;
        const Geo gl = geo_of(S / 2);
        float* chunk = sm + op.chunk;
        zero_halo_planes(chunk, op.up, g, tid, kFusedThreads);
        long long t_up = 0, t_acc = 0, t_a = 0, t_b = 0;
        (void)t_a; (void)t_b;
        for (int c0 = 0; c0 < op.Cin0; c0 += op.up) {
            const int nch = min(op.up, op.Cin0 - c0);
#ifndef PAIG_EMU
            if (tm) t_a = clock64();
#endif
            upsample_chunk(sm + op.in0 + c0 * gl.plane, gl, chunk, g, nch,
                           op.gup ? op.gup + (long)f * op.gup_bs + (long)c0 * S * S : nullptr, tid);
            __syncthreads();
#ifndef PAIG_EMU
            if (tm) { t_b = clock64(); t_up += t_b - t_a; }
#endif
            if (active) conv_accumulate<CO, COUT, PY>(acc, chunk, nch, g, w + c0 * 9 * COUT + cg * CO, y, qx);
            __syncthreads();
#ifndef PAIG_EMU
            if (tm) t_acc += clock64() - t_b;
#endif
        }
        if (tm) { tm[2] = t_up; tm[3] = t_acc; }
        if (active) conv_epilogue<CO, PY>(acc, op, g, sm, bias, cg, y, qx, f);
    }
}

template <int COUT>
__device__ __forceinline__ void run_conv_co(const FusedOp& op, float* sm, int f, int tid, long long* tm) {
    // (8 x 2 tiles were measured too: never the fastest; 16 x 1 stays for the single-pass upsampling convs at 36 px)
    if (op.py_tile == 2) run_conv<4, COUT, 2>(op, sm, f, tid, tm);
    else if (COUT >= 16 && op.co_tile == 16) run_conv<(COUT >= 16 ? 16 : 4), COUT, 1>(op, sm, f, tid, tm);
    else if (op.co_tile == 8) run_conv<8, COUT, 1>(op, sm, f, tid, tm);
    else run_conv<4, COUT, 1>(op, sm, f, tid, tm);
}


// ---- tensor-core variant of the accumulate loop (32-px frames: levels of 32, 16 and 8 px) ------------------------------
// The same convolutions as implicit GEMMs on the warp-level tensor-core path (mma.sync.m16n8k8, SASS HMMA.1688.F32.TF32),
// fp32-accurate through the 3xTF32 split (x = hi + lo, hi = rn_tf32(x), lo = rn_tf32(x - hi); a.b ~ hi.hi + hi.lo + lo.hi):
//
//     D[co, px] = sum_{tap, ci} W[co, ci, tap] * in[ci, px + tap]        M = 16 output channels, N = 8 pixels of a row, K = 8 input channels
//
// Why mma.sync and not tcgen05 here: with 8..32 output channels a tcgen05.mma (M = 128 pixels) reads a 4 KB pixel slab from
// shared memory per K = 8 step whatever N is, which is what bounds it (conv_tc.cu at N = 16: a tie with the FMA loop,
// profiles/r2k_conv_tc_sweep.txt) and it needs the activations re-laid out as hi and lo tiles -- twice the footprint of a
// frame that fills the SM already.  The warp-level path takes the B fragments straight out of the existing [C][S+2][P]
// planes (lane (g, t) reads channel t / t+4 at pixel g: 4 planes x 8 consecutive floats = 32 distinct banks because
// plane = 8 or 24 (mod 32) at S = 32 / 16 / 8), splits them in registers, and keeps the weights of one tap in 8 registers
// for all the n-tiles of the warp.  Measured rate: 512 MAC/clk/SM (tools/ubench/mma_sync_rate.cu), i.e. 170 useful
// MAC/clk/SM after the split against the ~55 the FMA loop reaches (it is bound by shared-memory wavefronts: 9.3 per
// 1024 MAC; this loop needs 2).
//
// Accuracy: the tensor core sums the 8 products of an MMA with truncation.  hi.hi goes through an MMA with a ZERO
// accumulator and is added to a register sum with round-to-nearest (no accumulation chain inside the tensor core); the
// expected truncation loss of one MMA (kappa x result, see conv_tc.cu) is added back once, next to the small correction
// sum (hi.lo + lo.hi, accumulated inside the MMA), before the single final addition.
#ifndef PAIG_EMU
// hi = x rounded to TF32 (10 mantissa bits), nearest with ties away from zero -- what cvt.rna.tf32.f32 computes, but that
// instruction is emulated on sm_100a (FSETP + SEL + LOP3 + IADD, ~2.5 issue slots and it was a third of the loop);
// on the bit pattern it is one add and one mask (finite inputs).  lo = x - hi is exact in fp32 and goes to the tensor
// core as it is: the MMA reads only the upper 19 bits of a TF32 operand, a truncation of a term that is 2^-12 of x
// with a random sign.
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }
__device__ __forceinline__ void mma_tf32_acc(float (&d)[4], const float (&a)[4], float b0, float b1) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(__float_as_uint(a[0])), "r"(__float_as_uint(a[1])), "r"(__float_as_uint(a[2])), "r"(__float_as_uint(a[3])),
          "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)));
}
__device__ __forceinline__ void mma_tf32_zero(float (&d)[4], const float (&a)[4], float b0, float b1) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
        : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
        : "r"(__float_as_uint(a[0])), "r"(__float_as_uint(a[1])), "r"(__float_as_uint(a[2])), "r"(__float_as_uint(a[3])),
          "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)), "f"(0.f));
}

template <int S> struct MmaGeo {
    static constexpr int P = 4 * ((S + 3) / 4) + 4, PLANE = (S + 2) * P;
    static constexpr int TPR = S / 8;                         // n-tiles (8 pixels) per image row
    static constexpr int NT = S == 32 ? 4 : (S == 16 ? 2 : 1);   // n-tiles per warp and pass: one image row
    // correction accumulators per n-tile.  An accumulating HMMA needs the previous one on the same registers to have
    // finished (~33 cycles): hi.lo and lo.hi go to separate accumulators, and with fewer than 4 n-tiles per warp even
    // and odd taps alternate between two pairs, so that dependent MMAs are >= 6 MMA slots (48 cycles) apart
    static constexpr int SETS = NT >= 4 ? 2 : 4;               // (the two-row upsampling conv at 32 px has registers for 1 only)
};

// One K range (nkc chunks of 8 input channels starting at `planes`) into the warp's NT n-tiles of m-tile `mt`.
// pb: planes + t * PLANE + row * P + g (lane part folded in).  wf: the lane's float4 of chunk 0 / tap 0 / m-tile mt.
template <int S, int SETS>
__device__ __forceinline__ void conv_accumulate_mma(float (&sum)[MmaGeo<S>::NT][4], float (&corr)[MmaGeo<S>::NT][SETS][4],
                                                    const float* __restrict__ pb, int nkc, const float4* __restrict__ wf,
                                                    int wtap /* float4s between taps */) {
    constexpr int P = MmaGeo<S>::P, PLANE = MmaGeo<S>::PLANE, NT = MmaGeo<S>::NT;
#pragma unroll 1
    for (int kc = 0; kc < nkc; ++kc) {
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
            const int ky = tap / 3, kx = tap % 3;
            const float4 w4 = wf[tap * wtap];
            float ah[4], al[4];
            ah[0] = tf32_hi(w4.x); al[0] = w4.x - ah[0];
            ah[1] = tf32_hi(w4.y); al[1] = w4.y - ah[1];
            ah[2] = tf32_hi(w4.z); al[2] = w4.z - ah[2];
            ah[3] = tf32_hi(w4.w); al[3] = w4.w - ah[3];
            constexpr int kAlt = SETS == 4 ? 2 : 0, kSecond = SETS >= 2 ? 1 : 0;
            const int set = (tap & 1) * kAlt;
            float d[NT][4], h0[NT], h1[NT], l0[NT], l1[NT];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const float* q = pb + ky * P + nt * 8 + kx;
                const float b0 = q[0], b1 = q[4 * PLANE];
                h0[nt] = tf32_hi(b0); h1[nt] = tf32_hi(b1);
                l0[nt] = b0 - h0[nt]; l1[nt] = b1 - h1[nt];
                mma_tf32_zero(d[nt], ah, h0[nt], h1[nt]);
            }
            // (issue order: the two accumulating MMAs of an n-tile are NT slots apart even when they share registers)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) mma_tf32_acc(corr[nt][set], ah, l0[nt], l1[nt]);
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) mma_tf32_acc(corr[nt][set + kSecond], al, h0[nt], h1[nt]);
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                sum[nt][0] += d[nt][0]; sum[nt][1] += d[nt][1]; sum[nt][2] += d[nt][2]; sum[nt][3] += d[nt][3];
            }
        }
        pb += 8 * PLANE;
        wf += 9 * wtap;
    }
}

template <int S, int SETS>
__device__ __forceinline__ void mma_epilogue(float (&sum)[MmaGeo<S>::NT][4], const float (&corr)[MmaGeo<S>::NT][SETS][4],
                                             const FusedOp& op, float* sm, const float* bias, int mt, int y, int g, int t,
                                             int f, float kappa) {
    constexpr int P = MmaGeo<S>::P, PLANE = MmaGeo<S>::PLANE, NT = MmaGeo<S>::NT;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float c = corr[nt][0][i];
            if (SETS >= 2) c += corr[nt][1][i];
            if (SETS == 4) c += corr[nt][2][i] + corr[nt][3][i];
            sum[nt][i] += fmaf(sum[nt][i], kappa, c);                                        // (the corrections die here)
        }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int co = mt * 16 + g + 8 * h;
        if (co >= op.Cout) continue;                               // 8-channel layers: rows 8..15 of the m-tile are padding
        float2 gate[NT];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {                          // the gate loads go out before the first store
            gate[nt] = make_float2(1.f, 1.f);
            if (op.gmask)
                gate[nt] = *reinterpret_cast<const float2*>(op.gmask + (long)f * op.gmask_bs + ((long)co * S + y) * S + nt * 8 + 2 * t);
        }
        const float b = op.bias ? bias[co] : 0.f;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const int x = nt * 8 + 2 * t;
            float v0 = sum[nt][2 * h] + b, v1 = sum[nt][2 * h + 1] + b;
            if (op.relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
            v0 = gate[nt].x > 0.f ? v0 : 0.f;
            v1 = gate[nt].y > 0.f ? v1 : 0.f;
            if (op.out >= 0) {
                float* d = sm + op.out + co * PLANE + (y + 1) * P + x + 1;
                d[0] = v0; d[1] = v1;
            }
            if (op.gout)
                *reinterpret_cast<float2*>(op.gout + (long)f * op.gout_bs + ((long)co * S + y) * S + x) = make_float2(v0, v1);
        }
    }
}

template <int S>
__device__ __forceinline__ void run_conv_mma(const FusedOp& op, float* sm, int f, int tid, float kappa, long long* tm) {
    constexpr int P = MmaGeo<S>::P, PLANE = MmaGeo<S>::PLANE, NT = MmaGeo<S>::NT, SETS = MmaGeo<S>::SETS;
    const Geo geo = geo_of(S);
    const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int mtiles = (op.Cout + 15) >> 4;
    const int kc0 = (op.Cin0 + 7) >> 3, kc1 = op.Cin1 >> 3;
    const float4* w = reinterpret_cast<const float4*>(sm + op.wsm);
    const float* bias = sm + op.wsm + (kc0 + kc1) * 9 * mtiles * 128;
    if (op.out >= 0) zero_halo_planes(sm + op.out, op.Cout, geo, tid, kFusedThreads);
    PAIG_STAMP(tm, 1);
    // a warp item = (m-tile, image row): S * mtiles items over the 16 warps, m-tile the slowest index
    const int nitems = S * mtiles;
    const int wtap = mtiles * 32;
    float sum[NT][4], corr[NT][SETS][4];
    if (!op.up) {
        for (int item = warp; item < nitems; item += kFusedThreads / 32) {
            const int mt = item / S, y = item - mt * S;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    sum[nt][i] = 0.f;
#pragma unroll
                    for (int q = 0; q < SETS; ++q) corr[nt][q][i] = 0.f;
                }
            if (op.gmask) {
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                    for (int h = 0; h < 2; ++h)
                        if (mt * 16 + g + 8 * h < op.Cout)
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(op.gmask + (long)f * op.gmask_bs +
                                         ((long)(mt * 16 + g + 8 * h) * S + y) * S + nt * 8 + 2 * t));
            }
            const float4* wf = w + mt * 32 + lane;
            conv_accumulate_mma<S, SETS>(sum, corr, sm + op.in0 + t * PLANE + y * P + g, kc0, wf, wtap);
            if (kc1) conv_accumulate_mma<S, SETS>(sum, corr, sm + op.in1 + t * PLANE + y * P + g, kc1, wf + kc0 * 9 * wtap, wtap);
            mma_epilogue<S, SETS>(sum, corr, op, sm, bias, mt, y, g, t, f, kappa);
        }
        PAIG_STAMP(tm, 2);
        PAIG_STAMP(tm, 3);
    } else {
        // upsample fused into the conv: 8 channels (one K chunk) are built at a time.  One m-tile (planner); at S = 32 a
        // warp owns two rows, whose accumulators both persist across the chunks.
        constexpr int R = S >= 32 ? 2 : 1;                         // rows per warp: 2 at 32 px, 1 at 16 px (8 px: planner refuses)
        constexpr int US = R == 2 ? 1 : SETS;
        float sum2[R][NT][4], corr2[R][NT][US][4];
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    sum2[r][nt][i] = 0.f;
#pragma unroll
                    for (int q = 0; q < US; ++q) corr2[r][nt][q][i] = 0.f;
                }
        const Geo gl = geo_of(S / 2);
        float* chunk = sm + op.chunk;
        zero_halo_planes(chunk, 8, geo, tid, kFusedThreads);
        for (int c0 = 0; c0 < op.Cin0; c0 += 8) {
            upsample_chunk(sm + op.in0 + c0 * gl.plane, gl, chunk, geo, 8,
                           op.gup ? op.gup + (long)f * op.gup_bs + (long)c0 * S * S : nullptr, tid);
            __syncthreads();
#pragma unroll
            for (int r = 0; r < R; ++r)
                conv_accumulate_mma<S, US>(sum2[r], corr2[r], chunk + t * PLANE + (warp + 16 * r) * P + g, 1,
                                       w + lane + (c0 >> 3) * 9 * wtap, wtap);
            __syncthreads();
        }
#pragma unroll
        for (int r = 0; r < R; ++r) mma_epilogue<S, US>(sum2[r], corr2[r], op, sm, bias, 0, warp + 16 * r, g, t, f, kappa);
    }
    (void)sum; (void)corr;
}
#endif   // !PAIG_EMU

template <bool MMA>
__device__ __forceinline__ void run_conv_any(const FusedOp& op, float* sm, int f, int tid, float kappa, long long* tm) {
#ifndef PAIG_EMU
    if (MMA) {
        if (op.S == 32) run_conv_mma<32>(op, sm, f, tid, kappa, tm);
        else if (op.S == 16) run_conv_mma<16>(op, sm, f, tid, kappa, tm);
        else run_conv_mma<8>(op, sm, f, tid, kappa, tm);
        return;
    }
#endif
    (void)kappa;
    if (op.Cout == 8) run_conv_co<8>(op, sm, f, tid, tm);
    else if (op.Cout == 16) run_conv_co<16>(op, sm, f, tid, tm);
    else run_conv_co<32>(op, sm, f, tid, tm);
}

template <bool MMA>
__global__ void __launch_bounds__(kFusedThreads, 1) unet_fused_fwd_kernel(const FusedPlan P) {
    PAIG_DYN_SMEM(float, sm);
    __shared__ unsigned long long bars[2];
    __shared__ FusedOp s_ops[kFusedMaxOps];      // indexed kernel-parameter reads are constant-cache loads: copy once
    const int tid = threadIdx.x;
    {
        const int* src = reinterpret_cast<const int*>(P.ops);
        int* dst = reinterpret_cast<int*>(s_ops);
        for (int e = tid; e < (int)(P.nops * sizeof(FusedOp) / sizeof(int)); e += kFusedThreads) dst[e] = src[e];
    }
#ifndef PAIG_EMU
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&bars[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&bars[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
#endif
    __syncthreads();
    unsigned phase0 = 0, phase1 = 0;
    const Geo gx = geo_of(P.H);
    const int HW = P.H * P.H;
    for (int f = blockIdx.x; f < P.N; f += gridDim.x) {
        if (tid == kIssueThread && P.first_w >= 0) {
            const FusedOp& o = s_ops[P.first_w];
            bulk_issue(sm + o.wsm, P.wpack + o.wglob, (unsigned)o.wfloats * 4u, &bars[o.wbar]);
        }
        // ---- the input frame, zero halo ----
        float* X = sm + P.x_off;
        zero_halo_planes(X, 3, gx, tid, kFusedThreads);
        if (MMA)                              // the first conv's K chunk is 8 channels wide: planes 3..7 are zeros
            for (int e = tid; e < 5 * gx.plane; e += kFusedThreads) X[3 * gx.plane + e] = 0.f;
        const float* xf = P.x + (long)(f / P.fps) * P.seq_stride + (long)(f % P.fps) * 3 * HW;
        if ((P.H & 3) == 0) {
            const int rowq = P.H / 4;
            for (int e = tid; e < 3 * HW / 4; e += kFusedThreads) {
                const int q = e % rowq, y = (e / rowq) % P.H, c = e / (rowq * P.H);
                const float4 v = *reinterpret_cast<const float4*>(xf + (long)c * HW + y * P.H + 4 * q);
                float* d = X + c * gx.plane + (y + 1) * gx.P + 4 * q + 1;
                d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
            }
        } else {
            for (int e = tid; e < 3 * HW; e += kFusedThreads) {
                const int x = e % P.H, y = (e / P.H) % P.H, c = e / HW;
                X[c * gx.plane + (y + 1) * gx.P + x + 1] = xf[e];
            }
        }
        __syncthreads();
#ifndef PAIG_EMU
        if (P.timing && tid == 0 && f == blockIdx.x + gridDim.x) P.timing[(long)blockIdx.x * 160] = clock64();
#endif
        for (int t = 0; t < P.nops; ++t) {
            const FusedOp op = s_ops[t];
            if (tid == kIssueThread && op.next_w >= 0) {
                const FusedOp& o = s_ops[op.next_w];
                bulk_issue(sm + o.wsm, P.wpack + o.wglob, (unsigned)o.wfloats * 4u, &bars[o.wbar]);
            }
            long long* tm = nullptr;
#ifndef PAIG_EMU
            if (P.timing && tid == 0 && f == blockIdx.x + gridDim.x) tm = P.timing + (long)blockIdx.x * 160 + 32 + t * 4;
#endif
            if (op.wfloats) {
                if (op.wbar == 0) { bulk_wait(&bars[0], phase0); phase0 ^= 1u; }
                else { bulk_wait(&bars[1], phase1); phase1 ^= 1u; }
            }
            PAIG_STAMP(tm, 0);
            if (op.kind == F_CONV) {
                run_conv_any<MMA>(op, sm, f, tid, P.kappa, tm);
            } else if (op.kind == F_POOL) {
                const int So = op.S, C = op.Cin0;
                const Geo go = geo_of(So), gi = geo_of(2 * So);
                if (op.out >= 0) zero_halo_planes(sm + op.out, C, go, tid, kFusedThreads);
                for (int e = tid; e < C * So * So; e += kFusedThreads) {
                    const int x = e % So, y = (e / So) % So, c = e / (So * So);
                    const float* p = sm + op.in0 + c * gi.plane + (2 * y + 1) * gi.P + 2 * x + 1;
                    const float m = fmaxf(fmaxf(p[0], p[1]), fmaxf(p[gi.P], p[gi.P + 1]));
                    if (op.out >= 0) sm[op.out + c * go.plane + (y + 1) * go.P + x + 1] = m;
                    if (op.gout) op.gout[(long)f * op.gout_bs + ((long)c * So + y) * So + x] = m;
                }
            } else {   // 1x1 head: logits[o] = (relu)(b[o] + sum_c w[o][c] * in[c])
                const int S = op.S, Cin = op.Cin0;
                const Geo g = geo_of(S);
                const float* w = sm + op.wsm;
                const float* b = w + op.Cout * Cin;
                for (int e = tid; e < S * S; e += kFusedThreads) {
                    const int x = e % S, y = e / S;
                    const float* p = sm + op.in0 + (y + 1) * g.P + x + 1;
                    for (int co = 0; co < op.Cout; ++co) {
                        float s = b[co];
                        for (int c = 0; c < Cin; ++c) s += w[co * Cin + c] * p[c * g.plane];
                        if (op.relu) s = fmaxf(s, 0.f);
                        op.gout[(long)f * op.gout_bs + (long)co * S * S + e] = s;
                    }
                }
            }
            __syncthreads();
#ifndef PAIG_EMU
            if (P.timing && tid == 0 && f == blockIdx.x + gridDim.x) P.timing[(long)blockIdx.x * 160 + 1 + t] = clock64();
#endif
        }
    }
}

// ---- adjoint ops of the backward-data pass -----------------------------------------------------------------------
// upsample adjoint (gather): input (i,j) of a 2x bilinear upsample collects outputs 2i-1..2i+2 x 2j-1..2j+2 with
// weights (.25,.75,.75,.25), the clamped border taps folding back (conv.cu up_adj, same arithmetic).
__device__ __forceinline__ void up_adj_f(int i, int Si, float (&w)[4]) {
    w[0] = i > 0 ? 0.25f : 0.f;
    w[1] = i > 0 ? 0.75f : 1.f;
    w[2] = i < Si - 1 ? 0.75f : 1.f;
    w[3] = i < Si - 1 ? 0.25f : 0.f;
}

__device__ __forceinline__ void run_upT(const FusedOp& op, float* sm, int f, int tid) {
    const int Si = op.S, C = op.Cin0;
    const Geo gi = geo_of(Si), go = geo_of(2 * Si);
    if (op.out >= 0) zero_halo_planes(sm + op.out, C, gi, tid, kFusedThreads);
    // a thread owns one low-resolution position and a residue class of the channels
    const int per = Si * Si;
    const int nsub = per < kFusedThreads ? kFusedThreads / per : 1;
    for (int wk = tid; wk < per * nsub; wk += kFusedThreads) {
        const int e = wk % per, sub = wk / per;
        const int i = e / Si, j = e % Si;
        float wy[4], wx[4];
        up_adj_f(i, Si, wy);
        up_adj_f(j, Si, wx);
        // output (2i-1, 2j-1) sits at tile (2i, 2j); the halo rows/columns it may touch carry weight 0
        const float* g0 = sm + op.in0 + (2 * i) * go.P + 2 * j;
        for (int cb = sub; cb < C; cb += 4 * nsub) {                  // 4 channels per trip: their gate loads overlap
            float mv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int c = cb + u * nsub;
                mv[u] = (op.gmask && c < C) ? op.gmask[(long)f * op.gmask_bs + ((long)c * Si + i) * Si + j] : 1.f;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int c = cb + u * nsub;
                if (c >= C) continue;
                const float* g = g0 + c * go.plane;
                float s = 0.f;
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    if (wy[a] == 0.f) continue;
                    float r = 0.f;
#pragma unroll
                    for (int b = 0; b < 4; ++b)
                        if (wx[b] != 0.f) r += wx[b] * g[a * go.P + b];
                    s += wy[a] * r;
                }
                if (!(mv[u] > 0.f)) s = 0.f;
                if (op.out >= 0) sm[op.out + c * gi.plane + (i + 1) * gi.P + j + 1] = s;
                if (op.gout) op.gout[(long)f * op.gout_bs + ((long)c * Si + i) * Si + j] = s;
            }
        }
    }
}

// max-pool adjoint: each 2x2 window of the source X routes the pooled gradient to its first maximum (row-major, as
// ATen / conv.cu maxpool2_bwd_kernel), adds the gradient that reached X through its other consumer (in1, parked
// on chip) and applies X's own ReLU mask.  One thread per window and channel residue class; no atomics.
__device__ __forceinline__ void run_poolT(const FusedOp& op, float* sm, int f, int tid) {
    const int So = op.S, Si = 2 * So, C = op.Cin0;            // So: pooled side, Si: source side
    const Geo gp = geo_of(So), gx = geo_of(Si);
    if (op.out >= 0 && op.out != op.in1) zero_halo_planes(sm + op.out, C, gx, tid, kFusedThreads);
    const int per = So * So;
    const int nsub = per < kFusedThreads ? kFusedThreads / per : 1;
    for (int wk = tid; wk < per * nsub; wk += kFusedThreads) {
        const int e = wk % per, sub = wk / per;
        const int y = e / So, x = e % So;
        for (int cb = sub; cb < C; cb += 4 * nsub) {                  // 4 channels per trip: their source loads overlap
            float2 xa[4], xb[4], pa[4], pb[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int c = cb + u * nsub;
                pa[u] = pb[u] = make_float2(0.f, 0.f);
                if (c < C) {
                    const float* xs = op.gmask + (long)f * op.gmask_bs + ((long)c * Si + 2 * y) * Si + 2 * x;
                    xa[u] = *reinterpret_cast<const float2*>(xs);
                    xb[u] = *reinterpret_cast<const float2*>(xs + Si);
                    if (op.acc_gout) {                                 // share of the other reader, parked in HBM/L2
                        const float* ps = op.gout + (long)f * op.gout_bs + ((long)c * Si + 2 * y) * Si + 2 * x;
                        pa[u] = *reinterpret_cast<const float2*>(ps);
                        pb[u] = *reinterpret_cast<const float2*>(ps + Si);
                    }
                } else {
                    xa[u] = xb[u] = make_float2(0.f, 0.f);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int c = cb + u * nsub;
                if (c >= C) continue;
                const float v[4] = {xa[u].x, xa[u].y, xb[u].x, xb[u].y};
                int best = 0;
                float m = v[0];
                if (v[1] > m) { m = v[1]; best = 1; }
                if (v[2] > m) { m = v[2]; best = 2; }
                if (v[3] > m) { m = v[3]; best = 3; }
                const float g = sm[op.in0 + c * gp.plane + (y + 1) * gp.P + x + 1];
                const int t00 = c * gx.plane + (2 * y + 1) * gx.P + 2 * x + 1;
                const int toff[4] = {t00, t00 + 1, t00 + gx.P, t00 + gx.P + 1};
                const long g00 = (long)f * op.gout_bs + ((long)c * Si + 2 * y) * Si + 2 * x;
                const long goff[4] = {g00, g00 + 1, g00 + Si, g00 + Si + 1};
                const float parked[4] = {pa[u].x, pa[u].y, pb[u].x, pb[u].y};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    float r = op.in1 >= 0 ? sm[op.in1 + toff[k]] : parked[k];
                    if (k == best) r += g;
                    if (op.relu && !(v[k] > 0.f)) r = 0.f;
                    if (op.out >= 0) sm[op.out + toff[k]] = r;
                    if (op.gout) op.gout[goff[k]] = r;
                }
            }
        }
    }
}

// 1x1 head adjoint: d in[c] = sum_o w[o][c] * g[o], g = d logits gated by the head's own ReLU; result gated by the
// ReLU of the layer that produced `in`.
__device__ __forceinline__ void run_headT(const FusedOp& op, float* sm, int f, int tid) {
    const int S = op.S, C = op.Cout, NO = op.Cin0;            // C: channels of the head's input, NO: logits
    const Geo g = geo_of(S);
    const float* w = sm + op.wsm;                             // [NO][C]
    if (op.out >= 0) zero_halo_planes(sm + op.out, C, g, tid, kFusedThreads);
    for (int e = tid; e < S * S; e += kFusedThreads) {
        const int y = e / S, x = e % S;
        float gl[kMaxObjs], hm[kMaxObjs];
        // global loads first (d logits and the head's own ReLU gate), then the output gates 8 channels at a time
#pragma unroll
        for (int o = 0; o < kMaxObjs; ++o) {
            gl[o] = o < NO ? op.gsrc[(long)f * op.gsrc_bs + (long)o * S * S + e] : 0.f;
            hm[o] = (o < NO && op.gmask2) ? op.gmask2[(long)f * op.gmask2_bs + (long)o * S * S + e] : 1.f;
        }
#pragma unroll
        for (int o = 0; o < kMaxObjs; ++o)
            if (!(hm[o] > 0.f)) gl[o] = 0.f;
        for (int c0 = 0; c0 < C; c0 += 8) {
            float mk[8];
#pragma unroll
            for (int u = 0; u < 8; ++u)
                mk[u] = (c0 + u < C && op.gmask) ? op.gmask[(long)f * op.gmask_bs + (long)(c0 + u) * S * S + e] : 1.f;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int c = c0 + u;
                if (c < C) {
                    float d = 0.f;
#pragma unroll
                    for (int o = 0; o < kMaxObjs; ++o)
                        if (o < NO) d += w[o * C + c] * gl[o];
                    if (!(mk[u] > 0.f)) d = 0.f;
                    if (op.out >= 0) sm[op.out + c * g.plane + (y + 1) * g.P + x + 1] = d;
                    if (op.gout) op.gout[(long)f * op.gout_bs + (long)c * S * S + e] = d;
                }
            }
        }
    }
}

// Backward-data pass of the whole UNet for one frame at a time: the gradient of every conv output, already gated by
// that layer's ReLU, goes to the workspace (where the weight-gradient kernels read it) and stays on chip for the
// next transposed conv.  Same CTA shape, planner and weight pipeline as the forward kernel.
template <bool MMA>
__global__ void __launch_bounds__(kFusedThreads, 1) unet_fused_bwd_kernel(const FusedPlan P) {
    PAIG_DYN_SMEM(float, sm);
    __shared__ unsigned long long bars[2];
    __shared__ FusedOp s_ops[kFusedMaxOps];
    const int tid = threadIdx.x;
    {
        const int* src = reinterpret_cast<const int*>(P.ops);
        int* dst = reinterpret_cast<int*>(s_ops);
        for (int e = tid; e < (int)(P.nops * sizeof(FusedOp) / sizeof(int)); e += kFusedThreads) dst[e] = src[e];
    }
#ifndef PAIG_EMU
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&bars[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&bars[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
#endif
    __syncthreads();
    unsigned phase0 = 0, phase1 = 0;
    for (int f = blockIdx.x; f < P.N; f += gridDim.x) {
        if (tid == kIssueThread && P.first_w >= 0) {
            const FusedOp& o = s_ops[P.first_w];
            bulk_issue(sm + o.wsm, P.wpack + o.wglob, (unsigned)o.wfloats * 4u, &bars[o.wbar]);
        }
#ifndef PAIG_EMU
        if (P.timing && tid == 0 && f == blockIdx.x + gridDim.x) P.timing[(long)blockIdx.x * 160] = clock64();
#endif
        for (int t = 0; t < P.nops; ++t) {
            const FusedOp op = s_ops[t];
            if (tid == kIssueThread && op.next_w >= 0) {
                const FusedOp& o = s_ops[op.next_w];
                bulk_issue(sm + o.wsm, P.wpack + o.wglob, (unsigned)o.wfloats * 4u, &bars[o.wbar]);
            }
            long long* tm = nullptr;
#ifndef PAIG_EMU
            if (P.timing && tid == 0 && f == blockIdx.x + gridDim.x) tm = P.timing + (long)blockIdx.x * 160 + 32 + t * 4;
#endif
            if (op.wfloats) {
                if (op.wbar == 0) { bulk_wait(&bars[0], phase0); phase0 ^= 1u; }
                else { bulk_wait(&bars[1], phase1); phase1 ^= 1u; }
            }
            PAIG_STAMP(tm, 0);
            if (op.kind == F_CONV) {
                run_conv_any<MMA>(op, sm, f, tid, P.kappa, tm);
            } else if (op.kind == F_UPT) {
                run_upT(op, sm, f, tid);
            } else if (op.kind == F_POOLT) {
                run_poolT(op, sm, f, tid);
            } else {
                run_headT(op, sm, f, tid);
            }
            __syncthreads();
#ifndef PAIG_EMU
            if (P.timing && tid == 0 && f == blockIdx.x + gridDim.x) P.timing[(long)blockIdx.x * 160 + 1 + t] = clock64();
#endif
        }
    }
}

// ---- weight packing: [co][ci][tap] -> [ci][tap][co] | bias  (head: [co][ci] | bias kept as is) ----------------
__global__ void __launch_bounds__(256) pack_weights_kernel(const PackPlan P, float* __restrict__ dst) {
    const int l = blockIdx.y;
    if (l >= P.nlayers) return;
    const int Cout = P.Cout[l], Cin = P.Cin[l], taps = P.taps[l];
    const int nW = Cout * Cin * taps;
    float* d = dst + P.off[l];
    if (P.frag && taps == 9) {
        const int mtiles = (Cout + 15) >> 4, nkc = (Cin + 7) >> 3;
        const int nF = nkc * 9 * mtiles * 128;
        const bool tr = P.mode[l] == 1;
        for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nF + (tr ? 0 : Cout); e += gridDim.x * blockDim.x) {
            if (e >= nF) { d[e] = P.b[l][e - nF]; continue; }
            const int j = e & 3, lane = (e >> 2) & 31;
            int r = e >> 7;
            const int mt = r % mtiles; r /= mtiles;
            const int tap = r % 9, kc = r / 9;
            const int co = mt * 16 + (lane >> 2) + ((j & 1) ? 8 : 0), ci = kc * 8 + (lane & 3) + ((j & 2) ? 4 : 0);
            float v = 0.f;
            if (co < Cout && ci < Cin)
                v = tr ? P.w[l][((long)ci * P.cin_total[l] + P.ci0[l] + co) * 9 + (8 - tap)] : P.w[l][((long)co * Cin + ci) * 9 + tap];
            d[e] = v;
        }
        return;
    }
    if (P.mode[l] == 2) {                                    // head weights as they are, no bias
        for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nW; e += gridDim.x * blockDim.x) d[e] = P.w[l][e];
        return;
    }
    if (P.mode[l] == 1) {
        for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nW; e += gridDim.x * blockDim.x) {
            const int c = e % Cout, tap = (e / Cout) % 9, co = e / (Cout * 9);
            d[e] = P.w[l][((long)co * P.cin_total[l] + P.ci0[l] + c) * 9 + (8 - tap)];
        }
        return;
    }
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nW + Cout; e += gridDim.x * blockDim.x) {
        if (e >= nW) {
            d[e] = P.b[l][e - nW];
        } else if (taps == 1) {
            d[e] = P.w[l][e];
        } else {
            const int co = e % Cout, tap = (e / Cout) % taps, ci = e / (Cout * taps);
            d[e] = P.w[l][((long)co * Cin + ci) * taps + tap];
        }
    }
}

// ---- host: plan + launch -------------------------------------------------------------------------------------------
namespace {

struct Slice {
    int buf, c0, C, S;       // UNet buffer slice (buf -1: the input frame)
    int born, last;          // steps (fused-op indices); born -1: before step 0
    int off;                 // shared-memory offset (floats)
    int floats;
};

// (struct Planner: layout.h)

// Thread tile of a conv: CO output channels x PY rows x 4 pixels.  Cost model per input channel, for the busiest
// SM sub-partition: FMA issue slots vs shared-memory wavefronts (measured: 16-byte pixel load 4, 8-byte 4 (2-way
// conflict), broadcast 16-byte weight load 2), with a penalty when too few warps remain to hide latency.
bool choose_tile(FusedOp& fo, bool single_pass) {
    static const char* force = getenv("PAIG_TILE");          // "co,py": experiments
    int fco = 0, fpy = 0;
    if (force) sscanf(force, "%d,%d", &fco, &fpy);
    const Geo g = geo_of(fo.S);
    double best_cost = 0;
    int best_co = 0, best_py = 0;
    for (int py = 1; py <= 2; ++py) {
        if (fo.S % py) continue;
        for (int co = 4; co <= 16; co *= 2) {
            if (fo.Cout % co || (py == 2 && co >= 8)) continue;
            const int items = (fo.S / py) * g.nqx * (fo.Cout / co);
            const int passes = (items + kFusedThreads - 1) / kFusedThreads;
            if (single_pass && passes > 1) continue;
            const int warps = ((items < kFusedThreads ? items : kFusedThreads) + 31) / 32;
            // Calibrated on B200 (per-op cycle stamps, PAIG_DEBUG): an op takes (instructions issued by the busiest
            // sub-partition) / IPC, with IPC 0.55 / 0.67 / 0.75 / 0.80 at 1 / 2 / 3 / >=4 resident warps; the
            // shared-memory pipe (one wavefront per cycle per SM) is the second bound.
            const int wps = (warps + 3) / 4;
            const double per_warp = 36.0 * co * py + (py + 2) * 2 + 9 * co / 4 + 6;
            const double ipc = wps >= 4 ? 0.80 : (wps == 3 ? 0.75 : (wps == 2 ? 0.67 : 0.55));
            const double issue = (double)passes * wps * per_warp / ipc;
            const double lsu = (double)passes * warps * ((py + 2) * 8 + 9 * co / 4 * 2);
            double cost = issue > lsu ? issue : lsu;
            const bool forced = co == fco && py == fpy;
            if (forced) cost = 1.0;
            if (!best_co || cost < best_cost) { best_cost = cost; best_co = co; best_py = py; }
        }
    }
    if (!best_co) return false;
    fo.co_tile = best_co;
    fo.py_tile = best_py;
    return true;
}

// Tensor-core variant of the fused kernels (mma.sync 3xTF32, see run_conv_mma): 32-px frames only (every level a multiple of
// 8 pixels wide, plane strides of 8 / 24 mod 32 banks).  PAIG_FUSED_MMA=1 selects it.
bool fused_mma_enabled(int H) {
#ifdef PAIG_EMU
    (void)H;
    return false;
#else
    // measured (profiles/r2n_*): the HMMA pipe (one m16n8k8 per 8 cycles and scheduler) plus the split arithmetic leaves
    // the variant at 0.81 / 0.78 ms against 0.71 / 0.79 ms for the FMA loops on spring_color -- opt-in until it wins
    static const bool on = getenv("PAIG_FUSED_MMA") != nullptr && getenv("PAIG_NO_MMA") == nullptr;
    return on && H == 32;
#endif
}
float fused_mma_kappa() {
    static const float k = getenv("PAIG_MMA_KAPPA") ? (float)atof(getenv("PAIG_MMA_KAPPA")) : 0.20f;
    return k * 1.1920929e-7f;
}
inline int pad8(int c) { return (c + 7) & ~7; }
inline int pad16(int c) { return (c + 15) & ~15; }

int sm_count() {
#ifdef PAIG_EMU
    return 2;
#else
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
#endif
}

}  // namespace

size_t unet_wpack_floats(const UNetDesc& u, const paig_task* t) {
    (void)t;
    size_t total = 48 * 64;      // forward and backward-data packings, each entry padded to 256 bytes
    for (int i = 0; i < u.nops; ++i) {
        const Op& op = u.ops[i];
        // (fragment order of the tensor-core variant pads K to 8 and M to 16 channels, per concat slice in backward)
        if (op.kind == OP_CONV) total += align64((size_t)(pad8(op.in.C) + 16) * 9 * pad16(op.out.C) + op.out.C);
        else if (op.kind == OP_HEAD) total += align64((size_t)op.in.C * op.out.C + op.out.C);
    }
    return 2 * total;
}

#ifndef PAIG_EMU
// PAIG_DEBUG: cycles per op (second frame of every CTA, mean over CTAs) and thread 0's phases inside each conv
static long long* g_timing_buf = nullptr;
static long long* timing_buffer() {
    if (!g_timing_buf) cudaMalloc(&g_timing_buf, (size_t)160 * 160 * sizeof(long long));
    return g_timing_buf;
}
static void print_timing(const char* what, const FusedPlan& P, int grid, int N, cudaStream_t st) {
    if (N < 2 * grid) return;
    cudaStreamSynchronize(st);
    static long long host[160 * 160];
    cudaMemcpy(host, g_timing_buf, (size_t)grid * 160 * sizeof(long long), cudaMemcpyDeviceToHost);
    fprintf(stderr, "[paig] %s cycles per op (mean over %d CTAs):", what, grid);
    double total = 0;
    for (int k = 0; k < P.nops; ++k) {
        double s = 0;
        for (int b = 0; b < grid; ++b) s += (double)(host[b * 160 + 1 + k] - host[b * 160 + k]);
        fprintf(stderr, " op%d=%.0f", k, s / grid);
        total += s / grid;
    }
    fprintf(stderr, " total=%.0f\n", total);
    fprintf(stderr, "[paig]   thread 0 of CTA 0: op: wait | halo | compute | epilogue | sync\n");
    for (int k = 0; k < P.nops; ++k) {
        const long long* tmh = host + 32 + k * 4;
        if (P.ops[k].kind == F_CONV && P.ops[k].up)
            fprintf(stderr, "[paig]   op%-2d (up) wait %lld | halo %lld | upsample total %lld | accumulate total %lld\n", k,
                    tmh[0] - host[k], tmh[1] - tmh[0], tmh[2], tmh[3]);
        if (P.ops[k].kind == F_CONV && !P.ops[k].up)
            fprintf(stderr, "[paig]   op%-2d %6lld | %6lld | %6lld | %6lld | %6lld\n", k, tmh[0] - host[k], tmh[1] - tmh[0],
                    tmh[2] - tmh[1], tmh[3] - tmh[2], host[1 + k] - tmh[3]);
    }
}
#endif

// Returns 0 on success, 1 on error, -1 when the network does not fit on chip (caller uses the per-layer path).
int unet_fused_forward(const paig_task* t, const paig_params* p, const Layout& L, const float* x, long seq_stride, int fps,
                       float* ws, cudaStream_t st) {
    const UNetDesc& u = L.unet;
    const Dims& d = L.d;
    FusedPlan P;
    memset(&P, 0, sizeof(P));
    PackPlan K;
    memset(&K, 0, sizeof(K));

    for (int up_chunk = 8; up_chunk >= 2; up_chunk /= 2) {
        bool mma = fused_mma_enabled(d.H) && up_chunk == 8;
      replan:
        // ---- 1. fused op list (an UP is merged into the CONV that reads it) and the slices each op writes ----
        Slice sl[40];
        int ns = 0;
        sl[ns++] = Slice{-1, 0, 3, d.H, -1, 0, 0, 0};
        int nf = 0;
        int src_of[kFusedMaxOps];            // slice index of segment 0 / the low-res source
        int src1_of[kFusedMaxOps];
        int out_of[kFusedMaxOps];
        bool ok = true;
        long woff = 0;
        int nl = 0;
        auto find_slices = [&](const Ref& r, int* a, int* b) {
            *a = *b = -1;
            for (int i = 0; i < ns; ++i) {
                if (sl[i].buf != r.buf) continue;
                if (r.buf == -1) { *a = i; return true; }
                if (sl[i].c0 >= r.c0 && sl[i].c0 + sl[i].C <= r.c0 + r.C) {
                    if (*a < 0) *a = i;
                    else if (*b < 0) { if (sl[i].c0 < sl[*a].c0) { *b = *a; *a = i; } else *b = i; }
                    else return false;
                }
            }
            if (*a < 0) return false;
            const int covered = sl[*a].C + (*b >= 0 ? sl[*b].C : 0);
            return covered == r.C && sl[*a].c0 == r.c0;
        };
        for (int i = 0; i < u.nops && ok; ++i) {
            const Op& op = u.ops[i];
            if (nf >= kFusedMaxOps) { ok = false; break; }
            FusedOp& fo = P.ops[nf];
            memset(&fo, 0, sizeof(fo));
            fo.out = fo.chunk = fo.in1 = -1;
            fo.next_w = -1;
            if (op.kind == OP_UP) {
                // must be consumed, whole, by the next op (a conv)
                if (i + 1 >= u.nops || u.ops[i + 1].kind != OP_CONV || u.ops[i + 1].in.buf != op.out.buf ||
                    u.ops[i + 1].in.c0 != op.out.c0 || u.ops[i + 1].in.C != op.out.C) { ok = false; break; }
                continue;       // handled when the conv is visited
            }
            const int Sout = op.kind == OP_HEAD ? d.H : (d.H >> u.bufs[op.out.buf].shift);
            fo.S = Sout;
            fo.relu = op.relu;
            int a = -1, b = -1;
            if (op.kind == OP_CONV && i > 0 && u.ops[i - 1].kind == OP_UP && u.ops[i - 1].out.buf == op.in.buf) {
                const Op& upo = u.ops[i - 1];
                if (!find_slices(upo.in, &a, &b) || b >= 0) { ok = false; break; }
                fo.up = up_chunk;
                const BufDesc& bd = u.bufs[upo.out.buf];
                const int Su = d.H >> bd.shift;
                fo.gup = ws + L.act[upo.out.buf] + (long)upo.out.c0 * Su * Su;
                fo.gup_bs = (long)bd.C * Su * Su;
                fo.Cin0 = op.in.C;
            } else {
                if (!find_slices(op.in, &a, &b)) { ok = false; break; }
                fo.Cin0 = sl[a].C;
                fo.Cin1 = b >= 0 ? sl[b].C : 0;
            }
            src_of[nf] = a;
            src1_of[nf] = b;
            sl[a].last = nf;
            if (b >= 0) sl[b].last = nf;
            if (op.kind == OP_POOL) {
                fo.kind = F_POOL;
                fo.Cout = op.out.C;
            } else {
                fo.kind = op.kind == OP_CONV ? F_CONV : F_HEAD;
                fo.Cout = op.out.C;
                const int taps = op.kind == OP_CONV ? 9 : 1;
                const int cin = fo.Cin0 + fo.Cin1;
                size_t wf = (size_t)cin * taps * fo.Cout + fo.Cout;
                if (mma && taps == 9) {
                    // K chunks of 8 channels must not straddle the two concat segments; a narrower first segment is the
                    // input frame (3 channels, zero planes behind it); an upsampling conv keeps one m-tile of accumulators
                    if ((fo.Cin1 && (fo.Cin0 % 8 || fo.Cin1 % 8)) || (fo.Cin0 % 8 && a != 0) || (fo.up && fo.Cout > 16) ||
                        (Sout != 32 && Sout != 16 && Sout != 8) || (fo.up && Sout < 16)) { mma = false; goto replan; }
                    wf = (size_t)(pad8(cin) / 8) * 9 * (pad16(fo.Cout) / 16) * 128 + fo.Cout;
                }
                fo.wfloats = ((int)wf + 3) & ~3;
                fo.wglob = woff;
                fo.bias = 1;
                K.w[nl] = p->conv[op.layer].w; K.b[nl] = p->conv[op.layer].b;
                K.Cout[nl] = fo.Cout; K.Cin[nl] = cin; K.taps[nl] = taps; K.off[nl] = woff;
                ++nl;
                woff += (long)align64(wf);
                if (fo.kind == F_CONV && fo.Cout != 8 && fo.Cout != 16 && fo.Cout != 32) { ok = false; break; }
            }
            if (op.kind == OP_HEAD) {
                fo.gout = ws + L.logits;
                fo.gout_bs = (long)d.n * d.HW;
                out_of[nf] = -1;
            } else {
                const BufDesc& bd = u.bufs[op.out.buf];
                fo.gout = ws + L.act[op.out.buf] + (long)op.out.c0 * Sout * Sout;
                fo.gout_bs = (long)bd.C * Sout * Sout;
                sl[ns] = Slice{op.out.buf, op.out.c0, op.out.C, Sout, nf, -1, 0, 0};
                out_of[nf] = ns++;
            }
            ++nf;
        }
        if (!ok) return -1;
        if (t->flags & PAIG_FLAG_INFERENCE) {
            // forward only: no activation leaves the SM except the logits (354 MB of stores per 100 sequences saved)
            for (int k = 0; k < nf; ++k) {
                if (P.ops[k].kind != F_HEAD) P.ops[k].gout = nullptr;
                P.ops[k].gup = nullptr;
            }
        }
        K.nlayers = nl;
        K.frag = mma ? 1 : 0;
        P.nops = nf;
        P.kappa = fused_mma_kappa();
        // ---- 2. thread tiling of each conv ----
        for (int k = 0; k < nf; ++k)
            if (!mma && P.ops[k].kind == F_CONV && !choose_tile(P.ops[k], P.ops[k].up != 0)) return -1;
        // ---- 3. weight prefetch chain + barriers ----
        int prev = -1, widx = 0;
        P.first_w = -1;
        for (int k = 0; k < nf; ++k) {
            if (!P.ops[k].wfloats) continue;
            P.ops[k].wbar = widx++ & 1;
            if (prev < 0) P.first_w = k;
            else P.ops[prev].next_w = k;       // issued at the start of op `prev`... (see below)
            prev = k;
        }
        // the prefetch for op k is issued when the previous *weighted* op starts; its buffer must be free from then on
        int issue_at[kFusedMaxOps];
        {
            int last_w = -1;
            for (int k = 0; k < nf; ++k) {
                issue_at[k] = -1;
                if (!P.ops[k].wfloats) continue;
                issue_at[k] = last_w;          // -1: frame start
                last_w = k;
            }
        }
        // ---- 4. shared-memory offsets by lifetime ----
        Planner al;
        sl[0].floats = (mma ? 8 : 3) * geo_of(d.H).plane;
        al.add(sl[0].floats, -1, sl[0].last, &sl[0].off);
        for (int k = 0; k < nf; ++k) {
            FusedOp& fo = P.ops[k];
            if (fo.wfloats) al.add(fo.wfloats, issue_at[k], k, &fo.wsm);        // in flight from the op that issues it
            if (fo.up) al.add(fo.up * geo_of(fo.S).plane, k, k, &fo.chunk);
            if (out_of[k] >= 0 && sl[out_of[k]].last >= 0) {
                Slice& s2 = sl[out_of[k]];
                s2.floats = s2.C * geo_of(s2.S).plane;
                al.add(s2.floats, k, s2.last, &s2.off);
            }
        }
        const int peak = al.place();
        P.x_off = sl[0].off;
        for (int k = 0; k < nf; ++k) {
            FusedOp& fo = P.ops[k];
            if (out_of[k] >= 0 && sl[out_of[k]].last >= 0) fo.out = sl[out_of[k]].off;
            fo.in0 = sl[src_of[k]].off;
            if (src1_of[k] >= 0) fo.in1 = sl[src1_of[k]].off;
        }
        static const bool debug = getenv("PAIG_DEBUG") != nullptr;
        if (debug) {
            fprintf(stderr, "[paig] fused UNet plan: H=%d ops=%d up_chunk=%d smem=%zu B\n", d.H, nf, up_chunk,
                    (size_t)peak * sizeof(float));
            for (int k = 0; k < nf; ++k) {
                const FusedOp& fo = P.ops[k];
                fprintf(stderr, "[paig]   op%-2d kind=%d S=%-2d Cin=%d+%d Cout=%-2d up=%d co=%-2d py=%d in0=%d in1=%d out=%d chunk=%d w@%d(%d)\n",
                        k, fo.kind, fo.S, fo.Cin0, fo.Cin1, fo.Cout, fo.up, fo.co_tile, fo.py_tile, fo.in0, fo.in1, fo.out, fo.chunk,
                        fo.wsm, fo.wfloats);
            }
        }
        if ((size_t)peak * sizeof(float) > kFusedSmemLimit) continue;       // try a smaller upsample chunk
        // ---- 5. launch ----
        P.N = L.N; P.fps = fps; P.H = d.H; P.seq_stride = seq_stride; P.x = x;
        float* wpack = ws + L.wpack;
        P.wpack = wpack;
        launch(pack_weights_kernel, dim3(4, K.nlayers), dim3(256), 0, st, K, wpack);
        int rc = check_launch("pack_weights");
        if (rc) return rc;
        int grid = L.N < sm_count() ? L.N : sm_count();
#ifndef PAIG_EMU
        P.timing = debug ? timing_buffer() : nullptr;
#endif
        if (mma) launch(unet_fused_fwd_kernel<true>, dim3(grid), dim3(kFusedThreads), (size_t)peak * sizeof(float), st, P);
        else launch(unet_fused_fwd_kernel<false>, dim3(grid), dim3(kFusedThreads), (size_t)peak * sizeof(float), st, P);
        (void)t;
        rc = check_launch("unet_fused_fwd");
#ifndef PAIG_EMU
        if (debug && !rc) print_timing("fused UNet forward", P, grid, L.N, st);
#endif
        return rc;
    }
    return -1;
}

// Offset (floats) where the backward-data packing starts inside the wpack region.
static long wpack_bwd_base(const UNetDesc& u) { return (long)(unet_wpack_floats(u, nullptr) / 2); }

// Backward-data pass of the UNet in one persistent kernel.  Expects d_logits in the workspace; writes the ReLU-gated
// gradient of every conv output into the workspace gradient buffers (what conv3x3_wgrad reads).  Returns -1 when
// the network or its pattern of skip connections is not supported (caller runs the per-layer kernels).
// park_global: a skip connection's gradient waits for the max-pool adjoint in the workspace gradient buffer (L2)
// instead of on chip -- 8-16 planes less shared memory, which is what lets the 36-px frames of 3bp fit.
// Returns -2 when the plan needs more shared memory than an SM has.
static int fused_backward_plan(const paig_task* t, const paig_params* p, const Layout& L, float* ws, cudaStream_t st,
                               bool park_global, BwdOps* info = nullptr) {
    const UNetDesc& u = L.unet;
    const Dims& d = L.d;
    FusedPlan P;
    memset(&P, 0, sizeof(P));
    PackPlan K;
    memset(&K, 0, sizeof(K));
    const bool mma = fused_mma_enabled(d.H);

    // forward facts: who produced each slice, with a ReLU or not, and how many ops read it
    struct Prod { int buf, c0, C, op, kind, relu, readers; };
    Prod pr[40];
    int npr = 0;
    for (int i = 0; i < u.nops; ++i) {
        const Op& op = u.ops[i];
        if (op.kind == OP_HEAD) continue;
        pr[npr++] = Prod{op.out.buf, op.out.c0, op.out.C, i, op.kind, op.relu, 0};
    }
    auto prod_of = [&](int buf, int c0, int C) -> int {
        for (int k = 0; k < npr; ++k)
            if (pr[k].buf == buf && pr[k].c0 == c0 && pr[k].C == C) return k;
        return -1;
    };
    auto prods_in = [&](const Ref& r, int* a, int* b) {          // producer slices tiling the range r, by channel
        *a = *b = -1;
        for (int k = 0; k < npr; ++k) {
            if (pr[k].buf != r.buf || pr[k].c0 < r.c0 || pr[k].c0 + pr[k].C > r.c0 + r.C) continue;
            if (*a < 0) *a = k;
            else if (*b < 0) { if (pr[k].c0 < pr[*a].c0) { *b = *a; *a = k; } else *b = k; }
            else return false;
        }
        if (*a < 0) return false;
        return pr[*a].C + (*b >= 0 ? pr[*b].C : 0) == r.C && pr[*a].c0 == r.c0;
    };
    for (int i = 0; i < u.nops; ++i) {
        const Op& op = u.ops[i];
        if (op.in.buf < 0) continue;
        int a, b;
        if (!prods_in(op.in, &a, &b)) return -1;
        pr[a].readers++;
        if (b >= 0) pr[b].readers++;
    }
    auto side_of = [&](int buf) { return d.H >> u.bufs[buf].shift; };
    auto act_ptr = [&](const Prod& q, long* bs) {
        const int S = side_of(q.buf);
        *bs = (long)u.bufs[q.buf].C * S * S;
        return (const float*)(ws + L.act[q.buf] + (long)q.c0 * S * S);
    };
    auto grad_ptr = [&](const Prod& q, long* bs) {
        const int S = side_of(q.buf);
        *bs = (long)u.bufs[q.buf].C * S * S;
        return ws + L.grad[q.buf] + (long)q.c0 * S * S;
    };

    // gradient slices on chip, one per producer slice
    struct GS { int born, last, off, floats, parked, in_hbm; };
    GS gs[40];
    for (int k = 0; k < npr; ++k) gs[k] = GS{-1, -1, -1, 0, 0, 0};
    int in0_of[kFusedMaxOps], in1_of[kFusedMaxOps], out_of[kFusedMaxOps];
    int nf = 0, nl = 0;
    long woff = wpack_bwd_base(u);
    auto new_op = [&](int kind) -> FusedOp* {
        if (nf >= kFusedMaxOps) return nullptr;
        FusedOp& fo = P.ops[nf];
        memset(&fo, 0, sizeof(fo));
        fo.kind = kind;
        fo.out = fo.chunk = fo.in1 = -1;
        fo.next_w = -1;
        in0_of[nf] = in1_of[nf] = out_of[nf] = -1;
        return &fo;
    };
    // result slice k of the op being emitted: final (gated + written for wgrad) or parked for a later max-pool adjoint
    auto finish = [&](FusedOp* fo, int k, bool final_) {
        const Prod& q = pr[k];
        if (final_) {
            if (q.relu) fo->gmask = act_ptr(q, &fo->gmask_bs);
            if (q.kind == OP_CONV) fo->gout = grad_ptr(q, &fo->gout_bs);
        }
        if (gs[k].born < 0) gs[k].born = nf;
        out_of[nf] = k;
    };
    for (int i = u.nops - 1; i >= 0; --i) {
        const Op& op = u.ops[i];
        if (op.kind == OP_UP) continue;                 // emitted with the conv that reads it
        if (op.kind == OP_HEAD) {
            int a, b;
            if (!prods_in(op.in, &a, &b) || b >= 0 || pr[a].readers != 1) return -1;
            FusedOp* fo = new_op(F_HEADT);
            if (!fo) return -1;
            fo->S = d.H; fo->Cin0 = op.out.C; fo->Cout = op.in.C;
            fo->gsrc = ws + L.d_logits; fo->gsrc_bs = (long)d.n * d.HW;
            if (op.relu) { fo->gmask2 = ws + L.logits; fo->gmask2_bs = (long)d.n * d.HW; }
            fo->wfloats = (op.out.C * op.in.C + 3) & ~3;
            fo->wglob = woff;
            K.w[nl] = p->conv[op.layer].w; K.b[nl] = nullptr; K.Cout[nl] = op.out.C; K.Cin[nl] = op.in.C; K.taps[nl] = 1;
            K.mode[nl] = 2; K.off[nl] = woff; ++nl;
            woff += (long)align64((size_t)op.out.C * op.in.C);
            finish(fo, a, true);
            ++nf;
            continue;
        }
        if (op.kind == OP_POOL) {
            const int kp = prod_of(op.out.buf, op.out.c0, op.out.C);
            int a, b;
            if (kp < 0 || gs[kp].born < 0 || !prods_in(op.in, &a, &b) || b >= 0) return -1;
            FusedOp* fo = new_op(F_POOLT);
            if (!fo) return -1;
            fo->S = side_of(op.out.buf); fo->Cin0 = op.out.C; fo->Cout = op.out.C;
            fo->relu = pr[a].relu;
            fo->gmask = act_ptr(pr[a], &fo->gmask_bs);           // source values: arg-max and ReLU gate
            if (pr[a].kind == OP_CONV) fo->gout = grad_ptr(pr[a], &fo->gout_bs);
            in0_of[nf] = kp; gs[kp].last = nf;
            if (gs[a].born >= 0 && gs[a].in_hbm) { fo->acc_gout = 1; gs[a].born = nf; gs[a].in_hbm = 0; }
            else if (gs[a].born >= 0) { in1_of[nf] = a; gs[a].last = nf; }   // parked contribution of the other reader
            else gs[a].born = nf;
            out_of[nf] = a;
            gs[a].parked = 0;
            ++nf;
            continue;
        }
        // OP_CONV
        const int ko = prod_of(op.out.buf, op.out.c0, op.out.C);
        if (ko < 0) return -1;
        if (op.in.buf < 0) continue;                    // first layer: no gradient w.r.t. the frames
        if (gs[ko].born < 0) return -1;
        const bool via_up = i > 0 && u.ops[i - 1].kind == OP_UP && u.ops[i - 1].out.buf == op.in.buf &&
                            u.ops[i - 1].out.c0 == op.in.c0 && u.ops[i - 1].out.C == op.in.C;
        int a, b;
        if (!prods_in(op.in, &a, &b)) return -1;
        const int parts[2] = {a, b};
        for (int part = 0; part < 2; ++part) {
            const int k = parts[part];
            if (k < 0) continue;
            if (pr[k].C != 8 && pr[k].C != 16 && pr[k].C != 32) return -1;
            FusedOp* fo = new_op(F_CONV);
            if (!fo) return -1;
            fo->S = side_of(op.out.buf); fo->Cin0 = op.out.C; fo->Cout = pr[k].C;
            size_t wf = (size_t)op.out.C * 9 * pr[k].C;
            if (mma) {
                if (op.out.C % 8 || (fo->S != 32 && fo->S != 16 && fo->S != 8)) return -1;
                wf = (size_t)(op.out.C / 8) * 9 * (pad16(pr[k].C) / 16) * 128;
            }
            fo->wfloats = ((int)wf + 3) & ~3;
            fo->wglob = woff;
            K.w[nl] = p->conv[op.layer].w; K.b[nl] = nullptr; K.Cout[nl] = pr[k].C; K.Cin[nl] = op.out.C; K.taps[nl] = 9;
            K.mode[nl] = 1; K.ci0[nl] = pr[k].c0 - op.in.c0; K.cin_total[nl] = op.in.C; K.off[nl] = woff; ++nl;
            woff += (long)align64(wf);
            in0_of[nf] = ko; gs[ko].last = nf;
            if (via_up) {
                // k is the upsampled tensor: keep its gradient on chip only, then gather it back to the source
                if (pr[k].readers != 1 || b >= 0) return -1;
                finish(fo, k, false);
                ++nf;
                const Op& upo = u.ops[i - 1];
                int ua, ub;
                if (!prods_in(upo.in, &ua, &ub) || ub >= 0 || pr[ua].readers != 1) return -1;
                FusedOp* fu = new_op(F_UPT);
                if (!fu) return -1;
                fu->S = side_of(upo.in.buf); fu->Cin0 = upo.in.C; fu->Cout = upo.in.C;
                in0_of[nf] = k; gs[k].last = nf;
                finish(fu, ua, true);
                ++nf;
            } else if (pr[k].readers == 1) {
                finish(fo, k, true);
                ++nf;
            } else if (pr[k].readers == 2 && gs[k].born < 0) {
                finish(fo, k, false);                    // parked until the max-pool adjoint adds its share
                gs[k].parked = 1;
                if (park_global) {                       // ... in the gradient buffer the adjoint finalises in place
                    if (pr[k].kind != OP_CONV) return -1;
                    fo->gout = grad_ptr(pr[k], &fo->gout_bs);
                    gs[k].in_hbm = 1;
                    out_of[nf] = -1;
                }
                ++nf;
            } else {
                return -1;
            }
        }
    }
    if (nl > 24) return -1;
    K.nlayers = nl;
    K.frag = mma ? 1 : 0;
    P.nops = nf;
    P.kappa = fused_mma_kappa();
    for (int k = 0; k < npr; ++k)
        if (gs[k].parked) return -1;                     // a parked gradient nobody finalised
    if (info) {                                          // the op list only: unet_tc.cu plans and runs it on the tensor cores
        info->P = P; info->K = K; info->nslices = npr;
        for (int k = 0; k < nf; ++k) { info->in0_of[k] = in0_of[k]; info->in1_of[k] = in1_of[k]; info->out_of[k] = out_of[k]; }
        for (int k = 0; k < npr; ++k) { info->born[k] = gs[k].born; info->last[k] = gs[k].last; info->C[k] = pr[k].C; info->S[k] = side_of(pr[k].buf); }
        return 0;
    }
    // thread tiling of the transposed convs
    for (int k = 0; k < nf; ++k)
        if (!mma && P.ops[k].kind == F_CONV && !choose_tile(P.ops[k], false)) return -1;
    // weight pipeline
    int issue_at[kFusedMaxOps];
    {
        int prev = -1, widx = 0;
        P.first_w = -1;
        for (int k = 0; k < nf; ++k) {
            issue_at[k] = -1;
            if (!P.ops[k].wfloats) continue;
            P.ops[k].wbar = widx++ & 1;
            issue_at[k] = prev;
            if (prev < 0) P.first_w = k;
            else P.ops[prev].next_w = k;
            prev = k;
        }
    }
    // shared-memory plan
    Planner al;
    for (int k = 0; k < nf; ++k)
        if (P.ops[k].wfloats) al.add(P.ops[k].wfloats, issue_at[k], k, &P.ops[k].wsm);
    for (int k = 0; k < npr; ++k) {
        if (gs[k].born < 0 || gs[k].last < 0) continue;
        gs[k].floats = pr[k].C * geo_of(side_of(pr[k].buf)).plane;
        al.add(gs[k].floats, gs[k].born, gs[k].last, &gs[k].off);
    }
    const int peak = al.place();
    for (int k = 0; k < nf; ++k) {
        FusedOp& fo = P.ops[k];
        if (in0_of[k] >= 0) fo.in0 = gs[in0_of[k]].off;
        if (in1_of[k] >= 0) fo.in1 = gs[in1_of[k]].off;
        if (out_of[k] >= 0 && gs[out_of[k]].last >= 0) fo.out = gs[out_of[k]].off;
    }
    static const bool debug = getenv("PAIG_DEBUG") != nullptr;
    if (debug) {
        fprintf(stderr, "[paig] fused UNet backward-data plan: H=%d ops=%d smem=%zu B\n", d.H, nf, (size_t)peak * sizeof(float));
        for (int k = 0; k < nf; ++k) {
            const FusedOp& fo = P.ops[k];
            fprintf(stderr, "[paig]   op%-2d kind=%d S=%-2d Cin=%-2d Cout=%-2d co=%-2d in0=%d in1=%d out=%d mask=%d gout=%d w@%d(%d)\n", k,
                    fo.kind, fo.S, fo.Cin0, fo.Cout, fo.co_tile, fo.in0, fo.in1, fo.out, fo.gmask != nullptr,
                    fo.gout != nullptr, fo.wsm, fo.wfloats);
        }
    }
    if ((size_t)peak * sizeof(float) > kFusedSmemLimit) return -2;
    P.N = L.N; P.fps = 1; P.H = d.H;
    float* wpack = ws + L.wpack;
    P.wpack = wpack;
    launch(pack_weights_kernel, dim3(4, K.nlayers), dim3(256), 0, st, K, wpack);
    int rc = check_launch("pack_weights");
    if (rc) return rc;
    const int grid = L.N < sm_count() ? L.N : sm_count();
#ifndef PAIG_EMU
    P.timing = debug ? timing_buffer() : nullptr;
#endif
    if (mma) launch(unet_fused_bwd_kernel<true>, dim3(grid), dim3(kFusedThreads), (size_t)peak * sizeof(float), st, P);
    else launch(unet_fused_bwd_kernel<false>, dim3(grid), dim3(kFusedThreads), (size_t)peak * sizeof(float), st, P);
    (void)t;
    rc = check_launch("unet_fused_bwd");
#ifndef PAIG_EMU
    if (debug && !rc) print_timing("fused UNet backward-data", P, grid, L.N, st);
#endif
    return rc;
}

int unet_fused_backward(const paig_task* t, const paig_params* p, const Layout& L, float* ws, cudaStream_t st) {
    int rc = fused_backward_plan(t, p, L, ws, st, false);
    if (rc == -2) rc = fused_backward_plan(t, p, L, ws, st, true);
    return rc == -2 ? -1 : rc;
}

// The backward-data op list of the network (transposed convs, adjoints, who gates / parks / writes what) without planning
// shared memory or launching anything: the tensor-core backward (unet_tc.cu) runs the same list in its own layout.
int unet_backward_ops(const paig_task* t, const paig_params* p, const Layout& L, float* ws, bool park_global, BwdOps* info) {
    return fused_backward_plan(t, p, L, ws, nullptr, park_global, info);
}

}  // namespace paig
