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

#include <cstdint>
#include <cstdlib>
#include <cstring>

namespace paig {

constexpr int kFusedThreads = 512;
constexpr int kFusedMaxOps = 24;
constexpr size_t kFusedSmemLimit = 227 * 1024 - 4096;   // dynamic part; the op table and barriers are static

enum { F_CONV = 0, F_POOL = 1, F_HEAD = 2 };

struct FusedOp {
    int kind, S, Cin0, Cin1, Cout, relu;
    int up;                        // >0: segment 0 is the 2x upsample of a half-resolution buffer, built `up` channels at a time
    int co_tile;                   // output channels per thread: 4, 8 or 16
    int in0, in1, out, chunk;      // shared-memory offsets (floats); out < 0: result is not read on chip
    int wsm, wfloats, wbar;        // weights: shared offset, packed floats (incl. bias), mbarrier index
    int next_w;                    // index of the next op that has weights (prefetched while this op runs), or -1
    long wglob;                    // offset of this layer in the packed weight buffer
    float* gout; long gout_bs;     // global destination of the result (kept for backward)
    float* gup; long gup_bs;       // global destination of the upsampled input
};
struct FusedPlan {
    int nops, N, fps, H, first_w;
    long seq_stride;
    int x_off;
    const float* x;
    const float* wpack;
    long long* timing;             // debug (PAIG_DEBUG): per-CTA cycle stamps after every op of the CTA's first frames
    FusedOp ops[kFusedMaxOps];
};

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

// acc[c][p] += sum_{ci < nci} sum_taps w[ci][tap][c] * in[ci][y + ky - 1][4 qx + p + kx - 1]
// COUT is a template parameter so that every weight load is [pointer + immediate]; the three row pointers advance
// by one plane per input channel.
template <int CO, int COUT>
__device__ __forceinline__ void conv_accumulate(float (&acc)[CO][4], const float* __restrict__ planes, int nci,
                                                const Geo g, const float* __restrict__ w, int y, int qx) {
    const float* r0 = planes + y * g.P + 4 * qx;
    const float* r1 = r0 + g.P;
    const float* r2 = r1 + g.P;
    const int plane = g.plane;
#pragma unroll 2
    for (int ci = 0; ci < nci; ++ci) {
        float v[3][6];
        {
            const float4 a4 = *reinterpret_cast<const float4*>(r0);
            const float2 a2 = *reinterpret_cast<const float2*>(r0 + 4);
            const float4 b4 = *reinterpret_cast<const float4*>(r1);
            const float2 b2 = *reinterpret_cast<const float2*>(r1 + 4);
            const float4 c4 = *reinterpret_cast<const float4*>(r2);
            const float2 c2 = *reinterpret_cast<const float2*>(r2 + 4);
            v[0][0] = a4.x; v[0][1] = a4.y; v[0][2] = a4.z; v[0][3] = a4.w; v[0][4] = a2.x; v[0][5] = a2.y;
            v[1][0] = b4.x; v[1][1] = b4.y; v[1][2] = b4.z; v[1][3] = b4.w; v[1][4] = b2.x; v[1][5] = b2.y;
            v[2][0] = c4.x; v[2][1] = c4.y; v[2][2] = c4.z; v[2][3] = c4.w; v[2][4] = c2.x; v[2][5] = c2.y;
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
                    for (int c = 0; c < 4; ++c)
#pragma unroll
                        for (int p = 0; p < 4; ++p) acc[c4 + c][p] += wv[c] * v[ky][kx + p];
                }
            }
        r0 += plane; r1 += plane; r2 += plane;
        w += 9 * COUT;
    }
}

__device__ __forceinline__ void up_taps_f(int o, int Si, int& i0, int& i1, float& w0, float& w1) {
    const int k = o >> 1;
    if (o & 1) { i0 = k; i1 = min(k + 1, Si - 1); w0 = 0.75f; w1 = 0.25f; }
    else { i0 = max(k - 1, 0); i1 = k; w0 = 0.25f; w1 = 0.75f; }
}

template <int CO>
__device__ __forceinline__ void conv_epilogue(const float (&acc)[CO][4], const FusedOp& op, const Geo g, float* sm,
                                              const float* bias, int cg, int y, int qx, int f) {
    const int S = g.S, x0 = 4 * qx;
    const bool vec = (S & 3) == 0;
#pragma unroll
    for (int c = 0; c < CO; ++c) {
        const int co = cg * CO + c;
        const float b = bias[co];
        float o[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            o[p] = acc[c][p] + b;
            if (op.relu) o[p] = fmaxf(o[p], 0.f);
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

template <int CO, int COUT>
__device__ __forceinline__ void run_conv(const FusedOp& op, float* sm, int f, int tid, long long* tm) {
    const Geo g = geo_of(op.S);
    const int nq = op.S * g.nqx, nitems = nq * (COUT / CO);
    const float* w = sm + op.wsm;
    const float* bias = w + (op.Cin0 + op.Cin1) * 9 * COUT;
    if (op.out >= 0) zero_halo_planes(sm + op.out, COUT, g, tid, kFusedThreads);
    PAIG_STAMP(tm, 1);
    if (!op.up) {
        const bool pow2 = (nq & (nq - 1)) == 0 && (g.nqx & (g.nqx - 1)) == 0;
        const int lq = 31 - __clz(nq), lx = 31 - __clz(g.nqx);
        for (int item = tid; item < nitems; item += kFusedThreads) {
            int cg, q, y, qx;
            if (pow2) { cg = item >> lq; q = item & (nq - 1); y = q >> lx; qx = q & (g.nqx - 1); }
            else { cg = item / nq; q = item % nq; y = q / g.nqx; qx = q % g.nqx; }
            float acc[CO][4];
#pragma unroll
            for (int c = 0; c < CO; ++c)
#pragma unroll
                for (int p = 0; p < 4; ++p) acc[c][p] = 0.f;
            conv_accumulate<CO, COUT>(acc, sm + op.in0, op.Cin0, g, w + cg * CO, y, qx);
            if (op.Cin1) conv_accumulate<CO, COUT>(acc, sm + op.in1, op.Cin1, g, w + op.Cin0 * 9 * COUT + cg * CO, y, qx);
            PAIG_STAMP(tm, 2);
            conv_epilogue<CO>(acc, op, g, sm, bias, cg, y, qx, f);
        }
        PAIG_STAMP(tm, 3);
    } else {
        // the planner guarantees nitems <= kFusedThreads here: accumulators persist across the channel chunks
        const bool active = tid < nitems;
        const int cg = active ? tid / nq : 0, q = active ? tid % nq : 0, y = q / g.nqx, qx = q % g.nqx;
        float acc[CO][4];
#pragma unroll
        for (int c = 0; c < CO; ++c)
#pragma unroll
            for (int p = 0; p < 4; ++p) acc[c][p] = 0.f;
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
            if (active) conv_accumulate<CO, COUT>(acc, chunk, nch, g, w + c0 * 9 * COUT + cg * CO, y, qx);
            __syncthreads();
#ifndef PAIG_EMU
            if (tm) t_acc += clock64() - t_b;
#endif
        }
        if (tm) { tm[2] = t_up; tm[3] = t_acc; }
        if (active) conv_epilogue<CO>(acc, op, g, sm, bias, cg, y, qx, f);
    }
}

template <int COUT>
__device__ __forceinline__ void run_conv_co(const FusedOp& op, float* sm, int f, int tid, long long* tm) {
    if (COUT >= 16 && op.co_tile == 16) run_conv<(COUT >= 16 ? 16 : 4), COUT>(op, sm, f, tid, tm);
    else if (op.co_tile == 8) run_conv<8, COUT>(op, sm, f, tid, tm);
    else run_conv<4, COUT>(op, sm, f, tid, tm);
}

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
        if (tid == 0 && P.first_w >= 0) {
            const FusedOp& o = s_ops[P.first_w];
            bulk_issue(sm + o.wsm, P.wpack + o.wglob, (unsigned)o.wfloats * 4u, &bars[o.wbar]);
        }
        // ---- the input frame, zero halo ----
        float* X = sm + P.x_off;
        zero_halo_planes(X, 3, gx, tid, kFusedThreads);
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
            if (tid == 0 && op.next_w >= 0) {
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
                if (op.Cout == 8) run_conv_co<8>(op, sm, f, tid, tm);
                else if (op.Cout == 16) run_conv_co<16>(op, sm, f, tid, tm);
                else run_conv_co<32>(op, sm, f, tid, tm);
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

// ---- weight packing: [co][ci][tap] -> [ci][tap][co] | bias  (head: [co][ci] | bias kept as is) ----------------
struct PackPlan {
    int nlayers;
    const float* w[18];
    const float* b[18];
    int Cout[18], Cin[18], taps[18];
    long off[18];
};
__global__ void __launch_bounds__(256) pack_weights_kernel(const PackPlan P, float* __restrict__ dst) {
    const int l = blockIdx.y;
    if (l >= P.nlayers) return;
    const int Cout = P.Cout[l], Cin = P.Cin[l], taps = P.taps[l];
    const int nW = Cout * Cin * taps;
    float* d = dst + P.off[l];
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

// Static shared-memory planner: blocks with [born, dies] step intervals are placed largest first at the lowest
// offset that is free over their whole interval (interval-graph colouring heuristic; near-optimal here).
struct Planner {
    struct Blk { int size, born, dies, off; int* dst; };
    Blk b[96];
    int n = 0;
    void add(int size, int born, int dies, int* dst) {
        b[n++] = Blk{(size + 3) & ~3, born, dies, -1, dst};       // 16-byte granularity
    }
    int place() {                                                  // returns the peak (floats)
        int order[96];
        for (int i = 0; i < n; ++i) order[i] = i;
        for (int i = 1; i < n; ++i)                                // insertion sort, size descending
            for (int j = i; j > 0 && b[order[j]].size > b[order[j - 1]].size; --j) {
                const int t = order[j]; order[j] = order[j - 1]; order[j - 1] = t;
            }
        int peak = 0;
        for (int oi = 0; oi < n; ++oi) {
            Blk& x = b[order[oi]];
            int off = 0;
            for (;;) {
                bool moved = false;
                for (int pj = 0; pj < oi; ++pj) {
                    const Blk& y = b[order[pj]];
                    if (y.born > x.dies || x.born > y.dies) continue;             // never alive together
                    if (off < y.off + y.size && y.off < off + x.size) { off = y.off + y.size; moved = true; }
                }
                if (!moved) break;
            }
            x.off = off;
            *x.dst = off;
            if (off + x.size > peak) peak = off + x.size;
        }
        return peak;
    }
};

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
    size_t total = 0;
    for (int i = 0; i < u.nops; ++i) {
        const Op& op = u.ops[i];
        if (op.kind == OP_CONV) total += align64((size_t)op.in.C * 9 * op.out.C + op.out.C);
        else if (op.kind == OP_HEAD) total += align64((size_t)op.in.C * op.out.C + op.out.C);
    }
    return total;
}

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
                fo.wfloats = (cin * taps * fo.Cout + fo.Cout + 3) & ~3;
                fo.wglob = woff;
                K.w[nl] = p->conv[op.layer].w; K.b[nl] = p->conv[op.layer].b;
                K.Cout[nl] = fo.Cout; K.Cin[nl] = cin; K.taps[nl] = taps; K.off[nl] = woff;
                ++nl;
                woff += (long)align64((size_t)cin * taps * fo.Cout + fo.Cout);
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
        K.nlayers = nl;
        P.nops = nf;
        // ---- 2. thread tiling of each conv ----
        for (int k = 0; k < nf; ++k) {
            FusedOp& fo = P.ops[k];
            if (fo.kind != F_CONV) continue;
            const Geo g = geo_of(fo.S);
            const int nq = fo.S * g.nqx;
            int best = 0;
            long best_cost = 0;
            const int cands[3] = {8, 4, 16};
            for (int ci = 0; ci < 3; ++ci) {
                const int co = cands[ci];
                if (fo.Cout % co) continue;
                const int items = nq * (fo.Cout / co);
                const int passes = (items + kFusedThreads - 1) / kFusedThreads;
                if (fo.up && passes > 1) continue;
                // issue slots per input channel: passes x (36*co FMA + 6 + 9*co/4 loads) for the busiest thread
                const long cost = (long)passes * (36 * co + 6 + 9 * co / 4);
                if (!best || cost < best_cost) { best = co; best_cost = cost; }
            }
            if (!best) return -1;
            fo.co_tile = best;
        }
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
        sl[0].floats = 3 * geo_of(d.H).plane;
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
                fprintf(stderr, "[paig]   op%-2d kind=%d S=%-2d Cin=%d+%d Cout=%-2d up=%d co=%-2d in0=%d in1=%d out=%d chunk=%d w@%d(%d)\n",
                        k, fo.kind, fo.S, fo.Cin0, fo.Cin1, fo.Cout, fo.up, fo.co_tile, fo.in0, fo.in1, fo.out, fo.chunk,
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
        static long long* timing_buf = nullptr;
        if (debug && !timing_buf) cudaMalloc(&timing_buf, (size_t)sm_count() * 160 * sizeof(long long));
        P.timing = debug ? timing_buf : nullptr;
#endif
        launch(unet_fused_fwd_kernel, dim3(grid), dim3(kFusedThreads), (size_t)peak * sizeof(float), st, P);
        (void)t;
        rc = check_launch("unet_fused_fwd");
#ifndef PAIG_EMU
        if (debug && !rc && L.N >= 2 * grid) {          // cycles per op, second frame of every CTA, mean over CTAs
            cudaStreamSynchronize(st);
            static long long host[160 * 160];
            cudaMemcpy(host, timing_buf, (size_t)grid * 160 * sizeof(long long), cudaMemcpyDeviceToHost);
            fprintf(stderr, "[paig] fused UNet cycles per op (mean over %d CTAs):", grid);
            double total = 0;
            for (int k = 0; k < nf; ++k) {
                double s = 0;
                for (int b = 0; b < grid; ++b) s += (double)(host[b * 160 + 1 + k] - host[b * 160 + k]);
                fprintf(stderr, " op%d=%.0f", k, s / grid);
                total += s / grid;
            }
            fprintf(stderr, " total=%.0f\n", total);
            fprintf(stderr, "[paig]   thread 0 of CTA 0: op: wait | halo | compute | epilogue | sync\n");
            for (int k = 0; k < nf; ++k) {
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
        return rc;
    }
    return -1;
}

}  // namespace paig
