// fp32 CUDA-core convolution kernels for the encoder UNets (blocks.py:106-308): 3x3 "same" convolution
// forward, its data gradient (the same kernel run on the masked output gradient with transposed, flipped
// weights) and its weight/bias gradient; the 1x1 head; 2x2 max-pool and the exact-2x bilinear upsample
// (tvtrans.Resize, blocks.py:260,269) with their adjoints.
//
// Every tensor is a channel-sliced view of an NCHW buffer: element (n,c,y,x) = p[n*bs + (c*S + y)*S + x].
// Concats of the UNet are realised by producers writing into channel slices of one buffer.
//
// Why CUDA cores and not tcgen05 here: output-channel counts are 8..32 and the step must match the
// reference to 1e-4 (single-pass TF32 is ~1e-3); see DESIGN.md for the ncu evidence.
#include "common.cuh"
#include "internal.h"

#include <cstdint>
#include <cstdlib>

namespace paig {

constexpr int kConvThreads = 256;
constexpr int kCK = 8;                 // input channels staged per pass

// ---------------------------------------------------------------------------------------------------------
// conv3x3: each thread produces 4 horizontally adjacent pixels x CO_T output channels.
// Block = FPB frames x TH rows x QX quads (<= 256 threads); blockIdx = (cout group, row strip, frame group).
// ---------------------------------------------------------------------------------------------------------
// Row-wise staging of NCHW rows into a zero-haloed shared tile (image column x sits at tile column x+1).
// LOG_QX >= 0: S == 4 << LOG_QX and every pointer is 16-byte aligned, so a row is QX float4 loads and the
// (frame, channel, row) of a thread's item advances incrementally -- no per-element div/mod.  LOG_QX < 0: generic.
// rows: list index R -> (ff, ci, r) with r fastest; global row gy = y0 + r + yofs of channel cbase+ci, frame f0+ff,
// stored at tile column col0.
template <int LOG_QX>
__device__ __forceinline__ void stage_rows(float* __restrict__ sDst, int plane, int cslots, int PITCH, int RT, int nfr,
                                           int nc, const float* __restrict__ src, long src_bs,
                                           const float* __restrict__ msk, long msk_bs, int cbase, int f0, int N, int y0,
                                           int S, int tid, int nthr, int col0 = 1, int yofs = -1) {
    if (LOG_QX >= 0) {
        constexpr int QXc = LOG_QX >= 0 ? (1 << (LOG_QX >= 0 ? LOG_QX : 0)) : 1;
        const int q = tid & (QXc - 1);
        const int step = nthr >> (LOG_QX >= 0 ? LOG_QX : 0);
        int R = tid >> (LOG_QX >= 0 ? LOG_QX : 0);
        int r = R % RT, t = R / RT;
        int ci = t % nc, ff = t / nc;
        while (ff < nfr) {
            const int gy = y0 + r + yofs, gf = f0 + ff;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gf < N && (unsigned)gy < (unsigned)S) {
                const long off = ((long)(cbase + ci) * S + gy) * S + 4 * q;
                v = *reinterpret_cast<const float4*>(src + (long)gf * src_bs + off);
                if (msk) {
                    const float4 m = *reinterpret_cast<const float4*>(msk + (long)gf * msk_bs + off);
                    v.x = m.x > 0.f ? v.x : 0.f; v.y = m.y > 0.f ? v.y : 0.f;
                    v.z = m.z > 0.f ? v.z : 0.f; v.w = m.w > 0.f ? v.w : 0.f;
                }
            }
            float* d = sDst + (ff * cslots + ci) * plane + r * PITCH + col0 + 4 * q;
            d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
            r += step;
            while (r >= RT) {
                r -= RT;
                if (++ci == nc) { ci = 0; ++ff; }
            }
        }
    } else {
        // generic (S = 18, 9, 36 ...): a warp stages one row at a time -- the (frame, channel, row) decomposition is
        // done once per row instead of once per element, and the lanes read the row coalesced
        const int lane = tid & 31;
        const int rows_total = nfr * nc * RT;
        const int nw = nthr >> 5;
        constexpr int U = 8;                                  // rows in flight per warp: the loads are latency-bound
        for (int R0 = tid >> 5; R0 < rows_total; R0 += U * nw) {
            float v[U][2];
            float* d[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int R = R0 + u * nw;
                v[u][0] = v[u][1] = 0.f;
                d[u] = nullptr;
                if (R < rows_total) {
                    const int r = R % RT, t = R / RT, ci = t % nc, ff = t / nc;
                    const int gy = y0 + r + yofs, gf = f0 + ff;
                    d[u] = sDst + (ff * cslots + ci) * plane + r * PITCH + col0;
                    if (gf < N && (unsigned)gy < (unsigned)S) {
                        const long off = ((long)(cbase + ci) * S + gy) * S;
                        const float* sp = src + (long)gf * src_bs + off;
                        const float* mp = msk ? msk + (long)gf * msk_bs + off : nullptr;
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int col = lane + 32 * h;
                            if (col < S) {
                                float x = sp[col];
                                if (mp && !(mp[col] > 0.f)) x = 0.f;
                                v[u][h] = x;
                            }
                        }
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (d[u]) {
                    if (lane < S) d[u][lane] = v[u][0];
                    if (lane + 32 < S) d[u][lane + 32] = v[u][1];
                }
        }
    }
}

// zero the halo columns (tile column 0 and columns S+1 .. PITCH-1) of every staged row once per CTA
__device__ __forceinline__ void zero_halo_cols(float* sDst, int rows_total, int PITCH, int S, int tid, int nthr) {
    const int hc = PITCH - S;                   // 1 left + (PITCH - S - 1) right
    for (int e = tid; e < rows_total * hc; e += nthr) {
        const int row = e / hc, k = e % hc;
        sDst[row * PITCH + (k == 0 ? 0 : S + k)] = 0.f;
    }
}

// Asynchronous variant of the generic row staging (no ReLU mask): every element is one 4-byte cp.async (LDGSTS) into
// the same zero-haloed tile, rows outside the image are zero-filled through the src-size operand, nothing waits here.
// Used by the weight-gradient kernel to stage strip k+1 while strip k is multiplied (18- and 9-px levels of 3bp, whose
// 72- and 36-byte row pitch the TMA unit cannot address).
__device__ __forceinline__ void stage_rows_async(float* sDst, int plane, int PITCH, int RT, int nc, const float* src,
                                                 long src_bs, int f, int y0, int S, int tid, int nthr, int col0, int yofs) {
#ifndef PAIG_EMU
    // one row per thread: the (channel, row) decomposition and the address set-up cost ~40 instructions, the S copies
    // two each -- with a warp per row (18 active lanes, one copy each) staging issued as many instructions as the
    // multiply (ncu: FFMA 41 % of the mix).  Lanes read different rows at the same column; the 32-byte sectors they
    // touch are reused by the next seven columns out of L1.
    const int rows_total = nc * RT;
    for (int R = tid; R < rows_total; R += nthr) {
        const int r = R % RT, ci = R / RT;
        const int gy = y0 + r + yofs;
        const bool ok = (unsigned)gy < (unsigned)S;
        const float* sp = src + (long)f * src_bs + ((long)ci * S + (ok ? gy : 0)) * S;
        const unsigned d0 = (unsigned)__cvta_generic_to_shared(sDst + ci * plane + r * PITCH + col0);
        const unsigned nbytes = ok ? 4u : 0u;
#pragma unroll 9
        for (int col = 0; col < S; ++col)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d0 + 4u * col), "l"(sp + col), "r"(nbytes)
                         : "memory");
    }
#endif
}

template <int CO_T, int LOG_QX>
__global__ void __launch_bounds__(kConvThreads) conv3x3_kernel(ConvArgs a) {
    PAIG_DYN_SMEM(float, smem);
    const int S = a.S, QX = a.QX, TH = a.TH, FPB = a.FPB;
    const int PITCH = 4 * QX + 4;
    const int plane = (TH + 2) * PITCH;
    float* sIn = smem;                                    // [FPB][kCK][TH+2][PITCH]
    float* sW = smem + FPB * kCK * plane;                 // [kCK][9][CO_T]
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int qx = tid % QX, ty = (tid / QX) % TH, fb = tid / (QX * TH);
    // the output-channel group is the fastest grid dimension: the CTAs that re-read one input tile for different
    // output channels run together and share it in L2 (with the batch fastest the second group found 1 % L2 hits)
    const int f0 = blockIdx.z * FPB, y0 = blockIdx.y * TH, co0 = blockIdx.x * CO_T;
    const int y = y0 + ty, f = f0 + fb;
    const bool active = fb < FPB && f < a.N && y < S;

    float acc[CO_T][4];
#pragma unroll
    for (int c = 0; c < CO_T; ++c)
#pragma unroll
        for (int p = 0; p < 4; ++p) acc[c][p] = 0.f;

    zero_halo_cols(sIn, FPB * kCK * (TH + 2), PITCH, S, tid, nthr);     // plane == (TH+2)*PITCH: rows are contiguous

    for (int c0 = 0; c0 < a.Cin; c0 += kCK) {
        const int nc = min(kCK, a.Cin - c0);
        // ---- stage the input tile (zero halo), optionally masked by the producer's ReLU ----
        stage_rows<LOG_QX>(sIn, plane, kCK, PITCH, TH + 2, FPB, nc, a.in, a.in_bs, a.mask, a.mask_bs, c0, f0, a.N, y0, S,
                           tid, nthr);
        // ---- stage the weights of this (cin chunk, cout group): sW[ci][tap][co] ----
        for (int e = tid; e < nc * 9 * CO_T; e += nthr) {
            const int co = e % CO_T, tap = (e / CO_T) % 9, ci = e / (9 * CO_T);
            float v = 0.f;
            if (co0 + co < a.Cout) {
                v = a.transposed ? a.w[((long)(c0 + ci) * a.Cout + (co0 + co)) * 9 + (8 - tap)]
                                 : a.w[((long)(co0 + co) * a.Cin + (c0 + ci)) * 9 + tap];
            }
            sW[(ci * 9 + tap) * CO_T + co] = v;
        }
        __syncthreads();
        if (active) {
            for (int ci = 0; ci < nc; ++ci) {
                const float* row = sIn + (fb * kCK + ci) * plane + ty * PITCH + 4 * qx;
                float v[3][6];
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const float4 p4 = *reinterpret_cast<const float4*>(row + r * PITCH);
                    const float2 p2 = *reinterpret_cast<const float2*>(row + r * PITCH + 4);
                    v[r][0] = p4.x; v[r][1] = p4.y; v[r][2] = p4.z; v[r][3] = p4.w; v[r][4] = p2.x; v[r][5] = p2.y;
                }
                const float* wp = sW + ci * 9 * CO_T;
#pragma unroll
                for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        const float* w = wp + (ky * 3 + kx) * CO_T;
#pragma unroll
                        for (int c4 = 0; c4 < CO_T; c4 += 4) {
                            const float4 w4 = *reinterpret_cast<const float4*>(w + c4);
                            const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                            for (int c = 0; c < 4; ++c)
#pragma unroll
                                for (int p = 0; p < 4; ++p) acc[c4 + c][p] += wv[c] * v[ky][kx + p];
                        }
                    }
            }
        }
        __syncthreads();
    }
    if (!active) return;
    const int x0 = 4 * qx;
#pragma unroll
    for (int c = 0; c < CO_T; ++c) {
        const int co = co0 + c;
        if (co >= a.Cout) break;
        const float bias = a.b ? a.b[co] : 0.f;
        float o[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            o[p] = acc[c][p] + bias;
            if (a.relu) o[p] = fmaxf(o[p], 0.f);
        }
        float* dst = a.out + (long)f * a.out_bs + ((long)co * S + y) * S + x0;
        if (x0 + 3 < S && (S & 3) == 0 && (a.out_bs & 3) == 0) {
            *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
        } else {
#pragma unroll
            for (int p = 0; p < 4; ++p)
                if (x0 + p < S) dst[p] = o[p];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// The same convolution with the staging taken off the critical path (ncu on the kernel above, 64-px UNet: issue
// slots 65 % busy, long_scoreboard the top real stall -- every chunk of 8 input channels is stage, barrier,
// multiply, barrier).  Here chunk k+1 (input rows as 16-byte cp.async, weights as 4-byte cp.async) is in flight
// into the other buffer while chunk k is multiplied.  The image sits at tile column x+4 so the copies are
// 16-byte aligned; a thread reads its 6 pixels as 4 B + 16 B + 4 B (the TMA weight-gradient kernel's pattern).
// Needs S = 4 * 2^k, 16-byte aligned rows and no ReLU mask on the way in (the per-layer backward gates g in place).
// ---------------------------------------------------------------------------------------------------------
template <int CO_T, int LOG_QX>
__global__ void __launch_bounds__(kConvThreads) conv3x3_async_kernel(ConvArgs a) {
#ifndef PAIG_EMU
    PAIG_DYN_SMEM(float, smem);
    constexpr int QX = 1 << LOG_QX;
    const int S = a.S, TH = a.TH, FPB = a.FPB, CK = a.CK;
    constexpr int PITCH = 4 * QX + 8;
    const int RT = TH + 2;
    const int plane = RT * PITCH;
    const int in_fl = FPB * CK * plane, stage_fl = in_fl + CK * 9 * CO_T;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int qx = tid & (QX - 1), ty = (tid >> LOG_QX) % TH, fb = tid / (QX * TH);
    const int f0 = blockIdx.z * FPB, y0 = blockIdx.y * TH, co0 = blockIdx.x * CO_T;
    const int y = y0 + ty, f = f0 + fb;
    const bool active = fb < FPB && f < a.N && y < S;

    float acc[CO_T][4];
#pragma unroll
    for (int c = 0; c < CO_T; ++c)
#pragma unroll
        for (int p = 0; p < 4; ++p) acc[c][p] = 0.f;
    for (int e = tid; e < 2 * stage_fl; e += nthr) smem[e] = 0.f;      // halo columns are never written by the copies
    __syncthreads();

    auto issue = [&](int c0, int buf) {
        const int nc = min(CK, a.Cin - c0);
        float* sIn = smem + buf * stage_fl;
        float* sW = sIn + in_fl;
        // input rows: item = (frame, channel, row, quad); (channel, frame) advance without a division per item
        {
            const int step = nthr >> LOG_QX;
            int R = tid >> LOG_QX;
            int r = R % RT, t = R / RT;
            int ci = t % nc, ff = t / nc;
            while (ff < FPB) {
                const int gy = y0 + r - 1, gf = f0 + ff;
                const bool ok = gf < a.N && (unsigned)gy < (unsigned)S;
                const float* src = ok ? a.in + (long)gf * a.in_bs + ((long)(c0 + ci) * S + gy) * S + 4 * qx : a.in;
                const unsigned dst = (unsigned)__cvta_generic_to_shared(sIn + (ff * CK + ci) * plane + r * PITCH + 4 + 4 * qx);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(ok ? 16u : 0u) : "memory");
                r += step;
                while (r >= RT) {
                    r -= RT;
                    if (++ci == nc) { ci = 0; ++ff; }
                }
            }
        }
        // weights of this (cin chunk, cout group): sW[ci][tap][co]
        for (int e = tid; e < nc * 9 * CO_T; e += nthr) {
            const int co = e % CO_T, tap = (e / CO_T) % 9, ci = e / (9 * CO_T);
            const bool ok = co0 + co < a.Cout;
            const float* src = !ok ? a.w
                               : (a.transposed ? a.w + ((long)(c0 + ci) * a.Cout + (co0 + co)) * 9 + (8 - tap)
                                               : a.w + ((long)(co0 + co) * a.Cin + (c0 + ci)) * 9 + tap);
            const unsigned dst = (unsigned)__cvta_generic_to_shared(sW + (ci * 9 + tap) * CO_T + co);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(ok ? 4u : 0u) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    issue(0, 0);
    int buf = 0;
    for (int c0 = 0; c0 < a.Cin; c0 += CK, buf ^= 1) {
        const int nc = min(CK, a.Cin - c0);
        if (c0 + CK < a.Cin) {
            issue(c0 + CK, buf ^ 1);                   // the barrier that ended the previous trip freed that buffer
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        const float* sIn = smem + buf * stage_fl;
        const float* sW = sIn + in_fl;
        if (active) {
            for (int ci = 0; ci < nc; ++ci) {
                const float* row = sIn + (fb * CK + ci) * plane + ty * PITCH + 4 * qx;
                float v[3][6];
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const float4 p4 = *reinterpret_cast<const float4*>(row + r * PITCH + 4);
                    v[r][0] = row[r * PITCH + 3]; v[r][1] = p4.x; v[r][2] = p4.y; v[r][3] = p4.z; v[r][4] = p4.w;
                    v[r][5] = row[r * PITCH + 8];
                }
                const float* wp = sW + ci * 9 * CO_T;
#pragma unroll
                for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        const float* w = wp + (ky * 3 + kx) * CO_T;
#pragma unroll
                        for (int c4 = 0; c4 < CO_T; c4 += 4) {
                            const float4 w4 = *reinterpret_cast<const float4*>(w + c4);
                            const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                            for (int c = 0; c < 4; ++c)
#pragma unroll
                                for (int p = 0; p < 4; ++p) acc[c4 + c][p] += wv[c] * v[ky][kx + p];
                        }
                    }
            }
        }
        __syncthreads();
    }
    if (!active) return;
    const int x0 = 4 * qx;
#pragma unroll
    for (int c = 0; c < CO_T; ++c) {
        const int co = co0 + c;
        if (co >= a.Cout) break;
        const float bias = a.b ? a.b[co] : 0.f;
        float o[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            o[p] = acc[c][p] + bias;
            if (a.relu) o[p] = fmaxf(o[p], 0.f);
        }
        float* dst = a.out + (long)f * a.out_bs + ((long)co * S + y) * S + x0;
        if ((a.out_bs & 3) == 0 && ((uintptr_t)a.out & 15) == 0) {
            *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
        } else {
#pragma unroll
            for (int p = 0; p < 4; ++p) dst[p] = o[p];
        }
    }
#endif
}

static void conv_geometry(int S, int N, int* QX, int* TH, int* FPB, int* strips) {
    *QX = (S + 3) / 4;
    const int per_frame = *QX * S;
    if (per_frame <= kConvThreads) {
        *TH = S;
        *strips = 1;
        *FPB = kConvThreads / per_frame;
        if (*FPB > N) *FPB = N;
        if (*FPB > 16) *FPB = 16;
    } else {
        *strips = cdiv(per_frame, kConvThreads);
        *TH = cdiv(S, *strips);
        *FPB = 1;
    }
}

int conv3x3(const ConvArgs& in_args, cudaStream_t st) {
    ConvArgs a = in_args;
    if (a.N <= 0) return 0;
    if (a.tc_scratch) {               // dense layers (channels in multiples of 8 / 16): 3xTF32 implicit GEMM on tcgen05
        const int rc = conv3x3_tc(a, a.tc_scratch, st);
        if (rc >= 0) return rc;
    }
    int strips;
    conv_geometry(a.S, a.N, &a.QX, &a.TH, &a.FPB, &strips);
    // the frame groups ride on grid.z (<= 65535): very large batches (the 8192-sequence evaluation sweep is 81 920
    // frames) go out as several launches over consecutive frame ranges
    {
        const long max_frames = 65535L * a.FPB;
        if (a.N > max_frames) {
            for (long f0 = 0; f0 < in_args.N; f0 += max_frames) {
                ConvArgs part = in_args;
                part.N = (int)(in_args.N - f0 < max_frames ? in_args.N - f0 : max_frames);
                part.in = in_args.in + f0 * in_args.in_bs;
                part.out = in_args.out + f0 * in_args.out_bs;
                if (in_args.mask) part.mask = in_args.mask + f0 * in_args.mask_bs;
                // (conv_geometry may pick a different FPB for the tail; every range is a self-contained launch)
                const int rc = conv3x3(part, st);
                if (rc) return rc;
            }
            return 0;
        }
    }
    const int PITCH = 4 * a.QX + 4;
    const int threads = ((a.FPB * a.TH * a.QX + 31) / 32) * 32;
    const bool wide = (a.Cout % 16) == 0;
    const int CO_T = wide ? 16 : 8;
    const size_t smem = ((size_t)a.FPB * kCK * (a.TH + 2) * PITCH + (size_t)kCK * 9 * CO_T) * sizeof(float);
    dim3 grid(cdiv(a.Cout, CO_T), strips, cdiv(a.N, a.FPB));
    // vector staging needs S = 4 * 2^k and 16-byte aligned rows
    int lq = -1;
    const bool aligned = ((uintptr_t)a.in % 16 == 0) && (a.in_bs % 4 == 0) &&
                         (!a.mask || (((uintptr_t)a.mask % 16 == 0) && (a.mask_bs % 4 == 0)));
    if (aligned && a.S == 4 * a.QX) {
        for (int k = 0; k <= 4; ++k)
            if (a.QX == (1 << k)) lq = k;
    }
    dim3 blk(threads);
#ifndef PAIG_EMU
    // double-buffered cp.async variant: largest channel chunk (8, then 4) whose two stages leave room for a second CTA
    static const bool async_off = getenv("PAIG_CONV_SYNC") != nullptr;
    if (!async_off && lq >= 2 && !a.mask && a.Cin >= 8) {      // (8-px layers: 16 frames per CTA, measured 3 % slower)
        const int PITCHa = 4 * a.QX + 8;
        for (int ck = 8; ck >= 4; ck >>= 1) {
            const size_t stage = (size_t)a.FPB * ck * (a.TH + 2) * PITCHa + (size_t)ck * 9 * CO_T;
            const size_t smem2 = 2 * stage * sizeof(float);
            if (smem2 > 110 * 1024) continue;
            a.CK = ck;
#define PAIG_CONVA_CASE(LQ)                                                                       \
    case LQ:                                                                                      \
        if (wide) launch(conv3x3_async_kernel<16, LQ>, grid, blk, smem2, st, a);                  \
        else launch(conv3x3_async_kernel<8, LQ>, grid, blk, smem2, st, a);                        \
        break;
            switch (lq) {
                PAIG_CONVA_CASE(1)
                PAIG_CONVA_CASE(2)
                PAIG_CONVA_CASE(3)
                PAIG_CONVA_CASE(4)
            }
#undef PAIG_CONVA_CASE
            return check_launch(layer_name("conv3x3", a.Cin, a.Cout, a.S));
        }
    }
#endif
#define PAIG_CONV_CASE(LQ)                                                                  \
    case LQ:                                                                                \
        if (wide) launch(conv3x3_kernel<16, LQ>, grid, blk, smem, st, a);                   \
        else launch(conv3x3_kernel<8, LQ>, grid, blk, smem, st, a);                         \
        break;
    switch (lq) {
        PAIG_CONV_CASE(1)
        PAIG_CONV_CASE(2)
        PAIG_CONV_CASE(3)
        PAIG_CONV_CASE(4)
        default:
            if (wide) launch(conv3x3_kernel<16, -1>, grid, blk, smem, st, a);
            else launch(conv3x3_kernel<8, -1>, grid, blk, smem, st, a);
    }
#undef PAIG_CONV_CASE
    return check_launch(layer_name("conv3x3", a.Cin, a.Cout, a.S));
}

// ---------------------------------------------------------------------------------------------------------
// conv3x3 weight + bias gradient.  A thread owns the 3x3 taps of (4 output channels x 1 input channel) and a
// partition of the pixels; accumulators live in registers across every (frame, row strip) item the CTA
// walks; pixel partitions are folded through shared memory at the end and each CTA leaves one partial.
//   dW[co,ci,ky,kx] = sum_{n,y,x} g[n,co,y,x] * in[n,ci,y+ky-1,x+kx-1],   g = dOut * (act > 0)
// ---------------------------------------------------------------------------------------------------------
constexpr int kWgTH = 8;                  // rows per staged strip (default; small images are staged whole, WgradArgs::TH)

// COB: output channels per owner thread (4; 8 for the cp.async path, halving the input-pixel loads per FMA)
template <int LOG_QX, int COB>
__global__ void __launch_bounds__(kConvThreads, COB == 8 ? 2 : 1) conv3x3_wgrad_kernel(WgradArgs a) {
    PAIG_DYN_SMEM(float, smem);
    const int S = a.S, QX = (S + 3) / 4;
    const int PITCH = 4 * QX + 4;
    const int in_plane = a.in_plane, g_plane = a.g_plane;            // padded to == 4 (mod 32) floats
    const int TH = a.TH;                                              // rows per staged strip
    float* sIn = smem;                                                // [Cin][TH+2][PITCH]
    float* sG = smem + (size_t)a.Cin * in_plane;                      // [Cout][TH][PITCH-4 .. ] pitch 4*QX
    const int GP = 4 * QX;
    const int tid = threadIdx.x;
    const int cob_n = (a.Cout + COB - 1) / COB;
    constexpr int kRed = COB * 10;                                    // floats a thread leaves for the final fold
    const int G = cob_n * a.Cin;                                      // owner groups
    const int gsets = gridDim.y;
    const int G_per = (G + gsets - 1) / gsets;                        // groups handled by this CTA (<= 256)
    const int P = max(1, kConvThreads / G_per);                       // pixel partitions
    const int grp_local = tid % G_per, part = tid / G_per;
    const int grp = blockIdx.y * G_per + grp_local;
    const bool owner = part < P && grp < G;
    const int ci = owner ? grp % a.Cin : 0, cob = owner ? grp / a.Cin : 0;

    float acc[COB][9];
    float bacc[COB];
#pragma unroll
    for (int c = 0; c < COB; ++c) {
        bacc[c] = 0.f;
#pragma unroll
        for (int t = 0; t < 9; ++t) acc[c][t] = 0.f;
    }
    // halo columns / plane padding are never written by the row staging: clear the tile(s) once
    const int stage_floats = a.Cin * in_plane + a.Cout * g_plane;
    const bool async = LOG_QX < 0 && a.async2;                         // two stage buffers, cp.async staging (host)
    for (int e = tid; e < (async ? 2 : 1) * stage_floats; e += kConvThreads) smem[e] = 0.f;
    __syncthreads();
    const int strips = (S + TH - 1) / TH;
    const int items = a.N * strips;
    auto issue = [&](int item, int buf) {                              // cp.async staging of one strip (no masks)
        const int f = item / strips, y0 = (item % strips) * TH;
        float* b = smem + buf * stage_floats;
        stage_rows_async(b, in_plane, PITCH, TH + 2, a.Cin, a.in, a.in_bs, f, y0, S, tid, kConvThreads, 1, -1);
        stage_rows_async(b + (size_t)a.Cin * in_plane, g_plane, GP, TH, a.Cout, a.g, a.g_bs, f, y0, S, tid, kConvThreads, 0, 0);
#ifndef PAIG_EMU
        asm volatile("cp.async.commit_group;" ::: "memory");
#endif
    };
    if (async && (int)blockIdx.x < items) issue(blockIdx.x, 0);
    int kbuf = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x, kbuf ^= 1) {
        const int f = item / strips, y0 = (item % strips) * TH;
        const int rows = min(TH, S - y0);
        if (async) {
            // strip k+1 goes out before strip k is waited for; the barrier at the end of the previous trip freed its buffer
            sIn = smem + kbuf * stage_floats;
            sG = sIn + (size_t)a.Cin * in_plane;
#ifndef PAIG_EMU
            if (item + (int)gridDim.x < items) {
                issue(item + gridDim.x, kbuf ^ 1);
                asm volatile("cp.async.wait_group 1;" ::: "memory");
            } else {
                asm volatile("cp.async.wait_group 0;" ::: "memory");
            }
#endif
        } else {
            // ---- stage input strip (zero halo) and the masked output-gradient strip ----
            stage_rows<LOG_QX>(sIn, in_plane, 0, PITCH, TH + 2, 1, a.Cin, a.in, a.in_bs, a.in_mask, a.in_mask_bs, 0, f, a.N,
                               y0, S, tid, kConvThreads);
            stage_rows<LOG_QX>(sG, g_plane, 0, GP, TH, 1, a.Cout, a.g, a.g_bs, a.act, a.act_bs, 0, f, a.N, y0, S, tid,
                               kConvThreads, 0, 0);
        }
        __syncthreads();
        if (owner) {
            const int nq = rows * QX;
            // quads q = part, part + P, ...: (row, quad-in-row) advance without a division per step
            const int dr = P / QX, dq = P % QX;
            int r = part / QX, qxx = part % QX;
            for (int q = part; q < nq; q += P) {
                const int x0 = qxx * 4;
                float v[3][6];
                const float* ip = sIn + ci * in_plane + r * PITCH + x0;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const float4 p4 = *reinterpret_cast<const float4*>(ip + k * PITCH);
                    const float2 p2 = *reinterpret_cast<const float2*>(ip + k * PITCH + 4);
                    v[k][0] = p4.x; v[k][1] = p4.y; v[k][2] = p4.z; v[k][3] = p4.w; v[k][4] = p2.x; v[k][5] = p2.y;
                }
#pragma unroll
                for (int c = 0; c < COB; ++c) {
                    const int co = cob * COB + c;
                    if (COB == 8 || co < a.Cout) {             // (COB == 8 is only launched when it divides Cout)
                        const float4 g4 = *reinterpret_cast<const float4*>(sG + co * g_plane + r * GP + x0);
                        const float gv[4] = {g4.x, g4.y, g4.z, g4.w};
                        bacc[c] += (gv[0] + gv[1]) + (gv[2] + gv[3]);
#pragma unroll
                        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                            for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                                for (int p = 0; p < 4; ++p) acc[c][ky * 3 + kx] += gv[p] * v[ky][kx + p];
                    }
                }
                r += dr; qxx += dq;
                if (qxx >= QX) { qxx -= QX; ++r; }
            }
        }
        __syncthreads();
    }
    // ---- fold pixel partitions (fixed order) and write this CTA's partial ----
    float* sRed = smem;                                       // [P][G_per][kRed]  (reuses the tile area)
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
    for (int e = tid; e < G_per * kRed; e += kConvThreads) {
        const int k = e % kRed, gl = e / kRed;
        const int gg = blockIdx.y * G_per + gl;
        if (gg >= G) continue;
        float s = 0.f;
        for (int p = 0; p < P; ++p) s += sRed[((size_t)p * G_per + gl) * kRed + k];
        const int gci = gg % a.Cin, gcob = gg / a.Cin;
        if (k < COB * 9) {
            const int co = gcob * COB + k / 9;
            if (co < a.Cout) out[((size_t)co * a.Cin + gci) * 9 + (k % 9)] = s;
        } else if (gci == 0) {
            const int co = gcob * COB + (k - COB * 9);
            if (co < a.Cout) out[nW + co] = s;
        }
    }
}

// out[i] = sum_p partials[p*stride + i]; two destinations (weights, bias).  A block reduces 32 outputs: its 8 warps
// each sum a strided subset of the partials (coalesced 128-byte rows), then fold through shared memory in a fixed
// order, so the result is deterministic and the long dimension is read in parallel.
__global__ void __launch_bounds__(256) reduce_partials_kernel(const float* __restrict__ partials, int nparts, int stride,
                                                              int n0, float* __restrict__ out0, int n1,
                                                              float* __restrict__ out1) {
    __shared__ float part[8][33];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + lane;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    if (i < n0 + n1) {
        int p = w;
        for (; p + 24 < nparts; p += 32) {          // four independent loads in flight
            s0 += partials[(size_t)p * stride + i];
            s1 += partials[(size_t)(p + 8) * stride + i];
            s2 += partials[(size_t)(p + 16) * stride + i];
            s3 += partials[(size_t)(p + 24) * stride + i];
        }
        for (; p < nparts; p += 8) s0 += partials[(size_t)p * stride + i];
    }
    part[w][lane] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (w == 0 && i < n0 + n1) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) s += part[k][lane];
        if (i < n0) out0[i] = s;
        else if (out1) out1[i - n0] = s;
    }
}

// the same fold for a batch of independent reductions: blockIdx.y selects the entry
__global__ void __launch_bounds__(256) reduce_partials_batch_kernel(const ReduceBatch b) {
    __shared__ float part[8][33];
    const int k = blockIdx.y;
    const int n0 = b.n0[k], n1 = b.n1[k], nparts = b.nparts[k], stride = b.stride[k];
    const float* __restrict__ partials = b.partials[k];
    if ((int)blockIdx.x * 32 >= n0 + n1) return;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + lane;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    if (i < n0 + n1) {
        int p = w;
        for (; p + 24 < nparts; p += 32) {
            s0 += partials[(size_t)p * stride + i];
            s1 += partials[(size_t)(p + 8) * stride + i];
            s2 += partials[(size_t)(p + 16) * stride + i];
            s3 += partials[(size_t)(p + 24) * stride + i];
        }
        for (; p < nparts; p += 8) s0 += partials[(size_t)p * stride + i];
    }
    part[w][lane] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (w == 0 && i < n0 + n1) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) s += part[j][lane];
        if (i < n0) b.out0[k][i] = s;
        else if (b.out1[k]) b.out1[k][i - n0] = s;
    }
}

int reduce_partials_batch(const ReduceBatch& b, cudaStream_t st) {
    if (b.n <= 0) return 0;
    int widest = 0;
    for (int k = 0; k < b.n; ++k) widest = b.n0[k] + b.n1[k] > widest ? b.n0[k] + b.n1[k] : widest;
    launch(reduce_partials_batch_kernel, dim3(cdiv(widest, 32), b.n), dim3(256), 0, st, b);
    return check_launch("reduce_partials");
}

int reduce_partials(const float* partials, int nparts, int stride, int n0, float* out0, int n1, float* out1,
                    cudaStream_t st) {
    launch(reduce_partials_kernel, dim3(cdiv(n0 + n1, 32)), dim3(256), 0, st, partials, nparts, stride, n0, out0, n1,
           out1);
    return check_launch("reduce_partials");
}

static int pad_plane(int n) {          // smallest m >= n with m % 32 == 4: conflict-free 16-byte loads across channels
    int m = n + ((4 - n) % 32 + 32) % 32;
    return m;
}

size_t wgrad_partials_floats(int Cin, int Cout) { return (size_t)kWgradMaxCtas * ((size_t)Cout * Cin * 9 + Cout); }

int conv3x3_wgrad(const WgradArgs& in_args, float* dW, float* db, cudaStream_t st) {
    WgradArgs a = in_args;
    if (a.N <= 0) return 0;
    {
        const int rc = conv3x3_wgrad_tc(in_args, dW, db, st);      // tcgen05 3xTF32 (32 .. 128 channels both sides)
        if (rc >= 0) return rc;
    }
    {
        const int rc = conv3x3_wgrad_tma(in_args, dW, db, st);     // TMA-staged, double-buffered strips when eligible
        if (rc >= 0) return rc;
    }
    const int QX = (a.S + 3) / 4, PITCH = 4 * QX + 4;
    // images up to 20 px are staged whole (18 px in strips of 8 would leave a 2-row strip paying a full stage + two
    // barriers) as long as two CTAs still fit an SM
    a.TH = kWgTH;
    if (a.S <= 20) a.TH = a.S <= 10 ? a.S : (a.S + 1) / 2;        // 18 px: two strips of 9, not 8 + 8 + 2
    a.in_plane = pad_plane((a.TH + 2) * PITCH);
    a.g_plane = pad_plane(a.TH * 4 * QX);
    // eight output channels per thread on the cp.async path (decided below; needs Cout % 8 == 0)
    static const bool async_off = getenv("PAIG_WGRAD_SYNC") != nullptr;
#ifdef PAIG_EMU
    const bool want_async = false;
#else
    const bool want_async = !async_off && a.S != 4 * QX && !a.in_mask && !a.act;
#endif
    const int COB = want_async && a.Cout % 8 == 0 ? 8 : 4;
    const int G = ((a.Cout + COB - 1) / COB) * a.Cin;
    const int gsets = cdiv(G, kConvThreads);
    const int G_per = cdiv(G, gsets);
    const int P = kConvThreads / G_per > 0 ? kConvThreads / G_per : 1;
    size_t tile = (size_t)a.Cin * a.in_plane + (size_t)a.Cout * a.g_plane;
    // rows the TMA unit cannot address (pitch not a multiple of 16 B): double-buffered cp.async staging when the tile
    // fits twice next to a second CTA and no ReLU mask has to be applied on the way in
    a.async2 = 0;
    if (want_async && 2 * tile * sizeof(float) <= 110 * 1024) {
        a.async2 = 1;
        tile *= 2;
    }
    const size_t red = (size_t)P * G_per * COB * 10;
    const size_t smem = (tile > red ? tile : red) * sizeof(float);
    if (smem > 220 * 1024) {
        set_error("conv3x3_wgrad: %d->%d channels at %d px needs %zu B of shared memory", a.Cin, a.Cout, a.S, smem);
        return 1;
    }
    const int strips = cdiv(a.S, a.TH);
    int ctas = a.N * strips;
    if (ctas > kWgradMaxCtas) ctas = kWgradMaxCtas;
    int lq = -1;
    const bool aligned = ((uintptr_t)a.in % 16 == 0) && (a.in_bs % 4 == 0) && ((uintptr_t)a.g % 16 == 0) &&
                         (a.g_bs % 4 == 0) && (!a.act || (((uintptr_t)a.act % 16 == 0) && (a.act_bs % 4 == 0))) &&
                         !a.in_mask;
    if (aligned && a.S == 4 * QX)
        for (int k = 0; k <= 4; ++k)
            if (QX == (1 << k)) lq = k;
    const dim3 wg_grid(ctas, gsets), wg_blk(kConvThreads);
    switch (lq) {
        case 1: launch(conv3x3_wgrad_kernel<1, 4>, wg_grid, wg_blk, smem, st, a); break;
        case 2: launch(conv3x3_wgrad_kernel<2, 4>, wg_grid, wg_blk, smem, st, a); break;
        case 3: launch(conv3x3_wgrad_kernel<3, 4>, wg_grid, wg_blk, smem, st, a); break;
        case 4: launch(conv3x3_wgrad_kernel<4, 4>, wg_grid, wg_blk, smem, st, a); break;
        default:
            if (COB == 8) launch(conv3x3_wgrad_kernel<-1, 8>, wg_grid, wg_blk, smem, st, a);
            else launch(conv3x3_wgrad_kernel<-1, 4>, wg_grid, wg_blk, smem, st, a);
    }
    int rc = check_launch(layer_name("conv3x3_wgrad", a.Cin, a.Cout, a.S));
    if (rc) return rc;
    const int nW = a.Cout * a.Cin * 9;
    if (a.defer) return a.defer->add(a.partials, ctas, nW + a.Cout, nW, dW, a.Cout, db) ? 0 : 1;
    return reduce_partials(a.partials, ctas, nW + a.Cout, nW, dW, a.Cout, db, st);
}

// g *= (act > 0) in place over channel-sliced NCHW views: the ReLU adjoint applied once, so that the weight-gradient
// kernel can take g through TMA (wgrad_tma.cu needs an already gated gradient) and the data-gradient kernel's own
// gate becomes a no-op.
__global__ void __launch_bounds__(256) relu_gate_kernel(float* __restrict__ g, long g_bs, const float* __restrict__ act,
                                                        long act_bs, long per_frame, long total) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const long f = idx / per_frame, r = idx % per_frame;
    if (!(act[f * act_bs + r] > 0.f)) g[f * g_bs + r] = 0.f;
}

// 16 bytes per thread, the frame on grid.y (no division); g is only touched where some activation is not positive
__global__ void __launch_bounds__(256) relu_gate_vec_kernel(float* __restrict__ g, long g_bs, const float* __restrict__ act,
                                                            long act_bs, int per4) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= per4) return;
    const long f = blockIdx.y;
    const float4 a = reinterpret_cast<const float4*>(act + f * act_bs)[i];
    const bool zx = !(a.x > 0.f), zy = !(a.y > 0.f), zz = !(a.z > 0.f), zw = !(a.w > 0.f);
    if (!(zx | zy | zz | zw)) return;
    float4* gp = reinterpret_cast<float4*>(g + f * g_bs) + i;
    float4 v = *gp;
    if (zx) v.x = 0.f;
    if (zy) v.y = 0.f;
    if (zz) v.z = 0.f;
    if (zw) v.w = 0.f;
    *gp = v;
}

int relu_gate(float* g, long g_bs, const float* act, long act_bs, int C, int S, int N, cudaStream_t st) {
    const long per = (long)C * S * S, total = per * N;
    if (total <= 0) return 0;
    if (per % 4 == 0 && g_bs % 4 == 0 && act_bs % 4 == 0 && ((uintptr_t)g % 16) == 0 && ((uintptr_t)act % 16) == 0 &&
        N <= 65535) {
        launch(relu_gate_vec_kernel, dim3(cdiv(per / 4, 256), N), dim3(256), 0, st, g, g_bs, act, act_bs, (int)(per / 4));
        return check_launch("relu_gate");
    }
    launch(relu_gate_kernel, dim3(cdiv(total, 256)), dim3(256), 0, st, g, g_bs, act, act_bs, per, total);
    return check_launch("relu_gate");
}

// ---------------------------------------------------------------------------------------------------------
// 1x1 head (blocks.py:236,307): Cin in {8,16} -> Cout = n_objs.  One thread per pixel.
// ---------------------------------------------------------------------------------------------------------
constexpr int kMaxHeadIn = 16;

__global__ void __launch_bounds__(256) conv1x1_fwd_kernel(const float* __restrict__ in, long in_bs, int Cin,
                                                          const float* __restrict__ w, const float* __restrict__ b,
                                                          float* __restrict__ out, long out_bs, int Cout, int HW, int N,
                                                          int relu) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)N * HW) return;
    const int f = (int)(idx / HW), p = (int)(idx % HW);
    float v[kMaxHeadIn];
#pragma unroll
    for (int c = 0; c < kMaxHeadIn; ++c) v[c] = c < Cin ? in[(long)f * in_bs + (long)c * HW + p] : 0.f;
    for (int co = 0; co < Cout; ++co) {
        float s = b[co];
#pragma unroll
        for (int c = 0; c < kMaxHeadIn; ++c)
            if (c < Cin) s += w[co * Cin + c] * v[c];
        if (relu) s = fmaxf(s, 0.f);
        out[(long)f * out_bs + (long)co * HW + p] = s;
    }
}

// dIn written; per-block partials of dW [Cout*Cin] and db [Cout].
__global__ void __launch_bounds__(256) conv1x1_bwd_kernel(const float* __restrict__ in, long in_bs, int Cin,
                                                          const float* __restrict__ w, const float* __restrict__ dout,
                                                          long dout_bs, const float* __restrict__ act, long act_bs,
                                                          int Cout, int HW, int N, float* __restrict__ din, long din_bs,
                                                          float* __restrict__ partials) {
    __shared__ float red[8][kMaxObjs * (kMaxHeadIn + 1)];
    float aw[kMaxObjs][kMaxHeadIn], ab[kMaxObjs];
#pragma unroll
    for (int o = 0; o < kMaxObjs; ++o) {
        ab[o] = 0.f;
#pragma unroll
        for (int c = 0; c < kMaxHeadIn; ++c) aw[o][c] = 0.f;
    }
    const long total = (long)N * HW;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int f = (int)(idx / HW), p = (int)(idx % HW);
        float g[kMaxObjs];
#pragma unroll
        for (int o = 0; o < kMaxObjs; ++o) {
            g[o] = 0.f;
            if (o < Cout) {
                g[o] = dout[(long)f * dout_bs + (long)o * HW + p];
                if (act && !(act[(long)f * act_bs + (long)o * HW + p] > 0.f)) g[o] = 0.f;
                ab[o] += g[o];
            }
        }
#pragma unroll
        for (int c = 0; c < kMaxHeadIn; ++c) {
            if (c < Cin) {
                const float v = in[(long)f * in_bs + (long)c * HW + p];
#pragma unroll
                for (int o = 0; o < kMaxObjs; ++o)
                    if (o < Cout) aw[o][c] += g[o] * v;
                if (din) {                       // (the fused backward-data kernel computes d in itself: din == NULL)
                    float d = 0.f;
#pragma unroll
                    for (int o = 0; o < kMaxObjs; ++o)
                        if (o < Cout) d += w[o * Cin + c] * g[o];
                    din[(long)f * din_bs + (long)c * HW + p] = d;
                }
            }
        }
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 0; o < kMaxObjs; ++o) {
#pragma unroll
        for (int c = 0; c < kMaxHeadIn; ++c) {
            const float s = warp_sum(aw[o][c]);
            if (lane == 0) red[wid][o * (kMaxHeadIn + 1) + c] = s;
        }
        const float sb = warp_sum(ab[o]);
        if (lane == 0) red[wid][o * (kMaxHeadIn + 1) + kMaxHeadIn] = sb;
    }
    __syncthreads();
    const int nW = Cout * Cin;
    for (int e = threadIdx.x; e < nW + Cout; e += blockDim.x) {
        const int o = e < nW ? e / Cin : e - nW, c = e < nW ? e % Cin : kMaxHeadIn;
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) s += red[k][o * (kMaxHeadIn + 1) + c];
        partials[(size_t)blockIdx.x * (nW + Cout) + e] = s;
    }
}

// Weight / bias gradient only (the fused backward-data kernel owns d in): a CTA walks whole frames, a thread owns four
// consecutive pixels (16-byte loads, 32-bit indexing), per-CTA partial at the end as in conv1x1_bwd_kernel.
template <int CIN>
__global__ void __launch_bounds__(256) conv1x1_wgrad_kernel(const float* __restrict__ in, long in_bs,
                                                            const float* __restrict__ dout, long dout_bs,
                                                            const float* __restrict__ act, long act_bs, int Cout, int HW,
                                                            int N, float* __restrict__ partials) {
    __shared__ float red[8][kMaxObjs * (kMaxHeadIn + 1)];
    float aw[kMaxObjs][CIN], ab[kMaxObjs];
#pragma unroll
    for (int o = 0; o < kMaxObjs; ++o) {
        ab[o] = 0.f;
#pragma unroll
        for (int c = 0; c < CIN; ++c) aw[o][c] = 0.f;
    }
    const int q4 = HW / 4;
    for (int f = blockIdx.x; f < N; f += gridDim.x) {
        const float* inf = in + (long)f * in_bs;
        const float* gf = dout + (long)f * dout_bs;
        const float* af = act ? act + (long)f * act_bs : nullptr;
        for (int q = threadIdx.x; q < q4; q += 256) {
            float4 g[kMaxObjs];
#pragma unroll
            for (int o = 0; o < kMaxObjs; ++o) {
                g[o] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (o < Cout) {
                    g[o] = *reinterpret_cast<const float4*>(gf + o * HW + 4 * q);
                    if (af) {
                        const float4 a = *reinterpret_cast<const float4*>(af + o * HW + 4 * q);
                        g[o].x = a.x > 0.f ? g[o].x : 0.f; g[o].y = a.y > 0.f ? g[o].y : 0.f;
                        g[o].z = a.z > 0.f ? g[o].z : 0.f; g[o].w = a.w > 0.f ? g[o].w : 0.f;
                    }
                    ab[o] += (g[o].x + g[o].y) + (g[o].z + g[o].w);
                }
            }
            float4 v[CIN];
#pragma unroll
            for (int c = 0; c < CIN; ++c) v[c] = *reinterpret_cast<const float4*>(inf + c * HW + 4 * q);
#pragma unroll
            for (int c = 0; c < CIN; ++c)
#pragma unroll
                for (int o = 0; o < kMaxObjs; ++o)
                    if (o < Cout) aw[o][c] += (g[o].x * v[c].x + g[o].y * v[c].y) + (g[o].z * v[c].z + g[o].w * v[c].w);
        }
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 0; o < kMaxObjs; ++o) {
#pragma unroll
        for (int c = 0; c < CIN; ++c) {
            const float s = warp_sum(aw[o][c]);
            if (lane == 0) red[wid][o * (kMaxHeadIn + 1) + c] = s;
        }
        const float sb = warp_sum(ab[o]);
        if (lane == 0) red[wid][o * (kMaxHeadIn + 1) + kMaxHeadIn] = sb;
    }
    __syncthreads();
    const int nW = Cout * CIN;
    for (int e = threadIdx.x; e < nW + Cout; e += blockDim.x) {
        const int o = e < nW ? e / CIN : e - nW, c = e < nW ? e % CIN : kMaxHeadIn;
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) s += red[k][o * (kMaxHeadIn + 1) + c];
        partials[(size_t)blockIdx.x * (nW + Cout) + e] = s;
    }
}

int conv1x1_forward(const float* in, long in_bs, int Cin, const float* w, const float* b, float* out, long out_bs,
                    int Cout, int S, int N, int relu, cudaStream_t st) {
    if (Cin > kMaxHeadIn || Cout > kMaxObjs) {
        set_error("conv1x1: %d->%d channels unsupported", Cin, Cout);
        return 1;
    }
    launch(conv1x1_fwd_kernel, dim3(cdiv((long)N * S * S, 256)), dim3(256), 0, st, in, in_bs, Cin, w, b, out, out_bs,
           Cout, S * S, N, relu);
    return check_launch("conv1x1_fwd");
}

int conv1x1_backward(const float* in, long in_bs, int Cin, const float* w, const float* dout, long dout_bs,
                     const float* act, long act_bs, int Cout, int S, int N, float* din, long din_bs, float* dW,
                     float* db, float* partials, cudaStream_t st, ReduceBatch* defer) {
    if (Cin > kMaxHeadIn || Cout > kMaxObjs) {
        set_error("conv1x1: %d->%d channels unsupported", Cin, Cout);
        return 1;
    }
    int blocks = cdiv((long)N * S * S, 256 * 4);
    if (blocks > 592) blocks = 592;
    const bool vec = !din && (S * S) % 4 == 0 && (in_bs % 4) == 0 && (dout_bs % 4) == 0 && (!act || act_bs % 4 == 0) &&
                     ((uintptr_t)in % 16) == 0 && ((uintptr_t)dout % 16) == 0 && (!act || (uintptr_t)act % 16 == 0);
    if (vec && (Cin == 8 || Cin == 16)) {
        blocks = N < 592 ? N : 592;
        if (Cin == 8) launch(conv1x1_wgrad_kernel<8>, dim3(blocks), dim3(256), 0, st, in, in_bs, dout, dout_bs, act, act_bs, Cout,
                             S * S, N, partials);
        else launch(conv1x1_wgrad_kernel<16>, dim3(blocks), dim3(256), 0, st, in, in_bs, dout, dout_bs, act, act_bs, Cout,
                    S * S, N, partials);
        int rc = check_launch("conv1x1_bwd");
        if (rc) return rc;
        if (defer) return defer->add(partials, blocks, Cout * Cin + Cout, Cout * Cin, dW, Cout, db) ? 0 : 1;
        return reduce_partials(partials, blocks, Cout * Cin + Cout, Cout * Cin, dW, Cout, db, st);
    }
    launch(conv1x1_bwd_kernel, dim3(blocks), dim3(256), 0, st, in, in_bs, Cin, w, dout, dout_bs, act, act_bs, Cout,
           S * S, N, din, din_bs, partials);
    int rc = check_launch("conv1x1_bwd");
    if (rc) return rc;
    if (defer) return defer->add(partials, blocks, Cout * Cin + Cout, Cout * Cin, dW, Cout, db) ? 0 : 1;
    return reduce_partials(partials, blocks, Cout * Cin + Cout, Cout * Cin, dW, Cout, db, st);
}

// ---------------------------------------------------------------------------------------------------------
// 2x2 max-pool (blocks.py:281,284) and its adjoint (first maximum in row-major window order wins, as ATen).
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) maxpool2_kernel(const float* __restrict__ in, long in_bs, float* __restrict__ out,
                                                       long out_bs, int C, int So, int N) {
    // the frame rides on grid.y: the element index inside a frame is 32-bit (64-bit div/mod per element made these
    // elementwise kernels compute-bound)
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= C * So * So) return;
    const int x = idx % So, y = (idx / So) % So, c = idx / (So * So);
    const int Si = 2 * So;
    for (long f = blockIdx.y; f < N; f += gridDim.y) {
        const float* p = in + f * in_bs + ((long)c * Si + 2 * y) * Si + 2 * x;
        const float m = fmaxf(fmaxf(p[0], p[1]), fmaxf(p[Si], p[Si + 1]));
        out[f * out_bs + ((long)c * So + y) * So + x] = m;
    }
}

// din[window argmax] += dout   (din is the gradient of the pooled tensor's source; other entries untouched)
__global__ void __launch_bounds__(256) maxpool2_bwd_kernel(const float* __restrict__ in, long in_bs,
                                                           const float* __restrict__ dout, long dout_bs,
                                                           float* __restrict__ din, long din_bs, int C, int So,
                                                           int N) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= C * So * So) return;
    const int x = idx % So, y = (idx / So) % So, c = idx / (So * So);
    const int Si = 2 * So;
    const long o = ((long)c * Si + 2 * y) * Si + 2 * x;
    for (long f = blockIdx.y; f < N; f += gridDim.y) {
        const float* p = in + f * in_bs + o;
        int best = 0;
        float m = p[0];
        if (p[1] > m) { m = p[1]; best = 1; }
        if (p[Si] > m) { m = p[Si]; best = Si; }
        if (p[Si + 1] > m) { m = p[Si + 1]; best = Si + 1; }
        din[f * din_bs + o + best] += dout[f * dout_bs + ((long)c * So + y) * So + x];
    }
}

int maxpool2(const float* in, long in_bs, float* out, long out_bs, int C, int So, int N, cudaStream_t st) {
    const long total = (long)N * C * So * So;
    if (total <= 0) return 0;
    launch(maxpool2_kernel, dim3(cdiv((long)C * So * So, 256), N < 32768 ? N : 32768), dim3(256), 0, st, in, in_bs, out, out_bs,
           C, So, N);
    return check_launch("maxpool2");
}
int maxpool2_backward(const float* in, long in_bs, const float* dout, long dout_bs, float* din, long din_bs, int C,
                      int So, int N, cudaStream_t st) {
    const long total = (long)N * C * So * So;
    if (total <= 0) return 0;
    launch(maxpool2_bwd_kernel, dim3(cdiv((long)C * So * So, 256), N < 32768 ? N : 32768), dim3(256), 0, st, in, in_bs, dout,
           dout_bs, din, din_bs, C, So, N);
    return check_launch("maxpool2_bwd");
}

// ---------------------------------------------------------------------------------------------------------
// exact 2x bilinear upsample, align_corners=False (tvtrans.Resize on a tensor):
//   out[2i] = .25 in[i-1] + .75 in[i],  out[2i+1] = .75 in[i] + .25 in[i+1],  indices clamped; separable.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void up_taps(int o, int Si, int& i0, int& i1, float& w0, float& w1) {
    const int k = o >> 1;
    if (o & 1) { i0 = k; i1 = min(k + 1, Si - 1); w0 = 0.75f; w1 = 0.25f; }
    else { i0 = max(k - 1, 0); i1 = k; w0 = 0.25f; w1 = 0.75f; }
}

__global__ void __launch_bounds__(256) upsample2_kernel(const float* __restrict__ in, long in_bs, float* __restrict__ out,
                                                        long out_bs, int C, int Si, long total) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int So = 2 * Si;
    const int x = (int)(idx % So), y = (int)((idx / So) % So), c = (int)((idx / ((long)So * So)) % C);
    const long f = idx / ((long)So * So * C);
    int xa, xb, ya, yb;
    float wxa, wxb, wya, wyb;
    up_taps(x, Si, xa, xb, wxa, wxb);
    up_taps(y, Si, ya, yb, wya, wyb);
    const float* p = in + f * in_bs + (long)c * Si * Si;
    const float top = wxa * p[ya * Si + xa] + wxb * p[ya * Si + xb];      // W pass, then H pass (ATen's order)
    const float bot = wxa * p[yb * Si + xa] + wxb * p[yb * Si + xb];
    out[f * out_bs + ((long)c * So + y) * So + x] = wya * top + wyb * bot;
}

// The same upsample, four adjacent outputs per thread (So % 4 == 0): 32-bit index arithmetic once per quad instead of
// 64-bit div/mod per element, the four source columns of both source rows loaded once, one 16-byte store.  Per element
// the expression is the scalar kernel's (W pass with up_taps' weights, then H pass).
__global__ void __launch_bounds__(256) upsample2_quad_kernel(const float* __restrict__ in, long in_bs,
                                                             float* __restrict__ out, long out_bs, int C, int Si,
                                                             unsigned total_quads) {
    const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total_quads) return;
    const int So = 2 * Si, qpr = So >> 2;
    const unsigned q = idx % qpr, t1 = idx / qpr;
    const unsigned y = t1 % So, t2 = t1 / So;
    const unsigned c = t2 % C, f = t2 / C;
    int ya, yb;
    float wya, wyb;
    up_taps((int)y, Si, ya, yb, wya, wyb);
    const float* p = in + (long)f * in_bs + (long)c * Si * Si;
    const int k = 2 * (int)q;                                  // source column of output x0 / 2
    const int cm = max(k - 1, 0), cp2 = min(k + 2, Si - 1);
    const float* ra = p + ya * Si;
    const float* rb = p + yb * Si;
    const float a[4] = {ra[cm], ra[k], ra[k + 1], ra[cp2]};
    const float b[4] = {rb[cm], rb[k], rb[k + 1], rb[cp2]};
    // outputs x0 (even): .25 s[k-1] + .75 s[k]; x0+1: .75 s[k] + .25 s[k+1]; x0+2: .25 s[k] + .75 s[k+1]; x0+3: .75 s[k+1] + .25 s[k+2]
    float o[4];
    {
        const float w0[4] = {0.25f, 0.75f, 0.25f, 0.75f}, w1[4] = {0.75f, 0.25f, 0.75f, 0.25f};
        const int i0[4] = {0, 1, 1, 2}, i1[4] = {1, 2, 2, 3};
#pragma unroll
        for (int px = 0; px < 4; ++px) {
            const float top = w0[px] * a[i0[px]] + w1[px] * a[i1[px]];
            const float bot = w0[px] * b[i0[px]] + w1[px] * b[i1[px]];
            o[px] = wya * top + wyb * bot;
        }
    }
    *reinterpret_cast<float4*>(out + (long)f * out_bs + ((long)c * So + y) * So + 4 * q) = make_float4(o[0], o[1], o[2], o[3]);
}

// adjoint (gather, written): input i feeds outputs 2i-1 (.25), 2i (.75), 2i+1 (.75), 2i+2 (.25); at the borders the
// clamped taps fold back: output 0 reads input 0 with .25 + .75 = 1, output So-1 reads input Si-1 with 1.
__device__ __forceinline__ void up_adj(int i, int Si, int& o0, float (&w)[4]) {
    o0 = 2 * i - 1;
    w[0] = i > 0 ? 0.25f : 0.f;
    w[1] = i > 0 ? 0.75f : 1.f;
    w[2] = i < Si - 1 ? 0.75f : 1.f;
    w[3] = i < Si - 1 ? 0.25f : 0.f;
}

__global__ void __launch_bounds__(256) upsample2_bwd_kernel(const float* __restrict__ dout, long dout_bs,
                                                            float* __restrict__ din, long din_bs, int C, int Si,
                                                            int N) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= C * Si * Si) return;
    const int So = 2 * Si;
    const int j = idx % Si, i = (idx / Si) % Si, c = idx / (Si * Si);
    int oy0, ox0;
    float wy[4], wx[4];
    up_adj(i, Si, oy0, wy);
    up_adj(j, Si, ox0, wx);
    for (long f = blockIdx.y; f < N; f += gridDim.y) {
        const float* g = dout + f * dout_bs + (long)c * So * So;
        float s = 0.f;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int oy = oy0 + a;
            if (wy[a] == 0.f) continue;
            const float* row = g + oy * So + ox0;
            float r = 0.f;
#pragma unroll
            for (int b = 0; b < 4; ++b)
                if (wx[b] != 0.f) r += wx[b] * row[b];
            s += wy[a] * r;
        }
        din[f * din_bs + ((long)c * Si + i) * Si + j] = s;
    }
}

int upsample2(const float* in, long in_bs, float* out, long out_bs, int C, int Si, int N, cudaStream_t st) {
    const long total = (long)N * C * 4 * Si * Si;
    if (total <= 0) return 0;
    if ((2 * Si) % 4 == 0 && (out_bs % 4) == 0 && ((uintptr_t)out % 16) == 0 && total / 4 < 0xffffffffL) {
        launch(upsample2_quad_kernel, dim3(cdiv(total / 4, 256)), dim3(256), 0, st, in, in_bs, out, out_bs, C, Si,
               (unsigned)(total / 4));
        return check_launch("upsample2");
    }
    launch(upsample2_kernel, dim3(cdiv(total, 256)), dim3(256), 0, st, in, in_bs, out, out_bs, C, Si, total);
    return check_launch("upsample2");
}
int upsample2_backward(const float* dout, long dout_bs, float* din, long din_bs, int C, int Si, int N, cudaStream_t st) {
    const long total = (long)N * C * Si * Si;
    if (total <= 0) return 0;
    launch(upsample2_bwd_kernel, dim3(cdiv((long)C * Si * Si, 256), N < 32768 ? N : 32768), dim3(256), 0, st, dout, dout_bs,
           din, din_bs, C, Si, N);
    return check_launch("upsample2_bwd");
}

}  // namespace paig
