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
    const size_t red = (size_t)P * G_per * COB * 10;
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
    if (COB == 8) launch(conv3x3_wgrad_tma_kernel<8>, dim3(ctas, sets), dim3(kWtThreads), smem, st, a);
    else launch(conv3x3_wgrad_tma_kernel<4>, dim3(ctas, sets), dim3(kWtThreads), smem, st, a);
    int rc = check_launch(layer_name("conv3x3_wgrad", -a.Cin, a.Cout, a.S));   // (negative Cin marks the TMA kernel)
    if (rc) return rc;
    const int nW = a.Cout * a.Cin * 9;
    if (w.defer) return w.defer->add(a.partials, ctas, nW + a.Cout, nW, dW, a.Cout, db) ? 0 : 1;
    return reduce_partials(a.partials, ctas, nW + a.Cout, nW, dW, a.Cout, db, st);
}

}  // namespace paig
