// ShallowUNet forward (blocks.py:278-308) on the 5th-generation tensor cores: one persistent kernel, a frame's activations
// resident in shared memory (as in unet_fused.cu), every 3x3 convolution a handful of tcgen05.mma instructions.
//
// Why a second formulation.  As an implicit GEMM with one MMA per tap (conv_tc.cu: M = 128 pixels, N = Cout, K = 8 input
// channels) a layer with 8..32 output channels is bound by the pixel slab every MMA reads from shared memory, whatever N is
// (measured here: a tf32 MMA with M = 128 costs ~80-130 cycles from shared memory for any N <= 192, i.e. the operand-A fetch
// of 128 rows paces it): at N = 16 the tensor path only ties with the FMA loop.  Here the taps of a row
// AND the two halves of the weight split are batched along N:
//
//      D[pixel p, (kx, co)] = sum_{ky, ci} in[p + (ky - 1) rows, ci] * w[co, ci, ky, kx]        M = 128, K = 3 Cin
//      out[y, x, co]        = D[(y, x-1), 0, co] + D[(y, x), 1, co] + D[(y, x+1), 2, co] + b[co]
//
// The ky shift is a START ADDRESS (activations live as [channel quad][row][pixel][4 channels], a pixel row shift is S x 16
// bytes: the K-major "no swizzle" core-matrix layout, 8 pixels x 16 bytes contiguous); the kx shift is left to the epilogue,
// where pixel x-1 / x+1 of the same image row is the neighbouring lane of the warp that drains the accumulator: two shuffles
// and two additions per output instead of 9 Cin FMAs.  Every pixel slab is read 3 times per K chunk instead of 9 and
// multiplies 6x as many columns.  No halo columns, no im2col, no per-tap conversion.
//
// fp32 accuracy (3xTF32): the tensor core reads the upper 19 bits of an fp32 operand, so the resident fp32 plane IS the hi
// operand (hi = x with the low 13 mantissa bits dropped); the stager warps write lo = rn_tf32(x - hi) (exact difference,
// rounded to nearest) for the 4..16 rows of the tile in flight into a small staging ring -- the frame is not stored twice.
// Weights are split once per step (hi = rn_tf32(w), lo = rn_tf32(w - hi)) and packed per (K chunk, ky) as ONE operand block
// [w_hi rows | w_lo rows], so that per (tile, K chunk, ky) two MMAs do the three products:
//      [main | corr] (+)= hi x [w_hi | w_lo]      (N = 2 npad)          corr += lo x w_hi      (N = npad, same block)
// The epilogue adds main + corr in fp32 (round to nearest) together with the expected truncation loss of the main chain
// (the tensor core accumulates with truncation; 3 Cin/8 MMAs per chain; kappa per chain length calibrated against float64,
// tc_kappa() below -- the same device conv_tc.cu uses, which explains why a coherent 1e-7 bias matters here).
//
// Warp roles (448 threads, one CTA per SM, CTA b handles frames b, b + grid, ...): 0-7 epilogue in two groups that take the
// tiles in turn (TMEM lane quarter = warp % 4), 8-11 stagers (lo operand), 12-13 MMA issuers (even / odd tiles; warp 12 also
// allocates TMEM and prefetches the weights: one TMA bulk copy per layer, one layer ahead).  Max-pool, the 2x bilinear
// upsample, the frame load and the 1x1 head are elementwise passes of warps 0-11 between the convolutions.  Everything the
// backward pass reads leaves the SM as fire-and-forget NCHW stores, exactly where unet_fused_fwd_kernel puts it.
//
// Measured (B200, 1000 frames of 32 x 32, profiles/r2s_*): 0.49 ms against 0.72 ms for the FMA kernel; closer to float64
// than the FMA kernel on every layer (tools/unet_tc_check.py).  PAIG_UNET_TC=0 selects the FMA kernel.
#include "common.cuh"
#include "internal.h"
#include "layout.h"
#include "unet_plan.h"

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace paig {

static inline int tc_pad16(int c) { return (c + 15) & ~15; }

// packed weights of the 3x3 layers: [K chunk][ky][channel quad 2][hi rows | lo rows: 2 npad][4]; the forward packing first
// (unet_tc_wpack_fwd_floats), then the transposed slices of the backward-data pass (a concat layer's two parts pad separately)
static size_t unet_tc_wpack_fwd_floats(const UNetDesc& u) {
    size_t total = 64;
    for (int i = 0; i < u.nops; ++i)
        if (u.ops[i].kind == OP_CONV) total += align64((size_t)((u.ops[i].in.C + 7) / 8) * 3 * 2 * 2 * tc_pad16(3 * u.ops[i].out.C) * 4);
    return total;
}
size_t unet_tc_wpack_floats(const UNetDesc& u) {
    size_t total = unet_tc_wpack_fwd_floats(u);
    for (int i = 0; i < u.nops; ++i)
        if (u.ops[i].kind == OP_CONV) total += 2 * 64 + (size_t)((u.ops[i].out.C + 7) / 8) * 3 * 2 * 2 * (tc_pad16(3 * u.ops[i].in.C) + 16) * 4;
    return total;
}

#ifdef PAIG_EMU
int unet_tc_forward(const paig_task*, const paig_params*, const Layout&, const float*, long, int, float*, cudaStream_t) { return -1; }
int unet_tc_backward(const paig_task*, const paig_params*, const Layout&, float*, cudaStream_t) { return -1; }
#else

constexpr int kTcThreads = 448;          // 14 warps: 0-7 epilogue (two groups), 8-11 stagers, 12-13 MMA issuers (even / odd tiles)
constexpr int kTcWorkers = 384;          // warps 0-11: elementwise passes
constexpr int kTcMmaWarp = 12;
constexpr int kTcMaxOps = 22;
constexpr int kTcMaxChunks = 4;
constexpr size_t kTcSmemLimit = 227 * 1024 - 1024;      // dynamic part; the barriers are static

enum { T_CONV = 0, T_POOL = 1, T_UP = 2, T_HEAD = 3, T_HEADT = 4, T_UPT = 5, T_POOLT = 6 };

struct TcOp {
    int kind, S, Cin, Cout, relu;
    int nchunks, npad;                   // conv: K chunks of 8 input channels; N = 3 Cout padded to a multiple of 16
    int src[kTcMaxChunks];               // byte offset of each chunk's first channel quad (others: src[0] = the input buffer)
    int out;                             // byte offset of the result, -1: not kept on chip
    int stage, stage_q, stage_slots;     // conv: the lo staging slots (2, or 1 when shared memory is short); bytes between channel quads inside a slot
    int w, wbytes, wbar, next_w;         // conv: weights in shared memory, mbarrier, next conv (prefetched while this one runs)
    int late_w;                          // conv: no room to prefetch: the weights are fetched when the op starts
    long wglob;                          // float offset into the packed weights
    const float* bias;
    const float* w1;                     // head: [Cout][Cin]
    float* gout; long gout_bs;           // NCHW destination in the workspace (nullable)
    float comp;                          // conv: expected relative truncation loss of the main accumulation chain (3 nchunks MMAs)
    // backward-data pass (the op list of unet_fused.cu's fused_backward_plan, run in this kernel's layout)
    const float* gmask; long gmask_bs;   // conv / T_UPT / T_HEADT: activation whose sign gates the result; T_POOLT: the pool's source
    const float* gsrc; long gsrc_bs;     // T_HEADT: d logits
    const float* gmask2; long gmask2_bs; // T_HEADT: the logits (head ReLU), nullable
    int in1;                             // T_POOLT: the other reader's share parked on chip (byte offset), -1: none
    int acc_gout;                        // T_POOLT: ... or parked in gout (global / L2), finalised in place
};
struct TcPlan {
    int nops, N, fps, H, first_w, x_off;
    long seq_stride;
    const float* x;
    const float* wpack;
    long long* timing;                   // PAIG_DEBUG: per-CTA cycle stamps after every op of the CTA's last frame
    int dbg_op;                          // PAIG_DEBUG: the conv whose tiles CTA 0 stamps in detail (timing + 160 * 32 ...)
    TcOp ops[kTcMaxOps];
};

__device__ __forceinline__ unsigned tc_smem(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tc_bar_init(unsigned long long* b, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc_smem(b)), "r"(count));
}
__device__ __forceinline__ void tc_wait(unsigned long long* b, unsigned parity) {
    const unsigned a = tc_smem(b);
    unsigned done = 0;
    while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ void tc_arrive(unsigned long long* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc_smem(b)) : "memory");
}
__device__ __forceinline__ bool tc_elect() {
    unsigned p;
    asm volatile("{ .reg .pred q; elect.sync _|q, 0xffffffff; selp.u32 %0, 1, 0, q; }" : "=r"(p));
    return p != 0;
}
__device__ __forceinline__ void tc_commit(unsigned long long* b) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tc_smem(b)) : "memory");
}
// K-major operand, no swizzle: core matrix = 8 rows x 16 bytes, contiguous; lbo = bytes between the two 16-byte K chunks
// of one MMA (K = 8 tf32), sbo = bytes between 8-row groups (same encoding as conv_tc.cu)
__device__ __forceinline__ uint64_t tc_desc(unsigned addr, unsigned lbo, unsigned sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(lbo >> 4) << 16;
    d |= (uint64_t)(sbo >> 4) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__device__ __forceinline__ void tc_mma(unsigned tmem_d, uint64_t da, uint64_t db, unsigned idesc, unsigned accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
        "}\n"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
// tcgen05.ld is asynchronous: the registers are valid after tcgen05.wait::ld.  tc_ld_pin() after the wait makes the compiler
// treat them as produced there (it cannot know the hardware writes them late and could otherwise copy them early).
__device__ __forceinline__ void tc_ld8(unsigned taddr, float (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "r"(taddr));
}
// a global load the compiler may not wait for or reorder (volatile): a run of these is in flight together, the first use of a
// result is where the thread waits.  (Left to the compiler, the gate loads of the backward pass were issued one round trip
// at a time: 128 registers per thread at 448 threads leave it no room to batch them on its own.)
__device__ __forceinline__ float tc_ldg_async(const float* p) {
    float v;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_ld_pin(float (&v)[8]) {
    asm volatile("" : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]));
}
__device__ __forceinline__ float tc_rn_tf32(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }
// what the tensor core does NOT see of x: x minus its upper 19 bits (exact), rounded to TF32
__device__ __forceinline__ float tc_lo(float x) { return tc_rn_tf32(x - __uint_as_float(__float_as_uint(x) & 0xffffe000u)); }
__device__ __forceinline__ void tc_bulk(unsigned dst, const float* src, unsigned bytes, unsigned long long* bar) {
    const unsigned bar_a = tc_smem(bar);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(__cvta_generic_to_global(src)), "r"(bytes), "r"(bar_a) : "memory");
}

// rows 0 and S+1 of every channel quad of an activation buffer are the convolution's zero padding
__device__ __forceinline__ void tc_zero_halo(unsigned char* buf, int quads, int S, int t, int nthr) {
    const int qs = (S + 2) * S * 16, ls = 31 - __clz(S);    // S is a power of two
    for (int e = t; e < quads * 2 * S; e += nthr) {
        const int px = e & (S - 1), r = (e >> ls) & 1, q = e >> (ls + 1);
        *reinterpret_cast<float4*>(buf + q * qs + (r ? (S + 1) * S * 16 : 0) + px * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// Stager side of a convolution (warps 8-11): lo = rn_tf32(x - upper19(x)) of each tile's rows plus one above and one below,
// into the two-slot ring the MMA issuer reads the lo operand from; also clears the padding rows of the output buffer.
#define TC_STAMP(slot) do { if (tm && (threadIdx.x & 127) == 0) tm[(t * 8 + (slot))] = clock64(); } while (0)
template <int NCH>
__device__ __forceinline__ void tc_conv_stage(const TcOp& op, unsigned char* sm, unsigned it, int ntiles,
                                              unsigned long long* lo_full, unsigned long long* lo_empty, long long* tm) {
    const int gt = threadIdx.x & 127;
    const int S = op.S, R = 128 / S, Rr = R < S ? R : S;
    const int qs = (S + 2) * S * 16;
    const int slot_bytes = NCH * 2 * op.stage_q;
    if (op.out >= 0) tc_zero_halo(sm + op.out, op.Cout / 4, S, gt, 128);
    const int per_q = (Rr + 2) * S;                         // float4 items per channel quad of a staged tile
    for (int t = 0; t < ntiles; ++t) {
        const unsigned i = it + t, s = i & 1u, ph = (i >> 1) & 1u;
        const int y0 = t * R;
        TC_STAMP(0);
        tc_wait(&lo_empty[s], ph ^ 1u);                                                        // tile i - 2 has been multiplied
        if (op.stage_slots == 1 && t > 0) tc_wait(&lo_empty[s ^ 1u], ((i - 1) >> 1) & 1u);     // one slot: so has tile i - 1
        TC_STAMP(1);
        unsigned char* stg = sm + op.stage + (op.stage_slots == 2 ? s * slot_bytes : 0);
        for (int r = gt; r < per_q; r += 128) {             // r = row * S + px; all channel quads of the pixel in flight together
            float4 v[2 * NCH];
#pragma unroll
            for (int qi = 0; qi < 2 * NCH; ++qi)
                v[qi] = *reinterpret_cast<const float4*>(sm + op.src[qi >> 1] + (qi & 1) * qs + (y0 * S + r) * 16);
#pragma unroll
            for (int qi = 0; qi < 2 * NCH; ++qi) {
                float4 l;
                l.x = tc_lo(v[qi].x); l.y = tc_lo(v[qi].y); l.z = tc_lo(v[qi].z); l.w = tc_lo(v[qi].w);
                *reinterpret_cast<float4*>(stg + qi * op.stage_q + r * 16) = l;
            }
        }
        TC_STAMP(2);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        tc_arrive(&lo_full[s]);
        TC_STAMP(3);
    }
}

// Epilogue side of a convolution (warps 0-7).  Group g = warp / 4 drains the accumulator sets of the tiles whose running index
// is g (mod 2): thread = pixel (TMEM lane), the kx shift is the neighbouring lane.  While one group drains tile i the tensor
// core works on tile i + 1 into the other set.
__device__ __forceinline__ void tc_conv_drain(const TcOp& op, unsigned char* sm, unsigned tmem, unsigned it, int ntiles, int f,
                                              unsigned long long* acc_full, unsigned long long* acc_empty, long long* tm) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int grp = warp >> 2;
    const int S = op.S, R = 128 / S, Rr = R < S ? R : S;
    const int qs = (S + 2) * S * 16;
    const int cols = 2 * op.npad;
    const int q = warp & 3, m = q * 32 + lane;
    const unsigned lane_base = tmem + ((unsigned)(q * 32) << 16);
    const int x = m & (S - 1);
    const bool valid = m < Rr * S;
    const float comp = op.comp;
    for (int t = 0; t < ntiles; ++t) {
        const unsigned i = it + t, s = i & 1u, ph = (i >> 1) & 1u;
        if ((int)s != grp) continue;
        const int y = t * R + m / S;
        // ReLU adjoint (backward-data pass): the gates of this pixel's Cout outputs are fetched while the MMAs run and kept as bits
        unsigned gate_bits = 0xffffffffu;
        if (op.gmask && valid) {
            const float* gm = op.gmask + (long)f * op.gmask_bs + y * S + x;
            gate_bits = 0u;
            for (int c0 = 0; c0 < op.Cout; c0 += 16) {      // Cout is 8, 16 or 32: 8 or 16 loads in flight
                float gv[16];
#pragma unroll
                for (int c = 0; c < 8; ++c) gv[c] = tc_ldg_async(gm + (long)(c0 + c) * S * S);
                if (op.Cout > 8) {
#pragma unroll
                    for (int c = 8; c < 16; ++c) gv[c] = tc_ldg_async(gm + (long)(c0 + c) * S * S);
                } else {
#pragma unroll
                    for (int c = 8; c < 16; ++c) gv[c] = 0.f;
                }
#pragma unroll
                for (int c = 0; c < 16; ++c) gate_bits |= (gv[c] > 0.f ? 1u : 0u) << (c0 + c);
            }
        }
        tc_wait(&acc_full[s], ph);
        TC_STAMP(4);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const unsigned d0 = lane_base + s * cols;
        for (int co0 = 0; co0 < op.Cout; co0 += 8) {
            float v[3][2][8];                               // [kx][main | corr][co]
#pragma unroll
            for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                for (int a = 0; a < 2; ++a) tc_ld8(d0 + kx * op.Cout + co0 + a * op.npad, v[kx][a]);
            tc_ld_wait();
#pragma unroll
            for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                for (int a = 0; a < 2; ++a) tc_ld_pin(v[kx][a]);
            if (co0 + 8 >= op.Cout) {                       // everything of this set is in registers: the next tile may overwrite it
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                tc_arrive(&acc_empty[s]);
            }
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float sx[3];
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const float sum = v[kx][0][j];
                    sx[kx] = sum + fmaf(sum, comp, v[kx][1][j]);
                }
                float left = __shfl_up_sync(0xffffffffu, sx[0], 1);
                float right = __shfl_down_sync(0xffffffffu, sx[2], 1);
                if (x == 0) left = 0.f;
                if (x == S - 1) right = 0.f;
                float r = (left + sx[1]) + (right + (op.bias ? __ldg(op.bias + co0 + j) : 0.f));
                if (op.relu) r = fmaxf(r, 0.f);
                if (!((gate_bits >> (co0 + j)) & 1u)) r = 0.f;
                o[j] = r;
            }
            if (valid) {
                if (op.out >= 0) {
                    unsigned char* d = sm + op.out + (co0 >> 2) * qs + ((y + 1) * S + x) * 16;
                    *reinterpret_cast<float4*>(d) = make_float4(o[0], o[1], o[2], o[3]);
                    *reinterpret_cast<float4*>(d + qs) = make_float4(o[4], o[5], o[6], o[7]);
                }
                if (op.gout) {
                    float* g = op.gout + (long)f * op.gout_bs + (long)co0 * S * S + y * S + x;
#pragma unroll
                    for (int j = 0; j < 8; ++j) g[j * S * S] = o[j];
                }
            }
        }
        TC_STAMP(5);
    }
}

__global__ void __launch_bounds__(kTcThreads, 1) unet_tc_fwd_kernel(const __grid_constant__ TcPlan P) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ unsigned long long wbar[2], lo_full[2], lo_empty[2], acc_full[2], acc_empty[2];
    __shared__ unsigned tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            tc_bar_init(&wbar[i], 1);
            tc_bar_init(&lo_full[i], 128);
            tc_bar_init(&lo_empty[i], 1);
            tc_bar_init(&acc_full[i], 1);
            tc_bar_init(&acc_empty[i], 128);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kTcMmaWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem(&tmem_slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = tmem_slot;
    const unsigned smb = tc_smem(sm);
    const bool issuer = tid == kTcMmaWarp * 32;            // lane 0 of the MMA warp
    unsigned it = 0;                                       // conv tiles so far: every role walks the same sequence
    unsigned wph0 = 0, wph1 = 0;                           // (issuer) phases of the two weight barriers
    const int H = P.H, HW = H * H;

    for (int f = blockIdx.x; f < P.N; f += gridDim.x) {
        if (issuer && P.first_w >= 0 && !P.ops[P.first_w].late_w) {
            const TcOp& o = P.ops[P.first_w];
            tc_bulk(smb + o.w, P.wpack + o.wglob, (unsigned)o.wbytes, &wbar[o.wbar]);
        }
        if (tid < kTcWorkers && !P.x) {
            // backward-data pass: everything this frame's gates read (saved activations, 350 KB, long evicted from L2 by the
            // time the backward runs) is requested into L2 now; the head adjoint comes first, so ITS inputs are requested one
            // frame ahead.  A gate load then costs an L2 round trip instead of a DRAM one.
            for (int k2 = 0; k2 < P.nops; ++k2) {
                const TcOp& o = P.ops[k2];
                const bool head = o.kind == T_HEADT;
                const int ff = head ? f + (int)gridDim.x : f;
                if (!o.gmask || ff >= P.N) continue;
                const int side = o.kind == T_POOLT ? 2 * o.S : o.S;
                const char* base = reinterpret_cast<const char*>(o.gmask + (long)ff * o.gmask_bs);
                for (int i = tid; i < o.Cout * side * side / 32; i += kTcWorkers) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (long)i * 128));
                if (head) {
                    const char* b1 = reinterpret_cast<const char*>(o.gsrc + (long)ff * o.gsrc_bs);
                    for (int i = tid; i < o.Cin * side * side / 32; i += kTcWorkers) asm volatile("prefetch.global.L2 [%0];" ::"l"(b1 + (long)i * 128));
                    if (o.gmask2) {
                        const char* b2 = reinterpret_cast<const char*>(o.gmask2 + (long)ff * o.gmask2_bs);
                        for (int i = tid; i < o.Cin * side * side / 32; i += kTcWorkers) asm volatile("prefetch.global.L2 [%0];" ::"l"(b2 + (long)i * 128));
                    }
                }
            }
        }
        if (tid < kTcWorkers && P.x) {
            // the input frame: quad 0 = (r, g, b, 0), quad 1 = 0 (the first conv's K chunk is 8 channels wide)
            unsigned char* X = sm + P.x_off;
            const int qs = (H + 2) * H * 16;
            tc_zero_halo(X, 2, H, tid, kTcWorkers);
            const float* xf = P.x + (long)(f / P.fps) * P.seq_stride + (long)(f % P.fps) * 3 * HW;
            for (int e = tid; e < HW; e += kTcWorkers) {
                const float4 v = make_float4(__ldg(xf + e), __ldg(xf + HW + e), __ldg(xf + 2 * HW + e), 0.f);
                *reinterpret_cast<float4*>(X + H * 16 + e * 16) = v;
                *reinterpret_cast<float4*>(X + qs + H * 16 + e * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (P.timing && tid == 0) P.timing[(long)blockIdx.x * 32] = clock64();

        for (int k = 0; k < P.nops; ++k) {
            const TcOp& op = P.ops[k];
            const int S = op.S;
            if (issuer && op.kind == T_CONV) {
                if (op.late_w) tc_bulk(smb + op.w, P.wpack + op.wglob, (unsigned)op.wbytes, &wbar[op.wbar]);
                if (op.next_w >= 0 && !P.ops[op.next_w].late_w) {
                    const TcOp& o = P.ops[op.next_w];
                    tc_bulk(smb + o.w, P.wpack + o.wglob, (unsigned)o.wbytes, &wbar[o.wbar]);
                }
            }
            if (op.kind == T_CONV) {
                const int ntiles = S * S >= 128 ? S * S / 128 : 1;
                const int R = 128 / S;                      // image rows a 128-pixel tile spans (S = 8: 16, half of them beyond the frame)
                const int qs = (S + 2) * S * 16;            // bytes between channel quads of an activation buffer
                const int cols = 2 * op.npad;               // [main | corr] per accumulator set, two sets
                const int slot_bytes = op.nchunks * 2 * op.stage_q;
                if (warp >= kTcMmaWarp) {
                    // ===== MMA issuers (warp 12: even tiles, warp 13: odd tiles -- one warp's barrier round trips hide under the
                    // other's MMAs): the whole warp walks the loop (warp-uniform control flow keeps the descriptor arithmetic in
                    // uniform registers; with a single thread every operand took a register -> uniform-register move and one MMA
                    // cost ~130 issue cycles), one elected lane issues =====
                    {
                        const bool lead = tc_elect();
                        if (op.wbar == 0) { tc_wait(&wbar[0], wph0); wph0 ^= 1u; }
                        else { tc_wait(&wbar[1], wph1); wph1 ^= 1u; }
                        // MMA 1: [main | corr] (+)= hi x [w_hi | w_lo]  (N = 2 npad: one read of the pixel slab feeds both products);
                        // MMA 2: corr += lo x w_hi  (N = npad: the first npad rows of the same weight block)
                        const unsigned idesc1 = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(2 * op.npad >> 3) << 17) | ((unsigned)(128 >> 4) << 24);
                        const unsigned idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(op.npad >> 3) << 17) | ((unsigned)(128 >> 4) << 24);
                        const unsigned wblk = 2u * 2u * op.npad * 16u;      // one (chunk, ky) operand block: [quad][hi rows | lo rows][4]
                        for (int t = 0; t < ntiles; ++t) {
                            const unsigned i = it + t, s = i & 1u, ph = (i >> 1) & 1u;
                            if ((int)s != warp - kTcMmaWarp) continue;
                            const int y0 = t * R;
                            long long* tm = (P.timing && blockIdx.x == 0 && k == P.dbg_op && lane == 0) ? P.timing + 160 * 32 : nullptr;
                            if (tm) tm[t * 8 + 6] = clock64();
                            tc_wait(&lo_full[s], ph);                       // lo tile staged
                            tc_wait(&acc_empty[s], ph ^ 1u);                // accumulator set s drained (tile i - 2)
                            if (tm) tm[t * 8 + 7] = clock64();
                            if (tm) tm[256 + t * 4 + 0] = clock64();
                            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                            const unsigned d0 = tmem + s * cols;
                            const unsigned stg = smb + op.stage + (op.stage_slots == 2 ? s * slot_bytes : 0u);
                            // descriptors differ only in their start-address field (bits 0-13, units of 16 bytes): one full
                            // encoding per operand kind, then additions
                            const uint64_t row = (uint64_t)(S * 16 >> 4), wstep = (uint64_t)(wblk >> 4);
                            uint64_t bw = tc_desc(smb + op.w, 2u * op.npad * 16u, 128u);
                            const uint64_t al0 = tc_desc(stg, op.stage_q, 128u);
                            // (issuing all the hi MMAs first and the lo MMAs after them was measured: 5 % slower)
                            for (int kc = 0; kc < op.nchunks; ++kc) {
                                uint64_t ah = tc_desc(smb + op.src[kc] + y0 * S * 16, qs, 128u);
                                uint64_t al = al0 + (uint64_t)((2 * kc * op.stage_q) >> 4);
#pragma unroll
                                for (int dy = 0; dy < 3; ++dy) {
                                    if (lead) tc_mma(d0, ah, bw, idesc1, (kc > 0 || dy > 0) ? 1u : 0u);
                                    if (lead) tc_mma(d0 + op.npad, al, bw, idesc2, 1u);
                                    ah += row; al += row; bw += wstep;
                                }
                            }
                            if (tm) tm[256 + t * 4 + 1] = clock64();
                            if (lead) tc_commit(&lo_empty[s]);
                            if (tm) tm[256 + t * 4 + 2] = clock64();
                            if (lead) tc_commit(&acc_full[s]);
                            if (tm) tm[256 + t * 4 + 3] = clock64();
                        }
                    }
                    __syncwarp();
                } else {
                    long long* tm = (P.timing && blockIdx.x == 0 && k == P.dbg_op) ? P.timing + 160 * 32 : nullptr;
                    if (warp < 8) tc_conv_drain(op, sm, tmem, it, ntiles, f, acc_full, acc_empty, tm);
                    else switch (op.nchunks) {
                        case 1: tc_conv_stage<1>(op, sm, it, ntiles, lo_full, lo_empty, tm); break;
                        case 2: tc_conv_stage<2>(op, sm, it, ntiles, lo_full, lo_empty, tm); break;
                        case 3: tc_conv_stage<3>(op, sm, it, ntiles, lo_full, lo_empty, tm); break;
                        default: tc_conv_stage<4>(op, sm, it, ntiles, lo_full, lo_empty, tm);
                    }
                }
                it += ntiles;
            } else if (tid < kTcWorkers) {
                long long* tm = (P.timing && blockIdx.x == 0 && k == P.dbg_op && tid == 0) ? P.timing + 160 * 32 : nullptr;
                if (tm) tm[0] = clock64();
                const int ls = 31 - __clz(S);                               // S is a power of two: divisions become shifts
                if (op.kind == T_POOL) {
                    // 2x2 max-pool: S = output side
                    const int Si = 2 * S, quads = op.Cin / 4;
                    const int qsi = (Si + 2) * Si * 16, qso = (S + 2) * S * 16;
                    if (op.out >= 0) tc_zero_halo(sm + op.out, quads, S, tid, kTcWorkers);
                    if (tm) tm[1] = clock64();
                    for (int e = tid; e < quads * S * S; e += kTcWorkers) {
                        const int x = e & (S - 1), y = (e >> ls) & (S - 1), q = e >> (2 * ls);
                        const unsigned char* p = sm + op.src[0] + q * qsi + ((2 * y + 1) * Si + 2 * x) * 16;
                        const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 16);
                        const float4 c = *reinterpret_cast<const float4*>(p + Si * 16), d = *reinterpret_cast<const float4*>(p + Si * 16 + 16);
                        float4 mx;
                        mx.x = fmaxf(fmaxf(a.x, b.x), fmaxf(c.x, d.x));
                        mx.y = fmaxf(fmaxf(a.y, b.y), fmaxf(c.y, d.y));
                        mx.z = fmaxf(fmaxf(a.z, b.z), fmaxf(c.z, d.z));
                        mx.w = fmaxf(fmaxf(a.w, b.w), fmaxf(c.w, d.w));
                        if (op.out >= 0) *reinterpret_cast<float4*>(sm + op.out + q * qso + ((y + 1) * S + x) * 16) = mx;
                        if (op.gout) {
                            float* g = op.gout + (long)f * op.gout_bs + ((4 * q) << (2 * ls)) + (y << ls) + x;
                            g[0] = mx.x; g[S * S] = mx.y; g[2 * S * S] = mx.z; g[3 * S * S] = mx.w;
                        }
                    }
                } else if (op.kind == T_UP) {
                    // 2x bilinear upsample, align_corners = False (the arithmetic of conv.cu's upsample2_kernel: W pass, then H).
                    const int Si = S / 2, quads = op.Cin / 4;
                    const int qsi = (Si + 2) * Si * 16, qso = (S + 2) * S * 16;
                    if (op.out >= 0) tc_zero_halo(sm + op.out, quads, S, tid, kTcWorkers);
                    const int npx = S * S;
                    if (tm) tm[1] = clock64();
                    for (int e = tid; e < quads * npx; e += kTcWorkers) {
                        const int x = e & (S - 1), y = (e >> ls) & (S - 1), q = e >> (2 * ls);
                        const int ky = y >> 1, kx = x >> 1;
                        int ya, yb, xa, xb;
                        float wya, wyb, wxa, wxb;
                        if (y & 1) { ya = ky; yb = min(ky + 1, Si - 1); wya = 0.75f; wyb = 0.25f; }
                        else { ya = max(ky - 1, 0); yb = ky; wya = 0.25f; wyb = 0.75f; }
                        if (x & 1) { xa = kx; xb = min(kx + 1, Si - 1); wxa = 0.75f; wxb = 0.25f; }
                        else { xa = max(kx - 1, 0); xb = kx; wxa = 0.25f; wxb = 0.75f; }
                        const unsigned char* src = sm + op.src[0] + q * qsi;
                        const float4 aa = *reinterpret_cast<const float4*>(src + ((ya + 1) * Si + xa) * 16);
                        const float4 ab = *reinterpret_cast<const float4*>(src + ((ya + 1) * Si + xb) * 16);
                        const float4 ba = *reinterpret_cast<const float4*>(src + ((yb + 1) * Si + xa) * 16);
                        const float4 bb = *reinterpret_cast<const float4*>(src + ((yb + 1) * Si + xb) * 16);
                        float4 o;
                        o.x = wya * (wxa * aa.x + wxb * ab.x) + wyb * (wxa * ba.x + wxb * bb.x);
                        o.y = wya * (wxa * aa.y + wxb * ab.y) + wyb * (wxa * ba.y + wxb * bb.y);
                        o.z = wya * (wxa * aa.z + wxb * ab.z) + wyb * (wxa * ba.z + wxb * bb.z);
                        o.w = wya * (wxa * aa.w + wxb * ab.w) + wyb * (wxa * ba.w + wxb * bb.w);
                        if (op.out >= 0) *reinterpret_cast<float4*>(sm + op.out + q * qso + ((y + 1) * S + x) * 16) = o;
                        if (op.gout) {
                            float* g = op.gout + (long)f * op.gout_bs + ((4 * q) << (2 * ls)) + (y << ls) + x;
                            g[0] = o.x; g[npx] = o.y; g[2 * npx] = o.z; g[3 * npx] = o.w;
                        }
                    }
                } else if (op.kind == T_HEADT) {
                    // 1x1 head adjoint: d in[c] = sum_o w[o][c] * g[o], g = d logits gated by the head's own ReLU; the result is
                    // gated by the ReLU of the layer that produced `in`.  Cout = 8 channels of the head's input, Cin = logits.
                    const int qso = (S + 2) * S * 16, NO = op.Cin;
                    if (op.out >= 0) tc_zero_halo(sm + op.out, 2, S, tid, kTcWorkers);
                    float w[3][8];
#pragma unroll
                    for (int o = 0; o < 3; ++o)
#pragma unroll
                        for (int c = 0; c < 8; ++c) w[o][c] = o < NO ? __ldg(op.w1 + o * 8 + c) : 0.f;
                    // (a compact loop on purpose: this code runs once per frame, and straight-line code that does not fit the
                    // instruction cache is paid for line by line -- fully unrolled over its three passes this op took 30 k cycles)
#pragma unroll 1
                    for (int e = tid; e < S * S; e += kTcWorkers) {
                        float gl[3], hm[3], mk[8];
#pragma unroll
                        for (int o = 0; o < 3; ++o) {
                            gl[o] = o < NO ? tc_ldg_async(op.gsrc + (long)f * op.gsrc_bs + (long)o * S * S + e) : 0.f;
                            hm[o] = (o < NO && op.gmask2) ? tc_ldg_async(op.gmask2 + (long)f * op.gmask2_bs + (long)o * S * S + e) : 1.f;
                        }
#pragma unroll
                        for (int c = 0; c < 8; ++c) mk[c] = op.gmask ? tc_ldg_async(op.gmask + (long)f * op.gmask_bs + (long)c * S * S + e) : 1.f;
                        float d[8];
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            float a = 0.f;
#pragma unroll
                            for (int o = 0; o < 3; ++o)
                                if (o < NO && hm[o] > 0.f) a += w[o][c] * gl[o];
                            d[c] = mk[c] > 0.f ? a : 0.f;
                        }
                        if (op.out >= 0) {
                            *reinterpret_cast<float4*>(sm + op.out + (S + e) * 16) = make_float4(d[0], d[1], d[2], d[3]);
                            *reinterpret_cast<float4*>(sm + op.out + qso + (S + e) * 16) = make_float4(d[4], d[5], d[6], d[7]);
                        }
                        if (op.gout) {
#pragma unroll
                            for (int c = 0; c < 8; ++c) op.gout[(long)f * op.gout_bs + (long)c * S * S + e] = d[c];
                        }
                    }
                } else if (op.kind == T_UPT) {
                    // upsample adjoint (gather): input (i, j) of a 2x bilinear upsample collects outputs 2i-1..2i+2 x 2j-1..2j+2
                    // with weights (.25, .75, .75, .25), the clamped border taps folding back (unet_fused.cu run_upT, same
                    // arithmetic).  S = the low-resolution side; src[0] = the gradient of the upsampled tensor.
                    const int So = 2 * S, quads = op.Cin / 4;
                    const int qsi = (So + 2) * So * 16, qso = (S + 2) * S * 16;
                    if (op.out >= 0) tc_zero_halo(sm + op.out, quads, S, tid, kTcWorkers);
                    for (int e = tid; e < quads * S * S; e += kTcWorkers) {
                        const int j = e & (S - 1), i = (e >> ls) & (S - 1), q = e >> (2 * ls);
                        float wy[4], wx[4];
                        wy[0] = i > 0 ? 0.25f : 0.f; wy[1] = i > 0 ? 0.75f : 1.f; wy[2] = i < S - 1 ? 0.75f : 1.f; wy[3] = i < S - 1 ? 0.25f : 0.f;
                        wx[0] = j > 0 ? 0.25f : 0.f; wx[1] = j > 0 ? 0.75f : 1.f; wx[2] = j < S - 1 ? 0.75f : 1.f; wx[3] = j < S - 1 ? 0.25f : 0.f;
                        float mk[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            mk[u] = op.gmask ? tc_ldg_async(op.gmask + (long)f * op.gmask_bs + ((long)(4 * q + u) << (2 * ls)) + (i << ls) + j) : 1.f;
                        // output (2i-1+a, 2j-1+b): buffer row 2i+a (the padding rows carry weight 0 or zeros), pixel 2j-1+b
                        const unsigned char* g0 = sm + op.src[0] + q * qsi + (2 * i * So + 2 * j - 1) * 16;
                        float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                        for (int a = 0; a < 4; ++a) {
                            if (wy[a] == 0.f) continue;
                            float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                            for (int b = 0; b < 4; ++b) {
                                if (wx[b] == 0.f) continue;
                                const float4 g = *reinterpret_cast<const float4*>(g0 + (a * So + b) * 16);
                                r.x += wx[b] * g.x; r.y += wx[b] * g.y; r.z += wx[b] * g.z; r.w += wx[b] * g.w;
                            }
                            s4.x += wy[a] * r.x; s4.y += wy[a] * r.y; s4.z += wy[a] * r.z; s4.w += wy[a] * r.w;
                        }
                        if (!(mk[0] > 0.f)) s4.x = 0.f;
                        if (!(mk[1] > 0.f)) s4.y = 0.f;
                        if (!(mk[2] > 0.f)) s4.z = 0.f;
                        if (!(mk[3] > 0.f)) s4.w = 0.f;
                        if (op.out >= 0) *reinterpret_cast<float4*>(sm + op.out + q * qso + ((i + 1) * S + j) * 16) = s4;
                        if (op.gout) {
                            float* g = op.gout + (long)f * op.gout_bs + ((long)(4 * q) << (2 * ls)) + (i << ls) + j;
                            g[0] = s4.x; g[S * S] = s4.y; g[2 * S * S] = s4.z; g[3 * S * S] = s4.w;
                        }
                    }
                } else if (op.kind == T_POOLT) {
                    // max-pool adjoint: each 2x2 window of the source X routes the pooled gradient to its first maximum (row-major,
                    // as ATen), adds the gradient that reached X through its other consumer (in1, parked on chip) and applies X's
                    // own ReLU mask (unet_fused.cu run_poolT).  S = pooled side; gmask = X (global NCHW).
                    const int Si = 2 * S, quads = op.Cin / 4;
                    const int qsp = (S + 2) * S * 16, qsx = (Si + 2) * Si * 16;
                    if (op.out >= 0 && op.out != op.in1) tc_zero_halo(sm + op.out, quads, Si, tid, kTcWorkers);
                    for (int e = tid; e < quads * S * S; e += kTcWorkers) {
                        const int x = e & (S - 1), y = (e >> ls) & (S - 1), q = e >> (2 * ls);
                        float2 xa[4], xb[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const float* xs = op.gmask + (long)f * op.gmask_bs + ((long)(4 * q + u) * Si + 2 * y) * Si + 2 * x;
                            xa[u] = *reinterpret_cast<const float2*>(xs);
                            xb[u] = *reinterpret_cast<const float2*>(xs + Si);
                        }
                        const float4 g4 = *reinterpret_cast<const float4*>(sm + op.src[0] + q * qsp + ((y + 1) * S + x) * 16);
                        const float g[4] = {g4.x, g4.y, g4.z, g4.w};
                        const int t00 = q * qsx + ((2 * y + 1) * Si + 2 * x) * 16;
                        const int toff[4] = {t00, t00 + 16, t00 + Si * 16, t00 + Si * 16 + 16};
                        float r[4][4];                                              // [window position][channel]
#pragma unroll
                        for (int k2 = 0; k2 < 4; ++k2) {
                            float4 pk = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (op.in1 >= 0) pk = *reinterpret_cast<const float4*>(sm + op.in1 + toff[k2]);
                            r[k2][0] = pk.x; r[k2][1] = pk.y; r[k2][2] = pk.z; r[k2][3] = pk.w;
                        }
                        if (op.acc_gout) {
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const float* ps = op.gout + (long)f * op.gout_bs + ((long)(4 * q + u) * Si + 2 * y) * Si + 2 * x;
                                const float2 pa = *reinterpret_cast<const float2*>(ps), pb = *reinterpret_cast<const float2*>(ps + Si);
                                r[0][u] = pa.x; r[1][u] = pa.y; r[2][u] = pb.x; r[3][u] = pb.y;
                            }
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const float v[4] = {xa[u].x, xa[u].y, xb[u].x, xb[u].y};
                            int best = 0;
                            float m = v[0];
                            if (v[1] > m) { m = v[1]; best = 1; }
                            if (v[2] > m) { m = v[2]; best = 2; }
                            if (v[3] > m) { m = v[3]; best = 3; }
#pragma unroll
                            for (int k2 = 0; k2 < 4; ++k2) {
                                if (k2 == best) r[k2][u] += g[u];
                                if (op.relu && !(v[k2] > 0.f)) r[k2][u] = 0.f;
                            }
                        }
#pragma unroll
                        for (int k2 = 0; k2 < 4; ++k2)
                            if (op.out >= 0) *reinterpret_cast<float4*>(sm + op.out + toff[k2]) = make_float4(r[k2][0], r[k2][1], r[k2][2], r[k2][3]);
                        if (op.gout) {
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                float* go = op.gout + (long)f * op.gout_bs + ((long)(4 * q + u) * Si + 2 * y) * Si + 2 * x;
                                *reinterpret_cast<float2*>(go) = make_float2(r[0][u], r[1][u]);
                                *reinterpret_cast<float2*>(go + Si) = make_float2(r[2][u], r[3][u]);
                            }
                        }
                    }
                } else {
                    // 1x1 head: logits[o] = (relu)(b[o] + sum_c w[o][c] * in[c]); Cin <= 8, Cout <= 3 (planner)
                    const int qsi = (S + 2) * S * 16;
                    float w[3][8], bs[3];
#pragma unroll
                    for (int co = 0; co < 3; ++co) {
                        bs[co] = co < op.Cout ? __ldg(op.bias + co) : 0.f;
#pragma unroll
                        for (int c = 0; c < 8; ++c) w[co][c] = (co < op.Cout && c < op.Cin) ? __ldg(op.w1 + co * op.Cin + c) : 0.f;
                    }
                    for (int e = tid; e < S * S; e += kTcWorkers) {
                        const float4 v0 = *reinterpret_cast<const float4*>(sm + op.src[0] + (S + e) * 16);
                        const float4 v1 = op.Cin > 4 ? *reinterpret_cast<const float4*>(sm + op.src[0] + qsi + (S + e) * 16) : make_float4(0.f, 0.f, 0.f, 0.f);
                        const float in[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
                        for (int co = 0; co < 3; ++co) {
                            if (co >= op.Cout) break;
                            float s2 = bs[co];
#pragma unroll
                            for (int c = 0; c < 8; ++c) s2 += w[co][c] * in[c];
                            if (op.relu) s2 = fmaxf(s2, 0.f);
                            op.gout[(long)f * op.gout_bs + (long)co * S * S + e] = s2;
                        }
                    }
                }
            }
            if (P.timing && blockIdx.x == 0 && k == P.dbg_op && tid == 0 && op.kind != T_CONV) P.timing[160 * 32 + 2] = clock64();
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            if (P.timing && blockIdx.x == 0 && k == P.dbg_op && tid == 0 && op.kind != T_CONV) P.timing[160 * 32 + 3] = clock64();
            __syncthreads();
            if (P.timing && blockIdx.x == 0 && k == P.dbg_op && tid == 0 && op.kind != T_CONV) P.timing[160 * 32 + 4] = clock64();
            if (P.timing && tid == 0) P.timing[(long)blockIdx.x * 32 + 1 + k] = clock64();
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == kTcMmaWarp) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
    }
}

struct TcPack {
    int nlayers;
    const float* w[20];
    int Cin[20], Cout[20], nchunks[20], npad[20];
    int tr[20], ci0[20], cin_total[20];      // backward-data: W[k][ci0 + n'][2 - ky][2 - kx] of the layer's [Cout_l = K][cin_total][3][3]
    long off[20];
};
// dst[layer][kc][ky][quad][v][n][j] = split_v(W[co][ci = 8 kc + 4 quad + j][ky][kx]),  n = kx Cout + co  (zero beyond Cin / 3 Cout):
// per (chunk, ky) one operand block of 2 npad rows, the hi rows first, then the lo rows
__global__ void __launch_bounds__(256) unet_tc_pack_kernel(const TcPack K, float* __restrict__ dst) {
    const int l = blockIdx.y;
    const int Cin = K.Cin[l], Cout = K.Cout[l], npad = K.npad[l];
    const int total = K.nchunks[l] * 3 * 2 * 2 * npad * 4;
    const float* __restrict__ w = K.w[l];
    float* d = dst + K.off[l];
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int j = e & 3;
        int r = e >> 2;
        const int n2 = r % (2 * npad); r /= 2 * npad;
        const int quad = r & 1; r >>= 1;
        const int ky = r % 3;
        const int kc = r / 3;
        const int v = n2 >= npad, n = n2 - v * npad;
        const int ci = kc * 8 + quad * 4 + j, kx = n / Cout, co = n - kx * Cout;
        float val = 0.f;
        if (ci < Cin && n < 3 * Cout)
            val = K.tr[l] ? w[((size_t)ci * K.cin_total[l] + K.ci0[l] + co) * 9 + (2 - ky) * 3 + (2 - kx)] : w[((size_t)co * Cin + ci) * 9 + ky * 3 + kx];
        const float hi = __uint_as_float((__float_as_uint(val) + 0x1000u) & 0xffffe000u);
        const float lo = val - hi;
        d[e] = v ? __uint_as_float((__float_as_uint(lo) + 0x1000u) & 0xffffe000u) : hi;
    }
}

// Expected relative truncation loss (in units of 2^-23) of a main accumulator after `chain` MMAs: the tensor core adds the
// 8 products of an MMA and the fp32 accumulator with truncation.  Calibrated on the mean signed error of every layer's
// pre-activations against float64 (tools/unet_tc_check.py); PAIG_UNET_TC_KAPPA="k3,k6,k9,k12" overrides.
static float tc_kappa(int chain) {
    static float k[4] = {0.27f, 0.70f, 1.225f, 1.75f};
    static bool init = false;
    if (!init) {
        init = true;
        if (const char* e = getenv("PAIG_UNET_TC_KAPPA")) sscanf(e, "%f,%f,%f,%f", &k[0], &k[1], &k[2], &k[3]);
    }
    const int i = chain / 3 - 1;
    return k[i < 0 ? 0 : (i > 3 ? 3 : i)];
}

static int tc_sm_count() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

// PAIG_DEBUG: cycles per op (last frame of every CTA, mean over CTAs) and, for op PAIG_UNET_TC_DBGOP of CTA 0, the stamps of
// every role inside each tile
static long long* tc_timing_buffer() {
    static long long* tbuf = nullptr;
    if (!tbuf) cudaMalloc(&tbuf, (size_t)172 * 32 * sizeof(long long));
    return tbuf;
}
static void tc_print_timing(const char* what, const TcPlan& P, int grid, cudaStream_t st) {
    if (grid > 160) return;
    static long long host[172 * 32];
    cudaStreamSynchronize(st);
    cudaMemcpy(host, P.timing, sizeof(host), cudaMemcpyDeviceToHost);
    fprintf(stderr, "[paig] tcgen05 UNet %s cycles per op (last frame of each CTA, mean over %d CTAs):", what, grid);
    double total = 0;
    for (int k = 0; k < P.nops; ++k) {
        double sum = 0;
        for (int b = 0; b < grid; ++b) sum += (double)(host[b * 32 + 1 + k] - host[b * 32 + k]);
        fprintf(stderr, " op%d=%.0f", k, sum / grid);
        total += sum / grid;
    }
    fprintf(stderr, " total=%.0f\n", total);
    if (P.dbg_op < 0) return;
    const long long* tm = host + 160 * 32;
    fprintf(stderr, "[paig]   op%d of CTA 0, per tile (cycles since the op's first stamp): start | stage free | staged | arrived | acc full | drained || mma: start | lo full\n", P.dbg_op);
    for (int t2 = 0; t2 < 8; ++t2) {
        fprintf(stderr, "[paig]    tile %d:", t2);
        for (int j = 0; j < 8; ++j) fprintf(stderr, " %lld", tm[t2 * 8 + j] ? tm[t2 * 8 + j] - tm[0] : -1);
        fprintf(stderr, " | waits done, mmas issued, commit 1, commit 2:");
        for (int j = 0; j < 4; ++j) fprintf(stderr, " %lld", tm[256 + t2 * 4 + j] ? tm[256 + t2 * 4 + j] - tm[0] : -1);
        fprintf(stderr, "\n");
    }
}

// 0 ok, > 0 error, -1: not applicable (the caller runs unet_fused_forward)
int unet_tc_forward(const paig_task* t, const paig_params* p, const Layout& L, const float* x, long seq_stride, int fps,
                    float* ws, cudaStream_t st) {
    const char* sw = getenv("PAIG_UNET_TC");               // PAIG_UNET_TC=0: the FMA kernel (read per call: tests switch it inside one process)
    if ((sw && sw[0] == '0') || getenv("PAIG_NO_TCGEN05")) return -1;
    const UNetDesc& u = L.unet;
    const Dims& d = L.d;
    if (t->deep_unet || d.H != 32) return -1;              // sides 32 / 16 / 8: a tile is a whole number of image rows
    TcPlan P;
    memset(&P, 0, sizeof(P));
    TcPack K;
    memset(&K, 0, sizeof(K));

    struct Sl { int buf, c0, C, S, born, last, off, bytes; };
    Sl sl[32];
    int ns = 0;
    sl[ns++] = Sl{-1, 0, 8, d.H, -1, -1, 0, 0};
    int out_sl[kTcMaxOps], in_sl[kTcMaxOps][kTcMaxChunks], n_in[kTcMaxOps];
    int n = 0, nl = 0;
    long woff = 0;
    for (int i = 0; i < u.nops; ++i) {
        const Op& op = u.ops[i];
        if (n >= kTcMaxOps) return -1;
        TcOp& o = P.ops[n];
        o.out = -1;
        o.next_w = -1;
        const int Sout = op.kind == OP_HEAD ? d.H : (d.H >> u.bufs[op.out.buf].shift);
        o.S = Sout;
        o.relu = op.relu;
        // input slices, in channel order
        int found[kTcMaxChunks], nf = 0, covered = 0;
        if (op.in.buf == -1) { found[nf++] = 0; covered = op.in.C; }
        else {
            for (int c0 = op.in.c0; c0 < op.in.c0 + op.in.C;) {
                int hit = -1;
                for (int s = 1; s < ns; ++s)
                    if (sl[s].buf == op.in.buf && sl[s].c0 == c0) hit = s;      // the latest producer of that slice
                if (hit < 0 || nf >= kTcMaxChunks) return -1;
                found[nf++] = hit;
                c0 += sl[hit].C;
                covered += sl[hit].C;
            }
        }
        if (covered != op.in.C) return -1;
        n_in[n] = nf;
        for (int s = 0; s < nf; ++s) { in_sl[n][s] = found[s]; sl[found[s]].last = n; }
        int cin = 0;
        for (int s = 0; s < nf; ++s) cin += sl[found[s]].C;
        o.Cin = cin;
        o.Cout = op.out.C;
        switch (op.kind) {
            case OP_CONV: {
                o.kind = T_CONV;
                if (cin % 8 || cin / 8 > kTcMaxChunks) return -1;
                for (int s = 0; s < nf; ++s) if (sl[found[s]].C % 8) return -1;
                if (o.Cout != 8 && o.Cout != 16 && o.Cout != 32) return -1;
                if (Sout != 32 && Sout != 16 && Sout != 8) return -1;
                o.nchunks = cin / 8;
                o.npad = tc_pad16(3 * o.Cout);
                if (4 * o.npad > 512) return -1;                     // two accumulator sets of [main | corr]
                o.comp = tc_kappa(3 * o.nchunks) * 1.1920929e-7f;
                o.wbytes = o.nchunks * 3 * 2 * 2 * o.npad * 16;
                o.wglob = woff;
                o.bias = p->conv[op.layer].b;
                K.w[nl] = p->conv[op.layer].w; K.Cin[nl] = op.in.C; K.Cout[nl] = o.Cout; K.nchunks[nl] = o.nchunks; K.npad[nl] = o.npad;
                K.off[nl] = woff;
                ++nl;
                woff += (long)align64((size_t)o.wbytes / 4);
                break;
            }
            case OP_POOL: o.kind = T_POOL; if (nf != 1 || cin % 4) return -1; break;
            case OP_UP: o.kind = T_UP; if (nf != 1 || cin % 4) return -1; break;
            default:
                o.kind = T_HEAD;
                if (nf != 1 || cin > 8 || cin % 4 || o.Cout > 3) return -1;
                o.bias = p->conv[op.layer].b;
                o.w1 = p->conv[op.layer].w;
        }
        if (op.kind == OP_HEAD) {
            o.gout = ws + L.logits;
            o.gout_bs = (long)d.n * d.HW;
            out_sl[n] = -1;
        } else {
            const BufDesc& bd = u.bufs[op.out.buf];
            static const bool nostore = getenv("PAIG_UNET_TC_NOSTORE") != nullptr;      // timing experiment
            o.gout = ((t->flags & PAIG_FLAG_INFERENCE) || nostore) ? nullptr : ws + L.act[op.out.buf] + (long)op.out.c0 * Sout * Sout;
            o.gout_bs = (long)bd.C * Sout * Sout;
            if (ns >= 32) return -1;
            sl[ns] = Sl{op.out.buf, op.out.c0, op.out.C, Sout, n, -1, 0, 0};
            out_sl[n] = ns++;
        }
        ++n;
    }
    P.nops = n;
    K.nlayers = nl;
    if ((size_t)woff > unet_tc_wpack_fwd_floats(u)) { set_error("unet_tc: packed weights exceed their workspace region"); return 1; }
    // weight prefetch chain: conv k's weights are fetched while the previous conv runs
    int issue_at[kTcMaxOps];
    {
        int prev = -1, widx = 0;
        P.first_w = -1;
        for (int k = 0; k < n; ++k) {
            issue_at[k] = -1;
            if (P.ops[k].kind != T_CONV) continue;
            P.ops[k].wbar = widx++ & 1;
            issue_at[k] = prev;
            if (prev < 0) P.first_w = k; else P.ops[prev].next_w = k;
            prev = k;
        }
    }
    // shared-memory offsets by lifetime (Planner counts 4-byte units).  The lo staging ring is two slots deep and a layer's
    // weights arrive while the previous layer runs; when the frame does not fit, the widest full-resolution layer gives up its
    // prefetch, then the wide full-resolution layers (then every layer) fall back to one staging slot.
    auto buf_bytes = [](int C, int S) { return (C / 4) * (S + 2) * S * 16 + (S == 8 ? 1024 : 0); };   // S = 8: an M = 128 tile reads 8 rows past the frame
    int stage_off[kTcMaxOps], w_off[kTcMaxOps];
    size_t peak = 0;
    for (int attempt = 0; attempt < 4; ++attempt) {
        Planner al;
        sl[0].bytes = buf_bytes(8, d.H);
        al.add(sl[0].bytes / 4, -1, sl[0].last, &sl[0].off);
        for (int k = 0; k < n; ++k) {
            TcOp& o = P.ops[k];
            if (o.kind == T_CONV) {
                o.late_w = attempt >= 1 && o.S == d.H && o.nchunks >= 3;
                al.add(o.wbytes / 4, o.late_w ? k : issue_at[k], k, &w_off[k]);
                const int R = 128 / o.S;
                o.stage_q = (R + 2) * o.S * 16;
                o.stage_slots = (attempt == 3 || (attempt == 2 && o.S == d.H && o.nchunks >= 2)) ? 1 : 2;
                al.add(o.stage_slots * o.nchunks * 2 * o.stage_q / 4, k, k, &stage_off[k]);
            }
            if (out_sl[k] >= 0 && sl[out_sl[k]].last >= 0) {
                Sl& s2 = sl[out_sl[k]];
                s2.bytes = buf_bytes(s2.C, s2.S);
                al.add(s2.bytes / 4, k, s2.last, &s2.off);
            }
        }
        peak = (size_t)al.place_best() * 4;
        if (getenv("PAIG_DEBUG")) fprintf(stderr, "[paig] tcgen05 UNet forward: plan attempt %d needs %zu B\n", attempt, peak);
        if (peak <= kTcSmemLimit) break;
    }
    P.x_off = sl[0].off * 4;
    for (int k = 0; k < n; ++k) {
        TcOp& o = P.ops[k];
        if (out_sl[k] >= 0 && sl[out_sl[k]].last >= 0) o.out = sl[out_sl[k]].off * 4;
        const int qs = (o.kind == T_POOL ? (2 * o.S + 2) * 2 * o.S : (o.kind == T_UP ? (o.S / 2 + 2) * (o.S / 2) : (o.S + 2) * o.S)) * 16;
        int c = 0;
        for (int s = 0; s < n_in[k]; ++s) {
            const Sl& in = sl[in_sl[k][s]];
            if (o.kind == T_CONV) {
                for (int q8 = 0; q8 < in.C / 8; ++q8) o.src[c++] = in.off * 4 + 2 * q8 * qs;
            } else o.src[0] = in.off * 4;
        }
        if (o.kind == T_CONV) { o.w = w_off[k] * 4; o.stage = stage_off[k] * 4; }
    }
    static const bool debug = getenv("PAIG_DEBUG") != nullptr;
    if (debug) {
        fprintf(stderr, "[paig] tcgen05 UNet forward plan: H=%d ops=%d smem=%zu B\n", d.H, n, peak);
        for (int k = 0; k < n; ++k) {
            const TcOp& o = P.ops[k];
            fprintf(stderr, "[paig]   op%-2d kind=%d S=%-2d Cin=%-2d Cout=%-2d relu=%d chunks=%d npad=%d src=%d,%d,%d,%d out=%d stage=%d(+%d x%d) w=%d(%d) bar=%d next=%d late=%d\n",
                    k, o.kind, o.S, o.Cin, o.Cout, o.relu, o.nchunks, o.npad, o.src[0], o.src[1], o.src[2], o.src[3], o.out, o.stage,
                    o.stage_q, o.stage_slots, o.w, o.wbytes, o.wbar, o.next_w, o.late_w);
        }
    }
    if (peak > kTcSmemLimit) {
        if (debug) fprintf(stderr, "[paig] tcgen05 UNet forward: %zu B of shared memory needed, not taken\n", peak);
        return -1;
    }
    P.N = L.N; P.fps = fps; P.H = d.H; P.seq_stride = seq_stride; P.x = x;
    float* wpack = ws + L.wpack_tc;
    P.wpack = wpack;
    launch(unet_tc_pack_kernel, dim3(4, K.nlayers), dim3(256), 0, st, K, wpack);
    int rc = check_launch("pack_weights");
    if (rc) return rc;
    const int grid = L.N < tc_sm_count() ? L.N : tc_sm_count();
    P.timing = debug ? tc_timing_buffer() : nullptr;
    P.dbg_op = getenv("PAIG_UNET_TC_DBGOP") ? atoi(getenv("PAIG_UNET_TC_DBGOP")) : 0;
    launch(unet_tc_fwd_kernel, dim3(grid), dim3(kTcThreads), (size_t)((peak + 1023) & ~(size_t)1023), st, P);
    rc = check_launch("unet_tc_fwd");
    if (debug && !rc) tc_print_timing("forward", P, grid, st);
    return rc;
}
// Backward-data pass of the ShallowUNet on the tensor cores: the op list of unet_fused.cu's planner (transposed convs with the
// ReLU gate in the epilogue, max-pool / upsample / head adjoints, skip gradients parked on chip) run by the same kernel in
// its own layout.  0 ok, > 0 error, -1: not applicable (the caller runs unet_fused_backward).
int unet_tc_backward(const paig_task* t, const paig_params* p, const Layout& L, float* ws, cudaStream_t st) {
    const char* sw = getenv("PAIG_UNET_TC_BWD");           // PAIG_UNET_TC_BWD=0: the FMA kernel (read per call, like PAIG_UNET_TC)
    if ((sw && sw[0] == '0') || getenv("PAIG_NO_TCGEN05")) return -1;
    const UNetDesc& u = L.unet;
    const Dims& d = L.d;
    if (t->deep_unet || d.H != 32) return -1;
    static BwdOps B;                                       // (large: op table + pack plan)
    memset(&B, 0, sizeof(B));
    // skip-connection gradients wait for the max-pool adjoint in the workspace gradient buffer (L2), not on chip: 53 KB less
    // shared memory, which is what lets every layer keep two staging slots
    if (unet_backward_ops(t, p, L, ws, true, &B) != 0) return -1;
    const int n = B.P.nops;
    if (n > kTcMaxOps) return -1;
    TcPlan P;
    memset(&P, 0, sizeof(P));
    TcPack K;
    memset(&K, 0, sizeof(K));
    long woff = (long)unet_tc_wpack_fwd_floats(u);
    int nl = 0, kw = 0;                                    // kw: index into B.K (ops that have weights, in op order)
    for (int k = 0; k < n; ++k) {
        const FusedOp& fo = B.P.ops[k];
        TcOp& o = P.ops[k];
        o.out = o.in1 = -1;
        o.next_w = -1;
        o.S = fo.S;
        o.relu = fo.relu;
        o.gout = fo.gout; o.gout_bs = fo.gout_bs;
        o.gmask = fo.gmask; o.gmask_bs = fo.gmask_bs;
        o.acc_gout = fo.acc_gout;
        switch (fo.kind) {
            case F_CONV: {
                o.kind = T_CONV;
                o.Cin = fo.Cin0; o.Cout = fo.Cout;
                if (o.Cin % 8 || o.Cin / 8 > kTcMaxChunks || (o.Cout != 8 && o.Cout != 16 && o.Cout != 32)) return -1;
                o.nchunks = o.Cin / 8;
                o.npad = tc_pad16(3 * o.Cout);
                o.comp = tc_kappa(3 * o.nchunks) * 1.1920929e-7f;
                o.wbytes = o.nchunks * 3 * 2 * 2 * o.npad * 16;
                o.wglob = woff;
                if (nl >= 20 || B.K.mode[kw] != 1) return -1;
                K.w[nl] = B.K.w[kw]; K.Cin[nl] = o.Cin; K.Cout[nl] = o.Cout; K.nchunks[nl] = o.nchunks; K.npad[nl] = o.npad;
                K.tr[nl] = 1; K.ci0[nl] = B.K.ci0[kw]; K.cin_total[nl] = B.K.cin_total[kw]; K.off[nl] = woff;
                ++nl; ++kw;
                woff += (long)align64((size_t)o.wbytes / 4);
                break;
            }
            case F_HEADT:
                o.kind = T_HEADT;
                o.Cin = fo.Cin0; o.Cout = fo.Cout;         // logits, channels of the head's input
                if (o.Cout != 8 || o.Cin > 3 || B.K.mode[kw] != 2) return -1;
                o.w1 = B.K.w[kw]; ++kw;
                o.gsrc = fo.gsrc; o.gsrc_bs = fo.gsrc_bs; o.gmask2 = fo.gmask2; o.gmask2_bs = fo.gmask2_bs;
                break;
            case F_UPT: o.kind = T_UPT; o.Cin = o.Cout = fo.Cin0; if (o.Cin % 4) return -1; break;
            case F_POOLT: o.kind = T_POOLT; o.Cin = o.Cout = fo.Cin0; if (o.Cin % 4 || !fo.gmask) return -1; break;
            default: return -1;
        }
    }
    P.nops = n;
    K.nlayers = nl;
    if ((size_t)woff > unet_tc_wpack_floats(u)) { set_error("unet_tc backward: packed weights exceed their workspace region"); return 1; }
    int issue_at[kTcMaxOps];
    {
        int prev = -1, widx = 0;
        P.first_w = -1;
        for (int k = 0; k < n; ++k) {
            issue_at[k] = -1;
            if (P.ops[k].kind != T_CONV) continue;
            P.ops[k].wbar = widx++ & 1;
            issue_at[k] = prev;
            if (prev < 0) P.first_w = k; else P.ops[prev].next_w = k;
            prev = k;
        }
    }
    auto buf_bytes = [](int C, int S) { return (C / 4) * (S + 2) * S * 16 + (S == 8 ? 1024 : 0); };
    int stage_off[kTcMaxOps], w_off[kTcMaxOps], sl_off[40];
    size_t peak = 0;
    for (int attempt = 0; attempt < 4; ++attempt) {
        Planner al;
        for (int k = 0; k < n; ++k) {
            TcOp& o = P.ops[k];
            if (o.kind != T_CONV) continue;
            o.late_w = attempt >= 1 && o.S == d.H && o.nchunks >= 2;
            al.add(o.wbytes / 4, o.late_w ? k : issue_at[k], k, &w_off[k]);
            o.stage_q = (128 / o.S + 2) * o.S * 16;
            o.stage_slots = (attempt == 3 || (attempt == 2 && o.S == d.H)) ? 1 : 2;
            al.add(o.stage_slots * o.nchunks * 2 * o.stage_q / 4, k, k, &stage_off[k]);
        }
        for (int g = 0; g < B.nslices; ++g) {
            sl_off[g] = -1;
            if (B.born[g] < 0 || B.last[g] < 0) continue;
            al.add(buf_bytes(B.C[g], B.S[g]) / 4, B.born[g], B.last[g], &sl_off[g]);
        }
        peak = (size_t)al.place_best() * 4;
        if (getenv("PAIG_DEBUG")) fprintf(stderr, "[paig] tcgen05 UNet backward: plan attempt %d needs %zu B\n", attempt, peak);
        if (peak <= kTcSmemLimit) break;
    }
    for (int k = 0; k < n; ++k) {
        TcOp& o = P.ops[k];
        const int a = B.in0_of[k], b = B.in1_of[k], c = B.out_of[k];
        if (c >= 0 && B.last[c] >= 0 && sl_off[c] >= 0) o.out = sl_off[c] * 4;
        if (a >= 0) {
            if (sl_off[a] < 0) return -1;
            if (o.kind == T_CONV) {
                const int qs = (o.S + 2) * o.S * 16;
                for (int q8 = 0; q8 < o.nchunks; ++q8) o.src[q8] = sl_off[a] * 4 + 2 * q8 * qs;
            } else o.src[0] = sl_off[a] * 4;
        } else if (o.kind != T_HEADT) return -1;
        if (b >= 0) { if (sl_off[b] < 0) return -1; o.in1 = sl_off[b] * 4; }
        if (o.kind == T_CONV) { o.w = w_off[k] * 4; o.stage = stage_off[k] * 4; }
    }
    static const bool debug = getenv("PAIG_DEBUG") != nullptr;
    if (debug) {
        fprintf(stderr, "[paig] tcgen05 UNet backward plan: H=%d ops=%d smem=%zu B\n", d.H, n, peak);
        for (int k = 0; k < n; ++k) {
            const TcOp& o = P.ops[k];
            fprintf(stderr, "[paig]   op%-2d kind=%d S=%-2d Cin=%-2d Cout=%-2d relu=%d chunks=%d npad=%d src=%d in1=%d out=%d mask=%d gout=%d stage x%d late=%d\n",
                    k, o.kind, o.S, o.Cin, o.Cout, o.relu, o.nchunks, o.npad, o.src[0], o.in1, o.out, o.gmask != nullptr, o.gout != nullptr,
                    o.stage_slots, o.late_w);
        }
    }
    if (peak > kTcSmemLimit) return -1;
    P.N = L.N; P.fps = 1; P.H = d.H; P.x = nullptr;
    float* wpack = ws + L.wpack_tc;
    P.wpack = wpack;
    launch(unet_tc_pack_kernel, dim3(4, K.nlayers), dim3(256), 0, st, K, wpack);
    int rc = check_launch("pack_weights");
    if (rc) return rc;
    const int grid = L.N < tc_sm_count() ? L.N : tc_sm_count();
    P.timing = debug ? tc_timing_buffer() : nullptr;
    P.dbg_op = -1;
    launch(unet_tc_fwd_kernel, dim3(grid), dim3(kTcThreads), (size_t)((peak + 1023) & ~(size_t)1023), st, P);
    rc = check_launch("unet_tc_bwd");
    if (debug && !rc) tc_print_timing("backward-data", P, grid, st);
    return rc;
}
#endif

}  // namespace paig
