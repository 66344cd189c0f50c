// Fused spatial-transformer decoder: affine grid + bilinear sampling of the per-object templates and
// contents + softmax alpha-compositing over the background, for every object of every frame, with the
// reconstruction / prediction loss and the whole decoder backward optionally done in the same pass.
//
// Replaces PhysicsNet.conv_st_decoder (physics_models.py:151-199) + stn (stn.py:5-16) and the squared
// error sums of compute_loss (physics_models.py:122-131).
//
// Per output pixel (i,j) of frame f with object location (lx,ly) = loc[f, 2o:2o+2]   (SURVEY 8 a10):
//     theta2 = (H/2 - lx) / t  (fp32, physics_models.py:177)     xs = (2j+1)/H - 1 + theta2  (fp64 grid, Q5)
//     ix = ((float)xs + 1) * t/2 - 0.5   (grid_sample, align_corners=False)          likewise iy from ly, i
//     L_o   = bil0(template_o + 5; iy,ix) - 5        c_o = bil0(sigmoid(content_o); iy,ix)      (zeros padding)
//     w     = softmax([L_0..L_{n-1}, 1])             out = sum_o w_o c_o + w_n sigmoid(background)[i,j]
// Object layers first, background last (the reference's stack order).
//
// Data movement: the constants [template+5 | sigmoid(contents) | sigmoid(background)] (20 KB at 32 px) are
// staged once per CTA into shared memory with a TMA bulk copy; a CTA then walks frames f = blockIdx.x,
// +gridDim.x, ...  Each thread owns quads of 4 horizontally adjacent pixels: frame reads / writes are
// 16-byte vectors, coalesced along rows.  Gradients of the constants accumulate in shared memory across
// all frames of the CTA (background: exclusive owner, plain add; templates/contents: shared-memory
// atomics) and leave as one partial per CTA, reduced by decode_reduce_kernel in fixed order.
#include "common.cuh"
#include "internal.h"

#include <cstdlib>

namespace paig {

constexpr int kDecThreads = 256;

struct Tap {          // bilinear footprint along one axis
    int i0;           // floor(coordinate)
    float w1;         // weight of tap i0+1 ; tap i0 gets (i0+1) - coordinate
    float w0;
};

// fp64 grid exactly as F.affine_grid builds it for theta = [1,0,th; 0,1,tv], then the fp32 unnormalise of
// F.grid_sample(align_corners=False).
__device__ __forceinline__ Tap make_tap(int j, float l, int H, int t) {
    const float th = __fdiv_rn(__fsub_rn((float)H * 0.5f, l), (float)t);
    const double xt = (double)(2 * j + 1) / (double)H - 1.0;
    const float xs = (float)(xt + (double)th);
    const float ix = __fsub_rn(__fmul_rn(__fadd_rn(xs, 1.f), (float)t * 0.5f), 0.5f);
    const float fl = floorf(ix);
    Tap tp;
    tp.i0 = (int)fl;
    tp.w1 = ix - fl;
    tp.w0 = (fl + 1.f) - ix;
    return tp;
}

// GATHER (backward only): instead of scattering every pixel's template / content gradient through shared-memory float
// atomics (order-dependent, and heavily conflicting: a 2x zoom sends ~16 pixels to each texel), the per-pixel upstream
// values are parked in shared memory and every texel *gathers* its <= 4 x 4 contributors in two separable passes
// (along x, then y) in a fixed order: deterministic, and the atomics leave the critical path.  Needs
// NOBJ*4*(H*H + H*H/2) extra floats; larger frames (64 px) keep the atomic path.
template <int NOBJ, bool BWD, bool GATHER>
__global__ void __launch_bounds__(kDecThreads)
decode_kernel(int H, int band_rows, const float* __restrict__ consts, DecSeg segA, DecSeg segB,
              float* __restrict__ partials) {
    const int t = H / 2, tt = t * t, HW = H * H;
    const int CN = NOBJ * tt * 4 + 3 * HW;           // floats in the constants block
    const int tid = threadIdx.x;
    PAIG_DYN_SMEM(float, smem);
    float* sT5 = smem;                                // [NOBJ][t][t]
    float* sSC = sT5 + NOBJ * tt;                     // [NOBJ][3][t][t]
    float* sSB = sSC + NOBJ * 3 * tt;                 // [3][H][H]
    float* sG = smem + CN;                            // gradient accumulators, same layout (BWD only)
    float* sTab = BWD ? sG + CN : smem + CN;          // per-frame tap tables
    int* sX0 = reinterpret_cast<int*>(sTab);          // [NOBJ][H]
    float* sXw1 = sTab + NOBJ * H;
    float* sXw0 = sXw1 + NOBJ * H;
    int* sY0 = reinterpret_cast<int*>(sXw0 + NOBJ * H);
    float* sYw1 = reinterpret_cast<float*>(sY0) + NOBJ * H;
    float* sYw0 = sYw1 + NOBJ * H;
    float* sRed = sYw0 + NOBJ * H;                    // [8 warps][2*NOBJ+1]
    // GATHER works on bands of `band_rows` image rows (the whole frame at 32 / 36 px, 16 rows at 64 px where a full
    // frame of per-pixel upstream values does not fit next to the constants): bands are walked in order, so the
    // texel sums keep one fixed order
    const int BH = band_rows * H;
    float* sD = sRed + 8 * (2 * NOBJ + 1);            // GATHER: [NOBJ*4][band][H] per-pixel dL | dC0..2
    float* sE = sD + (GATHER ? NOBJ * 4 * BH : 0);    // GATHER: [NOBJ*4][band][t] after the x pass
    float* sInv = sE + (GATHER ? NOBJ * 4 * band_rows * t : 0);   // GATHER: [2 axes][NOBJ][t][8]: first contributor, 6 weights
    __shared__ __align__(8) unsigned long long bar;

    stage_bulk(smem, consts, (unsigned)(CN * sizeof(float)), &bar, 0);
    if (BWD) {
        for (int i = tid; i < CN; i += kDecThreads) sG[i] = 0.f;
    }
    __syncthreads();

    const int F = segA.nframes + segB.nframes;
    const int quads_per_row = H / 4;
    for (int f = blockIdx.x; f < F; f += gridDim.x) {
        const bool inA = f < segA.nframes;
        const DecSeg& sg = inA ? segA : segB;
        const int fl = inA ? f : f - segA.nframes;
        const int q = fl / sg.fps, r = fl % sg.fps;
        const float* loc = sg.loc + (long)q * sg.loc_seq_stride + (long)r * sg.loc_row_stride;
        const float* tgt = sg.target ? sg.target + (long)q * sg.tgt_seq_stride + (long)r * 3 * HW : nullptr;
        float* frame = sg.frames ? sg.frames + (long)fl * 3 * HW : nullptr;
        const float* dfr = sg.dframes ? sg.dframes + (long)fl * 3 * HW : nullptr;
        const bool has_scale = sg.scale != nullptr || sg.use_scale != 0;
        const float srow = sg.scale ? sg.scale[r] : (r < sg.scale_split ? sg.scale_lo : sg.scale_hi);
        const float gscale = (BWD && has_scale) ? 2.f * srow : 0.f;
        // a frame needs the backward math only if some gradient reaches it
        const bool do_bwd = BWD && (dfr != nullptr || (has_scale && gscale != 0.f));

        for (int k = tid; k < 2 * NOBJ * H; k += kDecThreads) {
            const int axis = k / (NOBJ * H), o = (k / H) % NOBJ, p = k % H;
            Tap tp = make_tap(p, loc[2 * o + axis], H, t);
            if (axis == 0) {
                sX0[o * H + p] = tp.i0; sXw1[o * H + p] = tp.w1; sXw0[o * H + p] = tp.w0;
            } else {
                sY0[o * H + p] = tp.i0; sYw1[o * H + p] = tp.w1; sYw0[o * H + p] = tp.w0;
            }
        }
        __syncthreads();

        float acc[2 * NOBJ + 1];                      // d lx, d ly per object; squared error
#pragma unroll
        for (int k = 0; k < 2 * NOBJ + 1; ++k) acc[k] = 0.f;

        if (GATHER && BWD && do_bwd) {
            // inverse tap tables: texel x of object o hears from columns j with x0(j) == x (weight w0) or x0(j) + 1 == x
            // (weight w1).  x0 is non-decreasing and grows by one every two columns, so the contributors are <= 6
            // consecutive columns starting at the first j with x0(j) >= x - 1.
            for (int e = tid; e < 2 * NOBJ * t; e += kDecThreads) {
                const int x = e % t, o = (e / t) % NOBJ, axis = e / (t * NOBJ);
                const int* i0 = (axis ? sY0 : sX0) + o * H;
                const float* w0 = (axis ? sYw0 : sXw0) + o * H;
                const float* w1 = (axis ? sYw1 : sXw1) + o * H;
                int j = 2 * (x - 1 - i0[0]) - 1;                       // analytic guess, then settle on the exact first column
                j = min(max(j, 0), H - 1);
                while (j > 0 && i0[j - 1] >= x - 1) --j;
                while (j < H && i0[j] < x - 1) ++j;
                float* dst = sInv + (size_t)e * 8;
                reinterpret_cast<int*>(dst)[0] = j;
#pragma unroll
                for (int k = 0; k < 6; ++k) {
                    const int jj = j + k;
                    float w = 0.f;
                    if (jj < H) w = i0[jj] == x ? w0[jj] : (i0[jj] + 1 == x ? w1[jj] : 0.f);
                    dst[1 + k] = w;
                }
            }
        }

        for (int band0 = 0; band0 < H; band0 += band_rows) {
        const int brows = min(band_rows, H - band0);
        for (int qd = tid; qd < brows * quads_per_row; qd += kDecThreads) {
            const int i = band0 + qd / quads_per_row, j0 = (qd % quads_per_row) * 4;
            float out[3][4];
            float wgt[4][NOBJ + 1];
            float col[4][NOBJ][3];
#pragma unroll
            for (int px = 0; px < 4; ++px) {
                const int j = j0 + px;
                float logit[NOBJ];
#pragma unroll
                for (int o = 0; o < NOBJ; ++o) {
                    const int x0 = sX0[o * H + j], y0 = sY0[o * H + i];
                    const float wx1 = sXw1[o * H + j], wx0 = sXw0[o * H + j];
                    const float wy1 = sYw1[o * H + i], wy0 = sYw0[o * H + i];
                    const bool vx0 = (unsigned)x0 < (unsigned)t, vx1 = (unsigned)(x0 + 1) < (unsigned)t;
                    const bool vy0 = (unsigned)y0 < (unsigned)t, vy1 = (unsigned)(y0 + 1) < (unsigned)t;
                    const float w00 = (vy0 && vx0) ? wy0 * wx0 : 0.f, w01 = (vy0 && vx1) ? wy0 * wx1 : 0.f;
                    const float w10 = (vy1 && vx0) ? wy1 * wx0 : 0.f, w11 = (vy1 && vx1) ? wy1 * wx1 : 0.f;
                    // clamp addresses so masked taps read something valid (their weight is zero)
                    const int xa = min(max(x0, 0), t - 1), xb = min(max(x0 + 1, 0), t - 1);
                    const int ya = min(max(y0, 0), t - 1), yb = min(max(y0 + 1, 0), t - 1);
                    const float* T = sT5 + o * tt;
                    logit[o] = (T[ya * t + xa] * w00 + T[ya * t + xb] * w01 + T[yb * t + xa] * w10 + T[yb * t + xb] * w11) - 5.f;
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const float* C = sSC + (o * 3 + c) * tt;
                        col[px][o][c] = C[ya * t + xa] * w00 + C[ya * t + xb] * w01 + C[yb * t + xa] * w10 + C[yb * t + xb] * w11;
                    }
                }
                float m = 1.f;                            // background logit is the constant 1
#pragma unroll
                for (int o = 0; o < NOBJ; ++o) m = fmaxf(m, logit[o]);
                float den = 0.f;
#pragma unroll
                for (int o = 0; o < NOBJ; ++o) {
                    wgt[px][o] = expf(logit[o] - m);
                    den += wgt[px][o];
                }
                wgt[px][NOBJ] = expf(1.f - m);
                den += wgt[px][NOBJ];
                const float inv = 1.f / den;
#pragma unroll
                for (int o = 0; o <= NOBJ; ++o) wgt[px][o] *= inv;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    float s = 0.f;
#pragma unroll
                    for (int o = 0; o < NOBJ; ++o) s += wgt[px][o] * col[px][o][c];
                    out[c][px] = s + wgt[px][NOBJ] * sSB[c * HW + i * H + j];
                }
            }
            if (frame) {
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    *reinterpret_cast<float4*>(frame + c * HW + i * H + j0) =
                        make_float4(out[c][0], out[c][1], out[c][2], out[c][3]);
            }
            if (!BWD && sg.layer_c) {
                // physics_models.py:190,196: per-layer sampled contents and softmax masks of this decoder call,
                // [n+1][F][3][H][H] each (object layers, then the background; masks repeated over the 3 channels)
#pragma unroll
                for (int o = 0; o <= NOBJ; ++o) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const long at = (long)o * sg.layer_stride + (long)fl * 3 * HW + c * HW + i * H + j0;
                        float4 cv;
                        if (o < NOBJ) cv = make_float4(col[0][o < NOBJ ? o : 0][c], col[1][o < NOBJ ? o : 0][c],
                                                       col[2][o < NOBJ ? o : 0][c], col[3][o < NOBJ ? o : 0][c]);
                        else cv = *reinterpret_cast<const float4*>(sSB + c * HW + i * H + j0);
                        *reinterpret_cast<float4*>(sg.layer_c + at) = cv;
                        *reinterpret_cast<float4*>(sg.layer_m + at) = make_float4(wgt[0][o], wgt[1][o], wgt[2][o], wgt[3][o]);
                    }
                }
            }
            float G[3][4];
            if (tgt) {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float4 tv = *reinterpret_cast<const float4*>(tgt + c * HW + i * H + j0);
                    const float d0 = out[c][0] - tv.x, d1 = out[c][1] - tv.y, d2 = out[c][2] - tv.z, d3 = out[c][3] - tv.w;
                    acc[2 * NOBJ] += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
                    G[c][0] = gscale * d0; G[c][1] = gscale * d1; G[c][2] = gscale * d2; G[c][3] = gscale * d3;
                }
            }
            if (BWD && do_bwd) {
                if (dfr) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const float4 gv = *reinterpret_cast<const float4*>(dfr + c * HW + i * H + j0);
                        G[c][0] = gv.x; G[c][1] = gv.y; G[c][2] = gv.z; G[c][3] = gv.w;
                    }
                }
#pragma unroll
                for (int px = 0; px < 4; ++px) {
                    const int j = j0 + px;
                    // composite + softmax backward
                    float dw[NOBJ + 1];
                    float dot = 0.f;
#pragma unroll
                    for (int o = 0; o < NOBJ; ++o) {
                        dw[o] = G[0][px] * col[px][o][0] + G[1][px] * col[px][o][1] + G[2][px] * col[px][o][2];
                        dot += wgt[px][o] * dw[o];
                    }
                    dw[NOBJ] = G[0][px] * sSB[i * H + j] + G[1][px] * sSB[HW + i * H + j] + G[2][px] * sSB[2 * HW + i * H + j];
                    dot += wgt[px][NOBJ] * dw[NOBJ];
                    float* gSB = sG + NOBJ * tt * 4;
#pragma unroll
                    for (int c = 0; c < 3; ++c) gSB[c * HW + i * H + j] += G[c][px] * wgt[px][NOBJ];   // exclusive owner
#pragma unroll
                    for (int o = 0; o < NOBJ; ++o) {
                        const float dL = wgt[px][o] * (dw[o] - dot);
                        float dC[3];
#pragma unroll
                        for (int c = 0; c < 3; ++c) dC[c] = G[c][px] * wgt[px][o];
                        const int x0 = sX0[o * H + j], y0 = sY0[o * H + i];
                        const float wx1 = sXw1[o * H + j], wx0 = sXw0[o * H + j];
                        const float wy1 = sYw1[o * H + i], wy0 = sYw0[o * H + i];
                        if (GATHER) {
                            sD[(o * 4 + 0) * BH + (i - band0) * H + j] = dL;
#pragma unroll
                            for (int c = 0; c < 3; ++c) sD[(o * 4 + 1 + c) * BH + (i - band0) * H + j] = dC[c];
                        }
                        float gix = 0.f, giy = 0.f;
#pragma unroll
                        for (int dy = 0; dy < 2; ++dy) {
#pragma unroll
                            for (int dx = 0; dx < 2; ++dx) {
                                const int y = y0 + dy, x = x0 + dx;
                                if ((unsigned)y < (unsigned)t && (unsigned)x < (unsigned)t) {
                                    const float wy = dy ? wy1 : wy0, wx = dx ? wx1 : wx0;
                                    const float w = wy * wx;
                                    const int a = y * t + x;
                                    // value-weighted upstream of this tap over the 1 mask + 3 content channels
                                    float vg = sT5[o * tt + a] * dL;
                                    if (!GATHER) atomicAdd(sG + o * tt + a, w * dL);
#pragma unroll
                                    for (int c = 0; c < 3; ++c) {
                                        vg += sSC[(o * 3 + c) * tt + a] * dC[c];
                                        if (!GATHER) atomicAdd(sG + NOBJ * tt + (o * 3 + c) * tt + a, w * dC[c]);
                                    }
                                    // d(bilinear)/d ix = sum_taps value * (+-1 along x) * wy ; same for iy
                                    gix += vg * (dx ? wy : -wy);
                                    giy += vg * (dy ? wx : -wx);
                                }
                            }
                        }
                        // d ix / d lx = (t/2) * (-1/t) = -1/2   (theta2 = (H/2 - lx)/t ; ix = (xs+1) t/2 - 1/2)
                        acc[2 * o] -= 0.5f * gix;
                        acc[2 * o + 1] -= 0.5f * giy;
                    }
                }
            }
        }
        if (GATHER && BWD && do_bwd) {
            __syncthreads();                          // this band's sD complete (and sInv, on the first band)
            // x pass: E[oc][i][x] = sum_k w[k] D[oc][i][jlo + k]; a thread keeps one (oc, x) and walks the band's rows
            for (int e = tid; e < NOBJ * 4 * t; e += kDecThreads) {
                const int x = e % t, oc = e / t, o = oc >> 2;
                const float* inv = sInv + (size_t)(o * t + x) * 8;
                const int jlo = reinterpret_cast<const int*>(inv)[0];
                float w[6];
#pragma unroll
                for (int k = 0; k < 6; ++k) w[k] = inv[1 + k];
                const int nk = min(6, H - jlo);
                for (int il = 0; il < brows; ++il) {
                    const float* row = sD + oc * BH + il * H + jlo;
                    float s = 0.f;
#pragma unroll
                    for (int k = 0; k < 6; ++k)
                        if (k < nk) s += w[k] * row[k];
                    sE[(oc * band_rows + il) * t + x] = s;
                }
            }
            __syncthreads();
            // y pass over the rows of this band, accumulated into this CTA's gradient block by the texel's only owner
            for (int e = tid; e < NOBJ * 4 * tt; e += kDecThreads) {
                const int x = e % t, y = (e / t) % t, oc = e / tt, o = oc >> 2, ch = oc & 3;
                const float* inv = sInv + (size_t)((NOBJ + o) * t + y) * 8;
                const int ilo = reinterpret_cast<const int*>(inv)[0];
                const int k0 = max(0, band0 - ilo), k1 = min(6, min(H, band0 + brows) - ilo);
                if (k0 < k1) {
                    float s = 0.f;
#pragma unroll
                    for (int k = 0; k < 6; ++k)
                        if (k >= k0 && k < k1) s += inv[1 + k] * sE[(oc * band_rows + (ilo + k - band0)) * t + x];
                    float* dst = ch == 0 ? sG + o * tt : sG + NOBJ * tt + (o * 3 + ch - 1) * tt;
                    dst[y * t + x] += s;
                }
            }
            if (band0 + band_rows < H) __syncthreads();   // the next band rewrites sD / sE
        }
        }   // bands
        // ---- per-frame reductions: d loc (2 per object) and the squared error ----
        constexpr int NR = 2 * NOBJ + 1;
        const int warp = tid >> 5, lane = tid & 31;
#pragma unroll
        for (int k = 0; k < NR; ++k) {
            float v = warp_sum(acc[k]);
            if (lane == 0) sRed[warp * NR + k] = v;
        }
        __syncthreads();
        if (tid < NR) {
            float s = 0.f;
            for (int w = 0; w < kDecThreads / 32; ++w) s += sRed[w * NR + tid];
            if (tid == 2 * NOBJ) {
                if (sg.sse) sg.sse[fl] = s;
            } else if (BWD && sg.dloc) {
                sg.dloc[(long)q * sg.dloc_seq_stride + (long)r * sg.dloc_row_stride + tid] = do_bwd ? s : 0.f;
            }
        }
        __syncthreads();      // tables and sRed are rewritten by the next frame
    }
    if (BWD) {
        float* dst = partials + (long)blockIdx.x * CN;
        for (int i = tid; i < CN; i += kDecThreads) dst[i] = sG[i];
    }
}

// d_consts[i] (+)= sum_p partials[p][i], p in fixed order.
__global__ void __launch_bounds__(256) decode_reduce_kernel(const float* __restrict__ partials, int nparts, int CN,
                                                            float* __restrict__ d_consts, int accumulate) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= CN) return;
    float s = 0.f;
    for (int p = 0; p < nparts; ++p) s += partials[(long)p * CN + i];
    d_consts[i] = accumulate ? d_consts[i] + s : s;
}

static int decode_ctas_per_sm() {
    static const int v = getenv("PAIG_DECODE_CTAS") ? atoi(getenv("PAIG_DECODE_CTAS")) : 2;
    return v < 1 ? 1 : (v > 5 ? 5 : v);
}
int decode_grid(int F) {
    const int cap = 148 * decode_ctas_per_sm();
    return F < cap ? (F < 1 ? 1 : F) : cap;
}

size_t decode_partials_floats(const paig_task* t) {
    const Dims d = dims_of(t);
    return (size_t)(148 * 5) * (size_t)(d.n * d.t * d.t * 4 + 3 * d.HW);
}

template <int NOBJ>
static int run_decode(const Dims& d, const float* consts, const DecSeg& a, const DecSeg& b, bool bwd, float* partials,
                      float* d_consts, int accumulate, cudaStream_t st) {
    const int CN = NOBJ * d.t * d.t * 4 + 3 * d.HW;
    const int F = a.nframes + b.nframes;
    if (F <= 0) return 0;
    const int grid = decode_grid(F);
    const size_t tab = (size_t)6 * NOBJ * d.H + 8 * (2 * NOBJ + 1);
    if (bwd) {
        static const bool atomics = getenv("PAIG_DECODE_ATOMICS") != nullptr;
        // Three objects need > 200 registers per thread: one CTA per SM whatever the shared-memory footprint, so the
        // gather tables may take the whole SM (3bp, 36 px: 162 KB) and the grid is one CTA per SM.  At 64 px the
        // constants and their gradient block alone are 160 KB: one CTA per SM as well, and the gather runs over bands
        // of rows (the largest band that fits), which keeps the decoder backward deterministic at every frame size.
        auto smem_for = [&](int band) {
            const size_t gather_fl = (size_t)NOBJ * 4 * band * (d.H + d.t) + (size_t)2 * NOBJ * d.t * 8;
            return ((size_t)2 * CN + tab + gather_fl) * sizeof(float);
        };
        int band = d.H;
        bool one_cta = NOBJ >= 3;
        if (smem_for(band) > (size_t)(one_cta ? 220 : 110) * 1024) {
            one_cta = true;
            while (band > 4 && smem_for(band) > (size_t)220 * 1024) band = (band + 1) / 2;
        }
        if (!atomics && smem_for(band) <= (size_t)(one_cta ? 220 : 110) * 1024) {
            const int g1 = one_cta && grid > 148 ? 148 : grid;
            launch(decode_kernel<NOBJ, true, true>, dim3(g1), dim3(kDecThreads), smem_for(band), st, d.H, band, consts, a, b,
                   partials);
            int rc = check_launch("decode_bwd");
            if (rc) return rc;
            launch(decode_reduce_kernel, dim3(cdiv(CN, 256)), dim3(256), 0, st, (const float*)partials, g1, CN, d_consts,
                   accumulate);
            return check_launch("decode_reduce");
        } else {
            const size_t smem = ((size_t)2 * CN + tab) * sizeof(float);
            launch(decode_kernel<NOBJ, true, false>, dim3(grid), dim3(kDecThreads), smem, st, d.H, d.H, consts, a, b, partials);
        }
        int rc = check_launch("decode_bwd");
        if (rc) return rc;
        launch(decode_reduce_kernel, dim3(cdiv(CN, 256)), dim3(256), 0, st, (const float*)partials, grid, CN, d_consts,
               accumulate);
        return check_launch("decode_reduce");
    }
    const size_t smem = ((size_t)CN + tab) * sizeof(float);
    launch(decode_kernel<NOBJ, false, false>, dim3(grid), dim3(kDecThreads), smem, st, d.H, d.H, consts, a, b, (float*)nullptr);
    return check_launch("decode_fwd");
}

int decode_run(const paig_task* t, const float* consts, const DecSeg& a, const DecSeg& b, bool bwd, float* partials,
               float* d_consts, int accumulate, cudaStream_t st) {
    const Dims d = dims_of(t);
    switch (d.n) {
        case 1: return run_decode<1>(d, consts, a, b, bwd, partials, d_consts, accumulate, st);
        case 2: return run_decode<2>(d, consts, a, b, bwd, partials, d_consts, accumulate, st);
        case 3: return run_decode<3>(d, consts, a, b, bwd, partials, d_consts, accumulate, st);
    }
    set_error("decode: n_objs=%d unsupported", d.n);
    return 1;
}

}  // namespace paig
