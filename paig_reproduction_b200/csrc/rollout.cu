// Fused ODE rollout: all time steps of every sequence in one launch, state in registers.
//
// Replaces the Python time loop physics_models.py:231-239 calling cells.py:31-51 (spring),
// :60-83 (bouncing), :96-106 (gravity); each call is 5 explicit-Euler substeps.
//
// Forward arithmetic restates the reference op for op with explicit round-to-nearest intrinsics
// (__fmul_rn/__fadd_rn: never contracted into FMA), so trajectories are bit-identical to the ATen
// sequence on fp32.  Reference quirk kept on purpose: torch.split(poss, 1, dim=1) yields
// single COLUMNS, so the spring / bouncing cells act on columns 0 and 1 only; the rest pass through.
//
// Backward is a plain reverse sweep over the discrete substeps (no adjoint ODE): each step's five
// substeps are recomputed from the saved step-start state (pos_vel_seq row), kept in registers, and
// differentiated in reverse.  One thread integrates one sequence; a warp covers 32 sequences and the
// per-warp physics-constant gradients are reduced with shuffles in fp64.
#include "common.cuh"

namespace paig {

struct Phys {
    float h;     // dt/5 in fp32 (dt is an fp32 0-dim tensor: cells.py:26,58,90)
    float a;     // spring: (float)exp(k)            gravity: (float)(exp(g)*exp(2m))  (recomputed each call)
    float b;     // spring: (float)(2*exp(equil))
};

// a_frozen > 0: the gravity cell uses this A instead of exp(g) exp(2m) (the reference evaluates A once, in the cell's
// constructor, cells.py:92-94 -- SURVEY Q3; paig_task.gravity_A)
__device__ __forceinline__ Phys load_phys(int cell, const float* dt, const double* p0, const double* p1, float a_frozen) {
    Phys ph;
    ph.h = __fdiv_rn(*dt, 5.0f);
    ph.a = 0.f;
    ph.b = 0.f;
    if (cell == PAIG_CELL_SPRING) {
        ph.a = (float)exp(*p0);
        ph.b = (float)(2.0 * exp(*p1));
    } else if (cell == PAIG_CELL_GRAVITY) {
        ph.a = a_frozen > 0.f ? a_frozen : (float)(exp(*p0) * exp(2.0 * (*p1)));
    }
    return ph;
}

// ---- substeps: forward ---------------------------------------------------------------------------

// cells.py:34-47.  p,v: columns 0 and 1.
__device__ __forceinline__ void spring_sub(float& p0, float& p1, float& v0, float& v1, const Phys& ph) {
    float d = __fsub_rn(p0, p1);
    float nrm = __fsqrt_rn(fabsf(__fmul_rn(d, d)));
    float dir = __fdiv_rn(d, __fadd_rn(nrm, 1e-4f));
    float F = __fmul_rn(__fmul_rn(ph.a, __fsub_rn(nrm, ph.b)), dir);
    float hF = __fmul_rn(ph.h, F);
    v0 = __fsub_rn(v0, hF);
    v1 = __fadd_rn(v1, hF);
    p0 = __fadd_rn(p0, __fmul_rn(ph.h, v0));
    p1 = __fadd_rn(p1, __fmul_rn(ph.h, v1));
}

// cells.py:64-79 for one column.  Returns the three branch bits for the reverse sweep.
__device__ __forceinline__ int bounce_sub(float& p, float& v, const Phys& ph) {
    float p1 = __fadd_rn(p, __fmul_rn(ph.h, v));
    bool c1 = __fadd_rn(p1, 2.f) > 32.f;          // upper wall, tested on the moved position
    bool c2 = 0.f > __fsub_rn(p1, 2.f);           // lower wall, also on the moved (pre-reflection) position
    float vv = c1 ? -v : v;
    vv = c2 ? -vv : vv;
    float p2 = c1 ? __fsub_rn(__fsub_rn(32.f, __fsub_rn(__fadd_rn(p1, 2.f), 32.f)), 2.f) : p1;
    bool c4 = 0.f > __fsub_rn(p2, 2.f);           // lower wall for the position sees the upper-reflected value
    float p3 = c4 ? __fadd_rn(-__fsub_rn(p2, 2.f), 2.f) : p2;
    p = p3;
    v = vv;
    return (c1 ? 1 : 0) | (c2 ? 2 : 0) | (c4 ? 4 : 0);
}

// cells.py:97-105.  P,V: [x0,y0,x1,y1,x2,y2].
__device__ __forceinline__ void gravity_sub(float* P, float* V, const Phys& ph) {
    float vec[3][2], f[3][2];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int a = 2 * i, b = 2 * ((i + 1) % 3);
        vec[i][0] = __fsub_rn(P[a], P[b]);
        vec[i][1] = __fsub_rn(P[a + 1], P[b + 1]);
        float s = __fadd_rn(__fmul_rn(vec[i][0], vec[i][0]), __fmul_rn(vec[i][1], vec[i][1]));
        float nrm = __fsqrt_rn(fminf(fmaxf(s, 1e-1f), 1e5f));
        float q = fminf(fmaxf(nrm, 1.f), 170.f);
        float r = __fmul_rn(__fmul_rn(q, q), q);
        f[i][0] = __fdiv_rn(vec[i][0], r);
        f[i][1] = __fdiv_rn(vec[i][1], r);
    }
    const float nA = -ph.a;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int j = (i + 2) % 3;   // F0=f0-f2, F1=f1-f0, F2=f2-f1
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            float F = __fmul_rn(nA, __fsub_rn(f[i][c], f[j][c]));
            V[2 * i + c] = __fadd_rn(V[2 * i + c], __fmul_rn(ph.h, F));
        }
    }
#pragma unroll
    for (int c = 0; c < 6; ++c) P[c] = __fadd_rn(P[c], __fmul_rn(ph.h, V[c]));
}

// ---- substeps: reverse ---------------------------------------------------------------------------

// Inputs: state at the START of the substep (p0,p1; velocities are not needed), upstream gradients of the
// substep outputs in (g*); on return g* hold gradients of the substep inputs.  ga, gb accumulate
// dL/d(exp k) and dL/d(2 exp equil).
__device__ __forceinline__ void spring_sub_bwd(float p0, float p1, const Phys& ph, float& gp0, float& gp1, float& gv0,
                                               float& gv1, double& ga, double& gb) {
    float d = p0 - p1;
    float s = d * d;
    float nrm = sqrtf(fabsf(s));
    float den = nrm + 1e-4f;
    float dir = d / den;
    float u = nrm - ph.b;
    gv0 += ph.h * gp0;                    // p0' = p0 + h v0'
    gv1 += ph.h * gp1;
    float gF = ph.h * (gv1 - gv0);        // v0' = v0 - hF ; v1' = v1 + hF
    ga += (double)(gF * u * dir);
    float g_u = gF * ph.a * dir;
    float g_dir = gF * ph.a * u;
    gb -= (double)g_u;
    float g_nrm = g_u - g_dir * d / (den * den);
    float g_d = g_dir / den;
    // nrm = sqrt(|s|), s = d^2 :  d nrm / d d = sign(s) * d / nrm   (0*inf = NaN at d == 0, as in the reference, Q13)
    float sgn = s > 0.f ? 1.f : (s < 0.f ? -1.f : 0.f);
    g_d += g_nrm * (0.5f / nrm) * sgn * (2.f * d);
    gp0 += g_d;
    gp1 -= g_d;
}

__device__ __forceinline__ void bounce_sub_bwd(int bits, const Phys& ph, float& gp, float& gv) {
    if (bits & 4) gp = -gp;
    if (bits & 1) gp = -gp;
    if (bits & 2) gv = -gv;
    if (bits & 1) gv = -gv;
    gv += ph.h * gp;                      // p1 = p + h v
}

__device__ __forceinline__ void gravity_sub_bwd(const float* P, const Phys& ph, float* gP, float* gV, double& gA) {
    float vec[3][2], r[3], q[3], nrm[3], s[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int a = 2 * i, b = 2 * ((i + 1) % 3);
        vec[i][0] = P[a] - P[b];
        vec[i][1] = P[a + 1] - P[b + 1];
        s[i] = vec[i][0] * vec[i][0] + vec[i][1] * vec[i][1];
        nrm[i] = sqrtf(fminf(fmaxf(s[i], 1e-1f), 1e5f));
        q[i] = fminf(fmaxf(nrm[i], 1.f), 170.f);
        r[i] = q[i] * q[i] * q[i];
    }
    float gD[3][2];
#pragma unroll
    for (int c = 0; c < 6; ++c) gV[c] += ph.h * gP[c];        // pos' = pos + h vel'
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int j = (i + 2) % 3;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            float gF = ph.h * gV[2 * i + c];                   // vel' = vel + h F
            float D = vec[i][c] / r[i] - vec[j][c] / r[j];
            gA -= (double)(gF * D);                            // F = -A * D
            gD[i][c] = -ph.a * gF;
        }
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int k = (i + 1) % 3;   // f_i appears in D_i (+) and D_{i+1} (-)
        float gf0 = gD[i][0] - gD[k][0];
        float gf1 = gD[i][1] - gD[k][1];
        float gvec0 = gf0 / r[i], gvec1 = gf1 / r[i];
        float g_r = -(gf0 * vec[i][0] + gf1 * vec[i][1]) / (r[i] * r[i]);
        float g_q = g_r * 3.f * q[i] * q[i];
        float g_nrm = (nrm[i] >= 1.f && nrm[i] <= 170.f) ? g_q : 0.f;
        float g_c = g_nrm * 0.5f / nrm[i];
        float g_s = (s[i] >= 1e-1f && s[i] <= 1e5f) ? g_c : 0.f;
        gvec0 += 2.f * g_s * vec[i][0];
        gvec1 += 2.f * g_s * vec[i][1];
        const int a = 2 * i, b = 2 * ((i + 1) % 3);
        gP[a] += gvec0;
        gP[a + 1] += gvec1;
        gP[b] -= gvec0;
        gP[b + 1] -= gvec1;
    }
}

// ---- kernels -------------------------------------------------------------------------------------

template <int CELL, int NS>   // NS = 2*n_objs
__global__ void __launch_bounds__(128) rollout_fwd_kernel(int B, int steps, const float* __restrict__ dt,
                                                          const double* __restrict__ p0, const double* __restrict__ p1,
                                                          float a_frozen, float* __restrict__ seq) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const Phys ph = load_phys(CELL, dt, p0, p1, a_frozen);
    float* row = seq + (long)b * (steps + 1) * 2 * NS;
    float P[NS], V[NS];
#pragma unroll
    for (int c = 0; c < NS; ++c) {
        P[c] = row[c];
        V[c] = row[NS + c];
    }
    for (int s = 1; s <= steps; ++s) {
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            if (CELL == PAIG_CELL_SPRING) {
                spring_sub(P[0], P[1], V[0], V[1], ph);
            } else if (CELL == PAIG_CELL_BOUNCING) {
                bounce_sub(P[0], V[0], ph);     // reference order: both columns move first, then reflect one by one;
                bounce_sub(P[1], V[1], ph);     // the columns never interact, so per-column order is equivalent
            } else {
                gravity_sub(P, V, ph);
            }
        }
        row += 2 * NS;
#pragma unroll
        for (int c = 0; c < NS; ++c) {
            row[c] = P[c];
            row[NS + c] = V[c];
        }
    }
}

// dpos / dvel: upstream gradient of row s (1..steps, and row 0 when with_row0) at
// base + b*batch_stride + s*row_stride (+ column); either may be NULL.
template <int CELL, int NS>
__global__ void __launch_bounds__(128) rollout_bwd_kernel(int B, int steps, const float* __restrict__ dt,
                                                          const double* __restrict__ p0, const double* __restrict__ p1,
                                                          float a_frozen, const float* __restrict__ seq,
                                                          const float* __restrict__ dpos,
                                                          const float* __restrict__ dvel, long batch_stride,
                                                          long row_stride, int with_row0, float* __restrict__ d_state0,
                                                          double* d_phys0, double* d_phys1, double* __restrict__ block_part,
                                                          unsigned* __restrict__ block_count) {
    __shared__ double red[2][4];
    const Phys ph = load_phys(CELL, dt, p0, p1, a_frozen);
    double ga = 0.0, gb = 0.0;
    // a thread walks sequences b, b + grid*block, ... (one sequence per thread unless the caller had no room for
    // per-block partials and asked for a single block)
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) {
        float gP[NS], gV[NS];
#pragma unroll
        for (int c = 0; c < NS; ++c) gP[c] = gV[c] = 0.f;
        for (int s = steps; s >= 1; --s) {
#pragma unroll
            for (int c = 0; c < NS; ++c) {
                if (dpos) gP[c] += dpos[(long)b * batch_stride + (long)s * row_stride + c];
                if (dvel) gV[c] += dvel[(long)b * batch_stride + (long)s * row_stride + c];
            }
            const float* row = seq + ((long)b * (steps + 1) + (s - 1)) * 2 * NS;   // state at the start of step s
            float P[NS], V[NS];
#pragma unroll
            for (int c = 0; c < NS; ++c) {
                P[c] = row[c];
                V[c] = row[NS + c];
            }
            if (CELL == PAIG_CELL_SPRING) {
                float sp0[5], sp1[5];
#pragma unroll
                for (int i = 0; i < 5; ++i) {
                    sp0[i] = P[0];
                    sp1[i] = P[1];
                    spring_sub(P[0], P[1], V[0], V[1], ph);
                }
#pragma unroll
                for (int i = 4; i >= 0; --i) spring_sub_bwd(sp0[i], sp1[i], ph, gP[0], gP[1], gV[0], gV[1], ga, gb);
            } else if (CELL == PAIG_CELL_BOUNCING) {
                int bits0[5], bits1[5];
#pragma unroll
                for (int i = 0; i < 5; ++i) {
                    bits0[i] = bounce_sub(P[0], V[0], ph);
                    bits1[i] = bounce_sub(P[1], V[1], ph);
                }
#pragma unroll
                for (int i = 4; i >= 0; --i) {
                    bounce_sub_bwd(bits0[i], ph, gP[0], gV[0]);
                    bounce_sub_bwd(bits1[i], ph, gP[1], gV[1]);
                }
            } else {
                float sP[5][NS];
#pragma unroll
                for (int i = 0; i < 5; ++i) {
#pragma unroll
                    for (int c = 0; c < NS; ++c) sP[i][c] = P[c];
                    gravity_sub(P, V, ph);
                }
#pragma unroll
                for (int i = 4; i >= 0; --i) gravity_sub_bwd(sP[i], ph, gP, gV, ga);
            }
        }
        if (with_row0) {
#pragma unroll
            for (int c = 0; c < NS; ++c) {
                if (dpos) gP[c] += dpos[(long)b * batch_stride + c];
                if (dvel) gV[c] += dvel[(long)b * batch_stride + c];
            }
        }
#pragma unroll
        for (int c = 0; c < NS; ++c) {
            d_state0[(long)b * 2 * NS + c] = gP[c];
            d_state0[(long)b * 2 * NS + NS + c] = gV[c];
        }
    }
    // physics-constant gradients: chain through exp (d e^k/dk = e^k; d 2e^q/dq = 2e^q; dA/dg = A), fp64 reduce
    ga *= (double)ph.a;
    gb *= (double)ph.b;
    ga = warp_sum(ga);
    gb = warp_sum(gb);
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) {
        red[0][w] = ga;
        red[1][w] = gb;
    }
    __syncthreads();
    if (threadIdx.x == 0 && CELL != PAIG_CELL_BOUNCING && (d_phys0 != nullptr || d_phys1 != nullptr)) {
        double sa = 0.0, sb = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) {
            sa += red[0][i];
            sb += red[1][i];
        }
        if (gridDim.x == 1) {
            if (d_phys0) *d_phys0 = sa;
            if (d_phys1) *d_phys1 = sb;
        } else {
            // several blocks: every block parks its partial, the last one to arrive folds them in block order, so the
            // result does not depend on which block finishes first (run-to-run bit-identical for any B)
            block_part[2 * blockIdx.x] = sa;
            block_part[2 * blockIdx.x + 1] = sb;
            __threadfence();
            const unsigned seen = atomicAdd(block_count, 1u);
            if (seen == gridDim.x - 1) {
                __threadfence();
                double ta = 0.0, tb = 0.0;
                for (unsigned i = 0; i < gridDim.x; ++i) {
                    ta += ((volatile double*)block_part)[2 * i];
                    tb += ((volatile double*)block_part)[2 * i + 1];
                }
                if (d_phys0) *d_phys0 = ta;
                if (d_phys1) *d_phys1 = tb;
            }
        }
    }
}

// ---- host wrappers -------------------------------------------------------------------------------

template <int CELL, int NS>
static void launch_fwd(int B, int steps, const float* dt, const double* p0, const double* p1, float aF, float* seq,
                       cudaStream_t st) {
    launch(rollout_fwd_kernel<CELL, NS>, dim3(cdiv(B, 128)), dim3(128), 0, st, B, steps, dt, p0, p1, aF, seq);
}
template <int CELL, int NS>
static void launch_bwd(int B, int steps, const float* dt, const double* p0, const double* p1, float aF, const float* seq,
                       const float* dpos, const float* dvel, long bs, long rs, int with_row0, float* d0, double* dphys0,
                       double* dphys1, double* scratch, cudaStream_t st) {
    // scratch (rollout_scratch_doubles(B) doubles, zero before first use): [0] arrival counter, [2..] per-block partials.
    // Without it the launch is one block, which needs none.
    const int blocks = scratch ? cdiv(B, 128) : 1;
    launch(rollout_bwd_kernel<CELL, NS>, dim3(blocks), dim3(128), 0, st, B, steps, dt, p0, p1, aF, seq, dpos, dvel,
           bs, rs, with_row0, d0, dphys0, dphys1, scratch ? scratch + 2 : nullptr, reinterpret_cast<unsigned*>(scratch));
}

int rollout_forward(int cell, int n, int B, int steps, const float* dt, const double* p0, const double* p1, float aF,
                    float* seq, cudaStream_t st) {
    if (B <= 0) return 0;
    if (cell == PAIG_CELL_SPRING && n == 2) launch_fwd<PAIG_CELL_SPRING, 4>(B, steps, dt, p0, p1, aF, seq, st);
    else if (cell == PAIG_CELL_BOUNCING && n == 2) launch_fwd<PAIG_CELL_BOUNCING, 4>(B, steps, dt, p0, p1, aF, seq, st);
    else if (cell == PAIG_CELL_GRAVITY && n == 3) launch_fwd<PAIG_CELL_GRAVITY, 6>(B, steps, dt, p0, p1, aF, seq, st);
    else {
        set_error("rollout: unsupported cell %d with %d objects (reference cells assume 2 / 2 / 3)", cell, n);
        return 1;
    }
    return check_launch("rollout_fwd");
}

int rollout_backward(int cell, int n, int B, int steps, const float* dt, const double* p0, const double* p1, float aF,
                     const float* seq, const float* dpos, const float* dvel, long bs, long rs, int with_row0,
                     float* d_state0, double* d_phys0, double* d_phys1, double* scratch, cudaStream_t st) {
    if (B <= 0) return 0;
    if (B <= 128) scratch = nullptr;                     // one block: no partials, no counter
    // the arrival counter re-arms itself after every launch, but the caller's workspace starts out uninitialised
    if (scratch && cudaMemsetAsync(scratch, 0, 2 * sizeof(double), st) != cudaSuccess) return check_launch("rollout_bwd memset");
    if (cell == PAIG_CELL_SPRING && n == 2)
        launch_bwd<PAIG_CELL_SPRING, 4>(B, steps, dt, p0, p1, aF, seq, dpos, dvel, bs, rs, with_row0, d_state0, d_phys0, d_phys1, scratch, st);
    else if (cell == PAIG_CELL_BOUNCING && n == 2)
        launch_bwd<PAIG_CELL_BOUNCING, 4>(B, steps, dt, p0, p1, aF, seq, dpos, dvel, bs, rs, with_row0, d_state0, d_phys0, d_phys1, scratch, st);
    else if (cell == PAIG_CELL_GRAVITY && n == 3)
        launch_bwd<PAIG_CELL_GRAVITY, 6>(B, steps, dt, p0, p1, aF, seq, dpos, dvel, bs, rs, with_row0, d_state0, d_phys0, d_phys1, scratch, st);
    else {
        set_error("rollout: unsupported cell %d with %d objects", cell, n);
        return 1;
    }
    return check_launch("rollout_bwd");
}

}  // namespace paig
