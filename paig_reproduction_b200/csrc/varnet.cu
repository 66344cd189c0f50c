// The three VariableFromNetwork generators (blocks.py:311-322; instances physics_models.py:106-108) and the
// decoder's constant preprocessing (physics_models.py:163-171,182):
//     raw    = l2(tanh(l1(ones[1,10])))                     for template / contents / background
//     consts = [template + 5 | sigmoid(contents) | sigmoid(background)]
// The reference re-evaluates these on every decoder call (27x per training forward); they do
// not depend on the batch, so the step evaluates them once.  Net order everywhere: 0 template, 1 contents,
// 2 background -- the layout of `raw`, `consts` and `hidden`.
#include "common.cuh"
#include "internal.h"

namespace paig {

constexpr int kVarThreads = 256;
constexpr int kRowsPerWarp = 4;
constexpr int kRowsPerBlock = (kVarThreads / 32) * kRowsPerWarp;

struct VarNets {
    const float* w1[3];
    const float* b1[3];
    const float* w2[3];
    const float* b2[3];
    float* gw1[3];
    float* gb1[3];
    float* gw2[3];
    float* gb2[3];
    int numel[3];
    int offset[3];
};

static VarNets make_nets(const paig_task* t, const paig_params* p, const paig_params* g) {
    const Dims d = dims_of(t);
    VarNets v;
    const paig_wb* l1[3] = {&p->template_l1, &p->content_l1, &p->background_l1};
    const paig_wb* l2[3] = {&p->template_l2, &p->content_l2, &p->background_l2};
    const int numel[3] = {d.n * d.t * d.t, 3 * d.n * d.t * d.t, 3 * d.HW};
    int off = 0;
    for (int k = 0; k < 3; ++k) {
        v.w1[k] = l1[k]->w; v.b1[k] = l1[k]->b; v.w2[k] = l2[k]->w; v.b2[k] = l2[k]->b;
        v.numel[k] = numel[k];
        v.offset[k] = off;
        off += numel[k];
        v.gw1[k] = v.gb1[k] = v.gw2[k] = v.gb2[k] = nullptr;
    }
    if (g) {
        const paig_wb* g1[3] = {&g->template_l1, &g->content_l1, &g->background_l1};
        const paig_wb* g2[3] = {&g->template_l2, &g->content_l2, &g->background_l2};
        for (int k = 0; k < 3; ++k) {
            v.gw1[k] = g1[k]->w; v.gb1[k] = g1[k]->b; v.gw2[k] = g2[k]->w; v.gb2[k] = g2[k]->b;
        }
    }
    return v;
}

__device__ __forceinline__ void hidden_to_smem(const float* __restrict__ w1, const float* __restrict__ b1, float* sh) {
    for (int j = threadIdx.x; j < kHidden; j += blockDim.x) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < kVarIn; ++i) s += w1[j * kVarIn + i];      // input is ones[1,10]
        sh[j] = tanhf(s + b1[j]);
    }
}

__global__ void __launch_bounds__(kVarThreads) varnet_fwd_kernel(VarNets v, float* __restrict__ raw,
                                                                 float* __restrict__ consts, float* __restrict__ hidden) {
    __shared__ float sh[kHidden];
    const int net = blockIdx.y;
    const int row0 = blockIdx.x * kRowsPerBlock;
    if (row0 >= v.numel[net]) return;
    hidden_to_smem(v.w1[net], v.b1[net], sh);
    __syncthreads();
    if (blockIdx.x == 0 && hidden)
        for (int j = threadIdx.x; j < kHidden; j += blockDim.x) hidden[net * kHidden + j] = sh[j];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int rr = 0; rr < kRowsPerWarp; ++rr) {
        const int r = row0 + warp * kRowsPerWarp + rr;
        if (r >= v.numel[net]) break;
        const float* w = v.w2[net] + (long)r * kHidden;
        float s = 0.f;
        for (int j = lane; j < kHidden; j += 32) s += w[j] * sh[j];
        s = warp_sum(s);
        if (lane == 0) {
            s += v.b2[net][r];
            const int o = v.offset[net] + r;
            if (raw) raw[o] = s;
            consts[o] = net == 0 ? s + 5.f : sigmoidf_(s);
        }
    }
}

// l2 backward: db2, dW2 (outer product with the hidden vector) and per-block partials of dh.
__global__ void __launch_bounds__(kVarThreads) varnet_bwd_l2_kernel(VarNets v, const float* __restrict__ consts,
                                                                    const float* __restrict__ hidden,
                                                                    const float* __restrict__ d_consts,
                                                                    float* __restrict__ dh_part, int max_blocks) {
    __shared__ float sh[kHidden];
    __shared__ float sdh[kVarThreads / 32][kHidden];
    const int net = blockIdx.y;
    const int row0 = blockIdx.x * kRowsPerBlock;
    if (row0 >= v.numel[net]) return;
    for (int j = threadIdx.x; j < kHidden; j += blockDim.x) sh[j] = hidden[net * kHidden + j];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float dh[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) dh[k] = 0.f;
    for (int rr = 0; rr < kRowsPerWarp; ++rr) {
        const int r = row0 + warp * kRowsPerWarp + rr;
        if (r >= v.numel[net]) break;
        const int o = v.offset[net] + r;
        float g = d_consts[o];
        if (net != 0) {
            const float s = consts[o];
            g *= s * (1.f - s);                                   // sigmoid'
        }
        if (lane == 0) v.gb2[net][r] = g;
        const float* w = v.w2[net] + (long)r * kHidden;
        float* gw = v.gw2[net] + (long)r * kHidden;
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            const int j = lane + 32 * k;
            if (j < kHidden) {
                gw[j] = g * sh[j];
                dh[k] += g * w[j];
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        const int j = lane + 32 * k;
        if (j < kHidden) sdh[warp][j] = dh[k];
    }
    __syncthreads();
    for (int j = threadIdx.x; j < kHidden; j += blockDim.x) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kVarThreads / 32; ++w) s += sdh[w][j];
        dh_part[((long)net * max_blocks + blockIdx.x) * kHidden + j] = s;
    }
}

// tanh and l1 backward (input is the constant ones vector, so dW1[j,:] = dpre[j]).
__global__ void __launch_bounds__(kVarThreads) varnet_bwd_l1_kernel(VarNets v, const float* __restrict__ hidden,
                                                                    const float* __restrict__ dh_part, int max_blocks) {
    const int net = blockIdx.x;
    const int nb = (v.numel[net] + kRowsPerBlock - 1) / kRowsPerBlock;
    for (int j = threadIdx.x; j < kHidden; j += blockDim.x) {
        float s = 0.f;
        for (int b = 0; b < nb; ++b) s += dh_part[((long)net * max_blocks + b) * kHidden + j];
        const float h = hidden[net * kHidden + j];
        const float dpre = s * (1.f - h * h);
        v.gb1[net][j] = dpre;
#pragma unroll
        for (int i = 0; i < kVarIn; ++i) v.gw1[net][j * kVarIn + i] = dpre;
    }
}

static int max_blocks_of(const VarNets& v) {
    int m = 0;
    for (int k = 0; k < 3; ++k) m = v.numel[k] > m ? v.numel[k] : m;
    return cdiv(m, kRowsPerBlock);
}

size_t templates_scratch_floats(const paig_task* t) {
    const Dims d = dims_of(t);
    return (size_t)3 * cdiv(3 * d.HW, kRowsPerBlock) * kHidden;
}

int templates_forward(const paig_task* t, const paig_params* p, float* raw, float* consts, float* hidden,
                      cudaStream_t st) {
    VarNets v = make_nets(t, p, nullptr);
    launch(varnet_fwd_kernel, dim3(max_blocks_of(v), 3), dim3(kVarThreads), 0, st, v, raw, consts, hidden);
    return check_launch("varnet_fwd");
}

int templates_backward(const paig_task* t, const paig_params* p, const paig_params* g, const float* consts,
                       const float* hidden, const float* d_consts, float* scratch, cudaStream_t st) {
    VarNets v = make_nets(t, p, g);
    const int mb = max_blocks_of(v);
    launch(varnet_bwd_l2_kernel, dim3(mb, 3), dim3(kVarThreads), 0, st, v, consts, hidden, d_consts, scratch, mb);
    int rc = check_launch("varnet_bwd_l2");
    if (rc) return rc;
    launch(varnet_bwd_l1_kernel, dim3(3), dim3(kVarThreads), 0, st, v, hidden, (const float*)scratch, mb);
    return check_launch("varnet_bwd_l1");
}

}  // namespace paig
