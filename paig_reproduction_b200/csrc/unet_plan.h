// Op lists of the fused UNet kernels (unet_fused.cu) -- shared with the tensor-core kernels of unet_tc.cu, which run the
// same backward-data op list in their own shared-memory layout.
#pragma once
#include "common.cuh"
#include "layout.h"

namespace paig {

constexpr int kFusedMaxOps = 24;

enum { F_CONV = 0, F_POOL = 1, F_HEAD = 2, F_UPT = 3, F_POOLT = 4, F_HEADT = 5 };

struct FusedOp {
    int kind, S, Cin0, Cin1, Cout, relu;
    int up;                        // >0: segment 0 is the 2x upsample of a half-resolution buffer, built `up` channels at a time
    int co_tile;                   // output channels per thread: 4, 8 or 16
    int py_tile;                   // output rows per thread: 1 or 2 (x 4 pixels)
    int in0, in1, out, chunk;      // shared-memory offsets (floats); out < 0: result is not read on chip
    int wsm, wfloats, wbar;        // weights: shared offset, packed floats (incl. bias), mbarrier index
    int next_w;                    // index of the next op that has weights (prefetched while this op runs), or -1
    long wglob;                    // offset of this layer in the packed weight buffer
    float* gout; long gout_bs;     // global destination of the result (kept for backward)
    float* gup; long gup_bs;       // global destination of the upsampled input
    // backward-data pass (unet_fused_bwd_kernel): transposed convs reuse F_CONV with bias = 0 and a ReLU mask
    int bias;                      // weights are followed by a bias vector
    const float* gmask; long gmask_bs;    // activation whose sign gates the result (ReLU adjoint); F_POOLT: the pooled tensor's source
    const float* gsrc; long gsrc_bs;      // F_HEADT: upstream gradient of the logits
    const float* gmask2; long gmask2_bs;  // F_HEADT: the logits (head ReLU), nullable
    int acc_gout;                         // F_POOLT: the other reader's share was parked in gout (global), not in1
};
struct FusedPlan {
    int nops, N, fps, H, first_w;
    long seq_stride;
    int x_off;
    const float* x;
    const float* wpack;
    long long* timing;             // debug (PAIG_DEBUG): per-CTA cycle stamps after every op of the CTA's first frames
    float kappa;                   // tensor-core variant: expected relative truncation loss of one MMA, added back in the epilogue
    FusedOp ops[kFusedMaxOps];
};


struct PackPlan {
    int nlayers;
    const float* w[24];
    const float* b[24];
    int Cout[24], Cin[24], taps[24];
    int mode[24];      // 0: forward [ci][tap][co] | bias.  1: transposed slice for the backward-data pass:
                       //    dst[(co*9 + tap)*Cout + c] = W[co][ci0 + c][8 - tap]   (Cin = number of co, no bias)
    int ci0[24], cin_total[24];
    long off[24];
    int frag;          // tensor-core variant: 3x3 layers go out in mma.sync A-fragment order
                       //    dst[(((kc*9 + tap)*mtiles + mt)*32 + lane)*4 + j] = W[co = 16mt + lane/4 + 8(j&1)][ci = 8kc + lane%4 + 4(j>>1)][tap]
                       //    (zero beyond Cout / Cin), bias after the last fragment
};

// backward-data op list + what the planner needs to know about every gradient slice (unet_backward_ops)
struct BwdOps {
    FusedPlan P;                   // ops[0 .. P.nops): kind, S, Cin0, Cout, relu, gmask / gout / gsrc / gmask2, acc_gout
    PackPlan K;                    // weights of the ops that have some, in op order (mode 1: transposed slice, 2: head)
    int in0_of[kFusedMaxOps], in1_of[kFusedMaxOps], out_of[kFusedMaxOps];     // gradient slice indices (-1: none)
    int nslices;
    int born[40], last[40], C[40], S[40];                                      // lifetime (op indices), channels, side
};
int unet_backward_ops(const paig_task* t, const paig_params* p, const Layout& L, float* ws, bool park_global, BwdOps* info);

}  // namespace paig
