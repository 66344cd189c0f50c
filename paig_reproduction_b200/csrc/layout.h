// Workspace layout of one training step.  The caller owns the buffer (paig_workspace_bytes); this header
// only assigns offsets (in floats, each region 256-byte aligned) so forward and backward agree on them.
#pragma once
#include "common.cuh"
#include "internal.h"

namespace paig {

// ---- UNet description (blocks.py:106-308) as data: buffers + a forward op list ------------------
enum OpKind { OP_CONV = 0, OP_POOL = 1, OP_UP = 2, OP_HEAD = 3 };
struct Ref {
    int buf;   // >= 0: UNet buffer index; -1: the input frames; -2: the logits
    int c0;    // first channel of the slice
    int C;     // channels in the slice
};
struct Op {
    int kind;
    int layer;   // conv index (0-based into paig_params.conv) for OP_CONV / OP_HEAD
    Ref in, out;
    int relu;
};
struct BufDesc {
    int C;       // total channels
    int shift;   // spatial side = H >> shift
};
constexpr int kMaxBufs = 24, kMaxOps = 32;
struct UNetDesc {
    int nbufs = 0, nops = 0;
    BufDesc bufs[kMaxBufs];
    Op ops[kMaxOps];
    int logits_relu = 0;
};
UNetDesc make_unet(int deep, int n_objs);

inline size_t align64(size_t floats) { return (floats + 63) & ~(size_t)63; }

struct Layout {
    Dims d;
    int B = 0, N = 0;            // local sequences, encoded frames (B * (in+pr))
    int K = 0;                   // encoder.l1 in_features (blocks.py:70-73)
    UNetDesc unet;
    size_t act[kMaxBufs], grad[kMaxBufs];
    size_t logits, d_logits, masks, A, dA, H1, H2, O3, dH1, dH2, dO3, enc_pos, d_enc_pos;
    size_t vin, v1, v2, vout, dvin, dv1, dv2, dvout;
    size_t seq, d_seq, d_state0;
    size_t raw, consts, hidden, d_consts, dec_partials, tmpl_scratch;
    size_t sse, scales, losses, dphys;
    size_t partials;             // conv wgrad / head partial sums; split-K partials of the MLP GEMMs
    size_t partials_floats;
    size_t partials2, partials2_floats;   // side-stream scratch of the encoder.l1 weight gradient
    size_t tc_scratch;           // transposed operands of the tcgen05 encoder.l1 GEMMs: W1^T | dH1^T | A^T
    size_t convtc;               // packed weights of the layer in flight on the tcgen05 conv path (deep UNet)
    size_t wpack;                // UNet weights re-packed [ci][tap][co]|bias for the fused forward kernel
    size_t wpack_tc;             // ... and as hi / lo TF32 operand blocks for the tcgen05 forward (unet_tc.cu)
    size_t frames;               // gathered encoder frames [N,3,H,H] (when they are a prefix of each sequence)
    size_t x_stage;              // device copy of the input for the *_host entry points (slot 0)
    size_t x_stage2;             // second slot: paig_stage_input_host(slot 1) lands here while a step reads slot 0
    size_t total;
};

Layout make_layout(const paig_task* t, int B);

// Static shared-memory planner: blocks with [born, dies] step intervals are placed largest first at the lowest
// offset that is free over their whole interval (interval-graph colouring heuristic; near-optimal here).
struct Planner {
    struct Blk { int size, born, dies, off; int* dst; };
    Blk b[96];
    int n = 0;
    void add(int size, int born, int dies, int* dst) {
        b[n++] = Blk{(size + 3) & ~3, born, dies, -1, dst};       // 16-byte granularity
    }
    int place(int mode = 0) {                                      // returns the peak (floats)
        int order[96];
        for (int i = 0; i < n; ++i) order[i] = i;
        // mode 0: size descending.  1: lifetime descending, then size.  2: birth ascending, then size descending.
        auto before = [&](const Blk& x, const Blk& y) {
            if (mode == 1 && (x.dies - x.born) != (y.dies - y.born)) return (x.dies - x.born) > (y.dies - y.born);
            if (mode == 2 && x.born != y.born) return x.born < y.born;
            return x.size > y.size;
        };
        for (int i = 1; i < n; ++i)                                // insertion sort
            for (int j = i; j > 0 && before(b[order[j]], b[order[j - 1]]); --j) {
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
    int place_best() {                                             // the ordering heuristic with the lowest peak
        int best = 0, best_peak = place(0);
        for (int m = 1; m < 3; ++m) {
            const int pk = place(m);
            if (pk < best_peak) { best_peak = pk; best = m; }
        }
        return place(best);
    }
};


// unet_fused.cu -- whole-UNet forward in one persistent kernel; returns -1 when the network does not fit on chip
size_t unet_wpack_floats(const UNetDesc& u, const paig_task* t);
int unet_fused_forward(const paig_task* t, const paig_params* p, const Layout& L, const float* x, long seq_stride, int fps,
                       float* ws, cudaStream_t st);
int unet_fused_backward(const paig_task* t, const paig_params* p, const Layout& L, float* ws, cudaStream_t st);
// unet_tc.cu -- the same forward on the tcgen05 tensor cores (3xTF32, taps batched along N); -1: not applicable / switched off
int unet_tc_forward(const paig_task* t, const paig_params* p, const Layout& L, const float* x, long seq_stride, int fps,
                    float* ws, cudaStream_t st);
int unet_tc_backward(const paig_task* t, const paig_params* p, const Layout& L, float* ws, cudaStream_t st);
size_t unet_tc_wpack_floats(const UNetDesc& u);

// encoder.cu
int encoder_forward(const paig_task* t, const paig_params* p, const Layout& L, const float* x, long seq_stride,
                    int fps, float* enc_pos_out, float* enc_masks_out, float* masked_out, float* ws, cudaStream_t st);
int encoder_backward(const paig_task* t, const paig_params* p, const paig_params* g, const Layout& L, const float* x,
                     long seq_stride, int fps, const float* d_enc_pos, float* ws, cudaStream_t st);
int velocity_forward(const paig_task* t, const paig_params* p, const Layout& L, const float* enc_pos, float* ws,
                     cudaStream_t st);
int velocity_backward(const paig_task* t, const paig_params* p, const paig_params* g, const Layout& L,
                      const float* d_state0, float* d_enc_pos, float* ws, cudaStream_t st);

}  // namespace paig
