// Internal (C++) entry points shared between the stage files and api.cu.
#pragma once
#include "common.cuh"

namespace paig {

bool valid_task(const paig_task* t);

// side.cu -- the fused step's side streams (nullptr: run everything in line on the caller's stream)
struct Side {
    cudaStream_t main = nullptr, s1 = nullptr, s2 = nullptr;
    void* set = nullptr;
    void after(cudaStream_t waiter, cudaStream_t signaler);     // waiter waits for all work enqueued on signaler so far
    void fork1() { after(s1, main); }
    void fork2() { after(s2, main); }
    void join1() { after(main, s1); }
    void join2() { after(main, s2); }
};
Side* side_begin(cudaStream_t main);
void side_end();
Side* side_cur();
// an event the caller wants recorded once every gradient EXCEPT the UNet conv layers' is final (data-parallel overlap)
extern thread_local void* g_early_event;

// rollout.cu
// a_frozen: 0, or the gravity cell's constructor-time A (paig_task.gravity_A)
int rollout_forward(int cell, int n, int B, int steps, const float* dt, const double* p0, const double* p1,
                    float a_frozen, float* seq, cudaStream_t st);
int rollout_backward(int cell, int n, int B, int steps, const float* dt, const double* p0, const double* p1,
                     float a_frozen, const float* seq, const float* dpos, const float* dvel, long batch_stride, long row_stride,
                     int with_row0, float* d_state0, double* d_phys0, double* d_phys1, double* scratch, cudaStream_t st);
inline size_t rollout_scratch_doubles(int B) { return 2 + 2 * (size_t)((B + 127) / 128); }

// decoder.cu -- a run of frames sharing one indexing rule (recons frames of all sequences, or rollout frames)
struct DecSeg {
    int nframes = 0;              // frames in this segment; frame fl has q = fl / fps, r = fl % fps
    int fps = 1;
    const float* loc = nullptr;   // object locations of frame (q,r) at loc + q*loc_seq_stride + r*loc_row_stride
    long loc_seq_stride = 0, loc_row_stride = 0;
    const float* target = nullptr;   // frame the output is compared with: target + q*tgt_seq_stride + r*3*H*H
    long tgt_seq_stride = 0;
    const float* scale = nullptr;    // [fps] loss weight s_r: in-kernel upstream gradient is 2*s_r*(out - target)
    int use_scale = 0;               // ... or by value: s_r = r < scale_split ? scale_lo : scale_hi
    int scale_split = 0;
    float scale_lo = 0.f, scale_hi = 0.f;
    const float* dframes = nullptr;  // [nframes,3,H,H] upstream gradient from memory (overrides scale)
    float* frames = nullptr;         // [nframes,3,H,H] decoded output (nullable)
    float* sse = nullptr;            // [nframes] sum of squared error vs target (nullable)
    float* dloc = nullptr;           // gradient wrt loc, indexed like loc with its own strides (nullable)
    long dloc_seq_stride = 0, dloc_row_stride = 0;
    float* layer_c = nullptr;        // forward only, nullable: [n+1][nframes][3][H][H] per-layer contents (transf_contents)
    float* layer_m = nullptr;        //   and softmax masks (transf_masks); layer_stride = nframes*3*H*H
    long layer_stride = 0;
};
int decode_run(const paig_task* t, const float* consts, const DecSeg& a, const DecSeg& b, bool bwd, float* partials,
               float* d_consts, int accumulate, cudaStream_t st);
size_t decode_partials_floats(const paig_task* t);

// varnet.cu
int templates_forward(const paig_task* t, const paig_params* p, float* raw, float* consts, float* hidden,
                      cudaStream_t st);
int templates_backward(const paig_task* t, const paig_params* p, const paig_params* g, const float* consts,
                       const float* hidden, const float* d_consts, float* scratch, cudaStream_t st);
size_t templates_scratch_floats(const paig_task* t);

// conv.cu -- all tensors are channel-sliced NCHW views: (n,c,y,x) = p[n*bs + (c*S + y)*S + x]
struct ConvArgs {
    const float* in = nullptr; long in_bs = 0; int Cin = 0;
    const float* mask = nullptr; long mask_bs = 0;    // nullable: input is multiplied by (mask > 0)  (ReLU adjoint)
    const float* w = nullptr;                          // [Cout][Cin][3][3]; transposed: layer weight [Cin][Cout][3][3], taps flipped
    const float* b = nullptr;                          // nullable
    float* out = nullptr; long out_bs = 0; int Cout = 0;
    int S = 0, N = 0;
    int relu = 0, transposed = 0;
    float* tc_scratch = nullptr;                       // nullable: 9*Cin*Cout floats; lets conv3x3() take the tcgen05 path
    int force_tc = 0;                                  // test hook: skip the profitability rule of conv3x3_tc()
    int QX = 0, TH = 0, FPB = 0, CK = 0;               // filled by conv3x3()
};
int conv3x3(const ConvArgs& a, cudaStream_t st);
// conv_tc.cu: tcgen05 3xTF32 implicit GEMM; -1 when the layer does not qualify.  scratch: conv_tc_scratch_floats() floats.
int conv3x3_tc(const ConvArgs& a, float* scratch, cudaStream_t st);
size_t conv_tc_scratch_floats(int Cin, int Cout);
bool conv_tc_enabled();
constexpr int kWgradMaxCtas = 296;
// Deferred fixed-order reductions of per-CTA partials: the weight-gradient kernels of a whole backward pass queue
// their (partials -> dW, db) folds here and one launch performs them all.
struct ReduceBatch {
    static constexpr int kMax = 24;
    int n = 0;
    const float* partials[kMax];
    int nparts[kMax], stride[kMax], n0[kMax], n1[kMax];
    float* out0[kMax];
    float* out1[kMax];
    bool add(const float* p, int np, int st, int a, float* o0, int b, float* o1) {
        if (n >= kMax) return false;
        partials[n] = p; nparts[n] = np; stride[n] = st; n0[n] = a; out0[n] = o0; n1[n] = b; out1[n] = o1;
        ++n;
        return true;
    }
};
int reduce_partials_batch(const ReduceBatch& b, cudaStream_t st);
struct WgradArgs {
    const float* in = nullptr; long in_bs = 0; int Cin = 0;       // layer input
    const float* in_mask = nullptr; long in_mask_bs = 0;          // nullable
    const float* g = nullptr; long g_bs = 0; int Cout = 0;        // gradient of the layer's (post-ReLU) output
    const float* act = nullptr; long act_bs = 0;                  // nullable: layer output, for the ReLU mask
    int S = 0, N = 0;
    float* partials = nullptr;                                    // >= wgrad_partials_floats(Cin, Cout)
    ReduceBatch* defer = nullptr;                                 // queue the final fold instead of launching it
    int in_plane = 0, g_plane = 0, TH = 0, async2 = 0;            // filled by conv3x3_wgrad()
};
int conv3x3_wgrad(const WgradArgs& a, float* dW, float* db, cudaStream_t st);
// wgrad_tma.cu: TMA-staged variant; -1 when the layer geometry does not qualify (then conv3x3_wgrad's own kernel runs)
int conv3x3_wgrad_tma(const WgradArgs& a, float* dW, float* db, cudaStream_t st);
// wgrad_tc.cu: tcgen05 3xTF32 variant for layers with 32 .. 128 channels on both sides; same return convention
int conv3x3_wgrad_tc(const WgradArgs& a, float* dW, float* db, cudaStream_t st);
size_t wgrad_partials_floats(int Cin, int Cout);
int reduce_partials(const float* partials, int nparts, int stride, int n0, float* out0, int n1, float* out1,
                    cudaStream_t st);
int relu_gate(float* g, long g_bs, const float* act, long act_bs, int C, int S, int N, cudaStream_t st);
int conv1x1_forward(const float* in, long in_bs, int Cin, const float* w, const float* b, float* out, long out_bs,
                    int Cout, int S, int N, int relu, cudaStream_t st);
int conv1x1_backward(const float* in, long in_bs, int Cin, const float* w, const float* dout, long dout_bs,
                     const float* act, long act_bs, int Cout, int S, int N, float* din, long din_bs, float* dW,
                     float* db, float* partials, cudaStream_t st, ReduceBatch* defer = nullptr);
int maxpool2(const float* in, long in_bs, float* out, long out_bs, int C, int So, int N, cudaStream_t st);
int maxpool2_backward(const float* in, long in_bs, const float* dout, long dout_bs, float* din, long din_bs, int C,
                      int So, int N, cudaStream_t st);
int upsample2(const float* in, long in_bs, float* out, long out_bs, int C, int Si, int N, cudaStream_t st);
int upsample2_backward(const float* dout, long dout_bs, float* din, long din_bs, int C, int Si, int N, cudaStream_t st);

// gemm.cu
enum { SPLIT_NONE = 0, SPLIT_FIXED = 1, SPLIT_AUTO = 2 };
enum { EPI_NONE = 0, EPI_RELU = 1, EPI_TANH = 2, EPI_MASK_RELU = 3, EPI_MASK_TANH = 4 };
struct GemmArgs {
    const float* A = nullptr; long sam = 0, sak = 0;     // A(m,k) = A[m*sam + k*sak]
    const float* B = nullptr; long sbk = 0, sbn = 0;     // B(k,n) = B[k*sbk + n*sbn]
    float* C = nullptr; long ldc = 0;                    // C[m*ldc + n]
    int M = 0, N = 0, K = 0;
    const float* bias = nullptr;                         // [N] added before the epilogue
    int epi = EPI_NONE;
    const float* aux = nullptr; long ldaux = 0;          // activation the MASK_* epilogues differentiate
    int accumulate = 0;                                  // C += result
    float* splitk_ws = nullptr; size_t splitk_floats = 0; // optional scratch for split-K partials ([splits][M][N])
    int split_mode = 0;                                  // SPLIT_*
    const char* tag = nullptr;                           // name in the per-launch profile (default "sgemm")
    int a_vec = 0, b_vec = 0, kper = 0;                  // filled by gemm(): 16-byte global loads legal; K per split
};
int gemm(const GemmArgs& g, cudaStream_t st);
int colsum(const float* X, int M, int N, int ld, float* out, cudaStream_t st);
// fold split-K partials g.splitk_ws[splits][M][N] into g.C with g's bias / epilogue (fixed order)
int gemm_fold_partials(const GemmArgs& g, int splits, cudaStream_t st);
// gemm_tc.cu -- 3xTF32 tcgen05 GEMM for K-major operands; see the file header for the return convention
int gemm_tc_partials(const float* A, const float* B, int M, int N, int K, bool fixed_split, float* partials,
                     size_t partial_floats, const char* tag, cudaStream_t st);
// short accumulation chains + truncation compensation: for products that feed activations (encoder.l1 forward)
int gemm_tc_partials_drained(const float* A, const float* B, int M, int N, int K, float* partials, size_t partial_floats,
                             const char* tag, cudaStream_t st);
int transpose(const float* src, float* dst, int R, int C, cudaStream_t st);
int linear_forward(const float* X, const float* W, const float* b, float* Y, int M, int K, int N, int epi,
                   cudaStream_t st, float* splitk_ws = nullptr, size_t splitk_floats = 0, const char* tag = nullptr);
int linear_dgrad(const float* dY, const float* W, float* dX, int M, int K, int N, int epi, const float* aux,
                 cudaStream_t st, float* splitk_ws = nullptr, size_t splitk_floats = 0, const char* tag = nullptr);
int linear_wgrad(const float* dY, const float* X, float* dW, float* db, int M, int K, int N, cudaStream_t st,
                 float* splitk_ws = nullptr, size_t splitk_floats = 0, const char* tag = nullptr);

}  // namespace paig
