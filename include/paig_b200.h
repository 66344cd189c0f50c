/* paig_b200.h -- C ABI of libpaig_b200.so: the PhysicsNet per-sequence training step on B200 (sm_100a).
 *
 * The reference (Luka140/paig_reproduction) has no FFI / operator layer: its boundary for this path is
 * the Python surface of PhysicsNet (nn/network/physics_models.py:40-245).  Every entry point below
 * therefore cites the reference METHOD whose arithmetic it replaces; the Python mirror of that surface
 * (paig_reproduction_b200/physics_models.py) binds them with ctypes (INTEGRATION.md shows the stub).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no torch / C++ types.  All tensor pointers are DEVICE
 *     pointers unless a name ends in _host.  The caller owns every buffer, including the workspace.
 *   - Every function returns 0 on success, non-zero on error (paig_last_error() gives the text);
 *     nothing throws across the ABI.  Work is enqueued on `stream` (a cudaStream_t passed as void*)
 *     and is asynchronous with respect to the host.
 *   - Tensors are dense row-major fp32 in the reference's own layouts (NCHW frames, [x0,y0,x1,y1,..]
 *     position vectors); the physics scalars k/equil/g/m are fp64 device scalars as in the reference
 *     (cells.py:27-29,91-93).
 */
#ifndef PAIG_B200_H
#define PAIG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PAIG_ABI_VERSION 2

enum { PAIG_CELL_SPRING = 0, PAIG_CELL_BOUNCING = 1, PAIG_CELL_GRAVITY = 2 };

/* One row of the runner's task table (runners/torch_run_physics.py:49-75) plus the constructor flags
 * that change arithmetic (physics_models.py:41-55). */
typedef struct paig_task {
    int32_t cell;         /* PAIG_CELL_* : cells.py spring / bouncing / gravity */
    int32_t n_objs;       /* COORD_UNITS/4, physics_models.py:31-37,95-96 */
    int32_t H;            /* square frame side (32, 36, 64) */
    int32_t seq_len;      /* frames per sequence given to forward (train or test length) */
    int32_t input_steps;
    int32_t pred_steps;
    int32_t alt_vel;      /* VelocityEncoder alt_vel branch, blocks.py:33-41 */
    int32_t deep_unet;    /* 0: ShallowUNet (H<40), 1: UNet (blocks.py:79-82) */
    float alpha;          /* --autoencoder_loss weight, physics_models.py:137-139 */
    int32_t batch_global; /* B of the whole job: loss normalisers use it (data-parallel shards pass the same value) */
    float gravity_A;      /* 0: the gravity cell evaluates A = exp(g) exp(2m) at every forward (default; dL/dg flows).
                           * > 0: use this value instead -- the reference computes A ONCE in the cell's constructor
                           * (cells.py:92-94), so a checkpoint with g != 0 still rolls out with the
                           * constructor-time A there; pass that A here to reproduce it. */
    int32_t flags;        /* PAIG_FLAG_* */
} paig_task;

enum { PAIG_FLAG_INFERENCE = 1 /* forward only: keep nothing for backward (eval_performance, base.py:174-218) */ };

/* weight/bias pair of one Linear or Conv2d; NULL where the layer is absent or (in a gradient table) dead. */
typedef struct paig_wb {
    float* w;
    float* b;
} paig_wb;

/* Parameter table: pointers into the caller's tensors, named after the reference state_dict.  The same
 * struct type is used for the gradient table (each entry receives dL/dparam; every entry on the step's graph is
 * WRITTEN, not accumulated, by paig_step_backward / paig_step_fused.  Not on the graph, hence left untouched: vel[*]
 * when input_steps == 1 (physics_models.py:222-223: the initial velocity is zeros), vel[*] and the physics constants
 * in STALE mode (d_output_seq == NULL), dt always; NULL entries are skipped). */
typedef struct paig_params {
    paig_wb content_l1, content_l2;         /* var_net_content.l1/.l2        blocks.py:311-322 */
    paig_wb background_l1, background_l2;   /* var_net_background.l1/.l2 */
    paig_wb template_l1, template_l2;       /* var_net_template.l1/.l2 */
    paig_wb conv[18];                       /* encoder.shallow_unet.c1..c13 or encoder.unet.c1..c18 */
    paig_wb enc_l1, enc_l2, enc_l3;         /* encoder.l1..l3               blocks.py:70-75 */
    paig_wb vel[3];                         /* velocity_encoder.init_vel_mlp.{0,2,4}, or [0]=init_vel_linear */
    float* dt;                              /* rollout_cell.dt (fp32 scalar; never receives a gradient) */
    double* phys0;                          /* spring: k      gravity: g      bouncing: NULL */
    double* phys1;                          /* spring: equil  gravity: m (no gradient: A is recomputed from g,m; dm unused) */
} paig_params;

/* Outputs of the forward pass the reference caches on the module (physics_models.py:204-245). Any
 * pointer may be NULL to skip materialising that tensor (paig_step_fused never writes frames). */
typedef struct paig_outputs {
    float* output_seq;    /* [B, T-in, 3, H, H]      conv_feedforward return value */
    float* recons_out;    /* [B, in+pr, 3, H, H]     self.recons_out */
    float* enc_pos;       /* [B, in+pr, 2n]          self.enc_pos */
    float* pos_vel_seq;   /* [B, T-in+1, 4n]         self.pos_vel_seq */
    float* enc_masks;     /* [B*(in+pr), n+1, H, H]  self.enc_masks */
    float* masked_objs;   /* [n, B*(in+pr), 3, H, H] self.masked_objs (object-major) */
    float* templates;     /* [n*t*t | n*3*t*t | 3*H*H] raw template, contents, background (pre-sigmoid) */
    float* losses;        /* [4]: train, pred, extrap, recons  (compute_loss, physics_models.py:119-142) */
} paig_outputs;

int paig_abi_version(void);
const char* paig_last_error(void);

/* Bytes of scratch the step functions need for `B` local sequences of task `t`. */
size_t paig_workspace_bytes(const paig_task* t, int B);

/* ---- whole step ------------------------------------------------------------------------------- */

/* conv_feedforward (physics_models.py:204-245) + compute_loss (:119-142), materialising every tensor
 * requested in `out`.  x: [B, seq_len, 3, H, H].  Keeps what backward needs in `workspace`. */
int paig_step_forward(const paig_task* t, const paig_params* p, const float* x, int B,
                      const paig_outputs* out, void* workspace, void* stream);

/* autograd backward of the step (base.py:151) given upstream gradients of the forward outputs.
 * d_output_seq / d_recons_out / d_enc_pos / d_pos_vel_seq may be NULL (treated as zero: STALE mode,
 * SURVEY Q1, passes d_output_seq = NULL).  Must follow paig_step_forward on the same workspace. */
int paig_step_backward(const paig_task* t, const paig_params* p, const paig_params* grads, const float* x, int B,
                       const float* d_output_seq, const float* d_recons_out, const float* d_enc_pos,
                       const float* d_pos_vel_seq, void* workspace, void* stream);

/* LIVE training step in one call: forward, the three losses and all parameter gradients of
 * train = pred + alpha*recons, with the decoder's loss gradient formed in-kernel (no frames written).
 * out->losses receives the four scalars; out->enc_pos / pos_vel_seq are written when non-NULL. */
int paig_step_fused(const paig_task* t, const paig_params* p, const paig_params* grads, const float* x, int B,
                    const paig_outputs* out, void* workspace, void* stream);

/* Data-parallel overlap hook: arm a cudaEvent_t (passed as void*) that the NEXT paig_step_fused* call on this thread
 * records as soon as every gradient except the UNet conv layers' (encoder.{shallow_unet,unet}.c*) is final -- that is
 * before the UNet backward, ~half of the step.  The caller makes a side stream wait for the event and all-reduces that
 * part of its flat gradient buffer there while the UNet backward runs (paig_reproduction_b200/parallel.py).  The
 * reference has no distributed code; SURVEY 8(e). */
void paig_set_early_grad_event(void* cuda_event);

/* Same as paig_step_fused but with HOST buffers: x_host (pinned or pageable) is copied to the device
 * staging area inside the workspace, and the four losses are copied back to losses_host.  This is the
 * end-to-end call bench.py times as `e2e`. */
int paig_step_fused_host(const paig_task* t, const paig_params* p, const paig_params* grads, const float* x_host,
                         int B, float* losses_host, void* workspace, void* stream);

/* The same end-to-end step with the input copy pipelined: paig_stage_input_host enqueues the host->device copy of a
 * batch into staging slot 0 or 1 of the workspace on `copy_stream` (and records an event); paig_step_fused_staged makes
 * `stream` wait for that slot and runs the fused step on it, copying the four losses to losses_host (nullable).  A
 * training loop stages batch k+1 into the other slot right before it launches step k: the copy hides under the step.
 * The caller must not re-stage a slot before the step that reads it has finished. */
int paig_stage_input_host(const paig_task* t, const float* x_host, int B, int slot, void* workspace, void* copy_stream);
int paig_step_fused_staged(const paig_task* t, const paig_params* p, const paig_params* grads, int B, int slot,
                           float* losses_host, void* workspace, void* stream);

/* ---- stages (also used by the parity tests) ---------------------------------------------------- */

/* cells.py:31-51 / 60-83 / 96-106 iterated `steps` times.  pos_vel_seq: [B, steps+1, 4n]; row 0 must
 * already hold the initial state (pos || vel).  dt: fp32 scalar; phys0/phys1 as in paig_params. */
int paig_rollout_forward(int cell, int n_objs, int B, int steps, const float* dt, const double* phys0,
                         const double* phys1, float* pos_vel_seq, void* stream);

/* Reverse sweep.  d_seq: [B, steps+1, 4n] upstream gradient of every row of pos_vel_seq (pos and vel parts).
 * d_state0: [B, 4n] receives dL/d(initial pos||vel).  d_phys: [2] fp64, WRITTEN (dk, dequil | dg, 0 | 0, 0).
 * The sum over sequences runs in a fixed order for any B (run-to-run bit-identical). */
int paig_rollout_backward(int cell, int n_objs, int B, int steps, const float* dt, const double* phys0,
                          const double* phys1, const float* pos_vel_seq, const float* d_seq, float* d_state0,
                          double* d_phys, void* stream);

/* VariableFromNetwork x3 (blocks.py:318-322) + the decoder's constant preprocessing
 * (physics_models.py:163-171,182): raw [n*t*t | 3n*t*t | 3*H*H], consts = [template+5 | sigmoid(contents) |
 * sigmoid(background)], hidden [3*200] tanh activations kept for backward. */
int paig_templates_forward(const paig_task* t, const paig_params* p, float* raw, float* consts, float* hidden,
                           void* stream);
int paig_templates_backward(const paig_task* t, const paig_params* p, const paig_params* grads, const float* consts,
                            const float* hidden, const float* d_consts, void* workspace, void* stream);

/* conv_st_decoder + stn (physics_models.py:151-199, stn.py:5-16) for F frames.  loc: [F, 2n].
 * frames (nullable): [F,3,H,H].  If target != NULL, sse[F] receives sum_chw (target-frame)^2 per frame;
 * target frame f lives at target + (f / frames_per_seq) * target_seq_stride + (f % frames_per_seq) * 3*H*H. */
int paig_decode_forward(const paig_task* t, const float* consts, const float* loc, int F, float* frames,
                        const float* target, long target_seq_stride, int frames_per_seq, float* sse, void* stream);

/* The decoder's per-layer intermediates the reference caches as self.transf_contents / self.transf_masks
 * (physics_models.py:190,196) for F frames: transf_contents, transf_masks: [n+1][F][3][H][H] -- layer o < n is
 * object o's bilinear-sampled sigmoid(content) / its softmax weight (repeated over the 3 channels, as the
 * reference's tiled template makes it), layer n the background. */
int paig_decode_layers(const paig_task* t, const float* consts, const float* loc, int F, float* transf_contents,
                       float* transf_masks, void* stream);

/* Decoder backward.  Either d_frames [F,3,H,H] is given, or (d_frames == NULL) the gradient is formed
 * in-kernel as 2*scale[f % frames_per_seq]*(frame - target) (scale: device [frames_per_seq]).  d_loc [F,2n] is
 * WRITTEN; d_consts (same layout as consts) is ACCUMULATED (caller zeroes it). sse nullable as above. */
int paig_decode_backward(const paig_task* t, const float* consts, const float* loc, int F, const float* d_frames,
                         const float* target, long target_seq_stride, int frames_per_seq, const float* scale,
                         float* d_loc, float* d_consts, float* sse, void* workspace, void* stream);

/* ConvolutionalEncoder.forward (blocks.py:77-103) on N frames; frame f is read at
 * x + (f / frames_per_seq) * seq_stride + (f % frames_per_seq) * 3*H*H.  enc_pos: [N, 2n]. */
int paig_encoder_forward(const paig_task* t, const paig_params* p, const float* x, long seq_stride,
                         int frames_per_seq, int N, float* enc_pos, float* enc_masks, float* masked_objs,
                         void* workspace, void* stream);
int paig_encoder_backward(const paig_task* t, const paig_params* p, const paig_params* grads, const float* x,
                          long seq_stride, int frames_per_seq, int N, const float* d_enc_pos, void* workspace,
                          void* stream);

/* VelocityEncoder.forward (blocks.py:31-49).  enc_pos: [B, enc_steps, 2n] -> vel [B, 2n]. */
int paig_velocity_forward(const paig_task* t, const paig_params* p, const float* enc_pos, int B, float* vel,
                          void* workspace, void* stream);
int paig_velocity_backward(const paig_task* t, const paig_params* p, const paig_params* grads, const float* enc_pos,
                           int B, const float* d_vel, float* d_enc_pos_accum, void* workspace, void* stream);

/* compute_loss's per-frame squared error (physics_models.py:122-131) for the drop-in compute_loss():
 * sse[b,f] = sum_chw (x[b, first+f] - pred[b,f])^2, and d_pred = 2 * d_sse[b,f] * (pred - x). */
int paig_frame_sse_forward(const float* x, long x_seq_stride, int first, const float* pred, int B, int F, int chw,
                           float* sse, void* stream);
int paig_frame_sse_backward(const float* x, long x_seq_stride, int first, const float* pred, int B, int F, int chw,
                            const float* d_sse, float* d_pred, void* stream);

/* ---- layer primitives (exported for unit tests of the encoder's building blocks) ------------------- */

/* nn.Conv2d(Cin, Cout, 3, padding="same") (+ReLU) on x [N,Cin,S,S] -> y [N,Cout,S,S]   (blocks.py:246-276) */
int paig_conv3x3_forward(const float* x, const float* w, const float* b, float* y, int N, int Cin, int Cout, int S,
                         int relu, void* stream);
/* its backward: dy is the gradient of the (post-ReLU) output y; dx (nullable) [N,Cin,S,S], dw, db are WRITTEN.
 * workspace: at least 296 * (Cout*Cin*9 + Cout) floats. */
int paig_conv3x3_backward(const float* x, const float* w, const float* y, const float* dy, float* dx, float* dw,
                          float* db, int N, int Cin, int Cout, int S, int relu, void* workspace, void* stream);

/* ---- the steps either side of the hot path in the training loop ---------------- */

/* optimizer.step() of base.py:152 over a flat buffer, torch.optim defaults of the base.py:12-17 table.
 * kind: 0 sgd, 1 momentum (0.9), 2 rmsprop (alpha .99, eps 1e-8), 3 adam (betas .9/.999, eps 1e-8).
 * state0 / state1: momentum buffer | square average | (exp_avg, exp_avg_sq); zero-initialised by the caller;
 * unused ones may be NULL.  step: 1-based step count (Adam bias correction, first momentum step). */
int paig_optimizer_step(int kind, float* params, const float* grads, float* state0, float* state1, long n, float lr,
                        int step, void* stream);
int paig_optimizer_step_f64(int kind, double* params, const double* grads, double* state0, double* state1, long n,
                            double lr, int step, void* stream);

/* get_batch (physics_models.py:113-117, iterators.py:26-40,64) from a device-resident uint8 dataset:
 * out[b][j] = data[idx[b]][j] / 255 as float32, j < seq_elems (= T*H*W*C; the reference's layout change is a reshape,
 * SURVEY Q15).  idx: device int64 [B]. */
int paig_gather_batch_u8(const uint8_t* data, long seq_elems, const long* idx, int B, float* out, void* stream);

/* Measurement hooks (bench.py): cumulative number of kernels this library launched; per-launch CUDA-event
 * timing between begin/end, reported as "kernel-name launches total_ms" lines. */
long paig_launch_count(void);
void paig_profile_begin(void);
int paig_profile_end(char* buf, size_t cap);

/* Test hook: C[M,N] = A[M,K] . B[N,K]^T on the tcgen05 3xTF32 path (csrc/gemm_tc.cu); scratch holds the split-K partials.
 * fixed_split: 0 = split chosen for occupancy, 1 = batch-invariant split by K only, 2 = the drained kernel (short TMEM
 * accumulation chains + truncation compensation) that encoder.l1's forward product uses. */
int paig_debug_gemm_tc(const float* A, const float* B, float* C, int M, int N, int K, int fixed_split, float* scratch,
                       long scratch_floats, void* stream);

/* Test hook: the tcgen05 implicit-GEMM convolution (csrc/conv_tc.cu) on its own.  x [N,Cin,S,S] -> y [N,Cout,S,S];
 * w: [Cout][Cin][3][3], or with transposed != 0 the layer weight [Cin][Cout][3][3] applied with flipped taps (the data
 * gradient of blocks.py's Conv2d).  scratch: 9*Cin*Cout floats.  Fails when the shape does not qualify
 * (Cin % 8, Cout in {16,32,48,64,96,128}, S in {8,16,32,64}). */
int paig_debug_conv3x3_tc(const float* x, const float* w, const float* b, float* y, int N, int Cin, int Cout, int S,
                          int relu, int transposed, float* scratch, void* stream);

/* Test hook: offset (in floats) of a named workspace region ("act", "grad" with a UNet buffer index; "logits",
 * "d_logits", "enc_pos", "d_enc_pos", "seq", "d_seq", "d_consts", "consts", "A", "dA"), or -1. */
long paig_debug_workspace_offset(const paig_task* t, int B, const char* region, int index);
/* Test hook: where the (post-ReLU) output of UNet conv `layer` (0-based) lives in the workspace:
 * view[0] = offset in floats, view[1] = batch stride, view[2] = channels, view[3] = side, view[4] = has ReLU. */
int paig_debug_unet_conv_view(const paig_task* t, int B, int layer, long view[5]);

#ifdef __cplusplus
}
#endif
#endif /* PAIG_B200_H */
