// The whole PhysicsNet step: conv_feedforward (physics_models.py:204-245) + compute_loss (:119-142) and the
// backward pass autograd would run (base.py:151), as a fixed sequence of kernel launches on one stream.
//
//   forward : templates -> encoder -> velocity -> rollout -> decode(recons + rollout frames) -> losses
//   backward: decode^T (in the same kernel as decode when the loss gradient is formed in-kernel)
//             -> rollout^T -> velocity^T -> encoder^T -> templates^T
#include "common.cuh"
#include "internal.h"

#include <mutex>
#include "layout.h"

#include <cstring>

namespace paig {

// losses = [train, pred, extrap, recons]  (physics_models.py:122-141; means over the GLOBAL batch, so
// data-parallel shards sum to the job's loss).  One block, fixed order.
__global__ void __launch_bounds__(256) loss_finalize_kernel(const float* __restrict__ sse, int B, int e, int steps, int pr,
                                                            float alpha, float Bg, float* __restrict__ losses) {
    __shared__ float scratch[33];
    float r = 0.f, p = 0.f, x = 0.f;
    for (int i = threadIdx.x; i < B * e; i += blockDim.x) r += sse[i];
    const float* s2 = sse + (long)B * e;
    for (int i = threadIdx.x; i < B * steps; i += blockDim.x) {
        if (i % steps < pr) p += s2[i];
        else x += s2[i];
    }
    r = block_sum(r, scratch);
    p = block_sum(p, scratch);
    x = block_sum(x, scratch);
    if (threadIdx.x == 0) {
        const float recons = r / (Bg * (float)e), pred = p / (Bg * (float)pr);
        const int ex = steps - pr;
        const float extrap = ex > 0 ? x / (Bg * (float)ex) : 0.f;
        losses[0] = alpha > 0.f ? pred + alpha * recons : pred;
        losses[1] = pred;
        losses[2] = extrap;
        losses[3] = recons;
    }
}

__global__ void __launch_bounds__(256) axpy_kernel(float* __restrict__ y, const float* __restrict__ x, long n) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] += x[i];
}

static float batch_global(const paig_task* t, int B) { return (float)(t->batch_global > 0 ? t->batch_global : B); }

static void segments(const paig_task* t, const Layout& L, float* ws, const float* x, DecSeg* A, DecSeg* R) {
    const Dims& d = L.d;
    A->nframes = L.B * d.e;
    A->fps = d.e;
    A->loc = ws + L.enc_pos;
    A->loc_row_stride = 2 * d.n;
    A->loc_seq_stride = (long)d.e * 2 * d.n;
    A->target = x;
    A->tgt_seq_stride = (long)d.T * d.CHW;
    A->sse = ws + L.sse;
    R->nframes = L.B * d.steps;
    R->fps = d.steps;
    R->loc = ws + L.seq + 4 * d.n;                    // row s+1 of pos_vel_seq holds the state decoded as frame s
    R->loc_row_stride = 4 * d.n;
    R->loc_seq_stride = (long)(d.steps + 1) * 4 * d.n;
    R->target = x + (long)d.in * d.CHW;
    R->tgt_seq_stride = (long)d.T * d.CHW;
    R->sse = ws + L.sse + (long)L.B * d.e;
    (void)t;
}

static int forward_common(const paig_task* t, const paig_params* p, const Layout& L, const float* x,
                          const paig_outputs* out, float* ws, cudaStream_t st) {
    const Dims& d = L.d;
    int rc;
    Side* sd = side_cur();                            // fused step: the VariableFromNetwork branch runs beside the encoder
    if (sd) sd->fork1();
    if ((rc = templates_forward(t, p, out && out->templates ? out->templates : ws + L.raw, ws + L.consts,
                                ws + L.hidden, sd ? sd->s1 : st)))
        return rc;
    if ((rc = encoder_forward(t, p, L, x, (long)d.T * d.CHW, d.e, out ? out->enc_pos : nullptr,
                              out ? out->enc_masks : nullptr, out ? out->masked_objs : nullptr, ws, st)))
        return rc;
    if ((rc = velocity_forward(t, p, L, ws + L.enc_pos, ws, st))) return rc;
    if ((rc = rollout_forward(t->cell, d.n, L.B, d.steps, p->dt, p->phys0, p->phys1, t->gravity_A, ws + L.seq, st))) return rc;
    if (out && out->pos_vel_seq &&
        cudaMemcpyAsync(out->pos_vel_seq, ws + L.seq, (size_t)L.B * (d.steps + 1) * 4 * d.n * sizeof(float),
                        cudaMemcpyDeviceToDevice, st) != cudaSuccess)
        return check_launch("copy pos_vel_seq");
    if (sd) sd->join1();                              // decoder constants (and the side work of the encoder) are ready
    return 0;
}

static int finalize_losses(const paig_task* t, const Layout& L, float* ws, float* losses_out, cudaStream_t st) {
    const Dims& d = L.d;
    launch(loss_finalize_kernel, dim3(1), dim3(256), 0, st, (const float*)(ws + L.sse), L.B, d.e, d.steps, d.pr,
           t->alpha, batch_global(t, L.B), ws + L.losses);
    int rc = check_launch("loss_finalize");
    if (rc) return rc;
    if (losses_out &&
        cudaMemcpyAsync(losses_out, ws + L.losses, 4 * sizeof(float), cudaMemcpyDeviceToDevice, st) != cudaSuccess)
        return check_launch("copy losses");
    return 0;
}

namespace {
struct SideScope {                                   // side streams live for exactly one step call (fused, forward or backward)
    Side* sd;
    explicit SideScope(cudaStream_t st) : sd(side_begin(st)) {}
    ~SideScope() { if (sd) side_end(); }
};
}  // namespace

int step_forward(const paig_task* t, const paig_params* p, const float* x, int B, const paig_outputs* out, float* ws,
                 cudaStream_t st) {
    const Layout L = make_layout(t, B);
    // the split forward / backward calls of the drop-in module take the same side streams as the fused step (the
    // VariableFromNetwork branch beside the encoder, the l1 transposes, the twelve weight-gradient launches ...): serial,
    // that path ran 0.15 ms behind (tools/dropin_probe.py)
    SideScope scope(st);
    Side* sd = scope.sd;
    int rc = forward_common(t, p, L, x, out, ws, st);
    if (rc) { if (sd) { sd->join1(); sd->join2(); } return rc; }
    DecSeg A, R;
    segments(t, L, ws, x, &A, &R);
    A.frames = out ? out->recons_out : nullptr;
    R.frames = out ? out->output_seq : nullptr;
    rc = decode_run(t, ws + L.consts, A, R, false, nullptr, nullptr, 0, st);
    if (!rc) rc = finalize_losses(t, L, ws, out ? out->losses : nullptr, st);
    if (sd) { sd->join1(); sd->join2(); }             // every fork is joined before the call returns
    return rc;
}

// Everything after the decoder's backward: d_seq / d_enc_pos / d_consts are complete in the workspace.
thread_local void* g_early_event = nullptr;

static int backward_tail(const paig_task* t, const paig_params* p, const paig_params* g, const Layout& L,
                         const float* x, float* ws, cudaStream_t st, bool templates_done = false) {
    const Dims& d = L.d;
    int rc;
    // physics-constant gradients go straight to the caller's fp64 slots (spring: k, equil; gravity: g; bouncing: none)
    double* dphys_scratch = reinterpret_cast<double*>(ws + L.dphys);
    const long rs = 4L * d.n;
    if ((rc = rollout_backward(t->cell, d.n, L.B, d.steps, p->dt, p->phys0, p->phys1, t->gravity_A, ws + L.seq, ws + L.d_seq,
                               ws + L.d_seq + 2 * d.n, (d.steps + 1) * rs, rs, 1, ws + L.d_state0,
                               t->cell != PAIG_CELL_BOUNCING ? (double*)g->phys0 : nullptr,
                               t->cell == PAIG_CELL_SPRING ? (double*)g->phys1 : nullptr, dphys_scratch, st)))
        return rc;
    if ((rc = velocity_backward(t, p, g, L, ws + L.d_state0, ws + L.d_enc_pos, ws, st))) return rc;
    if ((rc = encoder_backward(t, p, g, L, x, (long)d.T * d.CHW, d.e, ws + L.d_enc_pos, ws, st))) return rc;
    if (templates_done) return 0;
    return templates_backward(t, p, g, ws + L.consts, ws + L.hidden, ws + L.d_consts, ws + L.tmpl_scratch, st);
}

static void route_dloc(const Layout& L, float* ws, DecSeg* A, DecSeg* R) {
    const Dims& d = L.d;
    A->dloc = ws + L.d_enc_pos;
    A->dloc_row_stride = A->loc_row_stride;
    A->dloc_seq_stride = A->loc_seq_stride;
    R->dloc = ws + L.d_seq + 4 * d.n;                 // position half of rows 1..steps of d(pos_vel_seq)
    R->dloc_row_stride = R->loc_row_stride;
    R->dloc_seq_stride = R->loc_seq_stride;
}

static bool refuse_inference(const paig_task* t, const char* what) {
    if (!(t->flags & PAIG_FLAG_INFERENCE)) return false;
    set_error("%s: the task carries PAIG_FLAG_INFERENCE (its workspace holds nothing for a backward pass)", what);
    return true;
}

static int step_backward_body(const paig_task* t, const paig_params* p, const paig_params* g, const float* x, int B,
                              const float* d_output_seq, const float* d_recons_out, const float* d_enc_pos,
                              const float* d_pos_vel_seq, float* ws, cudaStream_t st, const Layout& L, Side* sd);

int step_backward(const paig_task* t, const paig_params* p, const paig_params* g, const float* x, int B,
                  const float* d_output_seq, const float* d_recons_out, const float* d_enc_pos,
                  const float* d_pos_vel_seq, float* ws, cudaStream_t st) {
    if (refuse_inference(t, "step_backward")) return 1;
    const Layout L = make_layout(t, B);
    const Dims& d = L.d;
    SideScope scope(st);
    Side* sd = scope.sd;
    const int rc_all = step_backward_body(t, p, g, x, B, d_output_seq, d_recons_out, d_enc_pos, d_pos_vel_seq, ws, st, L, sd);
    if (sd) { sd->join1(); sd->join2(); }
    return rc_all;
}

static int step_backward_body(const paig_task* t, const paig_params* p, const paig_params* g, const float* x, int B,
                              const float* d_output_seq, const float* d_recons_out, const float* d_enc_pos,
                              const float* d_pos_vel_seq, float* ws, cudaStream_t st, const Layout& L, Side* sd) {
    const Dims& d = L.d;
    const size_t seq_fl = (size_t)B * (d.steps + 1) * 4 * d.n, ep_fl = (size_t)L.N * 2 * d.n;
    if (cudaMemsetAsync(ws + L.d_seq, 0, seq_fl * sizeof(float), st) != cudaSuccess) return check_launch("memset");
    if (cudaMemsetAsync(ws + L.d_enc_pos, 0, ep_fl * sizeof(float), st) != cudaSuccess) return check_launch("memset");
    DecSeg A, R, none;
    segments(t, L, ws, x, &A, &R);
    route_dloc(L, ws, &A, &R);
    A.target = R.target = nullptr;                    // gradients come from memory; losses were done in forward
    A.sse = R.sse = nullptr;
    A.dframes = d_recons_out;
    R.dframes = d_output_seq;
    if (!d_recons_out) A = none;                      // STALE mode: that branch receives no gradient
    if (!d_output_seq) R = none;
    int rc;
    const size_t CN = (size_t)d.n * d.t * d.t * 4 + 3 * d.HW;
    if (A.nframes + R.nframes > 0) {
        if ((rc = decode_run(t, ws + L.consts, A.nframes ? A : R, A.nframes ? R : none, true, ws + L.dec_partials,
                             ws + L.d_consts, 0, st)))
            return rc;
    } else {
        if (cudaMemsetAsync(ws + L.d_consts, 0, CN * sizeof(float), st) != cudaSuccess) return check_launch("memset");
    }
    if (d_enc_pos) {
        launch(axpy_kernel, dim3(cdiv(ep_fl, 256)), dim3(256), 0, st, ws + L.d_enc_pos, d_enc_pos, (long)ep_fl);
        if ((rc = check_launch("axpy"))) return rc;
    }
    if (d_pos_vel_seq) {
        launch(axpy_kernel, dim3(cdiv(seq_fl, 256)), dim3(256), 0, st, ws + L.d_seq, d_pos_vel_seq, (long)seq_fl);
        if ((rc = check_launch("axpy"))) return rc;
    }
    // the VariableFromNetwork backward only needs the decoder's d_consts: beside the rollout / encoder chain
    if (sd) {
        sd->fork1();
        if ((rc = templates_backward(t, p, g, ws + L.consts, ws + L.hidden, ws + L.d_consts, ws + L.tmpl_scratch, sd->s1))) return rc;
    }
    return backward_tail(t, p, g, L, x, ws, st, sd != nullptr);
}

int step_fused(const paig_task* t, const paig_params* p, const paig_params* g, const float* x, int B,
               const paig_outputs* out, float* ws, cudaStream_t st) {
    if (refuse_inference(t, "step_fused")) return 1;
    const Layout L = make_layout(t, B);
    const Dims& d = L.d;
    SideScope scope(st);
    Side* sd = scope.sd;
    int rc = forward_common(t, p, L, x, out, ws, st);
    if (rc) return rc;
    if (cudaMemsetAsync(ws + L.d_seq, 0, (size_t)B * (d.steps + 1) * 4 * d.n * sizeof(float), st) != cudaSuccess)
        return check_launch("memset d_seq");
    DecSeg A, R;
    segments(t, L, ws, x, &A, &R);
    route_dloc(L, ws, &A, &R);
    // loss weights of the in-kernel upstream gradient 2 s_r (out - target): alpha/(Bg e) on every reconstruction
    // frame; 1/(Bg pr) on the first pr rollout frames, 0 on the extrapolation frames (physics_models.py:122-139)
    const float Bg = batch_global(t, B);
    A.scale_lo = A.scale_hi = t->alpha > 0.f ? t->alpha / (Bg * (float)d.e) : 0.f;
    A.scale_split = d.e;
    R.scale_lo = 1.f / (Bg * (float)d.pr);
    R.scale_hi = 0.f;
    R.scale_split = d.pr;
    A.use_scale = R.use_scale = 1;
    A.frames = out ? out->recons_out : nullptr;       // normally NULL: the fused step never writes frames
    R.frames = out ? out->output_seq : nullptr;
    if ((rc = decode_run(t, ws + L.consts, A, R, true, ws + L.dec_partials, ws + L.d_consts, 0, st))) return rc;
    // the loss scalars and the VariableFromNetwork backward only need the decoder's results: beside the main chain
    cudaStream_t ts = sd ? sd->s1 : st;
    if (sd) sd->fork1();
    if ((rc = finalize_losses(t, L, ws, out ? out->losses : nullptr, ts))) return rc;
    if (sd && (rc = templates_backward(t, p, g, ws + L.consts, ws + L.hidden, ws + L.d_consts, ws + L.tmpl_scratch, ts)))
        return rc;
    if ((rc = backward_tail(t, p, g, L, x, ws, st, sd != nullptr))) return rc;
    if (sd) { sd->join1(); sd->join2(); }
    g_early_event = nullptr;                          // armed for one step (paig_set_early_grad_event)
    return 0;
}

}  // namespace paig

using namespace paig;

extern "C" {

size_t paig_workspace_bytes(const paig_task* t, int B) {
    if (!t || B < 0 || t->n_objs < 1 || t->n_objs > kMaxObjs || t->H < 8 || t->seq_len <= t->input_steps + t->pred_steps) {
        set_error("paig_workspace_bytes: invalid task");
        return 0;
    }
    return make_layout(t, B > 0 ? B : 1).total * sizeof(float);
}

int paig_step_forward(const paig_task* t, const paig_params* p, const float* x, int B, const paig_outputs* out,
                      void* workspace, void* stream) {
    if (!valid_task(t)) return 1;
    return step_forward(t, p, x, B, out, (float*)workspace, (cudaStream_t)stream);
}

int paig_step_backward(const paig_task* t, const paig_params* p, const paig_params* grads, const float* x, int B,
                       const float* d_output_seq, const float* d_recons_out, const float* d_enc_pos,
                       const float* d_pos_vel_seq, void* workspace, void* stream) {
    if (!valid_task(t)) return 1;
    return step_backward(t, p, grads, x, B, d_output_seq, d_recons_out, d_enc_pos, d_pos_vel_seq, (float*)workspace,
                         (cudaStream_t)stream);
}

void paig_set_early_grad_event(void* cuda_event) { g_early_event = cuda_event; }

int paig_step_fused(const paig_task* t, const paig_params* p, const paig_params* grads, const float* x, int B,
                    const paig_outputs* out, void* workspace, void* stream) {
    if (!valid_task(t)) return 1;
    return step_fused(t, p, grads, x, B, out, (float*)workspace, (cudaStream_t)stream);
}

int paig_step_fused_host(const paig_task* t, const paig_params* p, const paig_params* grads, const float* x_host, int B,
                         float* losses_host, void* workspace, void* stream) {
    if (!valid_task(t) || refuse_inference(t, "step_fused_host")) return 1;
    const Layout L = make_layout(t, B);
    float* ws = (float*)workspace;
    cudaStream_t st = (cudaStream_t)stream;
    if (cudaMemcpyAsync(ws + L.x_stage, x_host, (size_t)B * L.d.T * L.d.CHW * sizeof(float), cudaMemcpyHostToDevice, st) != cudaSuccess)
        return check_launch("step_fused_host: input copy");
    int rc = step_fused(t, p, grads, ws + L.x_stage, B, nullptr, ws, st);
    if (rc) return rc;
    if (cudaMemcpyAsync(losses_host, ws + L.losses, 4 * sizeof(float), cudaMemcpyDeviceToHost, st) != cudaSuccess)
        return check_launch("step_fused_host: losses copy");
    return check_launch("step_fused_host");
}

// ---- pipelined host input: copy batch k+1 on a side stream while step k computes -----------------------------------
// One event per (workspace, slot): two workspaces / devices / nets in one process never share a staging event.
#ifndef PAIG_EMU
namespace {
struct StageEvents { const void* ws; cudaEvent_t ev[2]; };
constexpr int kMaxStagedWorkspaces = 64;
StageEvents g_stage[kMaxStagedWorkspaces];
int g_stage_n = 0;
std::mutex g_stage_mu;
// create = false: look up only (nullptr when the slot was never staged on this workspace)
cudaEvent_t stage_event(const void* ws, int slot, bool create) {
    std::lock_guard<std::mutex> lock(g_stage_mu);
    int k = 0;
    for (; k < g_stage_n; ++k)
        if (g_stage[k].ws == ws) break;
    if (k == g_stage_n) {
        if (!create || g_stage_n == kMaxStagedWorkspaces) return nullptr;
        g_stage[k] = StageEvents{ws, {nullptr, nullptr}};
        ++g_stage_n;
    }
    if (!g_stage[k].ev[slot] && create && cudaEventCreateWithFlags(&g_stage[k].ev[slot], cudaEventDisableTiming) != cudaSuccess)
        return nullptr;
    return g_stage[k].ev[slot];
}
}  // namespace
#endif

int paig_stage_input_host(const paig_task* t, const float* x_host, int B, int slot, void* workspace, void* copy_stream) {
    if (!valid_task(t) || refuse_inference(t, "stage_input_host")) return 1;
    if (slot != 0 && slot != 1) { set_error("stage_input_host: slot must be 0 or 1"); return 1; }
    const Layout L = make_layout(t, B);
    float* ws = (float*)workspace;
    cudaStream_t cs = (cudaStream_t)copy_stream;
    if (cudaMemcpyAsync(ws + (slot ? L.x_stage2 : L.x_stage), x_host, (size_t)B * L.d.T * L.d.CHW * sizeof(float),
                        cudaMemcpyHostToDevice, cs) != cudaSuccess)
        return check_launch("stage_input_host: copy");
#ifndef PAIG_EMU
    cudaEvent_t ev = stage_event(workspace, slot, true);
    if (!ev) { set_error("stage_input_host: no staging event (more than %d workspaces staged?)", kMaxStagedWorkspaces); return 1; }
    if (cudaEventRecord(ev, cs) != cudaSuccess) return check_launch("stage_input_host: event");
#endif
    return check_launch("stage_input_host");
}

int paig_step_fused_staged(const paig_task* t, const paig_params* p, const paig_params* grads, int B, int slot,
                           float* losses_host, void* workspace, void* stream) {
    if (!valid_task(t)) return 1;
    if (slot != 0 && slot != 1) { set_error("step_fused_staged: slot must be 0 or 1"); return 1; }
    const Layout L = make_layout(t, B);
    float* ws = (float*)workspace;
    cudaStream_t st = (cudaStream_t)stream;
#ifndef PAIG_EMU
    cudaEvent_t ev = stage_event(workspace, slot, false);
    if (!ev) { set_error("step_fused_staged: slot %d of this workspace was never staged", slot); return 1; }
    if (cudaStreamWaitEvent(st, ev, 0) != cudaSuccess) return check_launch("step_fused_staged: wait");
#endif
    int rc = step_fused(t, p, grads, ws + (slot ? L.x_stage2 : L.x_stage), B, nullptr, ws, st);
    if (rc) return rc;
    if (losses_host && cudaMemcpyAsync(losses_host, ws + L.losses, 4 * sizeof(float), cudaMemcpyDeviceToHost, st) != cudaSuccess)
        return check_launch("step_fused_staged: losses copy");
    return check_launch("step_fused_staged");
}

int paig_encoder_forward(const paig_task* t, const paig_params* p, const float* x, long seq_stride, int frames_per_seq,
                         int N, float* enc_pos, float* enc_masks, float* masked_objs, void* workspace, void* stream) {
    if (!valid_task(t)) return 1;
    if (N % (t->input_steps + t->pred_steps) != 0) {
        set_error("encoder: N=%d must be a multiple of input_steps+pred_steps", N);
        return 1;
    }
    const Layout L = make_layout(t, N / (t->input_steps + t->pred_steps));
    return encoder_forward(t, p, L, x, seq_stride, frames_per_seq, enc_pos, enc_masks, masked_objs, (float*)workspace,
                           (cudaStream_t)stream);
}

int paig_encoder_backward(const paig_task* t, const paig_params* p, const paig_params* grads, const float* x,
                          long seq_stride, int frames_per_seq, int N, const float* d_enc_pos, void* workspace,
                          void* stream) {
    if (!valid_task(t)) return 1;
    const Layout L = make_layout(t, N / (t->input_steps + t->pred_steps));
    return encoder_backward(t, p, grads, L, x, seq_stride, frames_per_seq, d_enc_pos, (float*)workspace,
                            (cudaStream_t)stream);
}

int paig_velocity_forward(const paig_task* t, const paig_params* p, const float* enc_pos, int B, float* vel,
                          void* workspace, void* stream) {
    if (!valid_task(t)) return 1;
    const Layout L = make_layout(t, B);
    float* ws = (float*)workspace;
    int rc = velocity_forward(t, p, L, enc_pos, ws, (cudaStream_t)stream);
    if (rc) return rc;
    // vel [B, 2n] = velocity half of row 0 of the rollout state
    const Dims& d = L.d;
    if (vel) {
        // strided copy: row 0 of sequence b starts at seq + b*(steps+1)*4n; take columns [2n, 4n)
        cudaMemcpy2DAsync(vel, 2 * d.n * sizeof(float), ws + L.seq + 2 * d.n, (size_t)(d.steps + 1) * 4 * d.n * sizeof(float),
                          2 * d.n * sizeof(float), B, cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
    }
    return check_launch("velocity_forward");
}

int paig_velocity_backward(const paig_task* t, const paig_params* p, const paig_params* grads, const float* enc_pos,
                           int B, const float* d_vel, float* d_enc_pos_accum, void* workspace, void* stream) {
    if (!valid_task(t)) return 1;
    (void)enc_pos;
    const Layout L = make_layout(t, B);
    const Dims& d = L.d;
    float* ws = (float*)workspace;
    cudaStream_t st = (cudaStream_t)stream;
    // d_state0 = [0 | d_vel]
    if (cudaMemsetAsync(ws + L.d_state0, 0, (size_t)B * 4 * d.n * sizeof(float), st) != cudaSuccess) return check_launch("memset");
    cudaMemcpy2DAsync(ws + L.d_state0 + 2 * d.n, 4 * d.n * sizeof(float), d_vel, 2 * d.n * sizeof(float),
                      2 * d.n * sizeof(float), B, cudaMemcpyDeviceToDevice, st);
    return velocity_backward(t, p, grads, L, ws + L.d_state0, d_enc_pos_accum, ws, st);
}

int paig_debug_gemm_tc(const float* A, const float* B, float* C, int M, int N, int K, int fixed_split, float* scratch,
                       long scratch_floats, void* stream) {
    const int sp = fixed_split == 2
        ? gemm_tc_partials_drained(A, B, M, N, K, scratch, (size_t)scratch_floats, "gemm_tf32x3_drained", (cudaStream_t)stream)
        : gemm_tc_partials(A, B, M, N, K, fixed_split != 0, scratch, (size_t)scratch_floats, "gemm_tf32x3",
                           (cudaStream_t)stream);
    if (sp < 0) { set_error("gemm_tc: shape %d x %d x %d does not qualify", M, N, K); return 1; }
    if (sp == 0) return 2;
    GemmArgs g;
    g.C = C; g.ldc = N; g.M = M; g.N = N; g.splitk_ws = scratch;
    return gemm_fold_partials(g, sp, (cudaStream_t)stream);
}

long paig_debug_workspace_offset(const paig_task* t, int B, const char* region, int index) {
    if (!valid_task(t) || !region) return -1;
    const Layout L = make_layout(t, B);
    auto is = [&](const char* n) { return strcmp(region, n) == 0; };
    if (is("act") || is("grad")) {
        if (index < 0 || index >= L.unet.nbufs) return -1;
        return (long)(is("act") ? L.act[index] : L.grad[index]);
    }
    if (is("logits")) return (long)L.logits;
    if (is("d_logits")) return (long)L.d_logits;
    if (is("enc_pos")) return (long)L.enc_pos;
    if (is("d_enc_pos")) return (long)L.d_enc_pos;
    if (is("seq")) return (long)L.seq;
    if (is("d_seq")) return (long)L.d_seq;
    if (is("consts")) return (long)L.consts;
    if (is("d_consts")) return (long)L.d_consts;
    if (is("H1")) return (long)L.H1;
    if (is("H2")) return (long)L.H2;
    if (is("A")) return (long)L.A;
    if (is("dA")) return (long)L.dA;
    return -1;
}

int paig_debug_unet_conv_view(const paig_task* t, int B, int layer, long view[5]) {
    if (!valid_task(t) || !view) return 1;
    const Layout L = make_layout(t, B);
    for (int i = 0; i < L.unet.nops; ++i) {
        const Op& op = L.unet.ops[i];
        if ((op.kind != OP_CONV && op.kind != OP_HEAD) || op.layer != layer) continue;
        if (op.kind == OP_HEAD) {
            view[0] = (long)L.logits; view[1] = (long)L.d.n * L.d.HW; view[2] = L.d.n; view[3] = L.d.H;
        } else {
            const BufDesc& b = L.unet.bufs[op.out.buf];
            const int S = L.d.H >> b.shift;
            view[0] = (long)L.act[op.out.buf] + (long)op.out.c0 * S * S;
            view[1] = (long)b.C * S * S; view[2] = op.out.C; view[3] = S;
        }
        view[4] = op.relu;
        return 0;
    }
    set_error("no conv layer %d", layer);
    return 1;
}

}  // extern "C"
