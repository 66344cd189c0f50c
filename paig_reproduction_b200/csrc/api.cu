// extern "C" surface of libpaig_b200.so (declared in include/paig_b200.h).
#include <cstring>
#include <cstdlib>
#include "internal.h"

#include <map>
#include <string>
#include <vector>

namespace paig {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

long g_launch_count = 0;
bool g_profiling = false;

#ifndef PAIG_EMU
struct ProfRec {
    cudaEvent_t a, b;
    const char* name;
};
static std::vector<ProfRec> g_recs;
static std::vector<cudaEvent_t> g_event_pool;
static cudaEvent_t take_event() {
    cudaEvent_t e;
    if (!g_event_pool.empty()) {
        e = g_event_pool.back();
        g_event_pool.pop_back();
    } else {
        cudaEventCreate(&e);
    }
    return e;
}
void prof_before(cudaStream_t st) {
    ProfRec r{take_event(), take_event(), nullptr};
    cudaEventRecord(r.a, st);
    g_recs.push_back(r);
}
void prof_after(cudaStream_t st) { cudaEventRecord(g_recs.back().b, st); }
static void prof_name(const char* what) {
    if (g_profiling && !g_recs.empty() && g_recs.back().name == nullptr) g_recs.back().name = what;
}
#else
void prof_before(cudaStream_t) {}
void prof_after(cudaStream_t) {}
static void prof_name(const char*) {}
#endif

// PAIG_PROFILE_LAYERS=1: per-layer names ("conv3x3_wgrad_tma 16->16@32") in the paig_profile_end report instead of
// one line per kernel family.  Names are interned so the pointers stay valid for the profile records.
const char* layer_name(const char* family, int Cin, int Cout, int S) {
    static const bool on = getenv("PAIG_PROFILE_LAYERS") != nullptr;
    if (!on) return family;
    static char pool[256][48];
    static int used = 0;
    char tmp[48];
    snprintf(tmp, sizeof(tmp), "%s_%d->%d@%d", family, Cin, Cout, S);
    for (int i = 0; i < used; ++i)
        if (!strcmp(pool[i], tmp)) return pool[i];
    if (used == 256) return family;
    strcpy(pool[used], tmp);
    return pool[used++];
}

int check_launch(const char* what) {
    prof_name(what);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return 2;
    }
    return 0;
}

bool valid_task(const paig_task* t) {
    if (!t) {
        set_error("task is NULL");
        return false;
    }
    if (t->n_objs < 1 || t->n_objs > kMaxObjs || t->H < 8 || (t->H % 4) != 0 || t->input_steps < 1 ||
        t->pred_steps < 1 || t->seq_len <= t->input_steps + t->pred_steps) {
        // physics_models.py:59,84-86 asserts seq_len > in+pr, in >= 1, pr >= 1
        set_error("invalid task: n_objs=%d H=%d seq_len=%d input_steps=%d pred_steps=%d", t->n_objs, t->H, t->seq_len,
                  t->input_steps, t->pred_steps);
        return false;
    }
    return true;
}

}  // namespace paig

using namespace paig;

extern "C" {

int paig_abi_version(void) { return PAIG_ABI_VERSION; }

const char* paig_last_error(void) { return g_err; }

long paig_launch_count(void) { return g_launch_count; }

void paig_profile_begin(void) {
#ifndef PAIG_EMU
    g_recs.clear();
    g_profiling = true;
#endif
}

/* Stops per-launch timing, waits for the device and writes "name launches total_ms" lines into buf. */
int paig_profile_end(char* buf, size_t cap) {
#ifndef PAIG_EMU
    g_profiling = false;
    if (cudaDeviceSynchronize() != cudaSuccess) {
        set_error("profile_end: %s", cudaGetErrorString(cudaGetLastError()));
        return 2;
    }
    std::map<std::string, std::pair<long, double>> agg;
    for (auto& r : g_recs) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r.a, r.b);
        auto& e = agg[r.name ? r.name : "?"];
        e.first += 1;
        e.second += ms;
        g_event_pool.push_back(r.a);
        g_event_pool.push_back(r.b);
    }
    g_recs.clear();
    std::string out;
    for (auto& kv : agg) {
        char line[256];
        snprintf(line, sizeof(line), "%s %ld %.6f\n", kv.first.c_str(), kv.second.first, kv.second.second);
        out += line;
    }
    if (buf && cap) {
        snprintf(buf, cap, "%s", out.c_str());
    }
#else
    if (buf && cap) buf[0] = 0;
#endif
    return 0;
}

int paig_rollout_forward(int cell, int n_objs, int B, int steps, const float* dt, const double* phys0,
                         const double* phys1, float* pos_vel_seq, void* stream) {
    return rollout_forward(cell, n_objs, B, steps, dt, phys0, phys1, 0.f, pos_vel_seq, (cudaStream_t)stream);
}

int paig_rollout_backward(int cell, int n_objs, int B, int steps, const float* dt, const double* phys0,
                          const double* phys1, const float* pos_vel_seq, const float* d_seq, float* d_state0,
                          double* d_phys, void* stream) {
    const long rs = 4L * n_objs;
    if (cell == PAIG_CELL_BOUNCING && d_phys) cudaMemsetAsync(d_phys, 0, 2 * sizeof(double), (cudaStream_t)stream);
    return rollout_backward(cell, n_objs, B, steps, dt, phys0, phys1, 0.f, pos_vel_seq, d_seq, d_seq + 2 * n_objs,
                            (steps + 1) * rs, rs, 1, d_state0, d_phys, d_phys ? d_phys + 1 : nullptr, nullptr,
                            (cudaStream_t)stream);   // no scratch: one block walks all sequences (fixed order)
}

int paig_templates_forward(const paig_task* t, const paig_params* p, float* raw, float* consts, float* hidden,
                           void* stream) {
    if (!valid_task(t)) return 1;
    return templates_forward(t, p, raw, consts, hidden, (cudaStream_t)stream);
}

int paig_templates_backward(const paig_task* t, const paig_params* p, const paig_params* grads, const float* consts,
                            const float* hidden, const float* d_consts, void* workspace, void* stream) {
    if (!valid_task(t)) return 1;
    return templates_backward(t, p, grads, consts, hidden, d_consts, (float*)workspace, (cudaStream_t)stream);
}

static DecSeg flat_segment(const paig_task* t, const float* loc, int F, const float* target, long target_seq_stride,
                           int frames_per_seq) {
    DecSeg s;
    s.nframes = F;
    s.fps = frames_per_seq > 0 ? frames_per_seq : 1;
    s.loc = loc;
    s.loc_row_stride = 2L * t->n_objs;
    s.loc_seq_stride = s.loc_row_stride * s.fps;
    s.target = target;
    s.tgt_seq_stride = target_seq_stride;
    return s;
}

int paig_decode_forward(const paig_task* t, const float* consts, const float* loc, int F, float* frames,
                        const float* target, long target_seq_stride, int frames_per_seq, float* sse, void* stream) {
    if (!valid_task(t)) return 1;
    DecSeg a = flat_segment(t, loc, F, target, target_seq_stride, frames_per_seq), b;
    a.frames = frames;
    a.sse = target ? sse : nullptr;
    return decode_run(t, consts, a, b, false, nullptr, nullptr, 0, (cudaStream_t)stream);
}

int paig_decode_layers(const paig_task* t, const float* consts, const float* loc, int F, float* transf_contents,
                       float* transf_masks, void* stream) {
    if (!valid_task(t)) return 1;
    if (!transf_contents || !transf_masks) { set_error("decode_layers: both outputs are required"); return 1; }
    DecSeg a = flat_segment(t, loc, F, nullptr, 0, 1), b;
    a.layer_c = transf_contents;
    a.layer_m = transf_masks;
    a.layer_stride = (long)F * 3 * t->H * t->H;
    return decode_run(t, consts, a, b, false, nullptr, nullptr, 0, (cudaStream_t)stream);
}

int paig_decode_backward(const paig_task* t, const float* consts, const float* loc, int F, const float* d_frames,
                         const float* target, long target_seq_stride, int frames_per_seq, const float* scale,
                         float* d_loc, float* d_consts, float* sse, void* workspace, void* stream) {
    if (!valid_task(t)) return 1;
    if (!d_frames && !(target && scale)) {
        set_error("decode_backward needs d_frames or (target and scale)");
        return 1;
    }
    DecSeg a = flat_segment(t, loc, F, target, target_seq_stride, frames_per_seq), b;
    a.dframes = d_frames;
    a.scale = d_frames ? nullptr : scale;
    a.sse = target ? sse : nullptr;
    a.dloc = d_loc;
    a.dloc_row_stride = a.loc_row_stride;
    a.dloc_seq_stride = a.loc_seq_stride;
    return decode_run(t, consts, a, b, true, (float*)workspace, d_consts, 1, (cudaStream_t)stream);
}

int paig_debug_conv3x3_tc(const float* x, const float* w, const float* b, float* y, int N, int Cin, int Cout, int S, int relu,
                          int transposed, float* scratch, void* stream) {
    ConvArgs a;
    a.in = x; a.in_bs = (long)Cin * S * S; a.Cin = Cin;
    a.w = w; a.b = b;
    a.out = y; a.out_bs = (long)Cout * S * S; a.Cout = Cout;
    a.S = S; a.N = N; a.relu = relu; a.transposed = transposed; a.force_tc = 1;
    const int rc = conv3x3_tc(a, scratch, (cudaStream_t)stream);
    if (rc < 0) { set_error("conv3x3_tc: %d -> %d channels at %d px does not qualify (or tcgen05 is switched off)", Cin, Cout, S); return 1; }
    return rc;
}

int paig_conv3x3_forward(const float* x, const float* w, const float* b, float* y, int N, int Cin, int Cout, int S,
                         int relu, void* stream) {
    ConvArgs a;
    a.in = x; a.in_bs = (long)Cin * S * S; a.Cin = Cin;
    a.w = w; a.b = b;
    a.out = y; a.out_bs = (long)Cout * S * S; a.Cout = Cout;
    a.S = S; a.N = N; a.relu = relu;
    return conv3x3(a, (cudaStream_t)stream);
}

int paig_conv3x3_backward(const float* x, const float* w, const float* y, const float* dy, float* dx, float* dw,
                          float* db, int N, int Cin, int Cout, int S, int relu, void* workspace, void* stream) {
    WgradArgs g;
    g.in = x; g.in_bs = (long)Cin * S * S; g.Cin = Cin;
    g.g = dy; g.g_bs = (long)Cout * S * S; g.Cout = Cout;
    g.act = relu ? y : nullptr; g.act_bs = g.g_bs;
    g.S = S; g.N = N; g.partials = (float*)workspace;
    int rc = conv3x3_wgrad(g, dw, db, (cudaStream_t)stream);
    if (rc || !dx) return rc;
    ConvArgs a;
    a.in = dy; a.in_bs = g.g_bs; a.Cin = Cout;
    a.mask = relu ? y : nullptr; a.mask_bs = g.g_bs;
    a.w = w; a.transposed = 1;
    a.out = dx; a.out_bs = g.in_bs; a.Cout = Cin;
    a.S = S; a.N = N;
    return conv3x3(a, (cudaStream_t)stream);
}

}  // extern "C"
