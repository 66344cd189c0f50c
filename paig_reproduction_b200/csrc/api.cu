// extern "C" surface of libpaig_b200.so (declared in include/paig_b200.h).
#include "internal.h"

namespace paig {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return 2;
    }
    return 0;
}

static bool valid_task(const paig_task* t) {
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

int paig_rollout_forward(int cell, int n_objs, int B, int steps, const float* dt, const double* phys0,
                         const double* phys1, float* pos_vel_seq, void* stream) {
    return rollout_forward(cell, n_objs, B, steps, dt, phys0, phys1, pos_vel_seq, (cudaStream_t)stream);
}

int paig_rollout_backward(int cell, int n_objs, int B, int steps, const float* dt, const double* phys0,
                          const double* phys1, const float* pos_vel_seq, const float* d_seq, float* d_state0,
                          double* d_phys, void* stream) {
    const long rs = 4L * n_objs;
    return rollout_backward(cell, n_objs, B, steps, dt, phys0, phys1, pos_vel_seq, d_seq, d_seq + 2 * n_objs,
                            (steps + 1) * rs, rs, 1, d_state0, d_phys, (cudaStream_t)stream);
}

}  // extern "C"
