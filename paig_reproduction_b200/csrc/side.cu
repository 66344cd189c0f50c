// Side streams of the fused training step.  The step is one dependent chain of kernels, but several short branches hang
// off it (VariableFromNetwork forward / backward, loss scalars, the encoder.l1 weight gradient, frame gather) and its
// twelve weight-gradient launches are independent of one another.  Each of those is 5-110 us of mostly idle GPU when
// run in line; forked onto a side stream they fill the SMs the main chain leaves empty (1-CTA rollout kernels, the tail
// waves of the persistent UNet kernels).  Fork = record an event on the main stream, make the side stream wait for it;
// join = the reverse.  Streams and events are created once per device and reused; everything stays ordered behind the
// caller's stream, so the C ABI's contract (work is enqueued on `stream`) is unchanged.
#include "common.cuh"
#include "internal.h"

#include <cstdlib>

namespace paig {

#ifndef PAIG_EMU
namespace {
constexpr int kMaxDev = 16, kEvents = 12;
struct SideSet {
    bool ready = false;
    cudaStream_t s[2] = {nullptr, nullptr};
    cudaEvent_t ev[kEvents];
    int next = 0;
};
SideSet g_sets[kMaxDev];
thread_local Side g_cur;
thread_local bool g_active = false;
}  // namespace

Side* side_begin(cudaStream_t main) {
    static const bool off = getenv("PAIG_NO_SIDE_STREAMS") != nullptr;
    if (off || g_profiling || g_active) return nullptr;          // per-launch timing wants one launch at a time
    // Under stream capture (the caller records the step into a CUDA graph) the side streams join the capture through
    // the fork / join events, which is legal as long as every fork is joined before the capture ends -- the fused step
    // does that.  PAIG_SIDE_IN_CAPTURE=0 keeps a captured step on one stream.
    static const bool in_capture = !(getenv("PAIG_SIDE_IN_CAPTURE") && getenv("PAIG_SIDE_IN_CAPTURE")[0] == '0');
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(main, &cap) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    if (cap != cudaStreamCaptureStatusNone && !in_capture) return nullptr;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDev) return nullptr;
    SideSet& set = g_sets[dev];
    if (!set.ready) {
        for (int i = 0; i < 2; ++i)
            if (cudaStreamCreateWithFlags(&set.s[i], cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        for (int i = 0; i < kEvents; ++i)
            if (cudaEventCreateWithFlags(&set.ev[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
        set.ready = true;
    }
    g_cur.main = main;
    g_cur.s1 = set.s[0];
    g_cur.s2 = set.s[1];
    g_cur.set = &set;
    g_active = true;
    return &g_cur;
}

void side_end() { g_active = false; }
Side* side_cur() { return g_active ? &g_cur : nullptr; }

// `waiter` continues only after everything enqueued so far on `signaler`
void Side::after(cudaStream_t waiter, cudaStream_t signaler) {
    SideSet* ss = static_cast<SideSet*>(set);
    cudaEvent_t e = ss->ev[ss->next];
    ss->next = (ss->next + 1) % kEvents;
    cudaEventRecord(e, signaler);
    cudaStreamWaitEvent(waiter, e, 0);
}
#else
Side* side_begin(cudaStream_t) { return nullptr; }
void side_end() {}
Side* side_cur() { return nullptr; }
void Side::after(cudaStream_t, cudaStream_t) {}
#endif

}  // namespace paig
