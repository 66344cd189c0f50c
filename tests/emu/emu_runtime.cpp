// Fiber scheduler behind emu_cuda.h (TEST TOOL ONLY; see the header).
// One host thread; the threads of a block are ucontext fibers run round-robin; a fiber yields
// only inside a barrier or a warp exchange, so execution is deterministic.
#include "emu_cuda.h"

#include <ucontext.h>

#include <vector>

namespace emu {
uint3 g_threadIdx, g_blockIdx;
dim3 g_blockDim, g_gridDim;
unsigned char* g_dyn_smem = nullptr;

namespace {
constexpr size_t kStack = 256 * 1024;
struct Fiber {
    ucontext_t ctx;
    uint3 tid;
    bool done;
};
struct Warp {
    uint64_t vals[32];
    int arrived = 0;
    unsigned gen = 0;
};
std::vector<Fiber> fibers;
std::vector<Warp> warps;
std::vector<unsigned char> stacks;
ucontext_t sched_ctx;
int cur = -1;
int n_alive = 0, blk_arrived = 0;
unsigned blk_gen = 0;
unsigned long progress = 0;
const std::function<void()>* cur_body = nullptr;

void yield() {
    swapcontext(&fibers[cur].ctx, &sched_ctx);
}
void trampoline() {
    (*cur_body)();
    fibers[cur].done = true;
    --n_alive;
    // a thread that exits releases a barrier the rest are waiting in (CUDA: exited threads do not count)
    if (n_alive > 0 && blk_arrived >= n_alive) {
        blk_arrived = 0;
        ++blk_gen;
    }
    swapcontext(&fibers[cur].ctx, &sched_ctx);
}
void warp_barrier(Warp& w, int lanes) {
    unsigned gen = w.gen;
    if (++w.arrived == lanes) {
        w.arrived = 0;
        ++w.gen;
        ++progress;
    } else {
        while (w.gen == gen) yield();
    }
}
int warp_lanes(int warp, int nthreads) {
    int lo = warp * 32;
    int hi = lo + 32 < nthreads ? lo + 32 : nthreads;
    return hi - lo;
}
}  // namespace

int lane_id() { return cur & 31; }

void syncthreads() {
    unsigned gen = blk_gen;
    if (++blk_arrived >= n_alive) {
        blk_arrived = 0;
        ++blk_gen;
    } else {
        while (blk_gen == gen) yield();
    }
}

void syncwarp() {
    int nthreads = (int)fibers.size();
    int w = cur >> 5;
    warp_barrier(warps[w], warp_lanes(w, nthreads));
}

uint64_t warp_exchange(uint64_t v, int src_lane) {
    int nthreads = (int)fibers.size();
    int w = cur >> 5;
    int lanes = warp_lanes(w, nthreads);
    Warp& W = warps[w];
    W.vals[cur & 31] = v;
    warp_barrier(W, lanes);
    uint64_t r = W.vals[src_lane & 31];
    warp_barrier(W, lanes);
    return r;
}

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
    const int nthreads = (int)(block.x * block.y * block.z);
    if (nthreads <= 0 || nthreads > 1024) {
        fprintf(stderr, "emu::launch: bad block size %d\n", nthreads);
        abort();
    }
    std::vector<unsigned char> dyn(smem + 64);
    g_dyn_smem = (unsigned char*)(((uintptr_t)dyn.data() + 63) & ~(uintptr_t)63);
    if (stacks.size() < kStack * (size_t)nthreads) stacks.resize(kStack * (size_t)nthreads);
    g_gridDim = grid;
    g_blockDim = block;
    cur_body = &body;
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                fibers.assign(nthreads, Fiber());
                warps.assign((nthreads + 31) / 32, Warp());
                n_alive = nthreads;
                blk_arrived = 0;
                int t = 0;
                for (unsigned tz = 0; tz < block.z; ++tz)
                    for (unsigned ty = 0; ty < block.y; ++ty)
                        for (unsigned tx = 0; tx < block.x; ++tx, ++t) {
                            Fiber& f = fibers[t];
                            f.tid = uint3{tx, ty, tz};
                            f.done = false;
                            getcontext(&f.ctx);
                            f.ctx.uc_stack.ss_sp = stacks.data() + kStack * (size_t)t;
                            f.ctx.uc_stack.ss_size = kStack;
                            f.ctx.uc_link = &sched_ctx;
                            makecontext(&f.ctx, trampoline, 0);
                        }
                g_blockIdx = uint3{bx, by, bz};
                int guard = 0;
                while (n_alive > 0) {
                    int before = n_alive;
                    unsigned gen_before = blk_gen;
                    unsigned long prog_before = progress;
                    for (int i = 0; i < nthreads; ++i) {
                        if (fibers[i].done) continue;
                        cur = i;
                        g_threadIdx = fibers[i].tid;
                        swapcontext(&sched_ctx, &fibers[i].ctx);
                    }
                    // progress check: a full pass with no exit and no barrier release, many times over, is a deadlock
                    if (n_alive == before && blk_gen == gen_before && progress == prog_before) {
                        if (++guard > 1000) {
                            fprintf(stderr, "emu::launch: deadlock (divergent barrier?) in block %u,%u,%u\n", bx, by, bz);
                            abort();
                        }
                    } else {
                        guard = 0;
                    }
                }
            }
    cur = -1;
    cur_body = nullptr;
    g_dyn_smem = nullptr;
}
}  // namespace emu
