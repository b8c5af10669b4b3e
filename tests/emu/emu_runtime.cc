// TEST INFRASTRUCTURE — scheduler of the host-side SIMT simulator (see emu_runtime.h).
#include "emu_runtime.h"
#undef threadIdx
#undef blockIdx
#undef blockDim
#undef gridDim

namespace emu {

static State g_state;
static long g_events = 0;   // bumped whenever a barrier/shuffle completes or a fiber exits
State& st() { return g_state; }

static const size_t kStack = 256 * 1024;
static std::vector<char*> g_stacks;

void yield() { swapcontext(&g_state.cur->ctx, &g_state.sched); }

static void fiber_entry() {
    State& s = g_state;
    s.body();
    s.cur->done = true;
    s.n_alive--;
    g_events++;
    // a thread that exits counts as arrived for any barrier the others are waiting on
    if (s.n_alive > 0 && s.bar_count >= s.n_alive) { s.bar_count = 0; s.bar_gen++; }
    swapcontext(&s.cur->ctx, &s.sched);
}

void syncthreads() {
    State& s = g_state;
    unsigned gen = s.bar_gen;
    s.bar_count++;
    if (s.bar_count >= s.n_alive) { s.bar_count = 0; s.bar_gen++; g_events++; return; }
    while (s.bar_gen == gen) yield();
}

float shfl(float v, int kind, int arg) {
    State& s = g_state;
    int lin = s.cur->tid.x + s.bdim.x * (s.cur->tid.y + s.bdim.y * s.cur->tid.z);
    int warp = lin >> 5, lane = lin & 31;
    int nthreads = s.bdim.x * s.bdim.y * s.bdim.z;
    int wsize = nthreads - warp * 32 < 32 ? nthreads - warp * 32 : 32;
    float* buf = &s.wbuf[(size_t)warp * 64];
    // phase 1: publish, wait for the whole warp
    unsigned gen = s.wgen[warp];
    buf[(gen & 1) * 32 + lane] = v;
    s.warrive[warp]++;
    if (s.warrive[warp] == wsize) { s.warrive[warp] = 0; s.wgen[warp]++; g_events++; }
    else while (s.wgen[warp] == gen) yield();
    int src = kind == 0 ? (lane ^ arg) : kind == 1 ? arg : lane + arg;
    if (src < 0 || src >= wsize) src = lane;
    return buf[(gen & 1) * 32 + src];   // double-buffered by generation parity
}

void launch(dim3 grid, dim3 block, size_t smem, std::function<void()> body) {
    State& s = g_state;
    int nthreads = block.x * block.y * block.z;
    if (nthreads <= 0 || nthreads > 1024) { s.error = "emu: bad block size"; return; }
    if (smem > 227 * 1024) { s.error = "emu: dynamic shared memory over 227 KB"; return; }
    while ((int)g_stacks.size() < nthreads) g_stacks.push_back((char*)malloc(kStack));
    if (s.dyn_cap < smem + 16) { free(s.dyn_smem); s.dyn_cap = smem + 16; s.dyn_smem = (unsigned char*)aligned_alloc(128, (s.dyn_cap + 127) / 128 * 128); }
    s.body = body;
    s.bdim = block;
    s.gdim = grid;
    s.fibers.resize(nthreads);
    int nwarps = (nthreads + 31) / 32;
    s.wbuf.assign((size_t)nwarps * 64, 0.f);
    for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
    for (unsigned bx = 0; bx < grid.x; ++bx) {
        s.bidx = dim3(bx, by, bz);
        s.n_alive = nthreads;
        s.bar_count = 0;
        s.warrive.assign(nwarps, 0);
        s.wgen.assign(nwarps, 0);
        memset(s.dyn_smem, 0xCD, smem);   // poison: uninitialised shared memory shows up as garbage
        for (int t = 0; t < nthreads; ++t) {
            Fiber& f = s.fibers[t];
            f.done = false;
            f.stack = g_stacks[t];
            f.tid = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = f.stack;
            f.ctx.uc_stack.ss_size = kStack;
            f.ctx.uc_link = &s.sched;
            makecontext(&f.ctx, fiber_entry, 0);
        }
        int remaining = nthreads;
        while (remaining > 0) {
            long before = g_events;
            for (int t = 0; t < nthreads; ++t) {
                Fiber& f = s.fibers[t];
                if (f.done) continue;
                s.cur = &f;
                swapcontext(&s.sched, &f.ctx);
                if (f.done) remaining--;
            }
            // a whole round in which no barrier/shuffle completed and no fiber exited can never make progress
            if (remaining > 0 && g_events == before) { s.error = "emu: deadlock (divergent barrier or shuffle)"; return; }
        }
    }
}

}  // namespace emu
