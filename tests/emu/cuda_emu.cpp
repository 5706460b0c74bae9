// Fiber scheduler of the CTA emulator (TEST INFRASTRUCTURE ONLY, see cuda_emu.h).
#include "cuda_emu.h"
#include <ucontext.h>
#include <chrono>
#include <cstdio>
#include <vector>

namespace emu {

uint3_emu tIdx, bIdx, bDim, gDim;
void* dyn_smem = nullptr;
size_t dyn_smem_bytes = 0;
double xd[1024];
int xi[1024];

namespace {
enum State { READY, WAIT_BLOCK, WAIT_WARP, DONE };
struct Fiber { ucontext_t ctx; State st; char* stack; };
constexpr size_t STACK = 512 * 1024;
std::vector<Fiber> fibers;
ucontext_t sched;
int cur = -1;
void (*g_thunk)(void*) = nullptr;
void* g_args = nullptr;

void entry() {
    g_thunk(g_args);
    fibers[cur].st = DONE;
}
void resume(int t) {
    cur = t;
    tIdx.x = (unsigned)t; tIdx.y = tIdx.z = 0;
    swapcontext(&sched, &fibers[t].ctx);
}
void yield(State s) {
    fibers[cur].st = s;
    swapcontext(&fibers[cur].ctx, &sched);
}
}  // namespace

void sync_block() { yield(WAIT_BLOCK); }
void sync_warp() { yield(WAIT_WARP); }

void run_grid(void (*thunk)(void*), void* args, dim3 grid, dim3 block, size_t smem) {
    const int NT = (int)block.x;
    g_thunk = thunk; g_args = args;
    gDim = {grid.x, grid.y, 1}; bDim = {block.x, 1, 1};
    if ((int)fibers.size() < NT) {
        size_t old = fibers.size();
        fibers.resize(NT);
        for (size_t k = old; k < fibers.size(); k++) fibers[k].stack = (char*)malloc(STACK);
    }
    std::vector<unsigned char> sm(smem + 16);
    dyn_smem_bytes = smem;
    for (unsigned bb = 0; bb < grid.x * grid.y; bb++) {
        const unsigned b = bb % grid.x;
        bIdx = {b, bb / grid.x, 0};
        memset(sm.data(), 0xff, sm.size());              // all-ones doubles are NaNs
        dyn_smem = (void*)(((uintptr_t)sm.data() + 15) & ~(uintptr_t)15);
        for (int t = 0; t < NT; t++) {
            Fiber& f = fibers[t];
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = f.stack; f.ctx.uc_stack.ss_size = STACK; f.ctx.uc_link = &sched;
            makecontext(&f.ctx, (void (*)())entry, 0);
            f.st = READY;
        }
        for (int t = 0; t < NT; t++) resume(t);
        for (;;) {
            bool progressed = false, all_done = true, all_block = true;
            for (int w = 0; w * 32 < NT; w++) {
                const int lo = w * 32, hi = std::min(NT, lo + 32);
                bool any = false, allw = true;
                for (int t = lo; t < hi; t++) if (fibers[t].st != DONE) { any = true; if (fibers[t].st != WAIT_WARP) allw = false; }
                if (any && allw) {
                    for (int t = lo; t < hi; t++) if (fibers[t].st == WAIT_WARP) { fibers[t].st = READY; }
                    for (int t = lo; t < hi; t++) if (fibers[t].st == READY) resume(t);
                    progressed = true;
                }
            }
            for (int t = 0; t < NT; t++) if (fibers[t].st != DONE) { all_done = false; if (fibers[t].st != WAIT_BLOCK) all_block = false; }
            if (all_done) break;
            if (all_block) {
                for (int t = 0; t < NT; t++) if (fibers[t].st == WAIT_BLOCK) fibers[t].st = READY;
                for (int t = 0; t < NT; t++) if (fibers[t].st == READY) resume(t);
                progressed = true;
            }
            if (!progressed) { fprintf(stderr, "cuda_emu: barrier deadlock in block %u\n", b); abort(); }
        }
    }
    dyn_smem = nullptr;
}

}  // namespace emu

double emu_now_ms() {
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}
