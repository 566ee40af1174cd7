// cuda_emu.cpp — the fiber scheduler and the host stand-ins for the CUDA runtime calls the product makes (see cuda_emu.h).
// TEST INFRASTRUCTURE ONLY.
#include "cuda_emu.h"

#include <setjmp.h>
#include <ucontext.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

namespace emu {
Idx tid{0, 0, 0}, bid{0, 0, 0}, bdim{1, 1, 1}, gdim{1, 1, 1};

namespace {
enum State { READY, WAIT_BLOCK, WAIT_WARP, DONE };
// Context switches: a fiber is ENTERED once with swapcontext (which installs its stack) and from then on switched with _setjmp / _longjmp:
// glibc's swapcontext saves and restores the signal mask with two system calls per switch, and a kernel with a few CTA barriers switches
// every one of its threads several times (the suite spent a third of its time in rt_sigprocmask).  Sanitizer builds keep swapcontext
// (AddressSanitizer follows it, but not a longjmp onto another stack).
#if defined(__SANITIZE_ADDRESS__) || defined(SPK_EMU_UCONTEXT_ONLY)
#define EMU_FAST_SWITCH 0
#else
#define EMU_FAST_SWITCH 1
#endif
struct Fiber {
    ucontext_t ctx;
    jmp_buf jb;
    bool entered = false;
    char* stack = nullptr;
    State state = DONE;
    unsigned wait_mask = 0;
};
jmp_buf sched_jb;
constexpr size_t STACK_BYTES = 512 * 1024;
std::vector<Fiber> fibers;
ucontext_t sched_ctx;
int cur = -1;
const std::function<void()>* body_fn = nullptr;
unsigned long long scratch[32][32];
alignas(128) unsigned char dyn_smem[232448];

void to_scheduler(Fiber& f) {
#if EMU_FAST_SWITCH
    if (_setjmp(f.jb) == 0) _longjmp(sched_jb, 1);
#else
    swapcontext(&f.ctx, &sched_ctx);
#endif
}
void to_fiber(Fiber& f) {
#if EMU_FAST_SWITCH
    if (_setjmp(sched_jb) == 0) {
        if (f.entered) _longjmp(f.jb, 1);
        f.entered = true;
        swapcontext(&sched_ctx, &f.ctx); // first entry: onto the fiber's own stack (it comes back through sched_jb)
    }
#else
    swapcontext(&sched_ctx, &f.ctx);
#endif
}
void trampoline() {
    (*body_fn)();
    fibers[cur].state = DONE;
    to_scheduler(fibers[cur]);
}
void park(State s, unsigned mask) {
    Fiber& f = fibers[cur];
    f.state = s;
    f.wait_mask = mask;
    to_scheduler(f);
}
void run_cta(unsigned nthreads) {
    if (fibers.size() < nthreads) fibers.resize(nthreads);
    for (unsigned t = 0; t < nthreads; ++t) {
        Fiber& f = fibers[t];
        if (!f.stack) f.stack = (char*)std::malloc(STACK_BYTES);
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = f.stack;
        f.ctx.uc_stack.ss_size = STACK_BYTES;
        f.ctx.uc_link = &sched_ctx;
        makecontext(&f.ctx, trampoline, 0);
        f.entered = false;
        f.state = READY;
    }
    // SPK_EMU_SCHEDULE=reverse runs the ready threads of a CTA (and the CTAs of a grid) in descending order: a read that is missing a
    // barrier after another thread's write passes in one order and fails in the other
    static const bool reverse = std::getenv("SPK_EMU_SCHEDULE") && std::string(std::getenv("SPK_EMU_SCHEDULE")) == "reverse";
    for (;;) {
        bool ran = false;
        for (unsigned k = 0; k < nthreads; ++k) {
            const unsigned t = reverse ? nthreads - 1 - k : k;
            if (fibers[t].state != READY) continue;
            cur = (int)t;
            tid = Idx{t, 0, 0};
            to_fiber(fibers[t]);
            ran = true;
        }
        // release warp-level waits: every live lane named by the mask has arrived with the same mask
        bool released = false;
        for (unsigned w = 0; w * 32 < nthreads; ++w) {
            for (unsigned l = 0; l < 32 && w * 32 + l < nthreads; ++l) {
                Fiber& f = fibers[w * 32 + l];
                if (f.state != WAIT_WARP) continue;
                const unsigned mask = f.wait_mask;
                bool all = true;
                for (unsigned k = 0; k < 32 && all; ++k) {
                    if (!(mask >> k & 1u) || w * 32 + k >= nthreads) continue;
                    const Fiber& o = fibers[w * 32 + k];
                    all = o.state == DONE || (o.state == WAIT_WARP && o.wait_mask == mask);
                }
                if (!all) continue;
                for (unsigned k = 0; k < 32; ++k)
                    if ((mask >> k & 1u) && w * 32 + k < nthreads && fibers[w * 32 + k].state == WAIT_WARP) fibers[w * 32 + k].state = READY;
                released = true;
            }
        }
        if (released) continue;
        // release the CTA barrier: every live thread waits on it
        bool any_live = false, all_block = true;
        for (unsigned t = 0; t < nthreads; ++t) {
            if (fibers[t].state == DONE) continue;
            any_live = true;
            if (fibers[t].state != WAIT_BLOCK) all_block = false;
        }
        if (!any_live) return;
        if (all_block) {
            for (unsigned t = 0; t < nthreads; ++t) if (fibers[t].state == WAIT_BLOCK) fibers[t].state = READY;
            continue;
        }
        if (!ran) {
            std::fprintf(stderr, "cuda_emu: deadlock in block %u (threads wait on different barriers)\n", bid.x);
            std::abort();
        }
    }
}
} // namespace

void launch(unsigned grid, unsigned block, const std::function<void()>& body) {
    if (grid == 0 || block == 0) return;
    body_fn = &body;
    gdim = Idx{grid, 1, 1};
    bdim = Idx{block, 1, 1};
    static const bool reverse = std::getenv("SPK_EMU_SCHEDULE") && std::string(std::getenv("SPK_EMU_SCHEDULE")) == "reverse";
    for (unsigned k = 0; k < grid; ++k) {
        bid = Idx{reverse ? grid - 1 - k : k, 0, 0};
        run_cta(block);
    }
    cur = -1;
    body_fn = nullptr;
}
void block_sync() { park(WAIT_BLOCK, 0); }
void warp_sync(unsigned mask) { park(WAIT_WARP, mask); }
unsigned long long* warp_scratch() { return scratch[(tid.x >> 5) & 31u]; }
unsigned char* dynamic_smem() { return dyn_smem; }
} // namespace emu

// ---- CUDA runtime stand-ins: one "device", memory is host memory, streams and events are inert ---------------------------------
extern "C" {
cudaError_t cudaSetDevice(int dev) { return dev == 0 ? cudaSuccess : cudaErrorInvalidDevice; }
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA error"; }
cudaError_t cudaGetDeviceProperties_v2(cudaDeviceProp* p, int) { std::memset(p, 0, sizeof(*p)); p->major = 10; p->minor = 0; p->multiProcessorCount = 148; return cudaSuccess; }
// device memory comes back zeroed by default; SPK_EMU_POISON=<byte> fills it with that byte instead (cudaMalloc returns garbage on a
// GPU: tests/test_emulated_cuda_cpu.py runs the parity bodies with poisoned allocations to catch reads of uninitialised memory)
cudaError_t cudaMalloc(void** p, size_t n) {
    static const char* poison = std::getenv("SPK_EMU_POISON");
    *p = std::calloc(n ? n : 1, 1);
    if (*p && poison) std::memset(*p, (int)std::strtol(poison, nullptr, 0), n ? n : 1);
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) { *p = std::calloc(n ? n : 1, 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaFreeHost(void* p) { std::free(p); return cudaSuccess; }
cudaError_t cudaHostGetDevicePointer(void** d, void* h, unsigned) { *d = h; return cudaSuccess; }
cudaError_t cudaMemset(void* p, int v, size_t n) { std::memset(p, v, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { std::memset(p, v, n); return cudaSuccess; }
cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { std::memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpyPeerAsync(void* d, int, const void* s, int, size_t n, cudaStream_t) { std::memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (cudaStream_t)std::malloc(8); return cudaSuccess; }
cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned, int) { *s = (cudaStream_t)std::malloc(8); return cudaSuccess; }
cudaError_t cudaDeviceGetStreamPriorityRange(int* least, int* greatest) { *least = 0; *greatest = -1; return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) { std::free(s); return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
struct EmuEvent { std::chrono::steady_clock::time_point t; };
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = (cudaEvent_t) new EmuEvent(); return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete (EmuEvent*)e; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { ((EmuEvent*)e)->t = std::chrono::steady_clock::now(); return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
    *ms = std::chrono::duration<float, std::milli>(((EmuEvent*)b)->t - ((EmuEvent*)a)->t).count();
    return cudaSuccess;
}
}
