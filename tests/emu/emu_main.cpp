// TEST INFRASTRUCTURE -- runs the product's CUDA kernels (agatha_b200/csrc/extend_kernel.cuh, pack_kernel.cuh) on the CPU,
// one fiber per CUDA thread, so that the kernel logic can be compared with the oracle without a GPU. See cuda_runtime.h in
// this directory for what is emulated. Built by tests/emu/emu.py into tests/emu/libagatha_emu.so; never shipped, never
// linked by anything under agatha_b200/.
//
// The variant selection (shape, band-edge template constant, kernel parameters) is the product's own code
// (extend_dispatch.h), so the emulation runs exactly the template instance the GPU would run.
#include <cstdarg>
#include <cstdio>
#include <cstdio>
#include <vector>

#include "cuda_runtime.h"          // the shim in this directory (found first through -I)
#include "extend_dispatch.h"
#include "pack_kernel.cuh"

// ---- fibers ------------------------------------------------------------------------------------------------------
extern "C" void emu_switch(void** save_sp, void* load_sp);
asm(R"(
.text
.globl emu_switch
.type emu_switch, @function
emu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size emu_switch, .-emu_switch
)");

namespace emu {

Thread* cur = nullptr;
Block blk;
emu_dim3 g_blockDim = {1, 1, 1}, g_gridDim = {1, 1, 1}, g_blockIdx = {0, 0, 0};

struct Fiber {
    void* sp = nullptr;
    char* stack = nullptr;
    bool done = false;
    Thread th;
};

static void* g_sched_sp = nullptr;
static Fiber* g_fiber = nullptr;
static void (*g_body)(void*) = nullptr;
static void* g_body_arg = nullptr;

void yield() { emu_switch(&g_fiber->sp, g_sched_sp); }

static void fiber_entry()
{
    g_body(g_body_arg);
    g_fiber->done = true;
    for (;;) yield();
}

// Run body(arg) once per thread of a CTA of `nthreads` threads (grid of one CTA: the kernels are persistent and pull
// their work from a queue, so one CTA processes everything).
static void run_block(int nthreads, void (*body)(void*), void* arg)
{
    constexpr size_t STACK = 256 << 10;
    const int nwarps = (nthreads + 31) / 32;
    std::vector<Fiber> fibers(nthreads);
    std::vector<Warp> warps(nwarps);
    blk = Block();
    blk.nthreads = nthreads;
    g_blockDim = {(unsigned)nthreads, 1, 1};
    g_gridDim = {1, 1, 1};
    g_body = body; g_body_arg = arg;
    for (int t = 0; t < nthreads; t++) {
        Fiber& f = fibers[t];
        f.stack = (char*)aligned_alloc(64, STACK);
        f.th.tid = {(unsigned)t, 0, 0};
        f.th.warp = &warps[t / 32];
        f.th.lane = t & 31;
        uintptr_t top = ((uintptr_t)(f.stack + STACK)) & ~(uintptr_t)15;
        void** sp = (void**)top;
        *--sp = nullptr;                       // keeps the entry function's frame 16-byte aligned
        *--sp = (void*)&fiber_entry;           // `ret` target of the first switch
        for (int r = 0; r < 6; r++) *--sp = nullptr;
        f.sp = sp;
    }
    // Order in which the fibers get their turn: forward (0, default), backward (1) or reshuffled every round (2+: the seed).
    // Shared-memory protocols between warps (hand-over slots, arrive / wait barriers) must not depend on it.
    const char* sched_env = getenv("AGATHA_EMU_SCHED");
    const int sched = sched_env ? atoi(sched_env) : 0;
    std::vector<int> order(nthreads);
    for (int t = 0; t < nthreads; t++) order[t] = sched == 1 ? nthreads - 1 - t : t;
    unsigned long long rng = 0x9e3779b97f4a7c15ull * (unsigned long long)(sched + 1);
    int live = nthreads;
    unsigned long long round = 0, stalled = 0;      // sched >= 2: warps that get no turn for a while (bit per warp), redrawn every 48 rounds
    while (live) {
        live = 0;
        if (sched >= 2 && (round++ % 48) == 0) {
            rng = rng * 6364136223846793005ull + 1442695040888963407ull;
            stalled = (rng >> 20) & (rng >> 40);                       // about a quarter of the warps
            if ((stalled & ((1ull << nwarps) - 1ull)) == ((1ull << nwarps) - 1ull)) stalled = 0;
        }
        if (sched >= 2) {
            for (int t = nthreads - 1; t > 0; t--) {
                rng = rng * 6364136223846793005ull + 1442695040888963407ull;
                std::swap(order[t], order[(int)((rng >> 33) % (unsigned long long)(t + 1))]);
            }
        }
        for (int ti = 0; ti < nthreads; ti++) {
            const int t = order[ti];
            Fiber& f = fibers[t];
            if (f.done) continue;
            if (sched >= 2 && ((stalled >> ((t / 32) & 63)) & 1ull)) { live++; continue; }
            g_fiber = &f; cur = &f.th;
            emu_switch(&g_sched_sp, f.sp);
            if (!f.done) live++;
        }
    }
    for (auto& f : fibers) free(f.stack);
    cur = nullptr; g_fiber = nullptr;
}

}  // namespace emu

// ---- what the product's translation units provide ---------------------------------------------------------------------
namespace agatha {
static char g_err[512] = "";
int set_error(int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
int cuda_error(cudaError_t, const char* what) { return set_error(AGATHA_ECUDA, "%s", what); }
void count_launch() {}
}  // namespace agatha

extern "C" int agatha_max_band_width(void) { return 8 * 32 * 32 - 1; }

using namespace agatha;

struct EmuLauncher {
    const JobArrays& ja; const KernelParams& kp;
    template <int C, int NW, bool WODD, int JWS> int run() const
    {
        struct Ctx { const JobArrays* ja; const KernelParams* kp; } ctx{&ja, &kp};
        constexpr bool HAS_REDO = WODD && JWS >= 0 && C >= 8;           // as in extend_launch.cuh
        if (ja.redo && !HAS_REDO) return set_error(AGATHA_EUNSUPPORTED, "no redo pass for this kernel variant");
        if constexpr (HAS_REDO) {
            if (ja.redo) {
                emu::run_block(KernelShape<C, NW>::threads, [](void* a) { Ctx* c = (Ctx*)a; extend_kernel<C, NW, WODD, JWS, true>(*c->ja, *c->kp); }, &ctx);
                return AGATHA_OK;
            }
        }
        emu::run_block(KernelShape<C, NW>::threads, [](void* a) { Ctx* c = (Ctx*)a; extend_kernel<C, NW, WODD, JWS, false>(*c->ja, *c->kp); }, &ctx);
        return AGATHA_OK;
    }
    template <int C, int NW, int JWS> int run16() const
    {
        struct Ctx { const JobArrays* ja; const KernelParams* kp; } ctx{&ja, &kp};
        emu::run_block(Shape16<C, NW>::threads, [](void* a) {
            Ctx* c = (Ctx*)a;
            extend16_kernel<C, NW, JWS>(*c->ja, *c->kp);
        }, &ctx);
        return AGATHA_OK;
    }
};

extern "C" {

const char* emu_last_error(void) { return g_err; }

static uint32_t g_redone = 0;
static int g_used16 = 0;
uint32_t emu_last_redo_count(void) { return g_redone; }     // pairs the packed kernel handed to the general kernel
int emu_last_used_packed(void) { return g_used16; }          // 1 when the last emu_extend ran the packed kernel first

void emu_set_s16_mode(int mode) { s16_mode() = mode; }

// pack_kernel on ASCII batches laid out like the reference stages them (multiples of 8, 'N' padding)
int emu_pack(const uint8_t* q, uint64_t qbytes, const uint8_t* t, uint64_t tbytes, uint32_t* qout, uint32_t* tout)
{
    struct Ctx { const uint8_t *q, *t; uint64_t qw, tw; uint32_t *qo, *to; } ctx{q, t, qbytes / 8, tbytes / 8, qout, tout};
    emu::run_block(256, [](void* a) {
        Ctx* c = (Ctx*)a;
        pack_kernel((const uint2*)c->q, c->qw, (const uint2*)c->t, c->tw, c->qo, c->to);
    }, &ctx);
    return 0;
}

int emu_apply_ops(const uint8_t* q, const uint8_t* t, const uint32_t* qoff, const uint32_t* toff, const uint32_t* qlen, const uint32_t* tlen,
                  const uint8_t* qop, const uint8_t* top, uint32_t n, uint32_t* qout, uint32_t* tout)
{
    struct Ctx { const uint8_t *q, *t; const uint32_t *qoff, *toff, *qlen, *tlen; const uint8_t *qop, *top; uint32_t n; uint32_t *qo, *to; }
        ctx{q, t, qoff, toff, qlen, tlen, qop, top, n, qout, tout};
    emu::run_block(256, [](void* a) {
        Ctx* c = (Ctx*)a;
        apply_ops_kernel(c->q, c->t, c->qoff, c->toff, c->qlen, c->tlen, c->qop, c->top, c->n, c->qo, c->to);
    }, &ctx);
    return 0;
}

// agatha_extend_device with host pointers
int emu_extend(const uint32_t* qpk, const uint32_t* tpk, const uint32_t* qoff, const uint32_t* toff,
               const uint32_t* qlen, const uint32_t* tlen, const uint32_t* order, uint32_t n, const agatha_params_t* params,
               int32_t* score, int32_t* qend, int32_t* tend, int32_t* stop, int32_t* dstop)
{
    KernelParams kp;
    int rc = make_kernel_params(params, &kp);
    if (rc) return rc;
    unsigned counter = 0, counter2 = 0;
    JobArrays ja;
    ja.qpk = qpk; ja.tpk = tpk; ja.qoff_w = qoff; ja.toff_w = toff; ja.qlen = qlen; ja.tlen = tlen; ja.order = order;
    ja.score = score; ja.qend = qend; ja.tend = tend; ja.stop = stop; ja.dstop = dstop;
    ja.counter = &counter;
    ja.n = (int)n;
    ja.redo = 0;
    const EmuLauncher l{ja, kp};
    g_used16 = 0; g_redone = 0;
    if (dispatch16_variant(kp, l, &rc)) {            // same sequence as agatha_extend_device (engine.cu)
        if (rc) return rc;
        g_used16 = 1;
        for (uint32_t i = 0; i < n; i++) g_redone += qend[i] == REDO_MARK;
        ja.counter = &counter2;
        ja.redo = 1;
    }
    if (dispatch_variant(kp, l, &rc)) return rc;
    return set_error(AGATHA_EUNSUPPORTED, "no kernel for band_width %d", kp.W);
}

}  // extern "C"
