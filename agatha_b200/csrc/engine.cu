// Device-level entry points of the C ABI (include/agatha_b200.h): packing and the extension kernel launch.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>

#include "extend_launch.cuh"
#include "pack_kernel.cuh"
#include "engine_internal.h"

namespace agatha {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

int set_error(int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int cuda_error(cudaError_t e, const char* what)
{
    return set_error(e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? AGATHA_ENODEV : AGATHA_ECUDA,
                     "%s: %s", what, cudaGetErrorString(e));
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// Kernel shape for a band width: a group of NW warps covers global cell indices g in [0, 32*NW*C); the band needs g <= W.
struct Shape { int C, NW; };
static Shape shape_for(int W)
{
    if (W < 32 * 2) return {2, 1};
    if (W < 32 * 4) return {4, 1};
    if (W < 32 * 8) return {8, 1};
    if (W < 32 * 16) return {16, 1};
    if (W < 32 * 24) return {24, 1};
    if (W < 32 * 32) return {32, 1};
    if (W < 2 * 32 * 32) return {32, 2};
    if (W < 4 * 32 * 32) return {32, 4};
    if (W < 8 * 32 * 32) return {32, 8};
    return {0, 0};
}

// The index of the band-edge cell inside its lane, JW = W % C, is a template constant for every band width that is
// 7 (mod 8) -- the only residue for which the reference's band is exact -- and a run-time value otherwise.
template <int C, int NW, int JW>
static int launch_static_jw(const JobArrays& ja, const KernelParams& kp, cudaStream_t st, bool& done)
{
    if constexpr (JW < C) {
        if (kp.JW == JW) { done = true; return launch_variant<C, NW, true, JW>(ja, kp, st); }
    }
    return AGATHA_OK;
}

template <int C, int NW>
static int launch_c(const JobArrays& ja, const KernelParams& kp, cudaStream_t st)
{
    const bool wodd = kp.W & 1;
    if (wodd) {
        bool done = false;
        int rc = launch_static_jw<C, NW, (C < 8 ? C - 1 : 7)>(ja, kp, st, done);   if (done) return rc;
        rc = launch_static_jw<C, NW, 15>(ja, kp, st, done);      if (done) return rc;
        rc = launch_static_jw<C, NW, 23>(ja, kp, st, done);      if (done) return rc;
        rc = launch_static_jw<C, NW, 31>(ja, kp, st, done);      if (done) return rc;
        return launch_variant<C, NW, true, -1>(ja, kp, st);
    }
    return launch_variant<C, NW, false, -1>(ja, kp, st);
}

int make_kernel_params(const agatha_params_t* p, KernelParams* kp)
{
    if (!p) return set_error(AGATHA_EINVAL, "params is NULL");
    if (p->band_width < 0) return set_error(AGATHA_EINVAL, "band_width < 0");
    if (p->slice_width < 1) return set_error(AGATHA_EINVAL, "slice_width < 1");
    const int C = shape_for(p->band_width).C;
    if (!C) return set_error(AGATHA_EUNSUPPORTED, "band_width %d > %d not supported by this build", p->band_width, agatha_max_band_width());
    kp->match = p->match; kp->mismatch = p->mismatch;
    kp->goe = p->gap_open + p->gap_extend;          // gasal_align.cu:301
    kp->ge = p->gap_extend;
    kp->sw = p->slice_width; kp->Z = p->z_threshold; kp->W = p->band_width;
    kp->LW = p->band_width / C; kp->JW = p->band_width % C;
    // PRMT table over x = query code ^ target code: 0 match, 1..3 mismatch, 4..7 N vs base (-N_PENALTY = -1);
    // x >= 8 selects the sign of entry x&7 replicated: 0xff == -1 for every negative entry (extend_kernel.cuh)
    const unsigned m = (unsigned)p->match & 0xffu, x = (unsigned)(-p->mismatch) & 0xffu;
    kp->tab_lo = m | (x << 8) | (x << 16) | (x << 24);
    kp->tab_hi = 0xffffffffu;
    kp->one = 1; kp->k32 = 32; kp->m16 = 0xffff;
    kp->force_generic = fast_table_ok(p) ? 0 : 1;
    // 16-bit packed steady state (extend_kernel.cuh run_fast16): needs small scoring values so that the per-window drift
    // bounds of its range monitor hold; AGATHA_S16=0 disables it, 1 = steady state only, 7 = also the tail (A/B measurements)
    const char* env = getenv("AGATHA_S16");
    kp->s16 = (!kp->force_generic && p->match >= 0 && p->match <= 100 && p->mismatch <= 100 && p->gap_open >= 0 && p->gap_extend >= 0 &&
               p->gap_open + 2 * p->gap_extend <= 2000 && !(env && env[0] == '0')) ? 1 : 0;
    // bit 1: the prologue (anti-diagonals 0..W) may run packed too, without a range monitor. On those anti-diagonals every
    // live value lies in [-(2*goe + ge*(W+1)) - mismatch*(W+2)/2 - goe, match*(W+2)/2] (a cell is at most (W+2)/2 diagonal
    // steps away from a matrix-edge value) and a dead cell creeps up by at most match*(W+2)/2 from the floor (-30000):
    // both must stay well apart and inside 16 bits.
    if (kp->s16) {
        const long long half = (p->band_width + 2) / 2;
        const long long depth = 3LL * kp->goe + (long long)kp->ge * (p->band_width + 1) + (long long)(p->mismatch + p->match) * half + 256;
        const bool only_steady = env && env[0] == '1' && env[1] == '\0', with_tail = env && env[0] == '7' && env[1] == '\0';
        if (depth < 24000 && !only_steady) kp->s16 |= 2;
        // bit 2: the tail (far matrix edges) packed as well (it keeps the range monitor, so it needs no bound of its own).
        // Opt-in (AGATHA_S16=7) and only present in builds with -DAGATHA_TAIL16=1: bit-exact and 12-14 % faster on equal-length
        // pairs, but on mixed-length batches the extra loop costs more in instruction fetch than it saves in issue slots
        // (C1: 21.2 -> 27.0 ms, ncu: no_instruction 1.6 -> 2.9 warps per issue).
        if (with_tail) kp->s16 |= 4;
    }
    return AGATHA_OK;
}

bool fast_table_ok(const agatha_params_t* p)
{
    return p->match >= -128 && p->match <= 127 && p->mismatch >= 1 && p->mismatch <= 128;
}

}  // namespace agatha

using namespace agatha;

extern "C" {

const char* agatha_last_error(void) { return g_err; }

int agatha_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int agatha_max_band_width(void) { return 8 * 32 * 32 - 1; }

uint64_t agatha_launch_count(void) { return g_launches.load(); }

int agatha_pack_device(const uint8_t* d_query_bases, uint64_t query_bytes, const uint8_t* d_target_bases, uint64_t target_bytes,
                       uint32_t* d_query_packed, uint32_t* d_target_packed, void* stream)
{
    if ((query_bytes & 7) || (target_bytes & 7)) return set_error(AGATHA_EINVAL, "batch bytes must be multiples of 8");
    if (query_bytes + target_bytes == 0) return AGATHA_OK;
    if (agatha_device_count() == 0) return set_error(AGATHA_ENODEV, "no CUDA device");
    const uint64_t words = (query_bytes + target_bytes) / 8;
    uint64_t blocks = (words + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    pack_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const uint2*)d_query_bases, query_bytes / 8,
                                                                     (const uint2*)d_target_bases, target_bytes / 8,
                                                                     d_query_packed, d_target_packed);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_error(e, "pack_kernel launch");
    return AGATHA_OK;
}

int agatha_apply_ops_device(const uint8_t* d_query_bases, const uint8_t* d_target_bases,
                            const uint32_t* d_query_offsets, const uint32_t* d_target_offsets,
                            const uint32_t* d_query_lens, const uint32_t* d_target_lens,
                            const uint8_t* d_query_ops, const uint8_t* d_target_ops, uint32_t n_alns,
                            uint32_t* d_query_packed, uint32_t* d_target_packed, void* stream)
{
    if (n_alns == 0) return AGATHA_OK;
    if (!d_query_bases || !d_target_bases || !d_query_offsets || !d_target_offsets || !d_query_lens || !d_target_lens ||
        !d_query_ops || !d_target_ops || !d_query_packed || !d_target_packed) return set_error(AGATHA_EINVAL, "NULL argument");
    if (agatha_device_count() == 0) return set_error(AGATHA_ENODEV, "no CUDA device");
    uint64_t blocks = (2ull * n_alns + 7) / 8;                       // 8 warps per block, one warp per sequence
    if (blocks > 148 * 8) blocks = 148 * 8;
    apply_ops_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d_query_bases, d_target_bases, d_query_offsets, d_target_offsets,
                                                                          d_query_lens, d_target_lens, d_query_ops, d_target_ops, n_alns,
                                                                          d_query_packed, d_target_packed);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_error(e, "apply_ops_kernel launch");
    return AGATHA_OK;
}

int agatha_extend_device(const uint32_t* d_query_packed, const uint32_t* d_target_packed,
                         const uint32_t* d_query_offsets, const uint32_t* d_target_offsets,
                         const uint32_t* d_query_lens, const uint32_t* d_target_lens,
                         const uint32_t* d_order, uint32_t n_alns, const agatha_params_t* params,
                         int32_t* d_score, int32_t* d_query_end, int32_t* d_target_end,
                         int32_t* d_stop, int32_t* d_dstop, void* d_workspace, void* stream)
{
    if (n_alns == 0) return set_error(AGATHA_EINVAL, "n_alns == 0");
    if (n_alns > 0x7fffffffu) return set_error(AGATHA_EINVAL, "n_alns too large");
    if (!d_query_packed || !d_target_packed || !d_query_offsets || !d_target_offsets || !d_query_lens || !d_target_lens ||
        !d_score || !d_query_end || !d_target_end || !d_workspace)
        return set_error(AGATHA_EINVAL, "NULL device pointer");
    KernelParams kp;
    int rc = make_kernel_params(params, &kp);
    if (rc) return rc;
    if (agatha_device_count() == 0) return set_error(AGATHA_ENODEV, "no CUDA device");
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(d_workspace, 0, AGATHA_WORKSPACE_BYTES, st);
    if (e != cudaSuccess) return cuda_error(e, "workspace memset");
    JobArrays ja;
    ja.qpk = d_query_packed; ja.tpk = d_target_packed;
    ja.qoff_w = d_query_offsets; ja.toff_w = d_target_offsets;
    ja.qlen = d_query_lens; ja.tlen = d_target_lens;
    ja.order = d_order;
    ja.score = d_score; ja.qend = d_query_end; ja.tend = d_target_end; ja.stop = d_stop; ja.dstop = d_dstop;
    ja.counter = (unsigned*)d_workspace;
    ja.n = (int)n_alns;
    const Shape sh = shape_for(kp.W);
    switch (sh.C * 100 + sh.NW) {
        case 201: return launch_c<2, 1>(ja, kp, st);
        case 401: return launch_c<4, 1>(ja, kp, st);
        case 801: return launch_c<8, 1>(ja, kp, st);
        case 1601: return launch_c<16, 1>(ja, kp, st);
        case 2401: return launch_c<24, 1>(ja, kp, st);
        case 3201: return launch_c<32, 1>(ja, kp, st);
        case 3202: return launch_c<32, 2>(ja, kp, st);
        case 3204: return launch_c<32, 4>(ja, kp, st);
        case 3208: return launch_c<32, 8>(ja, kp, st);
    }
    return set_error(AGATHA_EUNSUPPORTED, "no kernel for band_width %d", kp.W);
}

}  // extern "C"
