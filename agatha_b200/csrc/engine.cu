// Device-level entry points of the C ABI (include/agatha_b200.h): packing and the extension kernel launch.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>

#include "extend_launch.cuh"
#include "extend_dispatch.h"
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

struct CudaLauncher {
    const JobArrays& ja; const KernelParams& kp; cudaStream_t st;
    template <int C, int NW, bool WODD, int JWS> int run() const { return launch_variant<C, NW, WODD, JWS>(ja, kp, st); }
    template <int C, int NW, int JWS> int run16() const { return launch16_variant<C, NW, JWS>(ja, kp, st); }
};

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
    ja.redo = 0;
    const CudaLauncher l{ja, kp, st};
    // The packed kernel first (where it applies); it marks the pairs it cannot finish exactly and the general kernel,
    // second queue counter, aligns those. Otherwise the general kernel aligns everything.
    if (dispatch16_variant(kp, l, &rc)) {
        if (rc) return rc;
        ja.counter = (unsigned*)d_workspace + 1;
        ja.redo = 1;
    }
    if (dispatch_variant(kp, l, &rc)) return rc;
    return set_error(AGATHA_EUNSUPPORTED, "no kernel for band_width %d", kp.W);
}

}  // extern "C"
