// Launchers of the kernel variants; explicitly instantiated in extend_inst_*.cu / extend16_inst_*.cu so that the variants
// compile in parallel.
#pragma once
#include <atomic>

#include "extend_kernel.cuh"
#include "extend16_kernel.cuh"
#include "engine_internal.h"

namespace agatha {

template <int C, int NW, bool WODD, int JWS>
int launch_variant(const JobArrays& ja, const KernelParams& kp, cudaStream_t st);
template <int C, int NW, int JWS>
int launch16_variant(const JobArrays& ja, const KernelParams& kp, cudaStream_t st);

#ifdef AGATHA_DEFINE_LAUNCH
// Persistent grid = SMs x resident CTAs of the launching device. Cached per device ordinal and kernel variant: the job API
// launches on several devices from several threads, and a box may mix GPUs (or MIG slices) of different sizes.
constexpr int MAX_DEVICES = 64;
template <class K>
static int persistent_grid(K kernel, int threads)
{
    static std::atomic<int> cache[MAX_DEVICES];           // zero-initialised; 0 = not known yet
    int dev = 0;
    cudaGetDevice(&dev);
    const int slot = (dev >= 0 && dev < MAX_DEVICES) ? dev : -1;
    if (slot >= 0) { const int g = cache[slot].load(std::memory_order_relaxed); if (g > 0) return g; }
    int sms = 1, b = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kernel, threads, 0);
    const int g = sms * (b > 0 ? b : 1);
    if (slot >= 0) cache[slot].store(g, std::memory_order_relaxed);
    return g;
}

template <int C, int NW, bool WODD, int JWS>
int launch_variant(const JobArrays& ja, const KernelParams& kp, cudaStream_t st)
{
    constexpr int threads = KernelShape<C, NW>::threads;
    // persistent groups: never more groups than jobs, otherwise fill every SM
    const int groups_per_block = NW == 1 ? 4 : 1;
    long long want = ((long long)ja.n + groups_per_block - 1) / groups_per_block;
    // the redo pass exists for the variants the packed kernel precedes: W = 7 (mod 8), at least 8 cells per lane
    constexpr bool HAS_REDO = WODD && JWS >= 0 && C >= 8;
    if (ja.redo && !HAS_REDO) return set_error(AGATHA_EUNSUPPORTED, "no redo pass for this kernel variant");
    long long grid = persistent_grid(extend_kernel<C, NW, WODD, JWS, false>, threads);
    if (want < grid) grid = want;
    if (grid < 1) grid = 1;
    if constexpr (HAS_REDO) {
        if (ja.redo) extend_kernel<C, NW, WODD, JWS, true><<<(unsigned)grid, threads, 0, st>>>(ja, kp);
        else extend_kernel<C, NW, WODD, JWS, false><<<(unsigned)grid, threads, 0, st>>>(ja, kp);
    } else {
        extend_kernel<C, NW, WODD, JWS, false><<<(unsigned)grid, threads, 0, st>>>(ja, kp);
    }
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_error(e, "extend_kernel launch");
    return AGATHA_OK;
}

template <int C, int NW, int JWS>
int launch16_variant(const JobArrays& ja, const KernelParams& kp, cudaStream_t st)
{
    constexpr int threads = Shape16<C, NW>::threads;
    long long want = ((long long)ja.n + Shape16<C, NW>::groups - 1) / Shape16<C, NW>::groups;
    long long grid = persistent_grid(extend16_kernel<C, NW, JWS>, threads);
    if (want < grid) grid = want;
    if (grid < 1) grid = 1;
    extend16_kernel<C, NW, JWS><<<(unsigned)grid, threads, 0, st>>>(ja, kp);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_error(e, "extend16_kernel launch");
    return AGATHA_OK;
}
#define AGATHA_INSTANTIATE(C, NW, WODD, JWS) template int launch_variant<C, NW, WODD, JWS>(const JobArrays&, const KernelParams&, cudaStream_t);
#define AGATHA_INSTANTIATE16(C, NW, JWS) template int launch16_variant<C, NW, JWS>(const JobArrays&, const KernelParams&, cudaStream_t);
#endif

}  // namespace agatha
