// Launcher of one kernel variant; explicitly instantiated in extend_inst_*.cu so that the variants compile in parallel.
#pragma once
#include <mutex>

#include "extend_kernel.cuh"
#include "engine_internal.h"

namespace agatha {

template <int C, int NW, bool WODD, int JWS>
int launch_variant(const JobArrays& ja, const KernelParams& kp, cudaStream_t st);

#ifdef AGATHA_DEFINE_LAUNCH
template <int C, int NW, bool WODD, int JWS>
int launch_variant(const JobArrays& ja, const KernelParams& kp, cudaStream_t st)
{
    // resident CTAs per SM and SM count, once per process (all devices of a box are identical)
    static std::once_flag once;
    static int blocks_per_sm = 1, sms = 1;
    constexpr int threads = KernelShape<C, NW>::threads;
    std::call_once(once, [] {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        int b = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, extend_kernel<C, NW, WODD, JWS>, threads, 0);
        blocks_per_sm = b > 0 ? b : 1;
    });
    // persistent groups: never more groups than jobs, otherwise fill every SM
    const int groups_per_block = NW == 1 ? 4 : 1;
    long long want = ((long long)ja.n + groups_per_block - 1) / groups_per_block;
    long long grid = (long long)sms * blocks_per_sm;
    if (want < grid) grid = want;
    if (grid < 1) grid = 1;
    extend_kernel<C, NW, WODD, JWS><<<(unsigned)grid, threads, 0, st>>>(ja, kp);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_error(e, "extend_kernel launch");
    return AGATHA_OK;
}
#define AGATHA_INSTANTIATE(C, NW, WODD, JWS) template int launch_variant<C, NW, WODD, JWS>(const JobArrays&, const KernelParams&, cudaStream_t);
#endif

}  // namespace agatha
