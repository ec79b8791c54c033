// Measured integer issue rates of the device this library runs on: the denominators of bench.py's roofline.
// The extension kernels are bound by the ALU pipe (VIADDMNMX / VIMNMX3 / PRMT / LOP3 / SHF all issue there, 16 lanes per SM
// sub-partition per clock), with IMAD on the separate FMA pipe. MEASURED_PEAKS.json has HBM and bf16 numbers only
// (SURVEY.md section 8d), so bench.py measures the integer pipes itself, in the same run, through this entry point.
#include <cstdint>
#include <cuda_runtime.h>

#include "agatha_b200.h"
#include "engine_internal.h"

namespace agatha {

constexpr int PEAK_ILP = 8, PEAK_ITERS = 1 << 14;

// MODE 0: VIADDMNMX.U16x2 only (ALU pipe); 1: IMAD only (FMA pipe); 2: both interleaved 1:1
template <int MODE>
__global__ void __launch_bounds__(256) int_peak_kernel(unsigned* out, unsigned a0, unsigned b0, unsigned c0)
{
    unsigned v[PEAK_ILP];
#pragma unroll
    for (int i = 0; i < PEAK_ILP; i++) v[i] = a0 + threadIdx.x * (i + 1);
    const unsigned b = b0 + (threadIdx.x & 3), c = c0;
#pragma unroll 1
    for (int it = 0; it < PEAK_ITERS; it++) {
#pragma unroll
        for (int i = 0; i < PEAK_ILP; i++) {
            if (MODE == 0 || (MODE == 2 && (i & 1) == 0)) { v[i] = __viaddmax_u16x2(v[i], b, c); asm volatile("" : "+r"(v[i])); }
            else { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(v[i]) : "r"(b), "r"(c)); }
        }
    }
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < PEAK_ILP; i++) s ^= v[i];
    if (s == 0x12345678u) out[0] = s;            // keeps the chains alive
}

template <int MODE>
static int run_peak(int sms, unsigned* d_out, double* tera)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = sms * 8;                    // 8 CTAs x 8 warps: every sub-partition has 16 warps to pick from
    int_peak_kernel<MODE><<<blocks, 256>>>(d_out, 1u, 3u, 7u);   // warm-up
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        int_peak_kernel<MODE><<<blocks, 256>>>(d_out, 1u, 3u, 7u);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_error(e, "int_peak_kernel");
    *tera = (double)blocks * 256.0 * PEAK_ILP * PEAK_ITERS / (best * 1e-3) / 1e12;
    return AGATHA_OK;
}

}  // namespace agatha

using namespace agatha;

extern "C" int agatha_measure_int_peak(int device, double* alu_tera_lane_ops, double* fma_tera_lane_ops, double* mixed_tera_lane_ops)
{
    if (agatha_device_count() == 0) return set_error(AGATHA_ENODEV, "no CUDA device");
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return cuda_error(e, "cudaSetDevice");
    int sms = 1;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    unsigned* d_out = nullptr;
    if ((e = cudaMalloc((void**)&d_out, 256)) != cudaSuccess) return cuda_error(e, "cudaMalloc");
    double a = 0, f = 0, m = 0;
    int rc = run_peak<0>(sms, d_out, &a);
    if (!rc) rc = run_peak<1>(sms, d_out, &f);
    if (!rc) rc = run_peak<2>(sms, d_out, &m);
    cudaFree(d_out);
    if (rc) return rc;
    if (alu_tera_lane_ops) *alu_tera_lane_ops = a;
    if (fma_tera_lane_ops) *fma_tera_lane_ops = f;
    if (mixed_tera_lane_ops) *mixed_tera_lane_ops = m;
    return AGATHA_OK;
}
