// ASCII -> 4-bit packing for the extension kernel. Replaces gasal_pack_kernel (AGAThA/src/kernels/pack_rc_seqs.h:13-53).
//
// The reference keeps (ascii & 15) and packs 8 bases per word, first base in the top nibble. This library keeps the
// same information (a bijection of the 16 possible low nibbles, so equality of bases is preserved exactly) but
// renumbers it so that the five symbols real reads contain score with one PRMT lookup of the code XOR:
//      nibble(ascii&15):  1(A) 3(C) 7(G) 4(T) 14(N)   others
//      query  code        0    1    2    3    4       5..15
//      target code        0    1    2    3    13      4..12,14,15
// Query words are packed first-base-in-top-nibble (the query window of a lane shifts left), target words
// first-base-in-bottom-nibble (the target window shifts right); see extend_kernel.cuh.
#pragma once
#include <cstdint>

namespace agatha {

constexpr unsigned long long QCODE_LUT = 0xF4EDCBA928731605ull;   // nibble i of the LUT = query code of low-nibble i
constexpr unsigned long long TCODE_LUT = 0xFDE4CBA928731605ull;

__device__ __forceinline__ unsigned code_of(unsigned long long lut, unsigned byte)
{
    return (unsigned)(lut >> (4u * (byte & 15u))) & 15u;
}

// one thread = 8 bases = one packed word; grid-stride, 8-byte coalesced loads
__global__ void __launch_bounds__(256) pack_kernel(const uint2* __restrict__ qin, uint64_t qwords,
                                                   const uint2* __restrict__ tin, uint64_t twords,
                                                   uint32_t* __restrict__ qout, uint32_t* __restrict__ tout)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < qwords + twords; i += stride) {
        if (i < qwords) {
            const uint2 v = __ldg(qin + i);
            unsigned w = 0;
#pragma unroll
            for (int b = 0; b < 4; b++) {
                w |= code_of(QCODE_LUT, v.x >> (8 * b)) << (28 - 4 * b);
                w |= code_of(QCODE_LUT, v.y >> (8 * b)) << (12 - 4 * b);
            }
            qout[i] = w;
        } else {
            const uint64_t k = i - qwords;
            const uint2 v = __ldg(tin + k);
            unsigned w = 0;
#pragma unroll
            for (int b = 0; b < 4; b++) {
                w |= code_of(TCODE_LUT, v.x >> (8 * b)) << (4 * b);
                w |= code_of(TCODE_LUT, v.y >> (8 * b)) << (16 + 4 * b);
            }
            tout[k] = w;
        }
    }
}

}  // namespace agatha
