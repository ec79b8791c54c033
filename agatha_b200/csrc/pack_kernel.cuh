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

// nibble i of the LUT = complement of low-nibble i: A(1) <-> T(4), C(3) <-> G(7), everything else unchanged
// (the switch of gasal_reversecomplement_kernel, pack_rc_seqs.h:178-197)
constexpr unsigned long long COMPLEMENT_LUT = 0xFEDCBA9836517240ull;

// Per-sequence reverse / complement ("op": bit 0 = reverse, bit 1 = complement; test_prog.cpp:83-92), the job of
// gasal_reversecomplement_kernel (pack_rc_seqs.h:56-212, launched by gasal_aln_async when params->isReverseComplement,
// gasal_align.cu:199-212). Runs AFTER pack_kernel: the sequences that carry an op are packed again from the ASCII batch
// with the transformation applied to their real bases; the 'N' padding stays behind the sequence. One warp per sequence,
// one lane per packed word; sequences without an op cost one byte load.
__global__ void __launch_bounds__(256) apply_ops_kernel(const uint8_t* __restrict__ qin, const uint8_t* __restrict__ tin,
                                                        const uint32_t* __restrict__ qoff, const uint32_t* __restrict__ toff,
                                                        const uint32_t* __restrict__ qlen, const uint32_t* __restrict__ tlen,
                                                        const uint8_t* __restrict__ qop, const uint8_t* __restrict__ top, uint32_t n,
                                                        uint32_t* __restrict__ qout, uint32_t* __restrict__ tout)
{
    const unsigned lane = threadIdx.x & 31u;
    const uint64_t warps = (uint64_t)gridDim.x * (blockDim.x >> 5);
    for (uint64_t w = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < 2ull * n; w += warps) {
        const bool target = w >= n;
        const uint32_t i = (uint32_t)(target ? w - n : w);
        const unsigned op = target ? top[i] : qop[i];
        if (!(op & 3u)) continue;
        const uint32_t off = target ? toff[i] : qoff[i], len = target ? tlen[i] : qlen[i];
        const uint8_t* src = (target ? tin : qin) + off;
        uint32_t* dst = (target ? tout : qout) + (off >> 3);
        const unsigned long long lut = target ? TCODE_LUT : QCODE_LUT;
        for (uint32_t word = lane; word < (len + 7u) / 8u; word += 32u) {
            unsigned v = 0;
#pragma unroll
            for (unsigned b = 0; b < 8; b++) {
                const uint32_t pos = word * 8u + b;
                unsigned nib = 14u;                                  // 'N' padding (host_batch.cpp:143-146)
                if (pos < len) {
                    nib = src[(op & 1u) ? len - 1u - pos : pos] & 15u;
                    if (op & 2u) nib = (unsigned)(COMPLEMENT_LUT >> (4u * nib)) & 15u;
                }
                const unsigned code = code_of(lut, nib);
                v |= target ? code << (4u * b) : code << (28u - 4u * b);
            }
            dst[word] = v;
        }
    }
}

}  // namespace agatha
