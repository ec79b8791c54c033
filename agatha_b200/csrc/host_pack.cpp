// Host-side staging of a batch directly in the packed device format: 8 bases per 32-bit word, 4-bit codes, written into
// pinned memory while the sequences are copied -- half the H2D bytes of the reference's scheme (which uploads ASCII and packs
// on the device, host_batch.cpp:79-154 + kernels/pack_rc_seqs.h:13-53) and no pack kernel. Same layout and codes as
// pack_kernel (pack_kernel.cuh): every sequence starts at a multiple of 8 bases and is padded to a multiple of 8 with 'N';
// query words carry their first base in the top nibble, target words in the bottom nibble. The per-sequence op byte
// (bit 0 reverse, bit 1 complement; test_prog.cpp:83-92) is applied on the way, which replaces apply_ops_kernel on this path.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <immintrin.h>

#include "agatha_b200.h"
#include "engine_internal.h"

namespace agatha {

// nibble i = code of (ascii & 15) == i; the same tables as pack_kernel.cuh (kept in sync by tests/test_host_utils.py, which
// compares this packer with the emulated device kernel word by word)
static constexpr unsigned long long QCODE_LUT_H = 0xF4EDCBA928731605ull;
static constexpr unsigned long long TCODE_LUT_H = 0xFDE4CBA928731605ull;
static constexpr unsigned long long COMPLEMENT_LUT_H = 0xFEDCBA9836517240ull;

static inline unsigned code_of(unsigned long long lut, unsigned b) { return (unsigned)(lut >> (4u * (b & 15u))) & 15u; }

// scalar packer: any op, any length; also the tail of the vector packer
static void pack_scalar(const uint8_t* src, uint32_t len, uint32_t first_word, uint32_t n_words, bool target, unsigned op, uint32_t* dst)
{
    const unsigned long long lut = target ? TCODE_LUT_H : QCODE_LUT_H;
    for (uint32_t w = first_word; w < n_words; w++) {
        unsigned v = 0;
        for (unsigned b = 0; b < 8; b++) {
            const uint32_t pos = w * 8u + b;
            unsigned nib = 14u;                                          // 'N' padding (host_batch.cpp:143-146)
            if (pos < len) {
                nib = src[(op & 1u) ? len - 1u - pos : pos] & 15u;
                if (op & 2u) nib = (unsigned)(COMPLEMENT_LUT_H >> (4u * nib)) & 15u;
            }
            const unsigned c = code_of(lut, nib);
            v |= target ? c << (4u * b) : c << (28u - 4u * b);
        }
        dst[w] = v;
    }
}

// 32 bases -> 4 words per iteration; forward strand only (op == 0)
__attribute__((target("avx2"))) static uint32_t pack_avx2(const uint8_t* src, uint32_t len, bool target, uint32_t* dst)
{
    alignas(32) uint8_t lut[32];
    const unsigned long long l64 = target ? TCODE_LUT_H : QCODE_LUT_H;
    for (int i = 0; i < 16; i++) lut[i] = lut[16 + i] = (uint8_t)((l64 >> (4 * i)) & 15u);
    const __m256i vlut = _mm256_load_si256((const __m256i*)lut);
    const __m256i low4 = _mm256_set1_epi8(0x0f);
    // adjacent codes -> one byte: target = even | odd << 4, query = even << 4 | odd
    const __m256i mul = target ? _mm256_set1_epi16(0x1001) : _mm256_set1_epi16(0x0110);
    // query words want their first byte in the top byte of the word: reverse the bytes of every 32-bit group
    const __m256i rev = _mm256_setr_epi8(3, 2, 1, 0, 7, 6, 5, 4, 11, 10, 9, 8, 15, 14, 13, 12, 3, 2, 1, 0, 7, 6, 5, 4, 11, 10, 9, 8, 15, 14, 13, 12);
    const uint32_t full = len / 32u;
    for (uint32_t i = 0; i < full; i++) {
        const __m256i a = _mm256_loadu_si256((const __m256i*)(src + 32u * i));
        const __m256i c = _mm256_shuffle_epi8(vlut, _mm256_and_si256(a, low4));
        const __m256i h = _mm256_maddubs_epi16(c, mul);                 // 16 x 16-bit, each <= 255
        __m256i b = _mm256_packus_epi16(h, h);                          // per 128-bit lane: 8 bytes, duplicated
        if (!target) b = _mm256_shuffle_epi8(b, rev);
        const uint64_t lo = (uint64_t)_mm256_extract_epi64(b, 0), hi = (uint64_t)_mm256_extract_epi64(b, 2);
        std::memcpy(dst + 4u * i, &lo, 8);
        std::memcpy(dst + 4u * i + 2, &hi, 8);
    }
    return full * 4u;                                                   // words written
}

static const bool g_avx2 = __builtin_cpu_supports("avx2");

int pack_layout(const uint32_t* lens, const uint64_t* ids, uint64_t n, uint64_t dst_capacity_words, uint32_t* dst_offsets, uint32_t* dst_lens, uint64_t* bases_out)
{
    uint64_t o = 0;
    for (uint64_t j = 0; j < n; j++) {
        const uint64_t id = ids ? ids[j] : j;
        if (o > 0xfffffff8ull) return set_error(AGATHA_EINVAL, "batch exceeds 32-bit offsets");
        dst_offsets[j] = (uint32_t)o;
        if (dst_lens) dst_lens[j] = lens[id];
        o += ((uint64_t)lens[id] + 7) & ~7ull;
    }
    if (o == 0) o = 8;
    if (o / 8 > dst_capacity_words) return set_error(AGATHA_EINVAL, "packed staging buffer too small: need %llu words, have %llu", (unsigned long long)(o / 8), (unsigned long long)dst_capacity_words);
    if (bases_out) *bases_out = o;
    return AGATHA_OK;
}

void pack_range(const PackView& v, uint64_t j0, uint64_t j1)
{
    for (uint64_t j = j0; j < j1; j++) {
        const uint64_t id = v.ids ? v.ids[j] : j;
        const uint32_t len = v.lens[id];
        const unsigned op = v.ops ? (v.ops[id] & 3u) : 0u;
        const uint8_t* src = v.bases + v.offsets[id];
        uint32_t* d = v.dst_words + (v.dst_offsets[j] >> 3);
        uint32_t done = 0;
        if (op == 0 && g_avx2) done = pack_avx2(src, len, v.is_target != 0, d);
        pack_scalar(src, len, done, (len + 7u) / 8u, v.is_target != 0, op, d);
    }
}

void pack_empty(int is_target, uint32_t* dst_words) { pack_scalar(nullptr, 0, 0, 1, is_target != 0, 0, dst_words); }

}  // namespace agatha

using namespace agatha;

extern "C" int agatha_pack_batch(const uint8_t* bases, const uint64_t* offsets, const uint32_t* lens, const uint64_t* ids, const uint8_t* ops,
                                 uint64_t n, int32_t is_target, uint32_t* dst_words, uint64_t dst_capacity_words,
                                 uint32_t* dst_offsets, uint32_t* dst_lens, uint64_t* bases_out, int32_t n_threads)
{
    if (!bases || !offsets || !lens || !dst_words || !dst_offsets) return set_error(AGATHA_EINVAL, "NULL argument");
    uint64_t o = 0;
    int rc = pack_layout(lens, ids, n, dst_capacity_words, dst_offsets, dst_lens, &o);
    if (rc) return rc;
    if (n == 0 || (n == 1 && lens[ids ? ids[0] : 0] == 0)) pack_empty(is_target, dst_words);
    if (n_threads <= 0) n_threads = 4;
    const PackView v{bases, offsets, lens, ids, ops, is_target, dst_words, dst_offsets};
    const int64_t n_chunks = (int64_t)((n + 15) / 16);
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads)
    for (int64_t c = 0; c < n_chunks; c++) pack_range(v, (uint64_t)c * 16, std::min<uint64_t>(n, (uint64_t)c * 16 + 16));
    if (bases_out) *bases_out = o;
    return AGATHA_OK;
}
