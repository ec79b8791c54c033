// Host-side scheduling helpers of the C ABI: length-aware bucketing, cost-balanced sharding over devices and the
// cell accounting used for GCUPS. No GPU needed.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <vector>

#include "agatha_b200.h"
#include "engine_internal.h"

namespace agatha {

// Work estimate of one pair: cells inside the band (cost = min(qlen,tlen) * (2w+1), clipped to the matrix).
static inline uint64_t pair_cost(uint32_t ql, uint32_t tl, int32_t W)
{
    const uint64_t lo = std::min(ql, tl), hi = std::max(ql, tl);
    const uint64_t width = std::min<uint64_t>(2ull * (uint64_t)std::max(W, 0) + 1ull, hi);
    return lo * width + 64;   // + fixed per-job overhead so that empty pairs still count
}

}  // namespace agatha

using namespace agatha;

extern "C" {

// Replaces agatha_sort + host std::sort (agatha_kernel.h:434-458, gasal_align.cu:14-18). The reference sorts ascending by
// #block-anti-diagonals and interleaves 1 long + 3 short per warp; with a persistent work queue the equivalent is
// simply "most expensive first" (longest-processing-time order).
int agatha_bucket_order(const uint32_t* query_lens, const uint32_t* target_lens, uint32_t n, int32_t band_width, uint32_t* order_out)
{
    if (!query_lens || !target_lens || !order_out) return set_error(AGATHA_EINVAL, "NULL argument");
    std::vector<uint64_t> key(n);
    for (uint32_t i = 0; i < n; i++) {
        // cost in the high bits, inverted index in the low bits: one descending sort, stable in the input order
        const uint64_t c = std::min<uint64_t>(pair_cost(query_lens[i], target_lens[i], band_width), (1ull << 40) - 1);
        key[i] = (c << 24) | (uint64_t)(0xffffffu - (i & 0xffffffu));
    }
    if (n <= (1u << 24)) {
        std::sort(key.begin(), key.end(), std::greater<uint64_t>());
        for (uint32_t i = 0; i < n; i++) order_out[i] = 0xffffffu - (uint32_t)(key[i] & 0xffffffu);
    } else {
        std::vector<uint32_t> idx(n);
        std::iota(idx.begin(), idx.end(), 0u);
        std::stable_sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) { return (key[a] >> 24) > (key[b] >> 24); });
        std::copy(idx.begin(), idx.end(), order_out);
    }
    return AGATHA_OK;
}

int agatha_shard_pairs(const uint32_t* query_lens, const uint32_t* target_lens, uint64_t n, int32_t band_width,
                       int32_t n_shards, int32_t* shard_out)
{
    if (!query_lens || !target_lens || !shard_out) return set_error(AGATHA_EINVAL, "NULL argument");
    if (n_shards < 1) return set_error(AGATHA_EINVAL, "n_shards < 1");
    std::vector<uint64_t> idx(n);
    std::iota(idx.begin(), idx.end(), 0ull);
    std::vector<uint64_t> cost(n);
    for (uint64_t i = 0; i < n; i++) cost[i] = pair_cost(query_lens[i], target_lens[i], band_width);
    std::stable_sort(idx.begin(), idx.end(), [&](uint64_t a, uint64_t b) { return cost[a] > cost[b]; });
    std::vector<uint64_t> load((size_t)n_shards, 0);
    for (uint64_t k = 0; k < n; k++) {          // greedy LPT: next most expensive pair to the least loaded device
        int best = 0;
        for (int s = 1; s < n_shards; s++) if (load[s] < load[best]) best = s;
        shard_out[idx[k]] = best;
        load[best] += cost[idx[k]];
    }
    return AGATHA_OK;
}

int agatha_count_cells(const uint32_t* query_lens, const uint32_t* target_lens, const int32_t* dstop, uint64_t n,
                       int32_t band_width, uint64_t* cells_out, uint64_t* total_out)
{
    if (!query_lens || !target_lens) return set_error(AGATHA_EINVAL, "NULL argument");
    uint64_t total = 0;
    const int64_t W = band_width;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : total)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        const int64_t ql = query_lens[i], tl = target_lens[i];
        const int64_t ds = dstop ? (int64_t)dstop[i] : ql + tl;     // cells with q + r < ds
        uint64_t c = 0;
        for (int64_t q = 0; q < ql && q < ds; q++) {
            const int64_t lo = std::max<int64_t>(0, q - W);
            const int64_t hi = std::min<int64_t>(std::min<int64_t>(tl - 1, q + W), ds - q - 1);
            if (hi >= lo) c += (uint64_t)(hi - lo + 1);
        }
        if (cells_out) cells_out[i] = c;
        total += c;
    }
    if (total_out) *total_out = total;
    return AGATHA_OK;
}

uint64_t agatha_staged_bytes(const uint32_t* lens, const uint64_t* ids, uint64_t n)
{
    uint64_t tot = 0;
    for (uint64_t j = 0; j < n; j++) tot += ((uint64_t)lens[ids ? ids[j] : j] + 7) & ~7ull;
    return tot < 8 ? 8 : tot;
}

int agatha_stage_batch(const uint8_t* bases, const uint64_t* offsets, const uint32_t* lens, const uint64_t* ids, uint64_t n,
                       uint8_t* dst, uint64_t dst_capacity, uint32_t* dst_offsets, uint32_t* dst_lens, uint64_t* bytes_out, int32_t n_threads)
{
    if (!bases || !offsets || !lens || !dst || !dst_offsets) return set_error(AGATHA_EINVAL, "NULL argument");
    uint64_t o = 0;
    for (uint64_t j = 0; j < n; j++) {
        const uint64_t id = ids ? ids[j] : j;
        if (o > 0xfffffff8ull) return set_error(AGATHA_EINVAL, "batch exceeds 32-bit offsets");
        dst_offsets[j] = (uint32_t)o;
        if (dst_lens) dst_lens[j] = lens[id];
        o += ((uint64_t)lens[id] + 7) & ~7ull;
    }
    if (o == 0) o = 8;
    if (o > dst_capacity) return set_error(AGATHA_EINVAL, "staging buffer too small: need %llu bytes, have %llu", (unsigned long long)o, (unsigned long long)dst_capacity);
    if (n == 0) std::memset(dst, 'N', 8);
    if (n_threads <= 0) n_threads = 4;
#pragma omp parallel for schedule(static, 64) num_threads(n_threads)
    for (int64_t j = 0; j < (int64_t)n; j++) {
        const uint64_t id = ids ? ids[j] : (uint64_t)j;
        const uint32_t len = lens[id];
        uint8_t* d = dst + dst_offsets[j];
        std::memcpy(d, bases + offsets[id], len);
        std::memset(d + len, 'N', ((len + 7u) & ~7u) - len);
        if (len == 0 && n == 1) std::memset(d, 'N', 8);
    }
    if (bytes_out) *bytes_out = o;
    return AGATHA_OK;
}

}  // extern "C"
