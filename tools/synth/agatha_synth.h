/* BENCH / TEST TOOLING -- deterministic synthetic workloads (tools/synth/synth.cpp -> agatha_b200/lib/libagatha_synth.so).
 * Kept out of libagatha_b200.so on purpose: the reference arm of bench.py uses the generator and must not map the product. */
#ifndef AGATHA_SYNTH_H
#define AGATHA_SYNTH_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Deterministic synthetic read/reference pairs. profile: 1 = C1, 2 = ONT-like, 3 = HiFi-like,
 * 4 = heavy tail with early Z-drop. Two passes: sizes first (bases == NULL), then fill. Offsets in bytes, no padding.
 * Returns 0, -1 (bad argument) or -2 (base buffers too small). */
int agatha_synth_pairs(int32_t profile, uint64_t seed, uint64_t first_pair, uint64_t n_pairs,
                       uint32_t *query_lens, uint32_t *target_lens,
                       uint64_t *query_offsets, uint64_t *target_offsets,
                       uint8_t *query_bases, uint64_t query_capacity,
                       uint8_t *target_bases, uint64_t target_capacity, int32_t n_threads);

#ifdef __cplusplus
}
#endif
#endif
