// Internal helpers shared by the translation units of libagatha_b200 (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include "agatha_b200.h"

namespace agatha {
int set_error(int code, const char* fmt, ...);
int cuda_error(cudaError_t e, const char* what);
void count_launch();

// host packer (host_pack.cpp), split so that the job scheduler can spread one batch over its own packing threads:
// pack_layout lays the sequences out (each at a multiple of 8 bases) and checks the capacity, pack_range packs sequences
// [j0, j1) of that layout. agatha_pack_batch (ABI) is the two together.
struct PackView {
    const uint8_t* bases; const uint64_t* offsets; const uint32_t* lens; const uint64_t* ids; const uint8_t* ops;
    int is_target; uint32_t* dst_words; const uint32_t* dst_offsets;
};
int pack_layout(const uint32_t* lens, const uint64_t* ids, uint64_t n, uint64_t dst_capacity_words, uint32_t* dst_offsets, uint32_t* dst_lens, uint64_t* bases_out);
void pack_range(const PackView& v, uint64_t j0, uint64_t j1);
void pack_empty(int is_target, uint32_t* dst_words);          // the one padding word of a batch without bases
}  // namespace agatha
