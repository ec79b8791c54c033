// Internal helpers shared by the translation units of libagatha_b200 (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include "agatha_b200.h"

namespace agatha {
int set_error(int code, const char* fmt, ...);
int cuda_error(cudaError_t e, const char* what);
void count_launch();
}  // namespace agatha
