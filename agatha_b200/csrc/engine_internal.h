// Internal helpers shared by the translation units of libagatha_b200 (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include "agatha_b200.h"

namespace agatha {
struct KernelParams;
int set_error(int code, const char* fmt, ...);
int cuda_error(cudaError_t e, const char* what);
void count_launch();
int make_kernel_params(const agatha_params_t* p, KernelParams* kp);
bool fast_table_ok(const agatha_params_t* p);
}  // namespace agatha
