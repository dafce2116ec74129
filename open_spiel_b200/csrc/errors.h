// Thread-local error channel of the C ABI (b2s_last_error), shared by the translation units.
#pragma once
#include <cuda_runtime.h>

#include <string>

namespace b2s {
int fail(const std::string& m);                       // records the message, returns 1
int cuda_fail(cudaError_t e, const char* what);       // records "<what>: <cuda error>", returns 2
}  // namespace b2s
#define B2S_CU(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) return ::b2s::cuda_fail(_e, #x); } while (0)
