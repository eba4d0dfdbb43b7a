// Shared helpers for the mava_b200 kernels.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/mava_b200.h"

#define MAVA_CHECK_ARG(cond) \
  do {                       \
    if (!(cond)) return MAVA_E_BADARG; \
  } while (0)

#define MAVA_CHECK_PTR(p) \
  do {                    \
    if ((p) == nullptr) return MAVA_E_NULL; \
  } while (0)

namespace mava {

inline int launch_status() {
  cudaError_t e = cudaPeekAtLastError();
  return e == cudaSuccess ? 0 : (int)e;
}

inline cudaStream_t as_stream(mava_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

inline int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ inline int round_up(int a, int b) { return ceil_div(a, b) * b; }

}  // namespace mava
