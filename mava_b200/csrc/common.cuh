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
  static int n[32] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  const int slot = dev < 0 || dev >= 32 ? 0 : dev;
  if (n[slot] == 0) {
    cudaDeviceGetAttribute(&n[slot], cudaDevAttrMultiProcessorCount, dev);
    if (n[slot] <= 0) n[slot] = 148;
  }
  return n[slot];
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device function attribute: remember what has
// been configured per (call site, device), so that a second GPU driven by the same process gets its
// own cudaFuncSetAttribute (one process per GPU is the normal deployment; tests and population runs
// may differ).  `cache` is the call site's static array.
constexpr int kMaxDevices = 32;
template <typename K>
inline int ensure_dyn_smem(K kernel, size_t bytes, size_t (&cache)[kMaxDevices]) {
  int dev = 0;
  cudaGetDevice(&dev);
  const int slot = dev < 0 || dev >= kMaxDevices ? 0 : dev;
  if (bytes > 48 * 1024 && bytes > cache[slot]) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return (int)e;
    cache[slot] = bytes;
  }
  return 0;
}

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ inline int round_up(int a, int b) { return ceil_div(a, b) * b; }

}  // namespace mava
