// Development aid: TMEM read bandwidth of tcgen05.ld.32x32b.x32 with 4, 8 or 16 warps of one CTA
// reading a 128-lane x 128-column fp32 accumulator (warp w: lanes 32 (w % 4).., columns 32 (w / 4)..).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I mava_b200/csrc -I include \
//        scripts/tmem_rate.cu -o /tmp/tmem_rate && /tmp/tmem_rate
#include <cstdio>

#include "tc.cuh"

using namespace mava::tc;

__global__ void __launch_bounds__(512) tmem_kernel(int warps, int reps, long long* out, float* sink) {
  __shared__ uint32_t tmem_s;
  __shared__ long long tmax;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(&tmem_s);
  if (threadIdx.x == 0) tmax = 0;
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_s;
  float acc = 0.0f;
  __syncthreads();
  const long long t0 = clock64();
  if (warp < warps) {
    // with fewer than 16 warps each warp walks over several column blocks
    const int blocks = 16 / warps;
    for (int r = 0; r < reps; ++r)
      for (int b = 0; b < blocks; ++b) {
        float v[32];
        const int cb = (warp >> 2) * blocks + b;
        ld32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(cb * 32), v);
#pragma unroll
        for (int i = 0; i < 32; ++i) acc += v[i];
      }
  }
  const long long t1 = clock64();
  atomicMax((unsigned long long*)&tmax, (unsigned long long)(t1 - t0));
  __syncthreads();
  if (threadIdx.x == 0) out[0] = tmax;
  if (acc == 12345.678f) sink[0] = acc;
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

int main() {
  long long* d;
  float* sink;
  cudaMalloc(&d, 16);
  cudaMalloc(&sink, 16);
  for (int warps : {4, 8, 16})
    for (int reps : {1, 16}) {
      tmem_kernel<<<1, 512>>>(warps, reps, d, sink);
      long long h;
      cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      const double bytes = 128.0 * 128 * 4 * reps;
      printf("%2d warps, %2d x (128 x 128 fp32 = 64 KB): %6lld cycles -> %.1f B/clk (%s)\n", warps, reps, h,
             bytes / h, cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
