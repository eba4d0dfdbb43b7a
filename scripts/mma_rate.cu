// Development aid: issue-to-completion cost of tcgen05.mma (cta_group::1, M = 128, bf16, K = 16
// per instruction) on the no-swizzle core-matrix operand tiles of csrc/tc.cuh, per operand
// arrangement and N.  Build and run on a B200:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I mava_b200/csrc -I include \
//        scripts/mma_rate.cu -o /tmp/mma_rate && /tmp/mma_rate
#include <cstdio>

#include "tc.cuh"

using namespace mava::tc;

__global__ void __launch_bounds__(128) rate_kernel(int mode, int N, int ksteps, int reps,
                                                   long long* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_s;
  const int K = ksteps * 16;
  const int a_rows = mode == 2 ? K : 128, a_cols = mode == 2 ? 128 : K;
  const int b_rows = mode == 1 ? N : K, b_cols = mode == 1 ? K : N;
  Tile ta{smem_u32(smem), 128u, (uint32_t)(a_rows / 8) * 128u};
  Tile tb{ta.base + tile_bytes(a_rows, a_cols), 128u, (uint32_t)(b_rows / 8) * 128u};
  for (int i = threadIdx.x; i < (128 * K + K * N) / 2; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x < 32) tmem_alloc<512>(&tmem_s);
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_s;
  uint32_t phase = 0;
  long long t0 = 0, t1 = 0, t2 = 0;
  for (int it = 0; it < 3; ++it) {  // the last iteration is the measured one
    t0 = clock64();
    if (mma_issuer()) {
      const uint32_t idesc = instr_desc(128, N, mode == 2, mode != 1);
      // descriptors built once; the K steps differ by a constant in the start-address field
      const uint64_t ad0 = mode == 2 ? desc_mnmajor(ta, 0) : desc_kmajor(ta, 0);
      const uint64_t bd0 = mode == 1 ? desc_kmajor(tb, 0) : desc_mnmajor(tb, 0);
      const uint64_t ai = (mode == 2 ? desc_mnmajor(ta, 1) : desc_kmajor(ta, 1)) - ad0;
      const uint64_t bi = (mode == 1 ? desc_kmajor(tb, 1) : desc_mnmajor(tb, 1)) - bd0;
      for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int k = 0; k < 8; ++k) mma(tmem, ad0 + k * ai, bd0 + k * bi, idesc, k > 0);
      }
      commit(&bar);
    }
    t1 = clock64();
    mbar_wait(&bar, phase);
    phase ^= 1;
    t2 = clock64();
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<512>(tmem);
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const char* names[3] = {"A K-major, B MN-major (forward)", "A K-major, B K-major (backward)",
                          "A MN-major, B MN-major (wgrad)"};
  const int Ns[] = {16, 80, 128, 144, 256};
  for (int mode = 0; mode < 3; ++mode)
    for (int N : Ns) {
      for (int reps : {1, 8}) {
        const int ksteps = 8;
        rate_kernel<<<1, 128, 128 * 128 * 2 + 128 * 256 * 2 + 256>>>(mode, N, ksteps, reps, d);
        long long h[2];
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        cudaError_t e = cudaGetLastError();
        printf("%-34s N=%3d  %3d MMAs: issue %6lld cyc, done %6lld cyc (%.1f per MMA) %s\n",
               names[mode], N, reps * ksteps, h[0], h[1], (double)h[1] / (reps * ksteps),
               e == cudaSuccess ? "" : cudaGetErrorString(e));
      }
    }
  return 0;
}
