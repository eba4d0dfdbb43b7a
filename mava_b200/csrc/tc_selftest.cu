// Self-test of the tcgen05 operand conventions in tc.cuh: one 128 x N x K GEMM per launch in each
// of the three operand arrangements the MLP kernels use.  Exposed through the C ABI so the GPU
// tests can pin descriptor semantics on real hardware before trusting the fused kernels.
#include <cstdlib>

#include "common.cuh"
#include "tc.cuh"

namespace mava {
namespace {

using namespace tc;

// Convert a row-major f32 matrix [rows][cols] (leading dimension ld) into a bf16 core-matrix tile.
__device__ void fill_tile_f32(const Tile& t, const float* __restrict__ src, int rows, int cols,
                              int ld, int rows_valid, int cols_valid) {
  const int chunks = rows * (cols >> 3);
  for (int i = threadIdx.x; i < chunks; i += blockDim.x) {
    const int r = i % rows, cg = i / rows;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = cg * 8 + j;
      v[j] = (r < rows_valid && c < cols_valid) ? src[(size_t)r * ld + c] : 0.0f;
    }
    st_shared_v4(t.base + chunk_off(t, r, cg), pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]),
                 pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
  }
}

// mode 0: D = A[128][K] * B[K][N]          A K-major, B MN-major   (forward   X W)
// mode 1: D = A[128][K] * B[N][K]^T        A K-major, B K-major    (backward  dZ W^T)
// mode 2: D = A[K][128]^T * B[K][N]        A MN-major, B MN-major  (wgrad     H^T dZ)
__global__ void __launch_bounds__(128)
tc_selftest_kernel(int mode, const float* __restrict__ A, const float* __restrict__ B,
                   float* __restrict__ D, int N, int K, int dcol) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;

  // tile shapes [rows][cols] in memory
  const int a_rows = mode == 2 ? K : 128, a_cols = mode == 2 ? 128 : K;
  const int b_rows = mode == 1 ? N : K, b_cols = mode == 1 ? K : N;
  Tile ta{smem_u32(smem), 128u, (uint32_t)(a_rows / 8) * 128u};
  Tile tb{ta.base + tile_bytes(a_rows, a_cols), 128u, (uint32_t)(b_rows / 8) * 128u};

  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  fill_tile_f32(ta, A, a_rows, a_cols, a_cols, a_rows, a_cols);
  fill_tile_f32(tb, B, b_rows, b_cols, b_cols, b_rows, b_cols);
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_alloc_base = tmem_base_s;
  const uint32_t tmem = tmem_alloc_base + (uint32_t)dcol;

  if (threadIdx.x == 0) {
    const uint32_t idesc = instr_desc(128, N, mode == 2, mode != 1);
    for (int k = 0; k < K / 16; ++k) {
      const uint64_t ad = mode == 2 ? desc_mnmajor(ta, k) : desc_kmajor(ta, k);
      const uint64_t bd = mode == 1 ? desc_kmajor(tb, k) : desc_mnmajor(tb, k);
      mma(tmem, ad, bd, idesc, k > 0);
    }
    commit(&bar);
  }
  mbar_wait(&bar, 0);
  fence_after_sync();
  const int row = threadIdx.x;
  for (int c0 = 0; c0 < N; c0 += 16) {
    float v[16];
    ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
    for (int j = 0; j < 16; ++j) D[(size_t)row * N + c0 + j] = v[j];
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem_alloc_base);
}

}  // namespace
}  // namespace mava

extern "C" int mava_tc_selftest(int mode, const float* A, const float* B, float* D, int N, int K,
                                mava_stream_t s) {
  using namespace mava;
  MAVA_CHECK_PTR(A);
  MAVA_CHECK_PTR(B);
  MAVA_CHECK_PTR(D);
  MAVA_CHECK_ARG(mode >= 0 && mode <= 2);
  MAVA_CHECK_ARG(N >= 16 && N <= 128 && N % 16 == 0 && K >= 16 && K % 16 == 0 && K <= 320);
  const size_t smem = (size_t)128 * K * 2 + (size_t)K * N * 2 + 256;
  cudaError_t e = cudaFuncSetAttribute(tc_selftest_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  const char* dc = getenv("MAVA_TC_DCOL");
  tc_selftest_kernel<<<1, 128, smem, as_stream(s)>>>(mode, A, B, D, N, K, dc ? atoi(dc) : 0);
  return launch_status();
}
