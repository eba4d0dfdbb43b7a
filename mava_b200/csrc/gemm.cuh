// Argument block shared by the two dense-contraction kernels of the recurrent path:
// sgemm_kernel (fp32 SIMT, rnn_f32.cu) and tc_gemm_kernel (bf16 tcgen05, gemm_tc.cu).
//   C[M][N] (op)= alpha * opA(A)[M][K] * opB(B)[K][N] (+ bias) (relu) (zeroed where relu_ref <= 0)
#pragma once
#include "common.cuh"

namespace mava {

struct GemmArgs {
  const float* A;
  const float* B;
  float* C;
  int M, N, K;
  int64_t lda, ldb, ldc;
  int ta;  // 0: A(i,k) = A[i*lda + k]   1: A(i,k) = A[k*lda + i]
  int tb;  // 0: B(k,j) = B[k*ldb + j]   1: B(k,j) = B[j*ldb + k]
  const float* bias;      // [N] or null
  const float* relu_ref;  // [M][ldr] or null: result zeroed where relu_ref <= 0
  int64_t ldr;
  int relu;
  int mode;  // 0 store, 1 accumulate (+=), 2 atomicAdd
  float alpha;
  int kchunk;  // K range per blockIdx.z (multiple of 64)
};

// bf16 operands / fp32 accumulation on the tensor cores; same contract as the fp32 kernel.
int launch_tc_gemm(const GemmArgs& a, cudaStream_t s);
// fp32 SIMT kernel
int launch_sgemm(const GemmArgs& a, cudaStream_t s);

}  // namespace mava
