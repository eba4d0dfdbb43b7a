// Stable sort-by-key of (uint32 random key, int32 value) pairs: the sort inside
// jax.random.permutation (one round = sort_key_val(random_bits, x); ff_mappo.py:273,
// rec_mappo.py:350-352).  The keys are uniform 32-bit random numbers, so a two-level scheme is
// enough and needs no multi-pass radix machinery:
//   1. bucket by the top `kb` bits (histogram -> exclusive scan -> scatter; the order inside a
//      bucket does not matter at this point),
//   2. one CTA per bucket sorts its <= 1024 entries in shared memory on the 64-bit composite
//      (key << 32 | input position) -- unique, so any comparison sort is stable -- with a bitonic
//      network, and writes value[position] in sorted order.
// Buckets hold n / 2^kb = 256..512 entries on average; a bucket that does not fit (probability
// < 1e-30 for uniform keys) raises the overflow flag and the host API reports it.
#include "common.cuh"

namespace mava {
namespace {

constexpr int kSortCap = 1024;  // entries a bucket CTA can hold

__global__ void __launch_bounds__(256)
bucket_hist_kernel(const uint32_t* __restrict__ keys, int64_t n, int shift, uint32_t* __restrict__ count) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    atomicAdd(count + (keys[i] >> shift), 1u);
}

// exclusive scan of count[0..nb) -> offset[0..nb], cursor = offset (one CTA; nb <= 2^16)
__global__ void __launch_bounds__(1024)
bucket_scan_kernel(const uint32_t* __restrict__ count, int nb, uint32_t* __restrict__ offset,
                   uint32_t* __restrict__ cursor, int* __restrict__ overflow) {
  __shared__ uint32_t part[1024];
  const int per = (nb + 1023) / 1024;
  const int b0 = threadIdx.x * per;
  uint32_t s = 0;
  for (int b = b0; b < min(nb, b0 + per); ++b) {
    const uint32_t c = count[b];
    if (c > (uint32_t)kSortCap) *overflow = 1;
    s += c;
  }
  part[threadIdx.x] = s;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {  // Hillis-Steele inclusive scan
    const uint32_t v = threadIdx.x >= o ? part[threadIdx.x - o] : 0u;
    __syncthreads();
    part[threadIdx.x] += v;
    __syncthreads();
  }
  uint32_t run = part[threadIdx.x] - s;
  for (int b = b0; b < min(nb, b0 + per); ++b) {
    offset[b] = run;
    cursor[b] = run;
    run += count[b];
  }
  if (threadIdx.x == 1023) offset[nb] = part[1023];
}

__global__ void __launch_bounds__(256)
bucket_scatter_kernel(const uint32_t* __restrict__ keys, int64_t n, int shift,
                      uint32_t* __restrict__ cursor, unsigned long long* __restrict__ comp) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t k = keys[i];
    const uint32_t p = atomicAdd(cursor + (k >> shift), 1u);
    comp[p] = ((unsigned long long)k << 32) | (uint32_t)i;
  }
}

__global__ void __launch_bounds__(512)
bucket_sort_kernel(const unsigned long long* __restrict__ comp, const uint32_t* __restrict__ offset,
                   const int32_t* __restrict__ val_in, int32_t* __restrict__ val_out) {
  __shared__ unsigned long long s[kSortCap];
  const uint32_t lo = offset[blockIdx.x], hi = offset[blockIdx.x + 1];
  const int m = (int)min(hi - lo, (uint32_t)kSortCap);
  if (m == 0) return;
  int len = 2;
  while (len < m) len <<= 1;
  for (int i = threadIdx.x; i < len; i += blockDim.x) s[i] = i < m ? comp[lo + i] : ~0ull;
  __syncthreads();
  for (int k = 2; k <= len; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < len; i += blockDim.x) {
        const int l = i ^ j;
        if (l > i) {
          const unsigned long long a = s[i], b = s[l];
          const bool up = (i & k) == 0;
          if ((a > b) == up) {
            s[i] = b;
            s[l] = a;
          }
        }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < m; i += blockDim.x)
    val_out[lo + i] = val_in[(uint32_t)(s[i] & 0xffffffffull)];
}

int bucket_bits(int64_t n) {
  int kb = 1;
  while (((int64_t)384 << kb) < n) ++kb;  // ~ n / 2^kb in (192, 384]
  return kb < 4 ? 4 : kb;
}

}  // namespace
}  // namespace mava

using namespace mava;

extern "C" {

int64_t mava_sort_workspace_bytes(int64_t n) {
  if (n <= 0) return -1;
  const int64_t nb = (int64_t)1 << bucket_bits(n);
  return 8 * n + 4 * (3 * nb + 2) + 64;
}

int mava_sort_by_key(const uint32_t* keys, const int32_t* val_in, int32_t* val_out, int64_t n,
                     void* workspace, int* overflow_flag, mava_stream_t stream) {
  MAVA_CHECK_PTR(keys);
  MAVA_CHECK_PTR(val_in);
  MAVA_CHECK_PTR(val_out);
  MAVA_CHECK_PTR(workspace);
  MAVA_CHECK_PTR(overflow_flag);
  MAVA_CHECK_ARG(n > 0 && n < ((int64_t)1 << 31) && val_in != val_out);
  cudaStream_t s = as_stream(stream);
  const int kb = bucket_bits(n);
  if (kb > 16) return MAVA_E_UNSUPPORTED;
  const int nb = 1 << kb, shift = 32 - kb;
  unsigned long long* comp = static_cast<unsigned long long*>(workspace);
  uint32_t* count = reinterpret_cast<uint32_t*>(comp + n);
  uint32_t* offset = count + nb;
  uint32_t* cursor = offset + nb + 1;
  cudaError_t e = cudaMemsetAsync(count, 0, (size_t)nb * 4, s);
  if (e != cudaSuccess) return (int)e;
  const int blocks = (int)min((int64_t)sm_count() * 8, ceil_div64(n, 256));
  bucket_hist_kernel<<<blocks, 256, 0, s>>>(keys, n, shift, count);
  bucket_scan_kernel<<<1, 1024, 0, s>>>(count, nb, offset, cursor, overflow_flag);
  bucket_scatter_kernel<<<blocks, 256, 0, s>>>(keys, n, shift, cursor, comp);
  bucket_sort_kernel<<<nb, 512, 0, s>>>(comp, offset, val_in, val_out);
  return launch_status();
}

}  // extern "C"
