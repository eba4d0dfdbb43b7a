// Threefry2x32 counter PRNG with the key/counter conventions of the reference's jax.random
// (default threefry2x32 impl, jax==0.4.30, non-partitionable): split, random_bits, uniform, gumbel.
// Reference call sites: mava/systems/ppo/ff_mappo.py:81,204,269,273,392,417;
// mava/wrappers/auto_reset_wrapper.py:74; mava/wrappers/episode_metrics.py:61.
#pragma once
#include <cstdint>

namespace mava {

struct Key {
  uint32_t k0, k1;
};

__host__ __device__ __forceinline__ uint32_t rotl32(uint32_t x, int r) {
#ifdef __CUDA_ARCH__
  return __funnelshift_l(x, x, r);
#else
  return (x << r) | (x >> (32 - r));
#endif
}

// 20-round Threefry-2x32.
__host__ __device__ __forceinline__ void threefry2x32(Key k, uint32_t x0, uint32_t x1,
                                                      uint32_t& o0, uint32_t& o1) {
  const uint32_t ks0 = k.k0, ks1 = k.k1, ks2 = k.k0 ^ k.k1 ^ 0x1BD11BDAu;
  x0 += ks0;
  x1 += ks1;
#define MAVA_TF_R(r) { x0 += x1; x1 = rotl32(x1, r); x1 ^= x0; }
  MAVA_TF_R(13) MAVA_TF_R(15) MAVA_TF_R(26) MAVA_TF_R(6)
  x0 += ks1; x1 += ks2 + 1u;
  MAVA_TF_R(17) MAVA_TF_R(29) MAVA_TF_R(16) MAVA_TF_R(24)
  x0 += ks2; x1 += ks0 + 2u;
  MAVA_TF_R(13) MAVA_TF_R(15) MAVA_TF_R(26) MAVA_TF_R(6)
  x0 += ks0; x1 += ks1 + 3u;
  MAVA_TF_R(17) MAVA_TF_R(29) MAVA_TF_R(16) MAVA_TF_R(24)
  x0 += ks1; x1 += ks2 + 4u;
  MAVA_TF_R(13) MAVA_TF_R(15) MAVA_TF_R(26) MAVA_TF_R(6)
  x0 += ks2; x1 += ks0 + 5u;
#undef MAVA_TF_R
  o0 = x0;
  o1 = x1;
}

// jax.random.split(key) -> (first, second): counters iota(4) paired (0,2),(1,3).
__host__ __device__ __forceinline__ void split2(Key k, Key& first, Key& second) {
  uint32_t a0, a1, b0, b1;
  threefry2x32(k, 0u, 2u, a0, a1);
  threefry2x32(k, 1u, 3u, b0, b1);
  first = Key{a0, b0};
  second = Key{a1, b1};
}

// jax.random.split(key, num)[j]: counters iota(2*num), halves paired.
__host__ __device__ __forceinline__ Key split_n(Key k, uint32_t num, uint32_t j) {
  // flat output index 2j and 2j+1 of concat(y0[0..num), y1[0..num))
  uint32_t out[2];
  for (int w = 0; w < 2; ++w) {
    uint32_t f = 2u * j + (uint32_t)w;
    uint32_t p = f < num ? f : f - num;
    uint32_t y0, y1;
    threefry2x32(k, p, p + num, y0, y1);
    out[w] = f < num ? y0 : y1;
  }
  return Key{out[0], out[1]};
}

// Element i of jax's 32-bit random_bits(key, shape) with prod(shape) == size (flat index i).
__host__ __device__ __forceinline__ uint32_t random_bits_at(Key k, uint32_t i, uint32_t size) {
  const uint32_t half = (size + 1u) >> 1;
  const uint32_t p = i < half ? i : i - half;
  const uint32_t c1 = (p + half < size) ? p + half : 0u;  // odd sizes are padded with a zero counter
  uint32_t y0, y1;
  threefry2x32(k, p, c1, y0, y1);
  return i < half ? y0 : y1;
}

// Both outputs of pair p (elements p and p+half); caller checks p+half < size.
__host__ __device__ __forceinline__ void random_bits_pair(Key k, uint32_t p, uint32_t size,
                                                          uint32_t& lo, uint32_t& hi) {
  const uint32_t half = (size + 1u) >> 1;
  const uint32_t c1 = (p + half < size) ? p + half : 0u;
  threefry2x32(k, p, c1, lo, hi);
}

// How a device function that draws random bits wants the Threefry block emitted: inline
// (throughput kernels) or as ONE out-of-line copy per kernel (latency-bound kernels, where the
// instruction footprint of ~110 instructions per block, several blocks per draw, costs more in
// instruction-cache misses than a call).
#ifdef __CUDACC__
struct PrngInline {
  static __device__ __forceinline__ uint2 block(uint32_t k0, uint32_t k1, uint32_t x0, uint32_t x1) {
    uint32_t o0, o1;
    threefry2x32(Key{k0, k1}, x0, x1, o0, o1);
    return make_uint2(o0, o1);
  }
  static __device__ __forceinline__ void split(Key k, Key& first, Key& second) {
    split2(k, first, second);
  }
  static __device__ __forceinline__ uint32_t bits_at(Key k, uint32_t i, uint32_t size) {
    return random_bits_at(k, i, size);
  }
  static __device__ __forceinline__ void bits_pair(Key k, uint32_t p, uint32_t size, uint32_t& lo,
                                                   uint32_t& hi) {
    random_bits_pair(k, p, size, lo, hi);
  }
};
struct PrngCall {
  static __device__ __noinline__ uint2 block(uint32_t k0, uint32_t k1, uint32_t x0, uint32_t x1) {
    uint32_t o0, o1;
    threefry2x32(Key{k0, k1}, x0, x1, o0, o1);
    return make_uint2(o0, o1);
  }
  static __device__ __forceinline__ void split(Key k, Key& first, Key& second) {
    const uint2 a = block(k.k0, k.k1, 0u, 2u), b = block(k.k0, k.k1, 1u, 3u);
    first = Key{a.x, b.x};
    second = Key{a.y, b.y};
  }
  static __device__ __forceinline__ void bits_pair(Key k, uint32_t p, uint32_t size, uint32_t& lo,
                                                   uint32_t& hi) {
    const uint32_t half = (size + 1u) >> 1;
    const uint2 y = block(k.k0, k.k1, p, (p + half < size) ? p + half : 0u);
    lo = y.x;
    hi = y.y;
  }
  static __device__ __forceinline__ uint32_t bits_at(Key k, uint32_t i, uint32_t size) {
    const uint32_t half = (size + 1u) >> 1;
    uint32_t lo, hi;
    bits_pair(k, i < half ? i : i - half, size, lo, hi);
    return i < half ? lo : hi;
  }
};
#endif

// jax.random.uniform(minval=tiny, maxval=1) then gumbel = -log(-log(u)).
__device__ __forceinline__ float bits_to_gumbel(uint32_t bits) {
  const float tiny = 1.17549435e-38f;
  float u = __uint_as_float((bits >> 9) | 0x3F800000u) - 1.0f;
  u = fmaxf(tiny, u * (1.0f - tiny) + tiny);
  return -logf(-logf(u));
}

}  // namespace mava
