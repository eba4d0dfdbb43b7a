// Shared pieces of the tcgen05 MLP kernels: packed bf16 weight images, input-tile construction from
// the int8 observations, and the per-layer GEMM/epilogue building blocks.
#pragma once
#include "common.cuh"
#include "prng.cuh"
#include "tc.cuh"

namespace mava {
namespace tcmlp {

using namespace tc;

constexpr int TM = 128;        // rows per tile (= threads per CTA, thread t owns tile row t)
constexpr int HID = 128;       // hidden width the tensor-core path supports
constexpr int HCOLS = HID + 16;  // activation tiles carry a ones column (bias gradients for free)
constexpr int NHEAD = 16;      // head width padded to one MMA N step
constexpr float kF32Min = -3.402823466e38f;

__host__ __device__ constexpr int pad16(int x) { return (x + 15) / 16 * 16; }

// Byte layout of one network's packed weight image (bf16 core-matrix tiles, see tc.cuh):
//   W1 [K1p][HID] | W2 [HID][HID] | W3 [HID][NHEAD]
struct WImage {
  int k1p;
  __host__ __device__ uint32_t w1_bytes() const { return (uint32_t)k1p * HID * 2; }
  __host__ __device__ uint32_t w2_bytes() const { return HID * HID * 2; }
  __host__ __device__ uint32_t w3_bytes() const { return HID * NHEAD * 2; }
  __host__ __device__ uint32_t total() const { return w1_bytes() + w2_bytes() + w3_bytes(); }
};

// tiles over a weight image placed at `base` in shared memory
__device__ __forceinline__ Tile w1_tile(uint32_t base, int k1p) {
  return Tile{base, 128u, (uint32_t)(k1p / 8) * 128u};
}
__device__ __forceinline__ Tile w2_tile(uint32_t base, int k1p) {
  return Tile{base + (uint32_t)k1p * HID * 2, 128u, (uint32_t)(HID / 8) * 128u};
}
__device__ __forceinline__ Tile w3_tile(uint32_t base, int k1p) {
  return Tile{base + (uint32_t)k1p * HID * 2 + HID * HID * 2, 128u, (uint32_t)(HID / 8) * 128u};
}

struct NetDesc {
  int mode, add_id, A, FR, in_dim, k1p, out;  // k1p = pad16(in_dim + 1): room for the ones column
  const float *b1, *b2, *b3;                  // fp32 biases inside the flat parameter vector
};

// bulk (TMA) copy global -> shared, completion on an mbarrier (complete_tx bytes)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(dst),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}

// One thread streams a packed weight image into shared memory in <= 32 KB bulk copies.
__device__ __forceinline__ void load_weights(uint32_t dst, const unsigned char* image,
                                             uint32_t bytes, uint64_t* bar) {
  mbar_expect_tx(bar, bytes);
  for (uint32_t off = 0; off < bytes; off += 32768u) {
    const uint32_t n = bytes - off < 32768u ? bytes - off : 32768u;
    bulk_g2s(dst + off, image + off, n, bar);
  }
}

// Build the bf16 input tile X[TM][k1p] of one CTA from the int8 observations; thread t fills row t.
//   step_of_row(r) -> env-step index s of tile row r
// AGENT_VIEW rows are (s, agent = row % A): x = [onehot(agent) | view[s][agent][:] | 1 | 0...]
// GLOBAL rows are env-steps:                x = [view[s][0..A)[:]             | 1 | 0...]
// The trailing 1 at column in_dim makes the bias gradient fall out of the weight-gradient GEMM
// (the packed W1 image has a zero row there, so the forward pass is unaffected).
template <class RowFn>
__device__ __forceinline__ void build_x_tile(const NetDesc& d, const int8_t* __restrict__ view,
                                             const Tile& xt, int64_t row0, int64_t M,
                                             RowFn step_of_row) {
  const int t = threadIdx.x;
  const int64_t row = row0 + t;
  const bool valid = row < M;
  const int a = d.mode == MAVA_IN_GLOBAL ? 0 : (int)(row % d.A);
  const int id_cols = (d.mode == MAVA_IN_AGENT_VIEW && d.add_id) ? d.A : 0;
  const int8_t* src = view;
  if (valid) src = view + ((size_t)step_of_row(row) * d.A + a) * d.FR;
  for (int cg = 0; cg < d.k1p / 8; ++cg) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = cg * 8 + j;
      float x = 0.0f;
      if (valid) {
        if (k < id_cols) x = k == a ? 1.0f : 0.0f;
        else if (k < d.in_dim) x = (float)__ldg(src + (k - id_cols));
        else if (k == d.in_dim) x = 1.0f;
      }
      v[j] = x;
    }
    st_shared_v4(xt.base + chunk_off(xt, t, cg), pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]),
                 pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
  }
}

// One thread issues the K/16 MMAs of a GEMM whose A is K-major (rows x K) and commits to `bar`.
__device__ __forceinline__ void issue_gemm(uint32_t d_tmem, const Tile& a, bool a_mn, const Tile& b,
                                           bool b_mn, int N, int K, bool accumulate,
                                           uint64_t* bar) {
  const uint32_t idesc = instr_desc(TM, N, a_mn, b_mn);
  for (int k = 0; k < K / 16; ++k) {
    const uint64_t ad = a_mn ? desc_mnmajor(a, k) : desc_kmajor(a, k);
    const uint64_t bd = b_mn ? desc_mnmajor(b, k) : desc_kmajor(b, k);
    mma(d_tmem, ad, bd, idesc, accumulate || k > 0);
  }
  if (bar) commit(bar);
}

// Hidden-layer epilogue: TMEM accumulator row -> (+bias, relu) -> bf16 activation tile row, with a
// ones column at HID.  Returns the relu mask of the row (bit c set = unit c active).
struct RowMask {
  uint32_t w[HID / 32];
};

__device__ __forceinline__ RowMask hidden_epilogue(uint32_t tmem_acc, const float* __restrict__ bias,
                                                   const Tile& ht) {
  const int t = threadIdx.x, warp = t >> 5;
  RowMask m;
#pragma unroll
  for (int q = 0; q < HID / 32; ++q) {
    float v[32];
    ld32(tmem_acc + ((uint32_t)(warp * 32) << 16) + (uint32_t)(q * 32), v);
    uint32_t bits = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      v[j] = fmaxf(v[j] + __ldg(bias + q * 32 + j), 0.0f);
      bits |= (v[j] > 0.0f ? 1u : 0u) << j;
    }
    m.w[q] = bits;
#pragma unroll
    for (int cg = 0; cg < 4; ++cg)
      st_shared_v4(ht.base + chunk_off(ht, t, q * 4 + cg), pack_bf16(v[cg * 8], v[cg * 8 + 1]),
                   pack_bf16(v[cg * 8 + 2], v[cg * 8 + 3]), pack_bf16(v[cg * 8 + 4], v[cg * 8 + 5]),
                   pack_bf16(v[cg * 8 + 6], v[cg * 8 + 7]));
  }
  // ones column (bias gradient) + zero padding
  st_shared_v4(ht.base + chunk_off(ht, t, HID / 8), 0x00003F80u, 0u, 0u, 0u);
  st_shared_v4(ht.base + chunk_off(ht, t, HID / 8 + 1), 0u, 0u, 0u, 0u);
  return m;
}

}  // namespace tcmlp
}  // namespace mava
