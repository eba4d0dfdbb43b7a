// Shared pieces of the tcgen05 MLP kernels: packed bf16 weight images, input-tile construction from
// the int8 observations, and the per-layer GEMM/epilogue building blocks.
//
// Thread mapping of every kernel built from these pieces: one CTA = 512 threads = 16 warps works on
// one 128-row tile.  Warp w reads TMEM lanes 32*(w%4) .. +31 (the hardware restriction) and owns the
// column quarter w/4, so thread (row r = 32*(w%4)+lane, quarter q = w/4) handles 32 accumulator
// columns of tile row r in each epilogue.  Four warps per scheduler keep the epilogues from being
// issue-latency bound while a single elected thread feeds the tensor core.
#pragma once
#include "common.cuh"
#include "prng.cuh"
#include "tc.cuh"

namespace mava {
namespace tcmlp {

using namespace tc;

constexpr int TM = 128;          // rows per tile
constexpr int NT = 512;          // threads per CTA
constexpr int NWARPS = NT / 32;
constexpr int HID = 128;         // hidden width the tensor-core path supports
constexpr int HCOLS = HID + 16;  // activation tiles carry a ones column (bias gradients for free)
constexpr int NHEAD = 16;        // head width padded to one MMA N step
constexpr float kF32Min = -3.402823466e38f;

__host__ __device__ constexpr int pad16(int x) { return (x + 15) / 16 * 16; }

// Byte layout of one network's packed weight image (bf16 core-matrix tiles, see tc.cuh):
//   W1 [K1p][HID] | W2 [HCOLS][HID] | W3 [HCOLS][NHEAD]
// Each matrix carries its bias as one extra input row (row in_dim of W1, row HID of W2 and W3):
// the activation tiles have a ones column there, so the bias add is part of the GEMM.
struct WImage {
  int k1p;
  __host__ __device__ uint32_t w1_bytes() const { return (uint32_t)k1p * HID * 2; }
  __host__ __device__ uint32_t w2_bytes() const { return HCOLS * HID * 2; }
  __host__ __device__ uint32_t w3_bytes() const { return HCOLS * NHEAD * 2; }
  __host__ __device__ uint32_t total() const { return w1_bytes() + w2_bytes() + w3_bytes(); }
};

// tiles over a weight image placed at `base` in shared memory
__device__ __forceinline__ Tile w1_tile(uint32_t base, int k1p) {
  return Tile{base, 128u, (uint32_t)(k1p / 8) * 128u};
}
__device__ __forceinline__ Tile w2_tile(uint32_t base, int k1p) {
  return Tile{base + (uint32_t)k1p * HID * 2, 128u, (uint32_t)(HCOLS / 8) * 128u};
}
__device__ __forceinline__ Tile w3_tile(uint32_t base, int k1p) {
  return Tile{base + (uint32_t)k1p * HID * 2 + HCOLS * HID * 2, 128u, (uint32_t)(HCOLS / 8) * 128u};
}

// Byte offset, in a network's packed image, of the bf16 copy of flat parameter i (flax order
// W1 (in,H) | b1 | W2 (H,H) | b2 | W3 (H,out) | b3; biases are the extra input row of their matrix;
// tiles are grids of 8x8 core matrices, see tc.cuh / pack_kernel in mlp_tc.cu).
__device__ __forceinline__ uint32_t image_offset(int i, int in_dim, int k1p, int out) {
  int r, c, rows;
  uint32_t base;
  const int n_w1 = in_dim * HID;
  if (i < n_w1 + HID) {
    rows = k1p;
    base = 0;
    if (i < n_w1) { r = i / HID; c = i % HID; }
    else { r = in_dim; c = i - n_w1; }
  } else {
    const int i2 = i - n_w1 - HID;
    rows = HCOLS;
    if (i2 < HID * HID + HID) {
      base = (uint32_t)k1p * HID * 2;
      if (i2 < HID * HID) { r = i2 / HID; c = i2 % HID; }
      else { r = HID; c = i2 - HID * HID; }
    } else {
      const int i3 = i2 - HID * HID - HID;
      base = (uint32_t)k1p * HID * 2 + (uint32_t)HCOLS * HID * 2;
      if (i3 < HID * out) { r = i3 / out; c = i3 % out; }
      else { r = HID; c = i3 - HID * out; }
    }
  }
  return base + (uint32_t)(r >> 3) * 128 + (uint32_t)(c >> 3) * (rows / 8) * 128 + (r & 7) * 16 + (c & 7) * 2;
}

struct NetDesc {
  int mode, add_id, A, FR, in_dim, k1p, out;  // k1p = pad16(in_dim + 1): room for the ones column
};

// Position of the calling thread inside the tile.
struct Lane {
  int t, warp, lane;
  int r;  // tile row (= TMEM lane)
  int q;  // column quarter
  __device__ __forceinline__ Lane() {
    t = threadIdx.x;
    warp = t >> 5;
    lane = t & 31;
    r = (warp & 3) * 32 + lane;
    q = warp >> 2;
  }
  __device__ __forceinline__ uint32_t tmem_lane() const { return (uint32_t)((warp & 3) * 32) << 16; }
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

// Build the bf16 input tile X[TM][k1p] of one CTA from the int8 observations.
//   step_at(j) -> env-step index s of minibatch position j (rows of position j are its A agents
//   for AGENT_VIEW, or the single joint row for GLOBAL)
// AGENT_VIEW rows: x = [onehot(agent) | view[s][agent][:] | 1 | 0...]
// GLOBAL rows:     x = [view[s][0..A)[:]                  | 1 | 0...]
// The trailing 1 at column in_dim makes the bias gradient fall out of the weight-gradient GEMM
// (the packed W1 image has a zero row there, so the forward pass is unaffected).
//
// Two phases: (1) the env-steps the tile touches are gathered from HBM into the `stage` scratch
// area (shared memory, >= stage_bytes()) with asynchronous copies, all in flight at once - one
// env-step (A*FR contiguous bytes) per warp iteration; (2) thread (r, q) expands every fourth
// 8-column chunk of row r to bf16.  Contains two __syncthreads().
__host__ __device__ inline uint32_t stage_bytes(int A, int FR, int rows_per_step) {
  return ((uint32_t)(TM / rows_per_step + 2) * A * FR + 15) / 16 * 16 + TM * 4 + 16;
}

// Thread (r, q) expands every fourth 8-column chunk of tile row r from the int8 observation bytes
// at `mine` (shared memory) to bf16: [onehot(agent a) if add_id | view bytes | 1 | 0...].
__device__ __forceinline__ void expand_x_row(const NetDesc& d, const Tile& xt, const Lane& L,
                                             const signed char* mine, bool valid, int a,
                                             int cg0 = -1, int cg_step = 4) {
  const int id_cols = (d.mode == MAVA_IN_AGENT_VIEW && d.add_id) ? d.A : 0;
  const uint32_t mine_s = smem_u32(mine);
  for (int cg = cg0 < 0 ? L.q : cg0; cg < d.k1p / 8; cg += cg_step) {
    const int k0 = cg * 8;
    const uint32_t dst = xt.base + chunk_off(xt, L.r, cg);
    if (!valid || k0 > d.in_dim) {  // rows past the end of the minibatch, padding columns
      st_shared_v4(dst, 0u, 0u, 0u, 0u);
      continue;
    }
    float v[8];
    if (k0 >= id_cols && k0 + 8 <= d.in_dim) {
      // interior chunk: eight observation bytes as four 16-bit loads (rows are 2-byte aligned:
      // A * FR and FR are even for every supported env; odd FR falls back to byte loads below)
      const uint32_t src = mine_s + (uint32_t)(k0 - id_cols);
      if ((src & 1u) == 0u) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t h;
          asm volatile("ld.shared.u16 %0, [%1];" : "=r"(h) : "r"(src + 2u * j));
          v[2 * j] = (float)(int)(signed char)(h & 0xffu);
          v[2 * j + 1] = (float)(int)(signed char)(h >> 8);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          int b;
          asm volatile("ld.shared.s8 %0, [%1];" : "=r"(b) : "r"(src + (uint32_t)j));
          v[j] = (float)b;
        }
      }
    } else {
      // boundary chunk (agent-id columns in front, the ones column behind): clamped loads and
      // selects instead of per-element branches
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int k = k0 + j;
        const int idx = min(max(k - id_cols, 0), d.FR * (d.mode == MAVA_IN_GLOBAL ? d.A : 1) - 1);
        float x = (float)mine[idx];
        x = k < id_cols ? (k == a ? 1.0f : 0.0f) : x;
        x = k >= d.in_dim ? (k == d.in_dim ? 1.0f : 0.0f) : x;
        v[j] = x;
      }
    }
    st_shared_v4(dst, pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]),
                 pack_bf16(v[6], v[7]));
  }
}

// BAR = 0: called by the whole CTA (__syncthreads); else the named barrier the NT calling threads use
template <int BAR>
__device__ __forceinline__ void tile_sync() {
  if constexpr (BAR == 0) __syncthreads();
  else asm volatile("bar.sync %0, %1;" ::"n"(BAR), "n"(NT) : "memory");
}

template <int BAR = 0, class StepFn>
__device__ __forceinline__ void build_x_tile(const NetDesc& d, const int8_t* __restrict__ view,
                                             const Tile& xt, unsigned char* stage, int row0,
                                             int M, StepFn step_at) {
  const Lane L;
  const int rps = d.mode == MAVA_IN_GLOBAL ? 1 : d.A;
  const int step_bytes = d.A * d.FR;
  const int last = (row0 + TM - 1 < M ? row0 + TM - 1 : M - 1);
  const int j0 = row0 / rps;
  const int nsteps = last / rps - j0 + 1;
  // the env-step indices first (one coalesced load), so that the copies below are independent
  int* steps = reinterpret_cast<int*>(stage + ((size_t)(TM / rps + 2) * step_bytes + 15) / 16 * 16);
  if (L.t < nsteps) steps[L.t] = (int)step_at(j0 + L.t);
  tile_sync<BAR>();
  if ((step_bytes & 3) == 0) {
    const int unit = (step_bytes & 7) == 0 ? 8 : 4;
    const int units = step_bytes / unit;
    for (int js = L.warp; js < nsteps; js += NWARPS) {
      const int8_t* src = view + (size_t)steps[js] * step_bytes;
      const uint32_t dst = smem_u32(stage) + (uint32_t)js * step_bytes;
      for (int i = L.lane; i < units; i += 32) {
        if (unit == 8)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + i * 8), "l"(src + i * 8)
                       : "memory");
        else
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + i * 4), "l"(src + i * 4)
                       : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  } else {
    for (int js = L.warp; js < nsteps; js += NWARPS) {
      const int8_t* src = view + (size_t)steps[js] * step_bytes;
      unsigned char* dst = stage + (size_t)js * step_bytes;
      if ((step_bytes & 1) == 0) {
        for (int i = L.lane; i < (step_bytes >> 1); i += 32)
          reinterpret_cast<uint16_t*>(dst)[i] = __ldg(reinterpret_cast<const uint16_t*>(src) + i);
      } else {
        for (int i = L.lane; i < step_bytes; i += 32) dst[i] = (unsigned char)__ldg(src + i);
      }
    }
  }
  tile_sync<BAR>();
  const int row = row0 + L.r;
  const bool valid = row < M;
  const int a = d.mode == MAVA_IN_GLOBAL ? 0 : row % rps;
  const signed char* mine = reinterpret_cast<const signed char*>(stage) +
                            (valid ? (size_t)(row / rps - j0) * step_bytes + (size_t)a * d.FR : 0);
  expand_x_row(d, xt, L, mine, valid, a);
}

// Split form of build_x_tile for kernels that prefetch the next tile's observations while the
// current tile is in the tensor pipe: gather_issue() starts the asynchronous copies of the
// env-steps a tile touches into `stage` (no CTA barrier inside: every lane of a warp reads the
// warp's env-step index itself), gather_expand() is called after cp.async.wait_group 0 + a CTA
// barrier and builds the bf16 X tile.  Requires A * FR to be a multiple of 4 bytes.
template <class StepFn>
__device__ __forceinline__ void gather_issue(const NetDesc& d, const int8_t* __restrict__ view,
                                             unsigned char* stage, int row0, int M,
                                             StepFn step_at, int first_warp = 0,
                                             int num_warps = NWARPS) {
  const Lane L;
  if (L.warp < first_warp || L.warp >= first_warp + num_warps) return;
  const int rps = d.mode == MAVA_IN_GLOBAL ? 1 : d.A;
  const int step_bytes = d.A * d.FR;
  const int last = (row0 + TM - 1 < M ? row0 + TM - 1 : M - 1);
  const int j0 = row0 / rps;
  const int nsteps = last / rps - j0 + 1;
  const int unit = (step_bytes & 7) == 0 ? 8 : 4;
  const int units = step_bytes / unit;
  for (int js = L.warp - first_warp; js < nsteps; js += num_warps) {
    const int8_t* src = view + (size_t)step_at(j0 + js) * step_bytes;
    const uint32_t dst = smem_u32(stage) + (uint32_t)js * step_bytes;
    for (int i = L.lane; i < units; i += 32) {
      if (unit == 8)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + i * 8), "l"(src + i * 8)
                     : "memory");
      else
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + i * 4), "l"(src + i * 4)
                     : "memory");
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

__device__ __forceinline__ void gather_wait() {
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

__device__ __forceinline__ void gather_expand(const NetDesc& d, const Tile& xt,
                                              const unsigned char* stage, int row0, int M,
                                              int cg0 = -1, int cg_step = 4) {
  const Lane L;
  const int rps = d.mode == MAVA_IN_GLOBAL ? 1 : d.A;
  const int step_bytes = d.A * d.FR;
  const int j0 = row0 / rps;
  const int row = row0 + L.r;
  const bool valid = row < M;
  const int a = d.mode == MAVA_IN_GLOBAL ? 0 : row % rps;
  const signed char* mine = reinterpret_cast<const signed char*>(stage) +
                            (valid ? (size_t)(row / rps - j0) * step_bytes + (size_t)a * d.FR : 0);
  expand_x_row(d, xt, L, mine, valid, a, cg0, cg_step);
}

// ---- padded int8 staging rows ---------------------------------------------------------------------
// The int8 image of an X row, k1p bytes: [onehot(agent) | observation bytes | 1 | 0...].  With the
// rows staged like this the bf16 expansion is the same for every 8-column chunk: one 8-byte load,
// eight conversions, one 16-byte store.

// two int8 (bytes 0 and 1 of h) -> packed bf16 pair, exactly, without the conversion pipe:
// v = u7 - 128 s (u7 = low seven bits, s = sign bit); 0x4300 | u7 is the bf16 128 + u7 and
// 0x4300 | s << 7 the bf16 128 + 128 s; their difference v has at most 8 significant bits.
__device__ __forceinline__ uint32_t s8x2_bf16x2(uint32_t h) {
  const uint32_t x = __byte_perm(h, 0u, 0x4140);  // [b0, 0, b1, 0]
  uint32_t m = (x & 0x007f007fu) | 0x43004300u, c = (x & 0x00800080u) | 0x43004300u;
  __nv_bfloat162 r = __hsub2(*reinterpret_cast<__nv_bfloat162*>(&m), *reinterpret_cast<__nv_bfloat162*>(&c));
  return *reinterpret_cast<uint32_t*>(&r);
}

// ... and for bytes known to be in [0, 127] (RobotWarehouse observations: coordinates < 128, flags):
// 0x4300 | b is the bf16 128 + b, minus 128.
__device__ __forceinline__ uint32_t u7x2_bf16x2(uint32_t h) {
  uint32_t m = __byte_perm(h, 0x43004300u, 0x7150);  // [b0, 0x43, b1, 0x43]
  const uint32_t c = 0x43004300u;
  __nv_bfloat162 r = __hsub2(*reinterpret_cast<__nv_bfloat162*>(&m),
                             *reinterpret_cast<const __nv_bfloat162*>(&c));
  return *reinterpret_cast<uint32_t*>(&r);
}

// Thread (row L.r) expands chunks cg0, cg0 + cg_step, ... of its padded staging row into the bf16
// tile; rows that are not valid become zero rows.  NS chunks' loads are in flight at a time; NONNEG:
// every byte is in [0, 127].
template <int NS = 4, bool NONNEG = false>
__device__ __forceinline__ void expand_padded_row(const Tile& xt, const Lane& L, uint32_t stage_row,
                                                  bool valid, int nchunks, int cg0, int cg_step) {
  auto cv = [](uint32_t h) { return NONNEG ? u7x2_bf16x2(h) : s8x2_bf16x2(h); };
  for (int cgb = cg0; cgb < nchunks; cgb += NS * cg_step) {
    uint32_t w[NS][2];
#pragma unroll
    for (int i = 0; i < NS; ++i) {
      const int cg = cgb + i * cg_step;
      w[i][0] = w[i][1] = 0u;
      if (valid && cg < nchunks)
        asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];"
                     : "=r"(w[i][0]), "=r"(w[i][1])
                     : "r"(stage_row + 8u * cg));
    }
#pragma unroll
    for (int i = 0; i < NS; ++i) {
      const int cg = cgb + i * cg_step;
      if (cg < nchunks)
        st_shared_v4(xt.base + chunk_off(xt, L.r, cg), cv(w[i][0]), cv(w[i][0] >> 16), cv(w[i][1]),
                     cv(w[i][1] >> 16));
    }
  }
}

// Joint-observation rows (centralised critic: one env-step = one tile row of in_dim = A * FR bytes, a
// multiple of 8) are staged by bulk (TMA) copies, one per row, issued by the row's thread: no
// registers, no LSU queue, completion on an mbarrier initialised with TM arrivals -- so the gather of
// the NEXT tile can be in flight while this tile computes.  A bulk copy needs 16-byte aligned
// addresses and sizes and the rows are only 8-byte aligned: the copy takes the enclosing aligned
// range, which puts the row at offset `skew` (0 or 8) of its staging slot and drags up to 16
// foreign bytes along (they stay inside the rollout buffer: a minibatch never indexes its last time
// slot).  Staging slot = k1p + 16 bytes; after the copy has landed the row's thread writes the
// [1 | 0...] tail behind the observation bytes.
__host__ __device__ inline uint32_t grow_stride(int k1p) { return (uint32_t)k1p + 16u; }
__device__ __forceinline__ uint32_t grow_skew(const NetDesc& d, int step) {
  return (uint32_t)(((size_t)step * (size_t)d.in_dim) & 15u);
}
// thread t < TM: start the copy of row t (step index `step`, ignored when !active) and arrive
__device__ __forceinline__ void grow_issue(const NetDesc& d, const int8_t* __restrict__ view, int step,
                                           bool active, unsigned char* stg, int row, uint64_t* bar) {
  if (active) {
    const uint32_t skew = grow_skew(d, step);
    const uint32_t bytes = (skew + (uint32_t)d.in_dim + 15u) & ~15u;
    mbar_expect_tx(bar, bytes);
    bulk_g2s(smem_u32(stg) + (uint32_t)row * grow_stride(d.k1p),
             view + (size_t)step * (size_t)d.in_dim - skew, bytes, bar);
  } else {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
  }
}
// after the copies have landed: thread t < TM completes row t
__device__ __forceinline__ void grow_tail(const NetDesc& d, unsigned char* stg, int row, int step) {
  unsigned char* tail = stg + (size_t)row * grow_stride(d.k1p) + grow_skew(d, step) + d.in_dim;
  for (int k = 0; k < d.k1p - d.in_dim; ++k) tail[k] = k == 0 ? 1 : 0;
}

// One thread issues the K/16 MMAs of a GEMM (M = 128) and optionally commits to `bar`.
__device__ __forceinline__ void issue_gemm(uint32_t d_tmem, const Tile& a, bool a_mn, const Tile& b,
                                           bool b_mn, int N, int K, bool accumulate,
                                           uint64_t* bar) {
  // The K=16 steps of an operand differ only in the 14-bit start-address field of its descriptor:
  // build the descriptors once and step them by a constant.  The issuing thread is on the critical
  // path of every phase of every tile (one thread, dependent instruction stream), so the per-MMA
  // issue cost matters as much as the MMA itself.
  const uint32_t idesc = instr_desc(TM, N, a_mn, b_mn);
  uint64_t ad = a_mn ? desc_mnmajor(a, 0) : desc_kmajor(a, 0);
  uint64_t bd = b_mn ? desc_mnmajor(b, 0) : desc_kmajor(b, 0);
  const uint64_t a_inc = (uint64_t)(((a_mn ? a.s_r : a.s_c) * 2u) >> 4);
  const uint64_t b_inc = (uint64_t)(((b_mn ? b.s_r : b.s_c) * 2u) >> 4);
  const int steps = K / 16;
  for (int k = 0; k < steps; ++k) {
    mma(d_tmem, ad, bd, idesc, accumulate || k > 0);
    ad += a_inc;
    bd += b_inc;
  }
  if (bar) commit(bar);
}

__device__ __forceinline__ uint32_t relu_bf16x2(uint32_t x) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&x);
  v = __hmax2(v, __floats2bfloat162_rn(0.0f, 0.0f));
  return *reinterpret_cast<uint32_t*>(&v);
}

// x * (h > 0) on packed bf16 pairs (the relu derivative read back from the activation tile)
__device__ __forceinline__ uint32_t relu_grad_bf16x2(uint32_t x, uint32_t h) {
  const __nv_bfloat162 xv = *reinterpret_cast<__nv_bfloat162*>(&x);
  const __nv_bfloat162 hv = *reinterpret_cast<__nv_bfloat162*>(&h);
  __nv_bfloat162 o = __hmul2(xv, __hgt2(hv, __floats2bfloat162_rn(0.0f, 0.0f)));
  return *reinterpret_cast<uint32_t*>(&o);
}

__device__ __forceinline__ void ld_shared_v4(uint32_t addr, uint32_t (&w)[4]) {
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3])
               : "r"(addr));
}

// Hidden-layer epilogue for thread (r, q): 32 accumulator columns (bias already added by the GEMM
// through the ones column) -> relu -> bf16 into the activation tile; quarter 0 also writes the ones
// column at HID that feeds the next layer's bias row and the bias gradient.
__device__ __forceinline__ void hidden_epilogue(const Lane& L, uint32_t tmem_acc, const Tile& ht) {
  float v[32];
  ld32(tmem_acc + L.tmem_lane() + (uint32_t)(L.q * 32), v);
#pragma unroll
  for (int cg = 0; cg < 4; ++cg)
    st_shared_v4(ht.base + chunk_off(ht, L.r, L.q * 4 + cg),
                 relu_bf16x2(pack_bf16(v[cg * 8], v[cg * 8 + 1])),
                 relu_bf16x2(pack_bf16(v[cg * 8 + 2], v[cg * 8 + 3])),
                 relu_bf16x2(pack_bf16(v[cg * 8 + 4], v[cg * 8 + 5])),
                 relu_bf16x2(pack_bf16(v[cg * 8 + 6], v[cg * 8 + 7])));
  if (L.q == 0) {
    st_shared_v4(ht.base + chunk_off(ht, L.r, HID / 8), 0x00003F80u, 0u, 0u, 0u);
    st_shared_v4(ht.base + chunk_off(ht, L.r, HID / 8 + 1), 0u, 0u, 0u, 0u);
  }
}

// Only warp 0 polls the mbarrier; everybody else parks at the CTA barrier.
__device__ __forceinline__ void wait_mma(uint64_t* bar, uint32_t parity) {
  if (threadIdx.x < 32) mbar_wait(bar, parity);
  __syncthreads();
  fence_after_sync();
}

}  // namespace tcmlp
}  // namespace mava
