// Dense contraction on the tcgen05 tensor cores for the recurrent path: fp32 matrices in HBM, bf16
// operands staged in shared memory, fp32 accumulation in TMEM.
//
//   C[M][N] (op)= alpha * opA(A)[M][K] * opB(B)[K][N] (+ bias) (relu) (zeroed where relu_ref <= 0)
//
// One CTA (256 threads) computes a 128 x BN tile (BN <= 128) over a K range.  Per 64-wide K chunk
// the threads load the fp32 operands (16-byte loads when the leading dimensions allow it), convert
// to bf16 and store 16-byte core-matrix rows (tc.cuh operand format); whichever index is contiguous
// in HBM becomes the contiguous index of the staged tile, so no transposition happens on the way in
// and the operand major-ness (K-major / MN-major) is chosen in the MMA descriptors instead:
//     A(i,k) k-contiguous -> tile [i][k], K-major      A(i,k) i-contiguous -> tile [k][i], MN-major
//     B(k,j) j-contiguous -> tile [k][j], MN-major     B(k,j) k-contiguous -> tile [j][k], K-major
// Two staging buffers: the MMAs of chunk c (one elected thread, 4 x K=16 steps) run while the CTA
// stages chunk c + 1; an mbarrier per buffer (tcgen05.commit) says when it may be overwritten.
// Epilogue: tcgen05.ld -> bias / relu / relu' mask -> fp32 store, += or atomicAdd (split-K weight
// gradients).  Big-M contractions here are HBM-bound (K = 128: 0.5 FLOP per byte of C), the
// per-time-step GRU contractions are latency-bound; either way the tensor pipe is no longer the
// limit the fp32 SIMT kernel was.
#include "gemm.cuh"
#include "tc.cuh"

namespace mava {
namespace {

using namespace tc;

constexpr int TM = 128, BK = 64, NT = 256;

// Rows of the fp32 matrices are walked 32 bytes per thread, every lane of a warp on a different row:
// an access instruction then costs one LSU wavefront per lane whatever its width, so 256-bit accesses
// (sm_100) halve the wavefronts per byte against pairs of 128-bit ones.  The big-M contractions of the
// recurrent path are bound by exactly that.
__device__ __forceinline__ void ldg256(const float* p, float4& a, float4& b) {
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
               : "l"(p));
}
__device__ __forceinline__ void st256(float* p, const float4& a, const float4& b) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(a.x), "f"(a.y),
               "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w)
               : "memory");
}

struct Ctrl {
  uint64_t bar[2];
  uint32_t tmem;
};

// One operand tile of a K chunk on its way from HBM to shared memory.  load() issues ALL of the
// thread's global loads (ITERS x 32 bytes) before anything depends on them, store() converts to
// bf16 and writes the 16-byte core-matrix rows: with one CTA per SM (the per-time-step GRU
// contractions) the chunk costs one memory round trip instead of one per 8-column group.
// Element (r, c) of the [ROWS][COLS] tile = src[(r0 + r) * ld + c0 + c] for r0 + r < rmax and
// c0 + c < cmax, else 0; c is the HBM-contiguous index.  vec: rows of src are 16-byte aligned.
// GROUPS = rows * cols / 8 (8-column groups in the tile), cgs_log2 = log2(cols / 8).
template <int GROUPS>
struct Stager {
  static constexpr int ITERS = (GROUPS + NT - 1) / NT;
  float4 v[ITERS][2];

  // Eight consecutive threads take the same 8-column group of eight consecutive rows: their
  // 16-byte shared-memory stores fill one 128-byte core matrix (no bank conflicts), and every
  // 32-byte global load is a whole sector of its row.
  static __device__ __forceinline__ void map(int idx, int cgs_log2, int& r, int& cg) {
    const int rest = idx >> 3;
    cg = rest & ((1 << cgs_log2) - 1);
    r = ((rest >> cgs_log2) << 3) | (idx & 7);
  }

  __device__ __forceinline__ void load(const float* __restrict__ src, int64_t ld, int r0, int rmax,
                                       int c0, int cmax, bool vec, int cgs_log2) {
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      const int idx = it * NT + (int)threadIdx.x;
      int r, cg;
      map(idx, cgs_log2, r, cg);
      const int gr = r0 + r, gc = c0 + cg * 8;
      v[it][0] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      v[it][1] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      if (idx < GROUPS && gr < rmax && gc < cmax) {
        const float* p = src + (int64_t)gr * ld + gc;
        if (gc + 8 <= cmax && vec && ((size_t)p & 31) == 0) {
          ldg256(p, v[it][0], v[it][1]);
        } else if (gc + 8 <= cmax && vec) {
          v[it][0] = __ldg(reinterpret_cast<const float4*>(p));
          v[it][1] = __ldg(reinterpret_cast<const float4*>(p) + 1);
        } else {
          float e[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) e[j] = gc + j < cmax ? __ldg(p + j) : 0.0f;
          v[it][0] = make_float4(e[0], e[1], e[2], e[3]);
          v[it][1] = make_float4(e[4], e[5], e[6], e[7]);
        }
      }
    }
  }

  __device__ __forceinline__ void store(const Tile& t, int cgs_log2) const {
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      const int idx = it * NT + (int)threadIdx.x;
      if (idx < GROUPS) {
        int r, cg;
        map(idx, cgs_log2, r, cg);
        st_shared_v4(t.base + chunk_off(t, r, cg), pack_bf16(v[it][0].x, v[it][0].y),
                     pack_bf16(v[it][0].z, v[it][0].w), pack_bf16(v[it][1].x, v[it][1].y),
                     pack_bf16(v[it][1].z, v[it][1].w));
      }
    }
  }
};

template <int BN>
__global__ void __launch_bounds__(NT, 2) tc_gemm_kernel(const GemmArgs p) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ Ctrl ctrl;
  constexpr uint32_t A_BYTES = TM * BK * 2, B_BYTES = BN * BK * 2;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int m0 = blockIdx.x * TM, n0 = blockIdx.y * BN;
  const int kbeg = blockIdx.z * p.kchunk;
  const int kend = min(p.K, kbeg + p.kchunk);
  const uint32_t s0 = smem_u32(smem);

  if (warp == 0) tmem_alloc<(BN < 32 ? 32 : BN)>(&ctrl.tmem);
  if (t == 0) {
    mbar_init(&ctrl.bar[0], 1);
    mbar_init(&ctrl.bar[1], 1);
    fence_mbar_init();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = ctrl.tmem;

  // staged tile geometry: 8-row groups of one 8-column group are contiguous (s_r = 128)
  //   A k-contiguous: [TM][BK]  K-major        A i-contiguous: [BK][TM]  MN-major
  //   B j-contiguous: [BK][BN]  MN-major       B k-contiguous: [BN][BK]  K-major
  const bool a_mn = p.ta != 0, b_mn = p.tb == 0;
  const uint32_t a_sc = (uint32_t)((a_mn ? BK : TM) / 8) * 128u;
  const uint32_t b_sc = (uint32_t)((b_mn ? BK : BN) / 8) * 128u;
  const bool a_vec = (p.lda & 3) == 0 && ((size_t)p.A & 15) == 0;
  const bool b_vec = (p.ldb & 3) == 0 && ((size_t)p.B & 15) == 0;
  const uint32_t idesc = instr_desc(TM, BN, a_mn, b_mn);
  const int a_lg = a_mn ? 4 : 3;                                   // log2(TM / 8), log2(BK / 8)
  const int b_lg = b_mn ? (BN == 128 ? 4 : BN == 64 ? 3 : 1) : 3;  // log2(BN / 8), log2(BK / 8)

  uint32_t ph[2] = {0u, 0u};
  int used[2] = {0, 0};
  int c = 0;
  for (int k0 = kbeg; k0 < kend; k0 += BK, ++c) {
    const int buf = c & 1;
    const Tile at{s0 + (uint32_t)buf * (A_BYTES + B_BYTES), 128u, a_sc};
    const Tile bt{at.base + A_BYTES, 128u, b_sc};
    // all global loads of the chunk first (they fly while the buffer's previous MMAs finish) ...
    Stager<TM * BK / 8> sa;
    Stager<BN * BK / 8> sb;
    if (!a_mn) sa.load(p.A, p.lda, m0, p.M, k0, kend, a_vec, a_lg);   // [TM][BK], k contiguous
    else sa.load(p.A, p.lda, k0, kend, m0, p.M, a_vec, a_lg);         // [BK][TM], i contiguous
    if (b_mn) sb.load(p.B, p.ldb, k0, kend, n0, p.N, b_vec, b_lg);    // [BK][BN], j contiguous
    else sb.load(p.B, p.ldb, n0, p.N, k0, kend, b_vec, b_lg);         // [BN][BK], k contiguous
    if (used[buf]) {  // the MMAs that read this buffer two chunks ago must have completed
      mbar_wait(&ctrl.bar[buf], ph[buf]);
      ph[buf] ^= 1u;
    }
    // ... then bf16 conversion and the shared-memory stores
    sa.store(at, a_lg);
    sb.store(bt, b_lg);
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    if (mma_issuer()) {
      fence_after_sync();
#pragma unroll
      for (int k = 0; k < BK / 16; ++k) {
        const uint64_t ad = a_mn ? desc_mnmajor(at, k) : desc_kmajor(at, k);
        const uint64_t bd = b_mn ? desc_mnmajor(bt, k) : desc_kmajor(bt, k);
        mma(tmem, ad, bd, idesc, c > 0 || k > 0);
      }
      commit(&ctrl.bar[buf]);
    }
    used[buf] = 1;
  }
  // drain: the last commit on each used buffer (commits complete in issue order)
  for (int b = 0; b < 2; ++b) {
    if (used[b]) {
      mbar_wait(&ctrl.bar[b], ph[b]);
      ph[b] ^= 1u;
    }
  }
  fence_after_sync();

  // ---- epilogue: warp w reads TMEM lanes 32 (w % 4) .. + 31, column half w / 4
  const int r = (warp & 3) * 32 + lane;
  const int i = m0 + r;
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
  constexpr int HALF = BN >= 32 ? BN / 2 : BN;  // columns per warp group
  const int cbase = BN >= 32 ? (warp >> 2) * HALF : 0;
  const bool writer = BN >= 32 || (warp >> 2) == 0;
  if (c > 0 && writer) {
    // fast path: whole 16-column groups, 16-byte aligned rows, no atomics.  The old C values
    // (accumulate) and the relu' reference are fetched up front, all loads in flight at once.
    const int jb = n0 + cbase;
    float* crow = p.C + (int64_t)i * p.ldc + jb;
    const float* rrow = p.relu_ref ? p.relu_ref + (int64_t)i * p.ldr + jb : nullptr;
    const bool fast = i < p.M && jb + HALF <= p.N && p.mode != 2 && (((size_t)crow) & 15) == 0 &&
                      (rrow == nullptr || (((size_t)rrow) & 15) == 0);
    float4 cold[HALF / 4], rref[HALF / 4];
    const bool wide = fast && (((size_t)crow) & 31) == 0;  // 32-byte aligned rows: 256-bit stores
    if (fast) {
#pragma unroll
      for (int q = 0; q < HALF / 4; ++q) {
        cold[q] = p.mode == 1 ? *reinterpret_cast<const float4*>(crow + q * 4)
                              : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        rref[q] = rrow ? *reinterpret_cast<const float4*>(rrow + q * 4)
                       : make_float4(1.0f, 1.0f, 1.0f, 1.0f);
      }
    }
#pragma unroll
    for (int cc = 0; cc < HALF; cc += 16) {
      float v[16];
      ld16(tmem + lane_base + (uint32_t)(cbase + cc), v);
      const int j0 = jb + cc;
      if (fast) {
#pragma unroll
        for (int q8 = 0; q8 < 2; ++q8) {
          float4 o4[2];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int q4 = q8 * 2 + h;
            const float4 o = cold[cc / 4 + q4], rr = rref[cc / 4 + q4];
            const float ov[4] = {o.x, o.y, o.z, o.w}, rv[4] = {rr.x, rr.y, rr.z, rr.w};
            float x[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float y = p.alpha * v[q4 * 4 + q];
              if (p.bias != nullptr && blockIdx.z == 0) y += __ldg(p.bias + j0 + q4 * 4 + q);
              if (p.relu) y = fmaxf(y, 0.0f);
              if (!(rv[q] > 0.0f)) y = 0.0f;
              x[q] = y + ov[q];
            }
            o4[h] = make_float4(x[0], x[1], x[2], x[3]);
          }
          if (wide) {
            st256(crow + cc + q8 * 8, o4[0], o4[1]);
          } else {
            *reinterpret_cast<float4*>(crow + cc + q8 * 8) = o4[0];
            *reinterpret_cast<float4*>(crow + cc + q8 * 8 + 4) = o4[1];
          }
        }
    } else if (i < p.M && j0 < p.N) {
        float* dst = crow + cc;
        const float* ref = rrow ? rrow + cc : nullptr;
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          if (j0 + q < p.N) {
            float x = p.alpha * v[q];
            if (p.bias != nullptr && blockIdx.z == 0) x += __ldg(p.bias + j0 + q);
            if (p.relu) x = fmaxf(x, 0.0f);
            if (ref != nullptr && !(ref[q] > 0.0f)) x = 0.0f;
            if (p.mode == 0) dst[q] = x;
            else if (p.mode == 1) dst[q] += x;
            else atomicAdd(dst + q, x);
          }
        }
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<(BN < 32 ? 32 : BN)>(tmem);
}

template <int BN>
int launch(const GemmArgs& a, cudaStream_t s) {
  constexpr size_t smem = 2 * (size_t)(TM * BK * 2 + BN * BK * 2) + 128;
  static size_t configured[kMaxDevices] = {};
  if (int rc = ensure_dyn_smem(tc_gemm_kernel<BN>, smem, configured)) return rc;
  dim3 grid(ceil_div(a.M, TM), ceil_div(a.N, BN), ceil_div(a.K, a.kchunk));
  tc_gemm_kernel<BN><<<grid, NT, smem, s>>>(a);
  return launch_status();
}

}  // namespace

int launch_tc_gemm(const GemmArgs& a, cudaStream_t s) {
  if (a.M <= 0 || a.N <= 0 || a.K <= 0) return 0;
  if (a.kchunk % BK != 0) return MAVA_E_BADARG;
  if (a.N <= 16) return launch<16>(a, s);
  if (a.N <= 64) return launch<64>(a, s);
  return launch<128>(a, s);
}

}  // namespace mava
