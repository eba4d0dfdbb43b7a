// tcgen05 / TMEM / mbarrier primitives for sm_100a (inline PTX), plus the shared-memory operand
// format every tensor-core kernel here uses.
//
// Operand tiles live in shared memory as a grid of 8x8 bf16 "core matrices" (8 rows x 16 bytes,
// 128 contiguous bytes, no swizzle).  A logical tile T[r][c] (c contiguous) puts element (r, c) at
//     (r / 8) * S_r + (c / 8) * S_c + (r % 8) * 16 + (c % 8) * 2        [bytes]
// The same tile serves as
//   * a K-major operand  with MN index = r and K index = c  (SBO = S_r, LBO = S_c), or
//   * an MN-major operand with MN index = c and K index = r  (SBO = S_c, LBO = S_r),
// which is what lets one copy of an activation / weight tile feed the forward GEMM, the backward
// GEMM and the weight-gradient GEMM.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>

namespace mava {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- tile addressing ---------------------------------------------------------------------------
struct Tile {
  uint32_t base;  // shared-memory byte address
  uint32_t s_r;   // byte stride between 8-row groups
  uint32_t s_c;   // byte stride between 8-column groups
};

__host__ __device__ constexpr uint32_t tile_bytes(int rows, int cols) { return rows * cols * 2; }

// byte offset of the 16-byte chunk holding columns [8*cg, 8*cg+8) of row r
__device__ __forceinline__ uint32_t chunk_off(const Tile& t, int r, int cg) {
  return (uint32_t)(r >> 3) * t.s_r + (uint32_t)cg * t.s_c + (uint32_t)(r & 7) * 16u;
}

// ---- descriptors ---------------------------------------------------------------------------------
// Shared-memory matrix descriptor, SWIZZLE_NONE, version 1 (Blackwell).
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

// Instruction descriptor for kind::f16 with bf16 A/B and fp32 accumulation.
__host__ __device__ constexpr uint32_t instr_desc(int M, int N, bool a_mn_major, bool b_mn_major) {
  return (1u << 4)                       // D format f32
         | (1u << 7)                     // A format bf16
         | (1u << 10)                    // B format bf16
         | ((a_mn_major ? 1u : 0u) << 15)
         | ((b_mn_major ? 1u : 0u) << 16)
         | ((uint32_t)(N >> 3) << 17)
         | ((uint32_t)(M >> 4) << 24);
}

// A/B descriptor of the k-th K=16 slice of a tile used K-major (MN = r, K = c) ...
__device__ __forceinline__ uint64_t desc_kmajor(const Tile& t, int kstep, int mn_row0 = 0) {
  return smem_desc(t.base + (uint32_t)(mn_row0 >> 3) * t.s_r + (uint32_t)kstep * 2u * t.s_c, t.s_c,
                   t.s_r);
}
// ... or MN-major (MN = c, K = r)
__device__ __forceinline__ uint64_t desc_mnmajor(const Tile& t, int kstep, int mn_col0 = 0) {
  return smem_desc(t.base + (uint32_t)(mn_col0 >> 3) * t.s_c + (uint32_t)kstep * 2u * t.s_r, t.s_r,
                   t.s_c);
}

// ---- TMEM ----------------------------------------------------------------------------------------
template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_result)),
               "n"(COLS));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}

template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS));
}

__device__ __forceinline__ void fence_before_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void fence_after_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// generic-proxy smem writes -> visible to the async proxy (tensor core operand reads)
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// D[tmem] (+)= A[smem] * B[smem], issued by ONE thread.
__device__ __forceinline__ void mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                    bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}

// True in exactly one lane of warp 0 (call with warp 0 converged).  A warp-uniform condition plus
// a hardware election instead of `threadIdx.x == 0`: under a thread-index branch nvcc treats the
// MMA descriptors as per-lane values and rebuilds each one through an R2UR waterfall loop (~13
// dependent instructions per tcgen05.mma, on the critical path of every phase); here they stay in
// uniform registers.
__device__ __forceinline__ bool mma_issuer() {
  const int warp_u = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  if (warp_u != 0) return false;
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// Arrive on an mbarrier when all previously issued MMAs of this thread have completed.
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}

// 32 lanes x 32 columns of fp32: thread i of the warp gets row (lane base + i), 32 columns.
__device__ __forceinline__ void ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
      "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- mbarrier --------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// Spin until the phase with the given parity completes.  The spin is bounded: a kernel bug traps
// (an error the host sees) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  for (uint32_t spins = 0; !done; ++spins) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (spins > (1u << 24)) {
      printf("mava_b200: mbarrier wait timed out (block %d thread %d parity %u)\n", (int)blockIdx.x,
             (int)threadIdx.x, parity);
      __trap();
    }
  }
}

// ---- packing -----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c,
                                             uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c),
               "r"(d)
               : "memory");
}

}  // namespace tc
}  // namespace mava
