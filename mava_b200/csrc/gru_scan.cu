// Persistent GRU scan kernels for the recurrent systems (ScannedRNN, mava/networks.py:238-266, as
// rec_mappo.py:208-312 differentiates it): the whole time loop of a sequence minibatch in ONE launch
// per direction instead of two launches per time step.
//
// A CTA (512 threads) owns 128 sequences for all L positions.  W_h (H x 3H, H = 128) is packed once
// into bf16 operand tiles that stay in shared memory for the whole scan (96 KB); the hidden state of
// a sequence lives in fp32 registers of the thread that owns (row, 32 columns) and is mirrored as a
// bf16 tile in shared memory, the A operand of the step's contraction.
//
//   forward, per position l:    Gh = h_in W_h                        3 x (128x128x128) tcgen05 GEMMs -> TMEM
//                               r, z, n, h' (fp32, flax GRUCell)      epilogue on the accumulator
//                               h_in(l+1) = done_in(l+1) ? 0 : h'     ScannedRNN's reset
//   backward, per position l:   dh = dHout(l) + [no reset at l+1] (dh(l+1) z(l+1) + dT(l+1))
//                               gate derivatives -> dGx(l), dGh(l) (fp32 to HBM for the weight
//                               gradients, bf16 tile to shared memory)
//                               dT(l) = dGh(l) W_h^T                  3 chained GEMMs (K = 384) -> TMEM
//
// Same buffers as the per-step schedule of rnn_f32.cu (Gx, gate stash, Hin, Hout in fp32), so the
// dense contractions around the scan are unchanged; same arithmetic as its bf16 path (h and dGh are
// rounded to bf16 for the tensor core, everything else is fp32).
#include "gru_scan.cuh"
#include "mlp_tc.cuh"

namespace mava {
namespace {

using namespace tc;
using namespace tcmlp;  // Lane, issue_gemm, TM, NT

constexpr int H = 128;
constexpr uint32_t kTile = tile_bytes(TM, H);  // 32 KB: one 128 x 128 bf16 operand tile

struct ScanArgs {
  const float* Wh;    // [H][3H]
  const float* b_hn;  // [H]
  const int32_t* steps;    // [L][Senv] env-step of (position, env-sequence)
  const uint8_t* done_in;  // by env-step: the flag entering that step
  int64_t S;               // rows (sequences) per position
  int L, rpe;
  int64_t Senv;
  // forward
  const float* Gx;  // [L][S][3H]  x W_i + b_i
  float* gates;     // [L][S][4H]  stash (r, z, n, q) or null
  float* Hin;       // [L][S][H]   position 0 given, positions 1.. written
  float* Hout;      // [L][S][H]
  // backward
  const float* dHout;  // [L][S][H]
  float* dGx;          // [L][S][3H]
  float* dH0;          // [S][H] gradient flowing into the chunk-start state (dh z + dT at l = 0), or null
};

// Up to two networks (actor and critic of one minibatch) scanned by ONE launch: CTA b < tiles0 works
// on a[0], the others on a[1].  The critic of the MAPPO systems has 8x fewer sequences than the actor
// (16 CTAs against 128): launched alone it takes as long as the actor's scan, next to it it is free.
struct ScanArgs2 {
  ScanArgs a[2];
  int tiles0;
};

struct SCtrl {
  uint64_t mbar;
  uint32_t tmem;
};

__device__ __forceinline__ float sigmoid_(float x) { return 1.0f / (1.0f + expf(-x)); }

// A thread walks its row 32 bytes at a time (8 columns per pass): ask L2 for the whole 128-byte line
// on the first touch, and warm the lines of the NEXT position while this one is computed -- otherwise
// DRAM sees four separate 32-byte bursts per line and the scan runs at a third of the HBM rate.
// 8 consecutive floats (32 bytes) of a row in ONE 256-bit access.  Every lane of a warp is a different
// row, so each access instruction costs one LSU wavefront per lane whatever its width: 32-byte
// accesses halve the wavefronts per byte against 16-byte ones (the scan is bound by them, not by HBM).
__device__ __forceinline__ void ld8f(const float* p, float (&a)[8]) {
  asm volatile("ld.global.L2::128B.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(a[0]), "=f"(a[1]), "=f"(a[2]), "=f"(a[3]), "=f"(a[4]), "=f"(a[5]), "=f"(a[6]),
                 "=f"(a[7])
               : "l"(p));
}
__device__ __forceinline__ void st8f(float* p, const float (&a)[8]) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(a[0]), "f"(a[1]),
               "f"(a[2]), "f"(a[3]), "f"(a[4]), "f"(a[5]), "f"(a[6]), "f"(a[7])
               : "memory");
}
__device__ __forceinline__ void prefetch_line(const float* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// W_h [H][3H] fp32 -> three bf16 tiles [K = H rows][N = H columns] (one per gate) at `base`
__device__ __forceinline__ void pack_wh(const float* __restrict__ Wh, uint32_t base) {
  // 8-column groups: 3 gates x 128 rows x 16 groups
  for (int idx = threadIdx.x; idx < 3 * H * (H / 8); idx += blockDim.x) {
    const int g = idx / (H * (H / 8));
    const int rem = idx - g * (H * (H / 8));
    const int cg = rem / H, k = rem - cg * H;  // consecutive threads: consecutive rows k of one group
    // (scalar loads: the critic's parameters follow the actor's in one flat vector, so W_h is only
    // 4-byte aligned in general; this runs once per CTA)
    const float* src = Wh + (size_t)k * 3 * H + g * H + cg * 8;
    const float4 a = make_float4(__ldg(src), __ldg(src + 1), __ldg(src + 2), __ldg(src + 3));
    const float4 b = make_float4(__ldg(src + 4), __ldg(src + 5), __ldg(src + 6), __ldg(src + 7));
    const Tile t{base + (uint32_t)g * kTile, 128u, 2048u};
    st_shared_v4(t.base + chunk_off(t, k, cg), pack_bf16(a.x, a.y), pack_bf16(a.z, a.w),
                 pack_bf16(b.x, b.y), pack_bf16(b.z, b.w));
  }
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT, 1) gru_scan_fwd_kernel(const ScanArgs2 pp) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ SCtrl ctrl;
  const Lane L;
  const int t = L.t;
  const bool second = (int)blockIdx.x >= pp.tiles0;
  const ScanArgs& p = pp.a[second ? 1 : 0];
  const int tile = second ? blockIdx.x - pp.tiles0 : blockIdx.x;
  const uint32_t s_w = smem_u32(smem);
  const Tile ht{s_w + 3 * kTile, 128u, 2048u};
  const int64_t row = (int64_t)tile * TM + L.r;
  const bool valid = row < p.S;
  const int64_t q_env = valid ? row / p.rpe : 0;
  const int c0 = L.q * 32;  // this thread's 32 columns of every gate

  if (L.warp == 0) tmem_alloc<512>(&ctrl.tmem);
  if (t == 0) {
    mbar_init(&ctrl.mbar, 1);
    fence_mbar_init();
  }
  pack_wh(p.Wh, s_w);
  // chunk-start state (already reset-masked by the caller) -> registers and the bf16 operand tile
  float h[32];
#pragma unroll
  for (int cg = 0; cg < 4; ++cg) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (valid) {
      const float4* src = reinterpret_cast<const float4*>(p.Hin + row * H + c0 + cg * 8);
      a = src[0];
      b = src[1];
    }
    h[cg * 8 + 0] = a.x; h[cg * 8 + 1] = a.y; h[cg * 8 + 2] = a.z; h[cg * 8 + 3] = a.w;
    h[cg * 8 + 4] = b.x; h[cg * 8 + 5] = b.y; h[cg * 8 + 6] = b.z; h[cg * 8 + 7] = b.w;
    st_shared_v4(ht.base + chunk_off(ht, L.r, L.q * 4 + cg), pack_bf16(a.x, a.y),
                 pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w));
  }
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = ctrl.tmem;
  uint32_t phase = 0;

  for (int l = 0; l < p.L; ++l) {
    if (mma_issuer()) {
      fence_after_sync();
#pragma unroll
      for (int g = 0; g < 3; ++g) {
        const Tile wg{s_w + (uint32_t)g * kTile, 128u, 2048u};
        issue_gemm(tmem + (uint32_t)g * H, ht, false, wg, true, H, H, false, g == 2 ? &ctrl.mbar : nullptr);
      }
    }
    // while the tensor core works: reset flag of the next position
    bool reset_next = false;
    if (valid && l + 1 < p.L)
      reset_next = p.done_in[p.steps[(int64_t)(l + 1) * p.Senv + q_env]] != 0;
    const float* gx = p.Gx + ((int64_t)l * p.S + (valid ? row : 0)) * 3 * H + c0;
    float* go = p.gates ? p.gates + ((int64_t)l * p.S + (valid ? row : 0)) * 4 * H + c0 : nullptr;
    float* ho = p.Hout + ((int64_t)l * p.S + (valid ? row : 0)) * H + c0;
    float* hn = l + 1 < p.L ? p.Hin + ((int64_t)(l + 1) * p.S + (valid ? row : 0)) * H + c0 : nullptr;
    // x-side pre-activations of the first 8 columns, in flight under the MMAs
    float xr[8], xz[8], xn[8];
    auto load_x = [&](int cg) {
      if (valid) {
        ld8f(gx + cg * 8, xr);
        ld8f(gx + H + cg * 8, xz);
        ld8f(gx + 2 * H + cg * 8, xn);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) xr[j] = xz[j] = xn[j] = 0.0f;
      }
    };
    load_x(0);
    if (valid && l + 1 < p.L) {  // the three 128-byte lines this thread reads at the next position
      const float* nx = gx + p.S * 3 * H;
      prefetch_line(nx);
      prefetch_line(nx + H);
      prefetch_line(nx + 2 * H);
    }
    mbar_wait(&ctrl.mbar, phase);
    phase ^= 1u;
    fence_after_sync();
#pragma unroll
    for (int cg = 0; cg < 4; ++cg) {
      float gr[8], gz[8], gn[8];
      ld8(tmem + L.tmem_lane() + (uint32_t)(0 * H + c0 + cg * 8), gr);
      ld8(tmem + L.tmem_lane() + (uint32_t)(1 * H + c0 + cg * 8), gz);
      ld8(tmem + L.tmem_lane() + (uint32_t)(2 * H + c0 + cg * 8), gn);
      float r[8], z[8], n[8], qv[8], hv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        r[j] = xr[j] + gr[j];
        z[j] = xz[j] + gz[j];
        n[j] = xn[j];
      }
      if (cg + 1 < 4) load_x(cg + 1);  // next 8 columns while these are computed
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        r[j] = sigmoid_(r[j]);
        z[j] = sigmoid_(z[j]);
        qv[j] = gn[j] + __ldg(p.b_hn + c0 + cg * 8 + j);
        n[j] = tanhf(n[j] + r[j] * qv[j]);
        hv[j] = (1.0f - z[j]) * n[j] + z[j] * h[cg * 8 + j];
      }
      if (valid) {
        st8f(ho + cg * 8, hv);
        if (go) {
          st8f(go + cg * 8, r);
          st8f(go + H + cg * 8, z);
          st8f(go + 2 * H + cg * 8, n);
          st8f(go + 3 * H + cg * 8, qv);
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) hv[j] = reset_next ? 0.0f : hv[j];
      if (valid && hn) st8f(hn + cg * 8, hv);
#pragma unroll
      for (int j = 0; j < 8; ++j) h[cg * 8 + j] = hv[j];
      st_shared_v4(ht.base + chunk_off(ht, L.r, L.q * 4 + cg), pack_bf16(hv[0], hv[1]),
                   pack_bf16(hv[2], hv[3]), pack_bf16(hv[4], hv[5]), pack_bf16(hv[6], hv[7]));
    }
    // the next position's operand is complete and this accumulator has been read
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
  }
  if (L.warp == 0) tmem_dealloc<512>(tmem);
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT, 1) gru_scan_bwd_kernel(const ScanArgs2 pp) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ SCtrl ctrl;
  const Lane L;
  const int t = L.t;
  const bool second = (int)blockIdx.x >= pp.tiles0;
  const ScanArgs& p = pp.a[second ? 1 : 0];
  const int tile = second ? blockIdx.x - pp.tiles0 : blockIdx.x;
  const uint32_t s_w = smem_u32(smem);
  const uint32_t s_g = s_w + 3 * kTile;  // dGh as three [128 rows][128 k] tiles (r, z, q parts)
  const int64_t row = (int64_t)tile * TM + L.r;
  const bool valid = row < p.S;
  const int64_t q_env = valid ? row / p.rpe : 0;
  const int c0 = L.q * 32;

  if (L.warp == 0) tmem_alloc<128>(&ctrl.tmem);
  if (t == 0) {
    mbar_init(&ctrl.mbar, 1);
    fence_mbar_init();
  }
  pack_wh(p.Wh, s_w);
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = ctrl.tmem;
  uint32_t phase = 0;
  float cz[32];  // dh(l+1) * z(l+1): the part of the carry that needs no contraction
#pragma unroll
  for (int j = 0; j < 32; ++j) cz[j] = 0.0f;
  bool have_carry = false;  // position l+1 exists and did not reset

  for (int l = p.L - 1; l >= 0; --l) {
    const int64_t base = (int64_t)l * p.S + (valid ? row : 0);
    const float* dho = p.dHout + base * H + c0;
    const float* hin = p.Hin + base * H + c0;
    float* gs = p.gates + base * 4 * H + c0;
    float* dgx = p.dGx + base * 3 * H + c0;
    if (valid && l > 0) {  // the lines of position l - 1, fetched while this one is computed
      const int64_t pb = base - p.S;
      prefetch_line(p.dHout + pb * H + c0);
      prefetch_line(p.Hin + pb * H + c0);
#pragma unroll
      for (int g = 0; g < 4; ++g) prefetch_line(p.gates + pb * 4 * H + g * H + c0);
    }
    // dT(l+1) = dGh(l+1) W_h^T from the previous iteration's MMAs
    if (l + 1 < p.L) {
      mbar_wait(&ctrl.mbar, phase);
      phase ^= 1u;
      fence_after_sync();
    }
#pragma unroll
    for (int cg = 0; cg < 4; ++cg) {
      float dt[8];
      if (l + 1 < p.L) ld8(tmem + L.tmem_lane() + (uint32_t)(c0 + cg * 8), dt);
      float dh_[8], r[8], z[8], n[8], qv[8], hi[8];
      if (valid) {
        ld8f(dho + cg * 8, dh_);
        ld8f(gs + cg * 8, r);
        ld8f(gs + H + cg * 8, z);
        ld8f(gs + 2 * H + cg * 8, n);
        ld8f(gs + 3 * H + cg * 8, qv);
        ld8f(hin + cg * 8, hi);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) dh_[j] = r[j] = z[j] = n[j] = qv[j] = hi[j] = 0.0f;
      }
      float dar[8], daz[8], dan[8], dq[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float dh = dh_[j];
        if (have_carry) dh += cz[cg * 8 + j] + dt[j];
        const float dn = dh * (1.0f - z[j]);
        const float dz = dh * (hi[j] - n[j]);
        dan[j] = dn * (1.0f - n[j] * n[j]);
        dar[j] = dan[j] * qv[j] * r[j] * (1.0f - r[j]);
        daz[j] = dz * z[j] * (1.0f - z[j]);
        dq[j] = dan[j] * r[j];
        cz[cg * 8 + j] = dh * z[j];
      }
      if (valid) {
        // the stash row becomes dGh = [da_r | da_z | dq | .] (input of the W_h gradient), dGx gets
        // [da_r | da_z | da_n]
        st8f(gs + cg * 8, dar);
        st8f(gs + H + cg * 8, daz);
        st8f(gs + 2 * H + cg * 8, dq);
        st8f(dgx + cg * 8, dar);
        st8f(dgx + H + cg * 8, daz);
        st8f(dgx + 2 * H + cg * 8, dan);
      }
      // bf16 operand tiles of the contraction dT(l) = dGh(l) W_h^T (rows beyond S are zero)
      const Tile tr{s_g, 128u, 2048u}, tz{s_g + kTile, 128u, 2048u}, tq{s_g + 2 * kTile, 128u, 2048u};
      const uint32_t off = chunk_off(tr, L.r, L.q * 4 + cg);
      st_shared_v4(tr.base + off, pack_bf16(dar[0], dar[1]), pack_bf16(dar[2], dar[3]),
                   pack_bf16(dar[4], dar[5]), pack_bf16(dar[6], dar[7]));
      st_shared_v4(tz.base + off, pack_bf16(daz[0], daz[1]), pack_bf16(daz[2], daz[3]),
                   pack_bf16(daz[4], daz[5]), pack_bf16(daz[6], daz[7]));
      st_shared_v4(tq.base + off, pack_bf16(dq[0], dq[1]), pack_bf16(dq[2], dq[3]),
                   pack_bf16(dq[4], dq[5]), pack_bf16(dq[6], dq[7]));
    }
    // does position l hand its gradient to position l - 1?  (a reset at l cuts the carry)
    have_carry = l > 0;
    if (valid && l > 0) have_carry = p.done_in[p.steps[(int64_t)l * p.Senv + q_env]] == 0;
    if (!valid) have_carry = false;
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    if (l > 0 || p.dH0 != nullptr) {
      if (mma_issuer()) {
        fence_after_sync();
#pragma unroll
        for (int g = 0; g < 3; ++g) {
          const Tile ag{s_g + (uint32_t)g * kTile, 128u, 2048u};
          const Tile wg{s_w + (uint32_t)g * kTile, 128u, 2048u};
          // B = gate g of W_h used K-major: B(k = gate column, n = W_h row)
          issue_gemm(tmem, ag, false, wg, false, H, H, g > 0, g == 2 ? &ctrl.mbar : nullptr);
        }
      }
    }
  }
  if (p.dH0 != nullptr) {  // gradient into the chunk-start state (not used by the PPO systems)
    mbar_wait(&ctrl.mbar, phase);
    fence_after_sync();
#pragma unroll
    for (int cg = 0; cg < 4; ++cg) {
      float dt[8];
      ld8(tmem + L.tmem_lane() + (uint32_t)(c0 + cg * 8), dt);
      if (valid)
        for (int j = 0; j < 8; ++j) p.dH0[row * H + c0 + cg * 8 + j] = cz[cg * 8 + j] + dt[j];
    }
  }
  fence_before_sync();
  __syncthreads();
  if (L.warp == 0) tmem_dealloc<128>(tmem);
}

}  // namespace

// host entry points used by rnn_f32.cu (bf16 precision, H == 128); `n` = 1 or 2 networks
int launch_gru_scan_fwd(const GruScanNet* nets, int n, cudaStream_t s) {
  ScanArgs2 pp{};
  int tiles = 0;
  for (int i = 0; i < n; ++i) {
    ScanArgs& a = pp.a[i];
    const GruScanNet& g = nets[i];
    a.Wh = g.Wh; a.b_hn = g.b_hn; a.steps = g.steps; a.done_in = g.done_in; a.S = g.S; a.L = g.L;
    a.rpe = g.rpe; a.Senv = g.Senv; a.Gx = g.Gx; a.gates = g.gates; a.Hin = g.Hin; a.Hout = g.Hout;
    if (i == 0) pp.tiles0 = (int)ceil_div64(g.S, TM);
    tiles += (int)ceil_div64(g.S, TM);
  }
  const size_t smem = 4 * (size_t)kTile + 128;
  static size_t configured[kMaxDevices] = {};
  if (int rc = ensure_dyn_smem(gru_scan_fwd_kernel, smem, configured)) return rc;
  gru_scan_fwd_kernel<<<(unsigned)tiles, NT, smem, s>>>(pp);
  return launch_status();
}

int launch_gru_scan_bwd(const GruScanNet* nets, int n, cudaStream_t s) {
  ScanArgs2 pp{};
  int tiles = 0;
  for (int i = 0; i < n; ++i) {
    ScanArgs& a = pp.a[i];
    const GruScanNet& g = nets[i];
    a.Wh = g.Wh; a.steps = g.steps; a.done_in = g.done_in; a.S = g.S; a.L = g.L; a.rpe = g.rpe;
    a.Senv = g.Senv; a.dHout = g.Hout; a.Hin = g.Hin; a.gates = g.gates; a.dGx = g.Gx;
    a.dH0 = nullptr;
    if (i == 0) pp.tiles0 = (int)ceil_div64(g.S, TM);
    tiles += (int)ceil_div64(g.S, TM);
  }
  const size_t smem = 6 * (size_t)kTile + 128;
  static size_t configured[kMaxDevices] = {};
  if (int rc = ensure_dyn_smem(gru_scan_bwd_kernel, smem, configured)) return rc;
  gru_scan_bwd_kernel<<<(unsigned)tiles, NT, smem, s>>>(pp);
  return launch_status();
}

}  // namespace mava
