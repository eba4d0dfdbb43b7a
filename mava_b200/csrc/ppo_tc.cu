// Fused PPO minibatch kernel on the tcgen05 tensor cores (bf16 operands, fp32 TMEM accumulators).
//
// Replaces value_and_grad(_actor_loss_fn) + value_and_grad(_critic_loss_fn) + pmean("batch")
// (mava/systems/ppo/ff_mappo.py:150-234) for one minibatch:
//
//   ppo_fused_kernel   persistent CTAs, actor tiles and critic tiles as different CTAs of one launch.
//     per 128-row tile:   gather int8 obs rows through the shuffle index (no shuffled copy is ever
//                         materialised) -> X tile
//                         H1 = relu(X W1 + b1), H2 = relu(H1 W2 + b2), out = H2 W3 + b3
//                         loss epilogue in registers (softmax / clip / entropy | clipped value loss)
//                         dZ3 -> dH2 = dZ3 W3^T ; dW3 += H2^T dZ3          (accumulated in TMEM)
//                         dZ2 = dH2 * relu' -> dH1 = dZ2 W2^T ; dW2^T,db2 += dZ2^T [H1 | 1]  (TMEM)
//                         dZ1 = dH1 * relu' -> bf16 tile image in HBM
//     at the end:         TMEM weight-gradient accumulators -> fp32 atomics into the gradient buffer
//   ppo_wgrad1_kernel  first-layer gradient: dW1^T,db1 += dZ1^T [X | 1] over all tiles (TMEM), with the
//                      dZ1 tile images brought in by bulk (TMA) copies and X rebuilt from the int8 obs.
//
// Every activation / gradient tile is stored once in shared memory and consumed by up to three GEMMs
// (forward or backward as K-major A, weight gradient as MN-major A or B) - see tc.cuh.
//
// Roles inside a CTA of ppo_fused_kernel (640 threads, 20 warps x 96 registers):
//   warps 0-15   epilogue warps: thread (row r, column quarter q) of the tile, see mlp_tc.cuh.  They
//                turn accumulators into the next GEMM's operand (TMEM -> registers -> bf16 -> shared
//                memory) and finish the loss rows; they issue no MMA and (on the staged paths) no
//                global load.
//   warp 16      MMA-issue warp: waits on a named barrier for "operands ready", issues the GEMM the
//                next epilogue needs, commits, then issues the weight-gradient GEMMs behind the commit.
//   warps 17-19  loader warps: the next tile's observation bytes, loss inputs and index lists (plain
//                loads; they never execute a proxy fence, which would wait for their loads).
// Handshakes: epilogue -> issue: two alternating named barriers (bar.arrive / bar.sync);
// issue -> epilogue: tcgen05.commit on one mbarrier per kind of event; loaders <-> epilogue: FULL /
// EMPTY named barriers around the staging rows; bulk (TMA) copies complete on an mbarrier.
#include <cstdlib>

#include "mlp_tc.cuh"

namespace mava {
// from mlp_f32.cu
int launch_adv_stats(const float* adv, const int32_t* rows, int mb_size, int A, int num_replicas,
                     double* stats, cudaStream_t s);
int launch_finalize_loss(const double* acc, double denom, float ent_coef, float vf_coef, float* out5,
                         cudaStream_t s);

namespace tcmlp {
int make_net(const mava_mlp_desc* d, const float* params, NetDesc* n);

namespace {

// Phase timing (development aid): compile with -DMAVA_PROFILE_PHASES to have thread 0 of CTA 0
// record clock64() at the phase boundaries of its first tiles (read back with
// mava_debug_phases).  Compiled out by default.
#ifdef MAVA_PROFILE_PHASES
__device__ long long g_wg1_clock[8 * 12];  // ppo_wgrad1_kernel: CTA 0, first 8 tiles, 12 stamps
#define MAVA_WSTAMP(k)                                                             \
  do {                                                                             \
    if (t == 0 && blockIdx.x == 0 && it < 8) g_wg1_clock[it * 12 + (k)] = clock64(); \
  } while (0)
__device__ long long g_phase_clock[16 * 16];
__device__ long long g_phase_clock2[16 * 8 + 32];
#define MAVA_STAMP2(k)                                                             \
  do {                                                                             \
    if (t == 128 && blockIdx.x == 0 && it >= 2 && it < 18) g_phase_clock2[(it - 2) * 8 + (k)] = clock64(); \
  } while (0)
#ifdef MAVA_STAMP_CRITIC
#define MAVA_STAMP_CTA (gridDim.x - 1)
#else
#define MAVA_STAMP_CTA 0
#endif
#define MAVA_STAMP3(k)                                                             \
  do {                                                                             \
    if (t == 0 && blockIdx.x == MAVA_STAMP_CTA && it >= 2 && it < 18) g_phase_clock2[(it - 2) * 8 + 3 + (k)] = clock64(); \
  } while (0)
#define MAVA_STAMP(k)                                                              \
  do {                                                                             \
    if (t == 0 && blockIdx.x == MAVA_STAMP_CTA && it >= 2 && it < 18) g_phase_clock[(it - 2) * 16 + (k)] = clock64(); \
  } while (0)
#else
#define MAVA_STAMP(k) do { } while (0)
#define MAVA_WSTAMP(k) do { } while (0)
#define MAVA_STAMP2(k) do { } while (0)
#define MAVA_STAMP3(k) do { } while (0)
#endif

constexpr uint32_t kTmemCols = 512;
constexpr uint32_t COL_ACC = 0, COL_HEAD = 128, COL_DW3 = 144, COL_DW2 = 160;  // dW2: 144 columns
constexpr uint32_t COL_DW1 = 304;  // folded first-layer gradient of the actor: up to 208 columns
constexpr uint32_t COL_ACC1 = 384;  // layer-1 accumulator of the NEXT tile when k1p <= 80 leaves room
constexpr uint32_t kRegionMin = tile_bytes(TM, HCOLS) + tile_bytes(TM, HID);   // H2 + dZ2

struct TrainArgs {
  NetDesc actor, critic;
  const unsigned char *actor_img, *critic_img;
  const int8_t* view;
  const uint8_t* mask;
  const int8_t* action;
  const float *old_logp, *old_value, *adv, *targets;
  const int32_t* rows;
  const double* adv_stats;
  int R, mb_size, num_replicas;
  float clip_eps, ent_coef, vf_coef;
  int actor_ctas, critic_ctas;
  int fold_actor_w1;  // the actor's [dW1^T | db1] accumulates in TMEM inside the fused kernel
  int prefetch_actor; // fold mode: next tile's observation rows are gathered one tile ahead
  // the same two switches for a critic that reads the agent's own view (ff_ippo): it then runs on the
  // actor's pipeline (loader warps, folded first-layer gradient) and needs no ppo_wgrad1_kernel pass
  int fold_critic_w1, prefetch_critic;
  int pipe_layer1;    // prefetch mode: next tile's layer-1 GEMM issued behind this tile's backward pass
  float *grad_actor, *grad_critic;
  double* loss_acc;
  unsigned char *dz1_actor, *dz1_critic;  // [tiles][TM*HID*2] tile images
};

struct Ctrl {
  // weights landed | MMA chain | layer-1 accumulator | weight-gradient MMAs drained | gathered rows.
  // One barrier per kind of event: two commits in a row on ONE barrier would let a late thread miss
  // a phase (the parity it waits for comes round again).
  uint64_t wbar, mbar, mbar1, mbar2, gbar;
  uint32_t tmem;
  float db3[NHEAD];
  float adv_mean[8], adv_isd[8];  // per replica: mean and 1 / (std + 1e-8), from the fp64 sums
  int32_t steps[3][TM + 2];      // env-step index of each minibatch position of the tiles in flight
  // loss inputs of a tile's rows, written by the loader warps.  Actor: two buffers of
  // {mask | action << 8, old log-prob, advantage}; centralised critic: ONE buffer of 8 floats per
  // row, old values [0..A) and targets [4..4+A) of the env-step's agents (A <= 4)
  uint32_t lin[2][TM][4];
};

constexpr int kMaxReps = 8;  // agents a centralised-critic row stands for (mava: num_agents <= 8)

// per-row loss inputs, fetched at the start of a tile
struct LossIn {
  bool valid;
  int j;
  int64_t flat;
  uint32_t mk;
  int act;
  float f0[kMaxReps], f1[kMaxReps];  // actor: old_logp, adv in [0]; critic: old_value, targets
};

__host__ __device__ inline uint32_t region_bytes(int k1p) {
  const uint32_t x = tile_bytes(TM, k1p);
  return x > kRegionMin ? x : kRegionMin;
}

// dH (32 accumulator columns of row r) * relu'(H) -> bf16, written either into a shared-memory tile
// or into a tile image in global memory (same core-matrix layout).  The relu derivative is read back
// from the activation tile H (bf16 > 0), so no mask has to be carried from the forward pass.
template <bool TO_GLOBAL>
__device__ __forceinline__ void grad_epilogue(const Lane& L, uint32_t tmem_acc, const Tile& h,
                                              const Tile& dst, unsigned char* gdst) {
  float v[32];
  ld32(tmem_acc + L.tmem_lane() + (uint32_t)(L.q * 32), v);
#pragma unroll
  for (int cg = 0; cg < 4; ++cg) {
    uint32_t hw[4];
    ld_shared_v4(h.base + chunk_off(h, L.r, L.q * 4 + cg), hw);
    const uint32_t a = relu_grad_bf16x2(pack_bf16(v[cg * 8], v[cg * 8 + 1]), hw[0]),
                   b = relu_grad_bf16x2(pack_bf16(v[cg * 8 + 2], v[cg * 8 + 3]), hw[1]),
                   c = relu_grad_bf16x2(pack_bf16(v[cg * 8 + 4], v[cg * 8 + 5]), hw[2]),
                   d = relu_grad_bf16x2(pack_bf16(v[cg * 8 + 6], v[cg * 8 + 7]), hw[3]);
    const uint32_t off = chunk_off(dst, L.r, L.q * 4 + cg);
    if (TO_GLOBAL) *reinterpret_cast<uint4*>(gdst + off) = make_uint4(a, b, c, d);
    else st_shared_v4(dst.base + off, a, b, c, d);
  }
}

// One row of _actor_loss_fn (ff_mappo.py:159-180): masked log-softmax, clipped ratio, entropy.
// NLD = head columns read from TMEM (8 or 16), NO = columns looked at; EXACT: NO is the action count
// (instantiated for the action spaces of the supported envs), else columns q >= nout are skipped at
// run time.  dz = d(total loss)/d(logits) * wrow, also accumulated into db3 (head bias gradient).
// exp / log are the hardware approximations (the forward pass is bf16: 2^-8 relative).
template <int NLD, int NO, bool EXACT>
__device__ __forceinline__ void actor_loss_row(uint32_t tmem_head, bool valid, int nout, uint32_t mk,
                                               int a, float old_logp, float g, float clip_eps,
                                               float ent_coef, float wrow, float (&dz)[NHEAD],
                                               float (&db3)[NHEAD], float& l0f, float& l1f) {
  float out[NLD];
  if constexpr (NLD == 8) ld8(tmem_head, out);
  else ld16(tmem_head, out);
  if (!valid) return;
  float mx = kF32Min;
#pragma unroll
  for (int q = 0; q < NO; ++q) {
    out[q] = ((mk >> q) & 1u) ? out[q] : kF32Min;
    if (EXACT || q < nout) mx = fmaxf(mx, out[q]);
  }
  float se = 0.0f, ex[NO];
#pragma unroll
  for (int q = 0; q < NO; ++q) {
    ex[q] = (EXACT || q < nout) ? __expf(out[q] - mx) : 0.0f;
    se += ex[q];
  }
  const float lse = mx + __logf(se), inv_se = __fdividef(1.0f, se);
  float la = 0.0f, ent = 0.0f, logp[NO], pr[NO];
#pragma unroll
  for (int q = 0; q < NO; ++q) {
    logp[q] = out[q] - lse;
    pr[q] = ex[q] * inv_se;
    ent -= pr[q] != 0.0f ? pr[q] * logp[q] : 0.0f;
    la = q == a ? logp[q] : la;
  }
  const float ratio = __expf(la - old_logp);
  const float lo = 1.0f - clip_eps, hi = 1.0f + clip_eps;
  const float t1 = ratio * g, t2 = fminf(fmaxf(ratio, lo), hi) * g;
  const bool inside = ratio > lo && ratio < hi;
  float dr;
  if (t1 < t2) dr = -g;
  else if (t1 > t2) dr = inside ? -g : 0.0f;
  else dr = -g * (0.5f + (inside ? 0.5f : 0.0f));
  const float dla = dr * ratio * wrow, ec = ent_coef * wrow;
#pragma unroll
  for (int q = 0; q < NO; ++q) {
    float dl = dla * ((q == a ? 1.0f : 0.0f) - pr[q]);
    dl += pr[q] != 0.0f ? ec * pr[q] * (logp[q] + ent) : 0.0f;
    dz[q] = ((EXACT || q < nout) && ((mk >> q) & 1u)) ? dl : 0.0f;
    db3[q] += dz[q];
  }
  l0f += -fminf(t1, t2);
  l1f += ent;
}

// 16 epilogue warps (thread mapping of mlp_tc.cuh) + ONE MMA-issue warp.  tcgen05.mma issue blocks
// at the execution rate of the tensor pipe (scripts/mma_rate.cu: ~68 cycles per 128x128x16 MMA plus
// ~250 cycles from the first issue to the observed completion), so a thread that issues MMAs cannot
// also finish rows: the issue warp waits on a named barrier for "operands ready", issues the GEMM the
// next epilogue needs, commits, and then issues the weight-gradient GEMMs of the phase behind the
// commit, so that they run while the epilogue warps are already working on the accumulator.
constexpr int NLOAD = 96;             // three loader warps (20 warps x 96 registers fill the file)
constexpr int NT_F = NT + 32 + NLOAD;  // 16 epilogue warps + the MMA-issue warp + the loader warps
constexpr int NT_RDY = NT + 32;        // threads on the operands-ready barrier
constexpr int BAR_READY = 2, BAR_EPI = 3, BAR_FULL = 4, BAR_EMPTY = 5, BAR_LOAD = 6, BAR_READY2 = 7;
constexpr int kLdSlots = 22;           // words per loader thread: 22 x 96 = 2112 = 128 rows x 66 bytes / 4

// The operands-ready handshake alternates between two named barriers.  A named barrier cannot tell
// generations apart: with the next tile's layer 1 already in TMEM an epilogue warp passes from
// "X ready" to "H1 stored" without waiting for anything the issue warp does, so on ONE barrier its
// second arrival could complete the first generation while a slow warp is still storing.  Between
// two arrivals on the SAME barrier there is always a wait for an MMA the issue warp issued after the
// generation in between.
// epilogue side: my shared-memory / TMEM accesses of this phase are done
__device__ __forceinline__ void epi_arrive(uint32_t& rb) {
  fence_proxy_async();
  fence_before_sync();
  asm volatile("bar.arrive %0, %1;" ::"r"(rb ? BAR_READY2 : BAR_READY), "n"(NT_RDY) : "memory");
  rb ^= 1u;
}
// issue side: all 512 epilogue threads have arrived
__device__ __forceinline__ void issuer_wait(uint32_t& rb) {
  asm volatile("bar.sync %0, %1;" ::"r"(rb ? BAR_READY2 : BAR_READY), "n"(NT_RDY) : "memory");
  rb ^= 1u;
  fence_after_sync();
}
// barrier among the 16 epilogue warps only
__device__ __forceinline__ void epi_sync() {
  asm volatile("bar.sync %0, %1;" ::"n"(BAR_EPI), "n"(NT) : "memory");
}
// every epilogue thread polls the MMA-completion barrier itself (no CTA barrier behind it)
__device__ __forceinline__ void wait_acc(uint64_t* bar, uint32_t& phase) {
  mbar_wait(bar, phase);
  phase ^= 1u;
  fence_after_sync();
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(pred));
  return pred != 0;
}


__global__ void __launch_bounds__(NT_F, 1) ppo_fused_kernel(const TrainArgs p) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ Ctrl ctrl;
  const Lane L;
  const int t = L.t, warp = L.warp, lane = L.lane;
  // warp-uniform role (the shuffle tells the compiler so: descriptors stay in uniform registers)
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  const bool issue_warp = warp_u == NWARPS;
  const bool load_warp = warp_u > NWARPS;  // loader warps: every global load of the actor's tile loop
  const bool is_actor = (int)blockIdx.x < p.actor_ctas;
  const NetDesc& d = is_actor ? p.actor : p.critic;
  const int cta = is_actor ? blockIdx.x : blockIdx.x - p.actor_ctas;
  const int n_ctas = is_actor ? p.actor_ctas : p.critic_ctas;
  const int rps = d.mode == MAVA_IN_GLOBAL ? 1 : d.A;  // tile rows per env-step
  const int M = p.R * rps;  // < 2^31 (checked by the host entry)
  const int n_tiles = ceil_div(M, TM);
  const int step_bytes = d.A * d.FR;

  // shared memory: [weights][region: X, later H2 + dZ2][H1 (also the gather staging area)][dZ3]
  // fold mode (actor): [weights][X ping][X pong][region: H2 (later dZ1) + dZ2][H1][dZ3] -- X stays
  // alive to the end of the tile so that [dW1^T | db1] += dZ1^T [X | 1] runs here, in TMEM, while the
  // next tile is gathered into the other X buffer.
  const bool fold = (is_actor ? p.fold_actor_w1 : p.fold_critic_w1) != 0;
  const WImage wi{d.k1p};
  const uint32_t s_w = smem_u32(smem);
  const uint32_t x_bytes = tile_bytes(TM, d.k1p);
  const uint32_t s_x0 = s_w + wi.total();
  const uint32_t s_region = fold ? s_x0 + 2 * x_bytes : s_x0;
  const Tile h2t{s_region, 128u, 2048u};
  const Tile dz1t{s_region, 128u, 2048u};
  const Tile dz2t{s_region + tile_bytes(TM, HCOLS), 128u, 2048u};
  const Tile h1t{s_region + (fold ? kRegionMin : region_bytes(d.k1p)), 128u, 2048u};
  const Tile dz3t{h1t.base + tile_bytes(TM, HCOLS), 128u, 2048u};
  const Tile w1 = w1_tile(s_w, d.k1p), w2 = w2_tile(s_w, d.k1p), w3 = w3_tile(s_w, d.k1p);
  // fold mode also prefetches: the next tile's observation rows are gathered into a dedicated
  // staging buffer while this tile runs, so the two dependent HBM latencies of the gather (row
  // index -> observation bytes) leave the critical path
  const bool prefetch = fold && (is_actor ? p.prefetch_actor : p.prefetch_critic) != 0;
  // Layer 1 of the next tile is issued behind this tile's last backward GEMM, into its own
  // accumulator, when TMEM has 128 columns left: a tile then starts with its layer-1 result ready.
  const bool pipe1 = prefetch && p.pipe_layer1 != 0 && d.k1p <= (int)(COL_ACC1 - COL_DW1);
  const uint32_t col_acc1 = pipe1 ? COL_ACC1 : COL_ACC;
  unsigned char* pf_stage = smem + (dz3t.base - s_w) + tile_bytes(TM, NHEAD);

  // Tile geometry without divisions on the per-tile path: when a tile holds whole env-steps (rps
  // divides 128: every supported env) the step and agent of a tile row do not depend on the tile.
  const bool aligned = (TM % rps) == 0;
  const int spt = TM / rps;
  const int r_j = L.r / rps, r_a = L.r - r_j * rps;
  auto tile_span = [&](int tile_idx, int& j0, int& nsteps) {
    const int r0 = tile_idx * TM;
    const int rows = M - r0 < TM ? M - r0 : TM;
    if (aligned) {
      j0 = tile_idx * spt;
      nsteps = rows == TM ? spt : rows / rps;
    } else {
      j0 = r0 / rps;
      nsteps = (r0 + rows - 1) / rps - j0 + 1;
    }
  };
  auto load_steps = [&](int tile_idx) -> int32_t {  // this thread's entry of a tile's index list
    int j0, nsteps;
    tile_span(tile_idx, j0, nsteps);
    return t < nsteps ? __ldg(p.rows + j0 + t) : 0;
  };
  auto publish_steps = [&](int buf, int32_t my_step) {
    if (t < TM + 2) ctrl.steps[buf][t] = my_step;
  };
  // Prefetch by the loader warps: the raw observation rows of the next tile and its per-row loss
  // inputs are fetched with plain loads by three warps that do nothing else -- they never execute a
  // proxy fence (fence.proxy.async is a MEMBAR that waits for the thread's outstanding loads, and a
  // random row costs ~3 K cycles under this kernel), so a whole tile time hides the latency.
  // Staging layout: one padded byte row per tile row, [onehot(agent) | view bytes | 1 | 0...] of k1p
  // bytes -- the int8 image of the X row.  Agent-id, ones and padding bytes do not depend on the
  // tile and are written once; the loaders only replace the view bytes (two 16-bit stores per
  // 32-bit word: FR is even, so a half word never straddles two rows).  The expansion is then the
  // same for every 8-column chunk: one 8-byte load, eight conversions, one 16-byte store.
  // Slot metadata, one register: bit 0 = the word straddles two rows, bits 1..13 = destination half
  // word, bits 14..21 = word u of the step, bits 22..27 = step js of the tile.
  const int pf_units = step_bytes >> 2;
  const int id_cols = (d.mode == MAVA_IN_AGENT_VIEW && d.add_id) ? d.A : 0;
  const uint32_t pf_gap = d.mode == MAVA_IN_GLOBAL ? 0u : (uint32_t)(d.k1p - d.FR) >> 1;
  // bf16 X rows of tile `tile_idx` from the padded staging rows (thread: row L.r, chunks cg0,
  // cg0 + cg_step, ...)
  auto expand_rows = [&](int tile_idx, const Tile& xn, int cg0, int cg_step,
                         const unsigned char* stage) {
    const bool valid = tile_idx * TM + L.r < M;
    const uint32_t src = smem_u32(stage) + (uint32_t)(L.r * d.k1p);
    expand_padded_row(xn, L, src, valid, d.k1p >> 3, cg0, cg_step);
  };

  if (warp == 0) tmem_alloc<kTmemCols>(&ctrl.tmem);
  if (t == 0) {
    mbar_init(&ctrl.wbar, 1);
    mbar_init(&ctrl.mbar, 1);
    mbar_init(&ctrl.mbar1, 1);
    mbar_init(&ctrl.mbar2, 1);
    mbar_init(&ctrl.gbar, TM);
    fence_mbar_init();
  }
  if (t < NHEAD) ctrl.db3[t] = 0.0f;
  if (t < p.num_replicas && t < 8) {
    const double cnt = (double)p.mb_size * d.A;
    const double mean_d = p.adv_stats[2 * t] / cnt;
    const double var_d = fmax(p.adv_stats[2 * t + 1] / cnt - mean_d * mean_d, 0.0);
    ctrl.adv_mean[t] = (float)mean_d;
    ctrl.adv_isd[t] = 1.0f / ((float)sqrt(var_d) + 1e-8f);
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = ctrl.tmem;
  if (t == 0) load_weights(s_w, is_actor ? p.actor_img : p.critic_img, wi.total(), &ctrl.wbar);

  // plain path, joint-observation rows (centralised critic): rows fetched by bulk copies into padded
  // staging rows (see the tile loop); `cl`: its index lists and loss inputs come from the loader warps
  const bool padded_global = !prefetch && !fold && d.mode == MAVA_IN_GLOBAL && (step_bytes & 7) == 0 &&
                             TM * grow_stride(d.k1p) <= tile_bytes(TM, HCOLS) && d.k1p <= 288 &&
                             (reinterpret_cast<size_t>(p.view) & 15) == 0;
  const bool cl = padded_global && d.A <= 4;
  float* linc = reinterpret_cast<float*>(&ctrl.lin[0][0][0]);  // [TM][8]
  uint32_t phase = 0, phase1 = 0, phase2 = 0;
  uint32_t rb = 0;  // which of the two operands-ready barriers comes next (both sides alternate)
  float l0f = 0.0f, l1f = 0.0f;
  float db3_acc[NHEAD];
#pragma unroll
  for (int q = 0; q < NHEAD; ++q) db3_acc[q] = 0.0f;
  const bool any_tile = cta < n_tiles;

  if (issue_warp) {
    // ================================ MMA-issue warp ==========================================
    mbar_wait(&ctrl.wbar, 0);
    bool first = true;
    int it = 0;
    for (int tile = cta; tile < n_tiles; tile += n_ctas, first = false, ++it) {
      const Tile xt{fold ? s_x0 + (uint32_t)(it & 1) * x_bytes : s_x0, 128u, 2048u};
      const Tile xprev{s_x0 + (uint32_t)((it + 1) & 1) * x_bytes, 128u, 2048u};
      issuer_wait(rb);  // X built; the previous tile's dZ1 stored and its accumulator drained
      if (elect_one()) {
        // previous tile: [dW1^T | db1] += dZ1^T [X | 1] (A = dZ1 and B = X MN-major), then layer 1
        if (fold && !first)
          issue_gemm(tmem + COL_DW1, dz1t, true, xprev, true, d.k1p, TM, it > 1, nullptr);
        if (!pipe1 || first)
          issue_gemm(tmem + col_acc1, xt, false, w1, true, HID, d.k1p, false, &ctrl.mbar1);
      }
      __syncwarp();
      issuer_wait(rb);  // H1 stored
      if (elect_one())
        issue_gemm(tmem + COL_ACC, h1t, false, w2, true, HID, HCOLS, false, &ctrl.mbar);
      __syncwarp();
      issuer_wait(rb);  // H2 stored
      if (elect_one())
        issue_gemm(tmem + COL_HEAD, h2t, false, w3, true, NHEAD, HCOLS, false, &ctrl.mbar);
      __syncwarp();
      issuer_wait(rb);  // dZ3 stored
      if (elect_one()) {
        // dH2 = dZ3 W3^T is what the next epilogue waits for; dW3 += H2^T dZ3 runs behind it
        issue_gemm(tmem + COL_ACC, dz3t, false, w3, false, HID, NHEAD, false, &ctrl.mbar);
        issue_gemm(tmem + COL_DW3, h2t, true, dz3t, true, NHEAD, TM, !first, nullptr);
      }
      __syncwarp();
      issuer_wait(rb);  // dZ2 stored
      if (elect_one()) {
        // dH1 = dZ2 W2^T ; [dW2^T | db2] += dZ2^T [H1 | 1] behind it.  Without the prefetch pipeline
        // the next tile's X is built over H2 / dZ2 / H1 (staging): a second commit tells the
        // epilogue warps when the weight-gradient MMAs have stopped reading them.
        issue_gemm(tmem + COL_ACC, dz2t, false, w2, false, HID, HID, false, &ctrl.mbar);
        issue_gemm(tmem + COL_DW2, dz2t, true, h1t, true, HCOLS, TM, !first,
                   prefetch ? nullptr : &ctrl.mbar2);
        if (pipe1 && tile + n_ctas < n_tiles)  // the next tile's layer 1 (its X was built in the loss phase)
          issue_gemm(tmem + col_acc1, xprev, false, w1, true, HID, d.k1p, false, &ctrl.mbar1);
      }
      __syncwarp();
    }
    if (!first && fold) {
      const Tile xprev{s_x0 + (uint32_t)((it + 1) & 1) * x_bytes, 128u, 2048u};
      issuer_wait(rb);  // the last tile's dZ1
      if (elect_one())
        issue_gemm(tmem + COL_DW1, dz1t, true, xprev, true, d.k1p, TM, it > 1, &ctrl.mbar);
      __syncwarp();
    }
  } else if (load_warp) {
    // ================================ loader warps ============================================
    if (prefetch && cta + n_ctas < n_tiles) {
      const int lt = t - NT_RDY;  // 0 .. NLOAD-1
      // Thread lt walks words [lt * kLdSlots, (lt + 1) * kLdSlots) of the tile's observation bytes
      // (steps back to back, pf_units words each).  Source (step js, word u) and destination (padded
      // row, byte b) both advance incrementally -- a step is A rows of FR bytes, rows are numbered
      // js * A + a -- so the walk needs two divisions per kernel, not per word.
      uint32_t data[kLdSlots];
      const int w0 = lt * kLdSlots;
      const int js0 = w0 / pf_units, u0 = w0 - js0 * pf_units;
      int row0w = js0, b0w = 4 * u0;
      if (d.mode != MAVA_IN_GLOBAL) {
        const int a0 = b0w / d.FR;
        b0w -= a0 * d.FR;
        row0w = js0 * d.A + a0;
      }
      const int row_bytes = d.mode == MAVA_IN_GLOBAL ? step_bytes : d.FR;
      // index lists: ctrl.steps[1] / [2] belong to the loaders (slot 0 is the prologue's)
      {
        int j0, nsteps;
        tile_span(cta + n_ctas, j0, nsteps);
        if (lt < nsteps) ctrl.steps[1][lt] = __ldg(p.rows + j0 + lt);
      }
      asm volatile("bar.sync %0, %1;" ::"n"(BAR_LOAD), "n"(NLOAD) : "memory");
      int itl = 0;
      for (int tile = cta; tile + n_ctas < n_tiles; tile += n_ctas, ++itl) {
        const int nt = tile + n_ctas;  // the tile being fetched
        const int cur = 1 + (itl & 1), nxt = 1 + ((itl + 1) & 1);
        int j0, nsteps;
        tile_span(nt, j0, nsteps);
        // the index list of the tile after it: requested first, stored last
        int32_t idx_next = 0;
        if (nt + n_ctas < n_tiles) {
          int j0n, nstepsn;
          tile_span(nt + n_ctas, j0n, nstepsn);
          if (lt < nstepsn) idx_next = __ldg(p.rows + j0n + lt);
        }
        const int total = nsteps * pf_units;
        {
          int js = js0, u = u0;
          const uint32_t* src = reinterpret_cast<const uint32_t*>(
              p.view + (size_t)ctrl.steps[cur][js] * (size_t)step_bytes);
#pragma unroll
          for (int k = 0; k < kLdSlots; ++k) {
            if (w0 + k < total) data[k] = __ldg(src + u);
            if (++u == pf_units) {
              u = 0;
              ++js;
              src = reinterpret_cast<const uint32_t*>(
                  p.view + (size_t)ctrl.steps[cur][js < TM ? js : TM - 1] * (size_t)step_bytes);
            }
          }
        }
        // loss inputs of rows lt and lt + NLOAD
        uint32_t lm[2] = {0u, 0u}, la[2] = {0u, 0u};
        float lp[2] = {0.0f, 0.0f}, ladv[2] = {0.0f, 0.0f};
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int r = lt + h * NLOAD;
          if (r < TM && nt * TM + r < M) {
            const int rj = r / rps, ra = r - rj * rps;
            const size_t flat = (size_t)ctrl.steps[cur][rj] * d.A + ra;
            if (is_actor) {
              lm[h] = p.mask[flat];
              la[h] = (uint32_t)(uint8_t)p.action[flat];
              lp[h] = p.old_logp[flat];
              ladv[h] = p.adv[flat];
            } else {  // decentralised critic on this pipeline: old value and target of the row
              lp[h] = p.old_value[flat];
              ladv[h] = p.targets[flat];
            }
          }
        }
        // the staging rows and this parity's loss-input buffer have been consumed
        asm volatile("bar.sync %0, %1;" ::"n"(BAR_EMPTY), "n"(NT + NLOAD) : "memory");
        {
          // two 16-bit stores per word: the second half word may belong to the next row
          int b = b0w;
          uint32_t dst = smem_u32(pf_stage) + (uint32_t)(row0w * d.k1p + id_cols + b0w);
#pragma unroll
          for (int k = 0; k < kLdSlots; ++k) {
            const bool on = w0 + k < total;
            if (on) asm volatile("st.shared.u16 [%0], %1;" ::"r"(dst), "r"(data[k] & 0xffffu) : "memory");
            b += 2;
            dst += 2;
            if (b == row_bytes) {
              b = 0;
              dst += (uint32_t)(d.k1p - row_bytes);
            }
            if (on) asm volatile("st.shared.u16 [%0], %1;" ::"r"(dst), "r"(data[k] >> 16) : "memory");
            b += 2;
            dst += 2;
            if (b == row_bytes) {
              b = 0;
              dst += (uint32_t)(d.k1p - row_bytes);
            }
          }
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int r = lt + h * NLOAD;
          if (r < TM) {
            ctrl.lin[(itl + 1) & 1][r][0] = lm[h] | (la[h] << 8);
            ctrl.lin[(itl + 1) & 1][r][1] = __float_as_uint(lp[h]);
            ctrl.lin[(itl + 1) & 1][r][2] = __float_as_uint(ladv[h]);
          }
        }
        if (lt < TM + 2) ctrl.steps[nxt][lt] = idx_next;
        asm volatile("bar.arrive %0, %1;" ::"n"(BAR_FULL), "n"(NT + NLOAD) : "memory");
        asm volatile("bar.sync %0, %1;" ::"n"(BAR_LOAD), "n"(NLOAD) : "memory");
      }
    }
    if (cl && cta + n_ctas < n_tiles) {
      // centralised critic: index lists (list of tile k in ctrl.steps[k % 3]) and the per-row loss
      // inputs (old values and targets of the env-step's agents) of the tiles after the first
      const int lt = t - NT_RDY;
      {
        int j0, nsteps;
        tile_span(cta + n_ctas, j0, nsteps);
        for (int r = lt; r < TM; r += NLOAD) ctrl.steps[1][r] = r < nsteps ? __ldg(p.rows + j0 + r) : 0;
      }
      asm volatile("bar.sync %0, %1;" ::"n"(BAR_LOAD), "n"(NLOAD) : "memory");
      asm volatile("bar.arrive %0, %1;" ::"n"(BAR_FULL), "n"(NT + NLOAD) : "memory");  // list of tile 1
      int itl = 0, cur = 1;
      for (int tile = cta; tile + n_ctas < n_tiles; tile += n_ctas, ++itl, cur = cur == 2 ? 0 : cur + 1) {
        const int nt = tile + n_ctas;  // the tile whose loss inputs are fetched; its list is in `cur`
        const int nxt = cur == 2 ? 0 : cur + 1;
        int32_t idx_next[2] = {0, 0};
        if (nt + n_ctas < n_tiles) {
          int j0n, nstepsn;
          tile_span(nt + n_ctas, j0n, nstepsn);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int r = lt + h * NLOAD;
            if (r < nstepsn) idx_next[h] = __ldg(p.rows + j0n + r);
          }
        }
        float f0[2][4], f1[2][4];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int r = lt + h * NLOAD;
          const bool on = r < TM && nt * TM + r < M;
          const size_t flat = on ? (size_t)ctrl.steps[cur][r] * d.A : 0;
#pragma unroll
          for (int a = 0; a < 4; ++a) {
            f0[h][a] = (on && a < d.A) ? p.old_value[flat + a] : 0.0f;
            f1[h][a] = (on && a < d.A) ? p.targets[flat + a] : 0.0f;
          }
        }
        // the previous tile's loss inputs have been read
        asm volatile("bar.sync %0, %1;" ::"n"(BAR_EMPTY), "n"(NT + NLOAD) : "memory");
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int r = lt + h * NLOAD;
          if (r < TM) {
#pragma unroll
            for (int a = 0; a < 4; ++a) {
              linc[r * 8 + a] = f0[h][a];
              linc[r * 8 + 4 + a] = f1[h][a];
            }
            ctrl.steps[nxt][r] = idx_next[h];
          }
        }
        asm volatile("bar.arrive %0, %1;" ::"n"(BAR_FULL), "n"(NT + NLOAD) : "memory");
        asm volatile("bar.sync %0, %1;" ::"n"(BAR_LOAD), "n"(NLOAD) : "memory");
      }
    }
  } else {
    // ================================ epilogue warps ==========================================
    // Prefetch pipeline (fold mode): while tile i is in its loss epilogue (four warps busy), the
    // other twelve warps expand the raw rows of tile i+1 -- requested during tile i's first GEMM --
    // into the other X buffer.  Prologue: tile 0 is built by everybody.
    if (prefetch && any_tile) {
      publish_steps(0, load_steps(cta));
      // tile-invariant bytes of the padded rows (aligned tiles: the agent of a row is r % A)
      for (int i = t; i < TM * (d.k1p >> 2); i += NT) reinterpret_cast<uint32_t*>(pf_stage)[i] = 0u;
      epi_sync();
      if (t < TM) {
        if (id_cols) pf_stage[t * d.k1p + r_a] = 1;
        pf_stage[t * d.k1p + d.in_dim] = 1;
      }
      {
        int j0, nsteps;
        tile_span(cta, j0, nsteps);
        const int halves = step_bytes >> 1, total = nsteps * halves;
        for (int idx = t; idx < total; idx += NT) {
          const int js = idx / halves, hw = idx - js * halves;
          int row = js, b0 = 2 * hw;
          if (d.mode != MAVA_IN_GLOBAL) {
            const int a0 = b0 / d.FR;
            b0 -= a0 * d.FR;
            row = js * d.A + a0;
          }
          *reinterpret_cast<uint16_t*>(pf_stage + row * d.k1p + id_cols + b0) = __ldg(
              reinterpret_cast<const uint16_t*>(p.view + (size_t)ctrl.steps[0][js] * (size_t)step_bytes) + hw);
        }
      }
      epi_sync();
      expand_rows(cta, Tile{s_x0, 128u, 2048u}, L.q, 4, pf_stage);
      if (L.q == 0 && cta * TM + L.r < M) {  // tile 0's loss inputs (later tiles: the loader warps)
        const size_t flat = (size_t)ctrl.steps[0][r_j] * d.A + r_a;
        if (is_actor) {
          ctrl.lin[0][L.r][0] = (uint32_t)p.mask[flat] | ((uint32_t)(uint8_t)p.action[flat] << 8);
          ctrl.lin[0][L.r][1] = __float_as_uint(p.old_logp[flat]);
          ctrl.lin[0][L.r][2] = __float_as_uint(p.adv[flat]);
        } else {
          ctrl.lin[0][L.r][0] = 0u;
          ctrl.lin[0][L.r][1] = __float_as_uint(p.old_value[flat]);
          ctrl.lin[0][L.r][2] = __float_as_uint(p.targets[flat]);
        }
      }
      epi_sync();
      // the staging rows are free: the loaders may bring in tile 1
      if (cta + n_ctas < n_tiles)
        asm volatile("bar.arrive %0, %1;" ::"n"(BAR_EMPTY), "n"(NT + NLOAD) : "memory");
    }
    if (cl && any_tile) {
      publish_steps(0, load_steps(cta));
      epi_sync();
      if (L.q == 0 && cta * TM + L.r < M) {  // tile 0's loss inputs (later tiles: the loader warps)
        const size_t flat = (size_t)ctrl.steps[0][L.r] * d.A;
        for (int a = 0; a < d.A; ++a) {
          linc[L.r * 8 + a] = p.old_value[flat + a];
          linc[L.r * 8 + 4 + a] = p.targets[flat + a];
        }
      }
      epi_sync();
    }
    mbar_wait(&ctrl.wbar, 0);
    const float wrow = 1.0f / ((float)p.R * (float)d.A);  // mean over rows, replicas and agents
    bool first = true;
    int32_t next_step = 0;
    uint32_t gphase = 0;
    // bulk copies are issued one lane at a time: eight rows per warp, spread over all 16 warps
    const bool grow_thread = lane < TM / NWARPS;
    const int grow_r = warp * (TM / NWARPS) + lane;
    int it = 0, sb = 0;  // sb = it % 3: slot of ctrl.steps holding this tile's index list
    for (int tile = cta; tile < n_tiles; tile += n_ctas, first = false, ++it, sb = sb == 2 ? 0 : sb + 1) {
      const int row0 = tile * TM;
      const Tile xt{fold ? s_x0 + (uint32_t)(it & 1) * x_bytes : s_x0, 128u, 2048u};
      MAVA_STAMP(0);
#ifdef MAVA_PROFILE_PHASES
      if (t == 0 && blockIdx.x == gridDim.x - 1 && it < 16) g_phase_clock2[128 + it] = clock64();
#endif
      const bool has_next = prefetch && tile + n_ctas < n_tiles;
      const bool has_next2 = prefetch && tile + 2 * n_ctas < n_tiles;
      if (!prefetch) {
        // the previous tile's weight-gradient MMAs still read the buffers X is built over
        if (!first) wait_acc(&ctrl.mbar2, phase2);
        MAVA_STAMP3(0);
        if (padded_global) {
          // joint-observation rows (centralised critic): a row is one whole env-step.  Its bytes were
          // fetched by bulk copies started during the previous tile's backward pass (first tile:
          // here) into the H2 part of the region X is built over, so the expansion goes through
          // registers: every thread reads its chunks, then all of them write.
          unsigned char* gstg = smem + (h2t.base - s_w);
          int j0, nsteps;
          tile_span(tile, j0, nsteps);
          const int ls = cl ? sb : (it & 1);  // slot of this tile's index list
          if (first && grow_thread) {
            const int32_t st = cl ? ctrl.steps[0][grow_r] : (grow_r < nsteps ? __ldg(p.rows + j0 + grow_r) : 0);
            if (!cl) ctrl.steps[0][grow_r] = st;
            grow_issue(d, p.view, st, grow_r < nsteps, gstg, grow_r, &ctrl.gbar);
          }
          if (!cl && tile + n_ctas < n_tiles && grow_thread) {
            int j0n, nstepsn;
            tile_span(tile + n_ctas, j0n, nstepsn);
            next_step = grow_r < nstepsn ? __ldg(p.rows + j0n + grow_r) : 0;
          }
          mbar_wait(&ctrl.gbar, gphase);
          gphase ^= 1u;
          MAVA_STAMP3(1);
          if (t < nsteps) grow_tail(d, gstg, t, ctrl.steps[ls][t]);
          epi_sync();
          MAVA_STAMP3(2);
          constexpr int kMaxChunks = 9;  // per thread: k1p <= 288
          uint32_t w[kMaxChunks][2];
          const bool valid = row0 + L.r < M;
          const uint32_t src = smem_u32(gstg) + (uint32_t)L.r * grow_stride(d.k1p) +
                               (valid ? grow_skew(d, ctrl.steps[ls][L.r]) : 0u);
          const int nchunks = d.k1p >> 3;
          // chunk residue rotated with the row so that the 8-byte loads of a half warp (rows 288 bytes
          // apart) fall into sixteen different bank pairs
          const int q_rot = (L.q + (L.r >> 2)) & 3;
#pragma unroll
          for (int i = 0; i < kMaxChunks; ++i) {
            const int cg = q_rot + 4 * i;
            w[i][0] = w[i][1] = 0u;
            if (valid && cg < nchunks)
              asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(w[i][0]), "=r"(w[i][1]) : "r"(src + 8u * cg));
          }
          epi_sync();  // staging has been read: X may overwrite it
          MAVA_STAMP3(3);
#pragma unroll
          for (int i = 0; i < kMaxChunks; ++i) {
            const int cg = q_rot + 4 * i;
            if (cg < nchunks)
              st_shared_v4(xt.base + chunk_off(xt, L.r, cg), s8x2_bf16x2(w[i][0]),
                           s8x2_bf16x2(w[i][0] >> 16), s8x2_bf16x2(w[i][1]), s8x2_bf16x2(w[i][1] >> 16));
          }
          MAVA_STAMP3(4);
        } else {
          // (H1 is not live before layer 1: it doubles as the staging area)
          build_x_tile<BAR_EPI>(d, p.view, xt, smem + (h1t.base - s_w), row0, M,
                                [&](int64_t jj) { return (int64_t)__ldg(p.rows + jj); });
        }
      }
      epi_arrive(rb);  // -> layer 1 (and the previous tile's first-layer gradient)
      MAVA_STAMP(1);
      MAVA_STAMP(14);
      // loss inputs of this row.  Prefetch path: the loader warps put them into ctrl.lin; plain
      // paths: fetched here, in flight during the forward pass
      LossIn li{};
      float g_norm_adv = 0.0f;  // (advantage - mean) / (std + 1e-8) of this row's replica (actor)
      // this tile's loss inputs were published by the loader warps during the previous tile (its
      // BAR_FULL): they are unpacked, and the advantage normalised, while layer 2 is in the tensor
      // pipe -- not here, where layer 1 is already waiting, nor between the head GEMM and dZ3
      auto unpack_loss_inputs = [&]() {
        li.valid = row0 + L.r < M;
        li.j = li.valid ? tile * spt + r_j : 0;
        const uint32_t w0 = ctrl.lin[it & 1][L.r][0];
        li.mk = w0 & 0xffu;
        li.act = (int)(signed char)(w0 >> 8);
        li.f0[0] = __uint_as_float(ctrl.lin[it & 1][L.r][1]);
        li.f1[0] = __uint_as_float(ctrl.lin[it & 1][L.r][2]);
      };
      auto normalise_adv = [&]() {
        int u = 0;  // replica of this minibatch position (at most 8: no division)
#pragma unroll
        for (int k = 1; k < 8; ++k) u += (k < p.num_replicas && li.j >= k * p.mb_size) ? 1 : 0;
        g_norm_adv = (li.f1[0] - ctrl.adv_mean[u]) * ctrl.adv_isd[u];
      };
      if (L.q == 0 && prefetch) {
        // (below, behind the layer-2 arrive)
      } else if (L.q == 0 && cl) {
        li.valid = row0 + L.r < M;
        li.j = li.valid ? row0 + L.r : 0;
      } else if (L.q == 0) {
        const int row = row0 + L.r;
        li.valid = row < M;
        int j, ag;
        if (aligned) {
          j = tile * spt + r_j;
          ag = r_a;
        } else {
          j = row / rps;
          ag = row - j * rps;
        }
        if (!li.valid) j = 0;
        const int j0 = aligned ? tile * spt : row0 / rps;
        // (both staged paths keep the tile's index list in shared memory: no dependent HBM access)
        const int64_t sidx = !li.valid      ? 0
                             : prefetch      ? (int64_t)ctrl.steps[sb][j - j0]
                             : padded_global ? (int64_t)ctrl.steps[it & 1][L.r]
                                             : (int64_t)__ldg(p.rows + j);
        li.j = j;
        li.flat = sidx * d.A + (d.mode == MAVA_IN_GLOBAL ? 0 : ag);
        if (li.valid) {
          if (is_actor) {
            li.mk = p.mask[li.flat];
            li.act = p.action[li.flat];
            li.f0[0] = p.old_logp[li.flat];
            li.f1[0] = p.adv[li.flat];
          } else {
            const int reps = d.mode == MAVA_IN_GLOBAL ? d.A : 1;
#pragma unroll
            for (int a = 0; a < kMaxReps; ++a) {
              if (a < reps) {
                li.f0[a] = p.old_value[li.flat + a];
                li.f1[a] = p.targets[li.flat + a];
              }
            }
          }
        }
      }
      if (L.q == 0 && is_actor && !prefetch) normalise_adv();
      MAVA_STAMP(2);
      // ---- forward
      wait_acc(&ctrl.mbar1, phase1);
      MAVA_STAMP(3);
      hidden_epilogue(L, tmem + col_acc1, h1t);
      MAVA_STAMP(4);
      epi_arrive(rb);  // -> layer 2
      MAVA_STAMP(5);
      if (L.q == 0 && prefetch) {
        unpack_loss_inputs();
        if (is_actor) normalise_adv();
      }
      wait_acc(&ctrl.mbar, phase);
      hidden_epilogue(L, tmem + COL_ACC, h2t);  // X is dead (not folded): H2 replaces it
      epi_arrive(rb);  // -> head
      MAVA_STAMP(6);
      wait_acc(&ctrl.mbar, phase);
      MAVA_STAMP(7);
      // ---- loss epilogue (first four warps, one thread per row): d(total loss)/d(head output);
      //      the other twelve warps build the next tile's X meanwhile
      // the loaders have filled the staging rows (tile i+1) and the loss inputs (tile i+1)
      if (has_next) asm volatile("bar.sync %0, %1;" ::"n"(BAR_FULL), "n"(NT + NLOAD) : "memory");
      // centralised critic: the loaders' previous round (this tile's loss inputs, the next tile's
      // index list); the round before the first tile only carries the list of tile 1
      const bool c_next = cl && tile + n_ctas < n_tiles;
      if (cl && (it > 0 || c_next))
        asm volatile("bar.sync %0, %1;" ::"n"(BAR_FULL), "n"(NT + NLOAD) : "memory");
      if (L.q != 0) {
        MAVA_STAMP2(0);
        if (has_next) {
          MAVA_STAMP2(1);
          const Tile xn{s_x0 + (uint32_t)((it + 1) & 1) * x_bytes, 128u, 2048u};
          expand_rows(tile + n_ctas, xn, L.q - 1, 3, pf_stage);
          MAVA_STAMP2(2);
        }
      } else {
        const bool valid = li.valid;
        float dz[NHEAD];
#pragma unroll
        for (int q = 0; q < NHEAD; ++q) dz[q] = 0.0f;
        if (!is_actor) {
          // _critic_loss_fn, ff_mappo.py:190-201
          float out[8];
          ld8(tmem + L.tmem_lane() + COL_HEAD, out);
          if (valid) {
            const float v = out[0];
            const int reps = d.mode == MAVA_IN_GLOBAL ? d.A : 1;
            float dv = 0.0f;
            if (cl) {
#pragma unroll
              for (int a = 0; a < 4; ++a) {
                li.f0[a] = linc[L.r * 8 + a];
                li.f1[a] = linc[L.r * 8 + 4 + a];
              }
            }
#pragma unroll
            for (int a = 0; a < kMaxReps; ++a) {
              if (a >= reps) break;
              const float vo = li.f0[a], tg = li.f1[a];
              const float diff = v - vo;
              const float vc = vo + fminf(fmaxf(diff, -p.clip_eps), p.clip_eps);
              const float e1 = v - tg, e2 = vc - tg;
              const float a1 = e1 * e1, a2 = e2 * e2;
              const bool inside = diff > -p.clip_eps && diff < p.clip_eps;
              float g;
              if (a1 > a2) g = e1;
              else if (a2 > a1) g = inside ? e2 : 0.0f;
              else g = 0.5f * e1 + (inside ? 0.5f * e2 : 0.0f);
              dv += g;
              l0f += 0.5f * fmaxf(a1, a2);
            }
            dz[0] = dv * wrow * p.vf_coef;
            db3_acc[0] += dz[0];
          }
        } else {
          const float g = g_norm_adv;
          const uint32_t th = tmem + L.tmem_lane() + COL_HEAD;
#define MAVA_LOSS_ARGS th, valid, d.out, li.mk, li.act, li.f0[0], g, p.clip_eps, p.ent_coef, wrow, dz, db3_acc, l0f, l1f
          if (d.out == 5) actor_loss_row<8, 5, true>(MAVA_LOSS_ARGS);        // RobotWarehouse
          else if (d.out == 6) actor_loss_row<8, 6, true>(MAVA_LOSS_ARGS);   // LevelBasedForaging
          else if (d.out <= 8) actor_loss_row<8, 8, false>(MAVA_LOSS_ARGS);
          else actor_loss_row<16, 16, false>(MAVA_LOSS_ARGS);
#undef MAVA_LOSS_ARGS
        }
        st_shared_v4(dz3t.base + chunk_off(dz3t, L.r, 0), pack_bf16(dz[0], dz[1]),
                     pack_bf16(dz[2], dz[3]), pack_bf16(dz[4], dz[5]), pack_bf16(dz[6], dz[7]));
        st_shared_v4(dz3t.base + chunk_off(dz3t, L.r, 1), pack_bf16(dz[8], dz[9]),
                     pack_bf16(dz[10], dz[11]), pack_bf16(dz[12], dz[13]), pack_bf16(dz[14], dz[15]));
        MAVA_STAMP(15);
        // (head bias gradient: column sums of dZ3, kept per thread across tiles, reduced at the end)
      }
      // staging rows expanded, loss inputs read: the loaders may bring in tile i+2
      if (has_next2 || c_next) asm volatile("bar.arrive %0, %1;" ::"n"(BAR_EMPTY), "n"(NT + NLOAD) : "memory");
      epi_arrive(rb);  // -> backward through the head
      MAVA_STAMP(8);
      wait_acc(&ctrl.mbar, phase);
      MAVA_STAMP(9);
      grad_epilogue<false>(L, tmem + COL_ACC, h2t, dz2t, nullptr);  // dZ2 = dH2 * relu'(layer 2)
      epi_arrive(rb);  // -> dH1, dW2
      MAVA_STAMP(10);
      wait_acc(&ctrl.mbar, phase);
      MAVA_STAMP(11);
      if (padded_global && tile + n_ctas < n_tiles && grow_thread) {
        // H2 is dead (dW3 completed before dH1): the next tile's rows start arriving there now
        int j0n, nstepsn;
        tile_span(tile + n_ctas, j0n, nstepsn);
        const int sb1 = sb == 2 ? 0 : sb + 1;
        if (!cl) ctrl.steps[(it + 1) & 1][grow_r] = next_step;
        const int32_t st_next = cl ? ctrl.steps[sb1][grow_r] : next_step;
        grow_issue(d, p.view, st_next, grow_r < nstepsn, smem + (h2t.base - s_w), grow_r, &ctrl.gbar);
      }
      if (fold) {
        // dZ1 stays on chip (H2 is dead: the dW3 MMAs completed before dH1): the first-layer gradient
        // is issued with the next tile's layer 1
        grad_epilogue<false>(L, tmem + COL_ACC, h1t, dz1t, nullptr);
      } else {
        unsigned char* gdst = p.dz1_critic + (size_t)tile * tile_bytes(TM, HID);
        if (is_actor) gdst = p.dz1_actor + (size_t)tile * tile_bytes(TM, HID);
        const Tile gimg{0u, 128u, 2048u};
        grad_epilogue<true>(L, tmem + COL_ACC, h1t, gimg, gdst);  // dZ1 tile image for the wgrad1 kernel
      }
      MAVA_STAMP(12);
    }
    if (!first) {
      // everything still in the tensor pipe: the last tile's weight-gradient MMAs
      if (!prefetch) wait_acc(&ctrl.mbar2, phase2);  // dW2 / dW3
      if (fold) {                                  // dW1
        epi_arrive(rb);
        wait_acc(&ctrl.mbar, phase);
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();

  // head bias gradient: one reduction per CTA
  if (!issue_warp && L.q == 0) {
#pragma unroll
    for (int q = 0; q < NHEAD; ++q) {
      if (q < d.out) {
        float sum = db3_acc[q];
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (lane == 0) atomicAdd(&ctrl.db3[q], sum);
      }
    }
  }
  __syncthreads();
  // ---- flush: TMEM weight-gradient accumulators -> global fp32 gradient (atomics)
  float* g = is_actor ? p.grad_actor : p.grad_critic;
  float* gw2 = g + (size_t)d.in_dim * HID + HID;
  float* gb2 = gw2 + (size_t)HID * HID;
  float* gw3 = gb2 + HID;
  float* gb3 = gw3 + (size_t)HID * d.out;
  if (any_tile && warp_u < NWARPS) {  // the sixteen epilogue warps
    // dW2^T: TMEM lane = output unit n, columns = input unit k (column HID = bias gradient)
    {
      float v[32];
      ld32(tmem + L.tmem_lane() + COL_DW2 + (uint32_t)(L.q * 32), v);
#pragma unroll
      for (int c = 0; c < 32; ++c) atomicAdd(gw2 + (size_t)(L.q * 32 + c) * HID + L.r, v[c]);
    }
    if (L.q == 0) {
      float v[NHEAD];
      ld16(tmem + L.tmem_lane() + COL_DW2 + HID, v);
      atomicAdd(gb2 + L.r, v[0]);
      // dW3: TMEM lane = hidden unit k, columns = outputs
      ld16(tmem + L.tmem_lane() + COL_DW3, v);
#pragma unroll
      for (int q = 0; q < NHEAD; ++q)
        if (q < d.out) atomicAdd(gw3 + (size_t)L.r * d.out + q, v[q]);
    }
    if (t < d.out) atomicAdd(gb3 + t, ctrl.db3[t]);
    if (fold) {
      // dW1^T: TMEM lane = hidden unit n, columns = input feature k (column in_dim = db1)
      float* gb1 = g + (size_t)d.in_dim * HID;
      for (int c0 = L.q * 16; c0 < d.k1p; c0 += 64) {
        float v[16];
        ld16(tmem + L.tmem_lane() + COL_DW1 + (uint32_t)c0, v);
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          const int k = c0 + c;
          if (k < d.in_dim) atomicAdd(g + (size_t)k * HID + L.r, v[c]);
          else if (k == d.in_dim) atomicAdd(gb1 + L.r, v[c]);
        }
      }
    }
  }
  // loss sums
  double l0 = (double)l0f, l1 = (double)l1f;
  for (int o = 16; o > 0; o >>= 1) {
    l0 += __shfl_xor_sync(0xffffffffu, l0, o);
    l1 += __shfl_xor_sync(0xffffffffu, l1, o);
  }
  if (lane == 0 && L.q == 0 && !issue_warp) {
    if (is_actor) {
      atomicAdd(p.loss_acc + 0, l0);
      atomicAdd(p.loss_acc + 1, l1);
    } else {
      atomicAdd(p.loss_acc + 2, l0);
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<kTmemCols>(tmem);
}

// ------------------------------------------------------------------------------------------------
// first-layer weight gradient: [dW1^T | db1] += dZ1^T [X | 1]
// ------------------------------------------------------------------------------------------------
struct Wg1Ctrl {
  uint64_t lbar[2], mbar, gbar[2];  // dZ1 tile images landed | MMAs done | gathered rows landed
  uint64_t hbar[2];                  // MMAs of column half 0 / half 1 of X done (bulk-copied rows path)
  uint32_t tmem;
  int32_t steps[3][TM];  // env-step index of each row of the tile being built / the two being fetched
};

__global__ void __launch_bounds__(NT, 1) ppo_wgrad1_kernel(const TrainArgs p) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ Wg1Ctrl ctrl;
  const Lane L;
  const int t = L.t, warp = L.warp;
  const bool is_actor = (int)blockIdx.x < p.actor_ctas;
  const NetDesc& d = is_actor ? p.actor : p.critic;
  const int cta = is_actor ? blockIdx.x : blockIdx.x - p.actor_ctas;
  const int n_ctas = is_actor ? p.actor_ctas : p.critic_ctas;
  const int rows_per_step = d.mode == MAVA_IN_GLOBAL ? 1 : d.A;
  const int M = p.R * rows_per_step;
  const int n_tiles = ceil_div(M, TM);
  const unsigned char* dz1 = is_actor ? p.dz1_actor : p.dz1_critic;

  // shared memory: [dZ1 tile x 2][X tile][gather staging x 2 (one when the rows are not bulk-copied)]
  const uint32_t s0 = smem_u32(smem);
  const uint32_t dz_bytes = tile_bytes(TM, HID);
  const Tile xt{s0 + 2 * dz_bytes, 128u, 2048u};
  unsigned char* stage0 = smem + 2 * dz_bytes + tile_bytes(TM, d.k1p);
  const uint32_t stage_bytes_g = (TM * grow_stride(d.k1p) + 127u) & ~127u;
  if (warp == 0) tmem_alloc<kTmemCols>(&ctrl.tmem);
  if (t == 0) {
    mbar_init(&ctrl.lbar[0], 1);
    mbar_init(&ctrl.lbar[1], 1);
    mbar_init(&ctrl.mbar, 1);
    mbar_init(&ctrl.gbar[0], TM);
    mbar_init(&ctrl.gbar[1], TM);
    mbar_init(&ctrl.hbar[0], 1);
    mbar_init(&ctrl.hbar[1], 1);
    fence_mbar_init();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = ctrl.tmem;
  const int n_lo = d.k1p > 256 ? 256 : d.k1p, n_hi = d.k1p - n_lo;
  // bulk-copied rows path: X is expanded and multiplied in two column halves, so that the MMAs of one
  // half run under the int8 -> bf16 expansion of the other (the expansion, 2.6 K cycles per tile, is
  // what bounds this kernel; the MMAs, 1.5 K, used to follow it)
  const int c_h0 = d.k1p >= 64 ? pad16(d.k1p / 2) : d.k1p, c_h1 = d.k1p - c_h0;
  uint32_t phase = 0;
  bool first = true;
  // one env-step per row (centralised critic): rows fetched by bulk copies TWO tiles ahead into two
  // staging buffers (a tile's rows are requested as soon as the buffer of the tile two before it
  // has been expanded: a whole tile time hides the HBM latency of the random rows), padded staging
  // rows, uniform expansion (mlp_tc.cuh)
  const bool padded_global = d.mode == MAVA_IN_GLOBAL && (d.in_dim & 7) == 0 &&
                             (reinterpret_cast<size_t>(p.view) & 15) == 0;
  auto load_dz1 = [&](int tile_idx, int buf) {  // one thread: tile image -> buffer `buf`
    mbar_expect_tx(&ctrl.lbar[buf], dz_bytes);
    bulk_g2s(s0 + (uint32_t)buf * dz_bytes, dz1 + (size_t)tile_idx * dz_bytes, dz_bytes, &ctrl.lbar[buf]);
  };
  // bulk copies are issued one lane at a time: eight rows per warp, spread over all 16 warps
  const bool grow_thread = L.lane < TM / NWARPS;
  const int grow_r = warp * (TM / NWARPS) + L.lane;
  auto step_of = [&](int tile_idx) -> int32_t {
    const int r = tile_idx * TM + grow_r;
    return (grow_thread && r < M) ? __ldg(p.rows + r) : 0;
  };
  // request the rows of the k-th tile of this CTA (tile index tile_idx) into staging buffer k & 1;
  // its index list goes to ctrl.steps[k % 3]
  auto request_rows = [&](int k, int tile_idx, int32_t st) {
    if (grow_thread) {
      ctrl.steps[k % 3][grow_r] = st;
      grow_issue(d, p.view, st, tile_idx * TM + grow_r < M, stage0 + (size_t)(k & 1) * stage_bytes_g,
                 grow_r, &ctrl.gbar[k & 1]);
    }
  };
  if (padded_global && cta < n_tiles) {
    request_rows(0, cta, step_of(cta));
    if (cta + n_ctas < n_tiles) request_rows(1, cta + n_ctas, step_of(cta + n_ctas));
    if (t == 0) load_dz1(cta, 0);
  }
  int it = 0;
  for (int tile = cta; tile < n_tiles; tile += n_ctas, first = false, ++it) {
    const int row0 = tile * TM;
    const int buf = padded_global ? (it & 1) : 0;
    const Tile dzt{s0 + (uint32_t)buf * dz_bytes, 128u, 2048u};
    if (padded_global) {
      const int nxt = tile + n_ctas, nxt2 = tile + 2 * n_ctas;
      unsigned char* stage = stage0 + (size_t)(it & 1) * stage_bytes_g;
      const uint32_t hpar = (uint32_t)(it + 1) & 1u;  // parity of the previous tile's half commits
      // the MMAs of a column half: D[n][k] += sum_rows dZ1[row][n] * X[row][k], A = dZ1 and B = X both
      // MN-major; called by ONE elected thread
      auto issue_half = [&](int half) {
        const int c0 = half ? c_h0 : 0, nc = half ? c_h1 : c_h0;
        if (nc > 0) {
          const uint32_t idesc = instr_desc(TM, nc, true, true);
          for (int k = 0; k < TM / 16; ++k)
            mma(tmem + (uint32_t)c0, desc_mnmajor(dzt, k), desc_mnmajor(xt, k, c0), idesc,
                !first || k > 0);
        }
        commit(&ctrl.hbar[half]);
      };
      MAVA_WSTAMP(0);
      if (t == 0 && nxt < n_tiles) {
        // the other dZ1 buffer was read by the MMAs of tile it - 1 (both halves)
        if (it > 0) mbar_wait(&ctrl.hbar[1], hpar);
        load_dz1(nxt, buf ^ 1);
      }
      const int32_t step2 = nxt2 < n_tiles ? step_of(nxt2) : 0;  // index list two tiles ahead
      mbar_wait(&ctrl.gbar[it & 1], (uint32_t)(it >> 1) & 1u);
      MAVA_WSTAMP(1);
      if (t < TM && row0 + t < M) grow_tail(d, stage, t, ctrl.steps[it % 3][t]);
      if (it > 0) mbar_wait(&ctrl.hbar[0], hpar);  // X columns [0, c_h0) are free again
      __syncthreads();
      MAVA_WSTAMP(2);
      const bool valid = row0 + L.r < M;
      const uint32_t srow = smem_u32(stage) + (uint32_t)L.r * grow_stride(d.k1p) +
                            (valid ? grow_skew(d, ctrl.steps[it % 3][L.r]) : 0u);
      const int rot = (L.q + (L.r >> 2)) & 3;  // rotation: bank spread of the 8-byte staging reads
      // ---- column half 0
      expand_padded_row(xt, L, srow, valid, c_h0 >> 3, rot, 4);
      MAVA_WSTAMP(3);
      fence_proxy_async();
      fence_before_sync();
      __syncthreads();
      if (mma_issuer()) {
        fence_after_sync();
        mbar_wait(&ctrl.lbar[buf], (uint32_t)(it >> 1) & 1u);  // this tile's dZ1 image
        issue_half(0);
      }
      MAVA_WSTAMP(4);
      // ---- column half 1 (under the MMAs of half 0)
      if (c_h1 > 0) {
        if (it > 0) mbar_wait(&ctrl.hbar[1], hpar);  // X columns [c_h0, k1p) are free again
        const Tile xh{xt.base + (uint32_t)(c_h0 >> 3) * xt.s_c, xt.s_r, xt.s_c};
        expand_padded_row(xh, L, srow + (uint32_t)c_h0, valid, c_h1 >> 3, rot, 4);
      }
      MAVA_WSTAMP(5);
      fence_proxy_async();
      fence_before_sync();
      __syncthreads();  // this staging buffer has been read: the rows of tile it + 2 may land in it
      MAVA_WSTAMP(6);
      if (mma_issuer()) {
        fence_after_sync();
        issue_half(1);
      }
      if (nxt2 < n_tiles) request_rows(it + 2, nxt2, step2);
      MAVA_WSTAMP(7);
      continue;
    } else {
      if (t == 0) load_dz1(tile, 0);
      build_x_tile(d, p.view, xt, stage0, row0, M,
                   [&](int64_t jj) { return (int64_t)__ldg(p.rows + jj); });
      fence_proxy_async();
      mbar_wait(&ctrl.lbar[0], phase);
    }
    fence_before_sync();
    __syncthreads();
    if (mma_issuer()) {
      fence_after_sync();
      // D[n][k] += sum_rows dZ1[row][n] * X[row][k]: A = dZ1 (MN-major), B = X (MN-major)
      const uint32_t idesc_lo = instr_desc(TM, n_lo, true, true);
      for (int k = 0; k < TM / 16; ++k)
        mma(tmem, desc_mnmajor(dzt, k), desc_mnmajor(xt, k), idesc_lo, !first || k > 0);
      if (n_hi > 0) {
        const uint32_t idesc_hi = instr_desc(TM, n_hi, true, true);
        for (int k = 0; k < TM / 16; ++k)
          mma(tmem + 256, desc_mnmajor(dzt, k), desc_mnmajor(xt, k, 256), idesc_hi, !first || k > 0);
      }
      commit(&ctrl.mbar);
    }
    MAVA_WSTAMP(7);
    wait_mma(&ctrl.mbar, phase);  // operands are free again once the MMAs have completed
    MAVA_WSTAMP(8);
    phase ^= 1;
  }
  if (padded_global && !first) {  // the last tile's MMAs (commits complete in issue order)
    mbar_wait(&ctrl.hbar[1], (uint32_t)(it + 1) & 1u);
  }
#ifdef MAVA_PROFILE_PHASES
  if (t == 0 && blockIdx.x == 0) g_wg1_clock[7 * 12 + 10] = clock64();
#endif
  if (!first) {
    float* g = is_actor ? p.grad_actor : p.grad_critic;
    float* gb1 = g + (size_t)d.in_dim * HID;
    fence_after_sync();
    // TMEM lane = hidden unit n, columns = input feature k (column in_dim = bias gradient)
    for (int c0 = L.q * 16; c0 < d.k1p; c0 += 64) {
      float v[16];
      ld16(tmem + L.tmem_lane() + (uint32_t)c0, v);
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        const int k = c0 + c;
        if (k < d.in_dim) atomicAdd(g + (size_t)k * HID + L.r, v[c]);
        else if (k == d.in_dim) atomicAdd(gb1 + L.r, v[c]);
      }
    }
  }
  fence_before_sync();
  __syncthreads();
#ifdef MAVA_PROFILE_PHASES
  if (t == 0 && blockIdx.x == 0) g_wg1_clock[7 * 12 + 11] = clock64();
#endif
  if (warp == 0) tmem_dealloc<kTmemCols>(tmem);
}

int64_t tile_count(const mava_mlp_desc* d, int rows_total) {
  const int64_t m = (int64_t)rows_total * (d->input_mode == MAVA_IN_GLOBAL ? 1 : d->num_agents);
  return ceil_div64(m, TM);
}

}  // namespace
}  // namespace tcmlp
}  // namespace mava

using namespace mava;
using namespace mava::tcmlp;

extern "C" {

#ifdef MAVA_PROFILE_PHASES
int mava_debug_phases(long long* out_host) {
  return (int)cudaMemcpyFromSymbol(out_host, g_phase_clock, sizeof(long long) * 256);
}
int mava_debug_wg1(long long* out_host) {
  return (int)cudaMemcpyFromSymbol(out_host, g_wg1_clock, sizeof(long long) * 96);
}
int mava_debug_phases2(long long* out_host) {
  return (int)cudaMemcpyFromSymbol(out_host, g_phase_clock2, sizeof(long long) * 160);
}
#endif

int64_t mava_ppo_workspace_bytes_bf16(const mava_mlp_desc* actor, const mava_mlp_desc* critic,
                                      int rows_total) {
  if (!actor || !critic || rows_total <= 0) return -1;
  return 256 + (tile_count(actor, rows_total) + tile_count(critic, rows_total)) *
                   (int64_t)tile_bytes(TM, HID);
}

static int ppo_loss_grad_bf16_impl(const mava_mlp_desc* actor, const float* actor_params,
                                   const void* actor_image, const mava_mlp_desc* critic,
                                   const float* critic_params, const void* critic_image,
                                   const mava_ppo_hyper* hyper, const int8_t* view,
                                   const uint8_t* mask, const int8_t* action, const float* old_logp,
                                   const float* old_value, const float* adv, const float* targets,
                                   const int32_t* rows, int num_replicas, int mb_size,
                                   float* grad_out, void* workspace, const double* adv_stats,
                                   mava_stream_t stream, bool acc = false) {
  TrainArgs a{};
  int rc = make_net(actor, actor_params, &a.actor);
  if (rc) return rc;
  rc = make_net(critic, critic_params, &a.critic);
  if (rc) return rc;
  MAVA_CHECK_PTR(hyper);
  MAVA_CHECK_PTR(actor_params);
  MAVA_CHECK_PTR(critic_params);
  MAVA_CHECK_PTR(actor_image);
  MAVA_CHECK_PTR(critic_image);
  MAVA_CHECK_PTR(view);
  MAVA_CHECK_PTR(mask);
  MAVA_CHECK_PTR(action);
  MAVA_CHECK_PTR(old_logp);
  MAVA_CHECK_PTR(old_value);
  MAVA_CHECK_PTR(adv);
  MAVA_CHECK_PTR(targets);
  MAVA_CHECK_PTR(rows);
  MAVA_CHECK_PTR(grad_out);
  MAVA_CHECK_PTR(workspace);
  MAVA_CHECK_ARG(num_replicas > 0 && num_replicas <= 8 && mb_size > 0 && critic->out_dim == 1);
  MAVA_CHECK_ARG((int64_t)num_replicas * mb_size * actor->num_agents < ((int64_t)1 << 31));
  MAVA_CHECK_ARG(actor->input_mode == MAVA_IN_AGENT_VIEW);
  cudaStream_t s = as_stream(stream);
  const int R = num_replicas * mb_size;
  const int64_t na = mava_mlp_param_count(actor), nc = mava_mlp_param_count(critic);
  unsigned char* w = static_cast<unsigned char*>(workspace);
  double* stats = reinterpret_cast<double*>(w);
  double* loss_acc = reinterpret_cast<double*>(w + 128);
  a.dz1_actor = w + 256;
  a.dz1_critic = a.dz1_actor + tile_count(actor, R) * (int64_t)tile_bytes(TM, HID);
  a.actor_img = static_cast<const unsigned char*>(actor_image);
  a.critic_img = static_cast<const unsigned char*>(critic_image);
  a.view = view;
  a.mask = mask;
  a.action = action;
  a.old_logp = old_logp;
  a.old_value = old_value;
  a.adv = adv;
  a.targets = targets;
  a.rows = rows;
  a.adv_stats = adv_stats ? adv_stats : stats;
  a.R = R;
  a.mb_size = mb_size;
  a.num_replicas = num_replicas;
  a.clip_eps = hyper->clip_eps;
  a.ent_coef = hyper->ent_coef;
  a.vf_coef = hyper->vf_coef;
  a.grad_actor = grad_out;
  a.grad_critic = grad_out + na;
  a.loss_acc = loss_acc;

  const int64_t ta = tile_count(actor, R), tcn = tile_count(critic, R);
  const int sms = sm_count();

  // (accumulating variant: the caller's optimiser kernel left grad_out and the loss accumulators
  //  zero, and finalises the loss metrics itself)
  if (!acc) {
    cudaError_t e = cudaMemsetAsync(workspace, 0, 256, s);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemsetAsync(grad_out, 0, (size_t)(na + nc + 8) * sizeof(float), s);
    if (e != cudaSuccess) return (int)e;
  }
  if (!adv_stats) {
    rc = launch_adv_stats(adv, rows, mb_size, actor->num_agents, num_replicas, stats, s);
    if (rc) return rc;
  }

  const int k1p_max = a.actor.k1p > a.critic.k1p ? a.actor.k1p : a.critic.k1p;
  size_t smem_fused = (size_t)WImage{k1p_max}.total() + region_bytes(k1p_max) +
                      tile_bytes(TM, HCOLS) + tile_bytes(TM, NHEAD) + 128;
  // actor first-layer gradient folded into the fused kernel when TMEM (208 free columns) and
  // shared memory (two X buffers on top of the activation tiles) have room for it
  const size_t smem_fold = (size_t)WImage{a.actor.k1p}.total() + 2 * tile_bytes(TM, a.actor.k1p) +
                           kRegionMin + tile_bytes(TM, HCOLS) + tile_bytes(TM, NHEAD) + 128;
  a.fold_actor_w1 = a.actor.k1p <= 208 && smem_fold <= 227 * 1024;
  if (a.fold_actor_w1 && smem_fold > smem_fused) smem_fused = smem_fold;
  // prefetch pipeline: padded staging rows (TM x k1p bytes); whole env-steps per tile, even rows,
  // word-sized steps, and the tile's words must fit the register slots of the twelve idle warps
  const size_t smem_pf = smem_fold + (size_t)TM * a.actor.k1p + 16;
  a.prefetch_actor = a.fold_actor_w1 && smem_pf <= 227 * 1024 && (TM % a.actor.A) == 0 &&
                     (a.actor.FR & 1) == 0 && ((a.actor.A * a.actor.FR) & 3) == 0 &&
                     (a.actor.A * a.actor.FR) / 4 <= 255 && TM * a.actor.k1p / 2 < 8192 &&
                     (TM / a.actor.A) * ((a.actor.A * a.actor.FR) >> 2) <= kLdSlots * NLOAD &&
                     TM / a.actor.A + 2 <= NLOAD;
  if (a.prefetch_actor && smem_pf > smem_fused) smem_fused = smem_pf;
  // a critic on the agent's own view (ff_ippo) has the actor's input: same pipeline, same conditions
  {
    const NetDesc& c = a.critic;
    const size_t c_fold = (size_t)WImage{c.k1p}.total() + 2 * tile_bytes(TM, c.k1p) + kRegionMin +
                          tile_bytes(TM, HCOLS) + tile_bytes(TM, NHEAD) + 128;
    const size_t c_pf = c_fold + (size_t)TM * c.k1p + 16;
    const bool ok = c.mode == MAVA_IN_AGENT_VIEW && c.k1p <= 208 && c_pf <= 227 * 1024 &&
                    (TM % c.A) == 0 && (c.FR & 1) == 0 && ((c.A * c.FR) & 3) == 0 &&
                    (c.A * c.FR) / 4 <= 255 && TM * c.k1p / 2 < 8192 &&
                    (TM / c.A) * ((c.A * c.FR) >> 2) <= kLdSlots * NLOAD && TM / c.A + 2 <= NLOAD &&
                    getenv("MAVA_NO_CRITIC_FOLD") == nullptr;
    a.fold_critic_w1 = a.prefetch_critic = ok ? 1 : 0;
    if (ok && c_pf > smem_fused) smem_fused = c_pf;
  }
  {
    static const int pipe_env = getenv("MAVA_NO_PIPE1") ? 0 : 1;  // development switch
    a.pipe_layer1 = pipe_env;
  }
  MAVA_CHECK_ARG(actor->num_agents <= kMaxReps);
  // Split the SMs between actor and critic tiles: the split that minimises the slower side's
  // whole-tile count x cost per tile (cycles: scripts/exp_phase_clock.sh, checked by a sweep of the
  // split with MAVA_ACTOR_CTAS).  An actor tile with the folded first-layer gradient and the
  // loader-warp pipeline ~8.8 K at k1p = 80; a tile on the plain path ~6.4 K + 30 per input column
  // (14.5 K for the 272-wide MAPPO critic, whose rows arrive by bulk copies).
  {
    const double tile_a = (a.prefetch_actor ? 6420.0 : 9050.0) + 30.0 * a.actor.k1p;
    const double tile_c = (a.prefetch_critic ? 6420.0 : 6360.0) + 30.0 * a.critic.k1p;
    int n_actor = 1;
    double best = 1e300;
    for (int na = 1; na < sms; ++na) {
      const int64_t ca = ta < na ? ta : na, cc = tcn < sms - na ? tcn : sms - na;
      const double t_a = (double)ceil_div64(ta, ca) * tile_a, t_c = (double)ceil_div64(tcn, cc) * tile_c;
      const double worst = t_a > t_c ? t_a : t_c;
      if (worst < best) {
        best = worst;
        n_actor = na;
      }
    }
    static const char* const split_override = getenv("MAVA_ACTOR_CTAS");  // development switch
    if (split_override) n_actor = atoi(split_override);
    n_actor = n_actor < 1 ? 1 : (n_actor > sms - 1 ? sms - 1 : n_actor);
    a.actor_ctas = (int)(ta < n_actor ? ta : n_actor);
    a.critic_ctas = (int)(tcn < sms - n_actor ? tcn : sms - n_actor);
  }

  const size_t stage_rows = (size_t)TM * grow_stride(k1p_max);
  const size_t stage_wg1 = stage_rows > tile_bytes(TM, HCOLS) ? stage_rows : tile_bytes(TM, HCOLS);
  // two dZ1 tile images, X tile, two staging buffers (rows requested two tiles ahead)
  const size_t stage_g = ((size_t)TM * grow_stride(k1p_max) + 127) & ~(size_t)127;
  const size_t smem_wg1 = 2 * (size_t)tile_bytes(TM, HID) + tile_bytes(TM, k1p_max) +
                          (2 * stage_g > stage_wg1 ? 2 * stage_g : stage_wg1) + 128;
  static size_t conf_fused[kMaxDevices] = {}, conf_wg1[kMaxDevices] = {};
  rc = ensure_dyn_smem(ppo_fused_kernel, smem_fused, conf_fused);
  if (rc) return rc;
  rc = ensure_dyn_smem(ppo_wgrad1_kernel, smem_wg1, conf_wg1);
  if (rc) return rc;
  ppo_fused_kernel<<<a.actor_ctas + a.critic_ctas, NT_F, smem_fused, s>>>(a);
  rc = launch_status();
  if (rc) return rc;
  // first-layer gradients that were not folded into the fused kernel: every SM to what is left
  if (a.fold_actor_w1 && a.fold_critic_w1) {
    a.actor_ctas = a.critic_ctas = 0;
  } else if (a.fold_actor_w1) {
    a.actor_ctas = 0;
    a.critic_ctas = (int)(tcn < sms ? tcn : sms);
  } else if (a.fold_critic_w1) {
    a.actor_ctas = (int)(ta < sms ? ta : sms);
    a.critic_ctas = 0;
  }
  if (a.actor_ctas + a.critic_ctas > 0) {
    ppo_wgrad1_kernel<<<a.actor_ctas + a.critic_ctas, NT, smem_wg1, s>>>(a);
    rc = launch_status();
    if (rc) return rc;
  }
  if (acc) return 0;
  return launch_finalize_loss(loss_acc, (double)R * actor->num_agents, hyper->ent_coef,
                              hyper->vf_coef, grad_out + na + nc, s);
}

int mava_ppo_loss_grad_bf16(const mava_mlp_desc* actor, const float* actor_params,
                            const void* actor_image, const mava_mlp_desc* critic,
                            const float* critic_params, const void* critic_image,
                            const mava_ppo_hyper* hyper, const int8_t* view, const uint8_t* mask,
                            const int8_t* action, const float* old_logp, const float* old_value,
                            const float* adv, const float* targets, const int32_t* rows,
                            int num_replicas, int mb_size, float* grad_out, void* workspace,
                            mava_stream_t stream) {
  return ppo_loss_grad_bf16_impl(actor, actor_params, actor_image, critic, critic_params,
                                 critic_image, hyper, view, mask, action, old_logp, old_value, adv,
                                 targets, rows, num_replicas, mb_size, grad_out, workspace, nullptr,
                                 stream);
}

int mava_ppo_adv_stats(const float* adv, const int32_t* rows, int num_replicas, int mb_size,
                       int num_agents, double* stats, mava_stream_t stream) {
  MAVA_CHECK_PTR(adv);
  MAVA_CHECK_PTR(rows);
  MAVA_CHECK_PTR(stats);
  MAVA_CHECK_ARG(num_replicas > 0 && num_replicas <= 8 && mb_size > 0 && num_agents > 0);
  cudaError_t e = cudaMemsetAsync(stats, 0, 16 * sizeof(double), as_stream(stream));
  if (e != cudaSuccess) return (int)e;
  return launch_adv_stats(adv, rows, mb_size, num_agents, num_replicas, stats, as_stream(stream));
}

int mava_ppo_loss_grad_bf16_stats(const mava_mlp_desc* actor, const float* actor_params,
                                  const void* actor_image, const mava_mlp_desc* critic,
                                  const float* critic_params, const void* critic_image,
                                  const mava_ppo_hyper* hyper, const int8_t* view,
                                  const uint8_t* mask, const int8_t* action, const float* old_logp,
                                  const float* old_value, const float* adv, const float* targets,
                                  const int32_t* rows, int num_replicas, int mb_size,
                                  const double* adv_stats, float* grad_out, void* workspace,
                                  mava_stream_t stream) {
  MAVA_CHECK_PTR(adv_stats);
  return ppo_loss_grad_bf16_impl(actor, actor_params, actor_image, critic, critic_params,
                                 critic_image, hyper, view, mask, action, old_logp, old_value, adv,
                                 targets, rows, num_replicas, mb_size, grad_out, workspace, adv_stats,
                                 stream);
}

int mava_ppo_loss_grad_bf16_acc(const mava_mlp_desc* actor, const float* actor_params,
                                const void* actor_image, const mava_mlp_desc* critic,
                                const float* critic_params, const void* critic_image,
                                const mava_ppo_hyper* hyper, const int8_t* view, const uint8_t* mask,
                                const int8_t* action, const float* old_logp, const float* old_value,
                                const float* adv, const float* targets, const int32_t* rows,
                                int num_replicas, int mb_size, const double* adv_stats,
                                float* grad_out, void* workspace, mava_stream_t stream) {
  MAVA_CHECK_PTR(adv_stats);
  return ppo_loss_grad_bf16_impl(actor, actor_params, actor_image, critic, critic_params,
                                 critic_image, hyper, view, mask, action, old_logp, old_value, adv,
                                 targets, rows, num_replicas, mb_size, grad_out, workspace, adv_stats,
                                 stream, /*acc=*/true);
}

}  // extern "C"
