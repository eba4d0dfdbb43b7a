// Fused persistent rollout for ff_ippo / ff_mappo on RobotWarehouse: the whole scan over
// `rollout_length` env steps of mava/systems/ppo/ff_mappo.py:76-106 in ONE kernel launch.
//
// A CTA owns up to TM / A environments (32 for 4 agents = one 128-row tile of the actor; 16 at 2048
// envs, so that 128 SMs take part) for the whole rollout.  Their packed records, their current
// observation rows and the actor's bf16 weight image stay in shared memory from the first step to
// the last.  640 threads: sixteen foreground warps walk the step loop, four background warps only
// generate spare records (see RCtrl).  Per step
//   1. the actor MLP runs on the tcgen05 tensor cores (three MMA chains into TMEM, as act_kernel);
//      warps whose tile rows are all dead skip their epilogues and their waits,
//   2. threads 0 .. TM-1 (thread r = lane g of its env = tile row r) finish their row in the head
//      epilogue -- action mask, Gumbel arg-max on the noise prepared one step ahead by the idle
//      warps, log-prob -- and advance their env (env_rware.cuh::step_group: whole-warp collectives,
//      sequential agent turns only where they matter, deliveries, episode metrics); a finished env
//      copies its spare record in on its own lanes,
//   3. four threads per row build the next observation row (a quarter each) into the int8
//      observation block and into a padded int8 image of the X row, and expand every fourth chunk
//      of the image into the bf16 X tile -- the row never takes a detour through a separate
//      expansion phase,
//   4. the idle warps write the step's outputs (staged in shared memory by the env lanes) and the
//      observation block to the rollout buffers with plain stores.
// No state leaves the SM between steps and there is no launch per step.  The kernel is latency
// bound by construction -- a step is a dependent chain MLP -> sample -> env step -> observation row
// and every env of the job is already in flight -- so everything here is about a shorter chain on
// the env lanes (DESIGN.md 4.4 has the phase clocks).
//
// The critic is not needed to act (ff_mappo.py:83 only records its value), so the values of all
// T + 1 observation slots are computed afterwards by one batched critic launch
// (mava_ff_act_bf16 with actor == NULL over [T+1][NE] observations) -- same numbers, and the tensor
// cores see 2049 x 128-row tiles instead of 16 per step.
#include "env_rware.cuh"
#include "mlp_tc.cuh"

namespace mava {
namespace tcmlp {
int make_net(const mava_mlp_desc* d, const float* params, NetDesc* n);

namespace {

// Phase timing (development aid, -DMAVA_PROFILE_PHASES): thread 0 of CTA 0 records clock64() at the
// phase boundaries of steps 16..31 (read back with mava_debug_rollout_phases).
#ifdef MAVA_PROFILE_PHASES
__device__ long long g_rollout_clock[16 * 16];
#define MAVA_RSTAMP(k)                                                                        \
  do {                                                                                        \
    if (t == 0 && blockIdx.x == 0 && step >= 16 && step < 32) g_rollout_clock[(step - 16) * 16 + (k)] = clock64(); \
  } while (0)
#else
#define MAVA_RSTAMP(k) do { } while (0)
#endif

struct RolloutArgs {
  RwareConst c;
  NetDesc actor;
  const unsigned char* actor_img;
  uint8_t* state;
  int8_t* view;                 // [T+1][NE][A][FR]
  uint8_t* mask;                // [T+1][NE][A]
  const uint32_t* policy_keys;  // [T][2]
  int8_t* action;               // [T][NE][A]
  float* logp;                  // [T][NE][A]
  float* reward;                // [T][NE][A]
  uint8_t* done;                // [T][NE]
  float* ep_return;             // [T][NE]
  int32_t* ep_length;           // [T][NE]
  int num_envs, envs_per_replica, T;
  int epc;  // envs per CTA (<= TM / A): fewer than a full tile when that fills more SMs
  // evaluator flavour (mava/evaluator.py:80-172): no AutoResetWrapper, pi.mode() instead of a
  // sample, and -- record == 0 -- only done / episode metrics are kept per step: observation, mask,
  // action, log-prob and reward of every step go to slot 0 of their buffers
  int auto_reset, greedy, record;
};

constexpr int NBG = 4;             // background warps (spare records, see RCtrl)
constexpr int NTR = NT + 32 * NBG;  // threads per CTA
constexpr int QN = 128;             // ticket-queue slots (> envs per CTA: one request per env in flight)

struct RCtrl {
  uint64_t wbar, mbar, rbar;
  uint32_t tmem;
  int rcount;
  uint16_t rlist[TM];
  // Speculative regeneration.  What a finished env is regenerated from is its State.key, which only
  // deliveries change, and a regeneration is a dependent chain of Threefry blocks (five block times
  // with a whole warp, env_rware.cuh::generate) -- so every env's NEXT initial state is generated
  // ahead of time into a spare record, tagged with the key it came from, by NBG background warps
  // that take no part in the step loop: one of the threads that build the env's observation rows
  // posts (env, key) into a ticket queue whenever the env has no spare for its current key, a
  // background warp generates it and marks it ready.  A finished env whose key still matches
  // copies its spare in on the spot (its own lanes, no CTA barrier); otherwise (a delivery since, or
  // the spare is still being generated) it is regenerated by a whole warp after the step's barrier.
  // Same bits either way.
  uint32_t q_head, q_tail;       // tickets handed to producers / consumers
  uint32_t q_seq[QN];            // ticket + 1 once the slot is written
  uint16_t q_el[QN];
  uint32_t q_key[QN][2];
  uint32_t quit;
  uint32_t sstate[TM];           // spare of env el: 0 none, 1 being generated, 2 ready
  uint32_t skey[TM][2];          // key the spare of env el was generated from
  // Outputs of a step, staged here by the env's own lanes and written to HBM by the warps that idle
  // (a thread that has stores to global memory in flight waits for them at its next proxy fence,
  // and the env lanes are the critical path of a step)
  float o_logp[TM], o_reward[TM], o_ret[TM];
  int32_t o_len[TM];
  int8_t o_action[TM];
  uint8_t o_mask[TM], o_done[TM];
  uint8_t rmode[TM];             // how the next observation row of tile row r is built this pass
};

__device__ __forceinline__ uint32_t ld_vol(const uint32_t* p) {
  return *reinterpret_cast<const volatile uint32_t*>(p);
}
__device__ __forceinline__ void st_vol(uint32_t* p, uint32_t v) {
  *reinterpret_cast<volatile uint32_t*>(p) = v;
}

// barrier of the NT foreground threads (the background warps join no CTA barrier once the step
// loop runs)
__device__ __forceinline__ void fg_sync() {
  asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
}
// bytes between the rows of the padded int8 images (k1p bytes + 8: 8-byte loads of a warp's 32 rows
// spread over the banks)
__host__ __device__ constexpr int img_stride(int k1p) { return k1p + 8; }
// Every thread that needs the accumulator polls the completion mbarrier itself (no barrier hop
// behind one polling warp).
__device__ __forceinline__ void fg_wait_mma(uint64_t* bar, uint32_t parity) {
  mbar_wait(bar, parity);
  fence_after_sync();
}

// The thread that issues, commits and waits for the bulk store of the records at the end (not in warp
// 0, the MMA issuer: a bulk copy issued next to an MMA chain was measured to delay it by ~500
// cycles, which is why the per-step outputs leave with plain stores).
constexpr int kStoreThread = TM;
// Staged block -> HBM by the NT foreground threads: one bulk store when size and address allow
// it, else a copy.  Returns true when a bulk store was issued (the caller commits / waits).
__device__ __forceinline__ bool store_block(const uint8_t* src, uint8_t* dst, int bytes, int t) {
  if ((((size_t)dst | (size_t)bytes) & 15) == 0) {
    if (t == kStoreThread) rware::bulk_s2g(dst, src, (uint32_t)bytes);
    return true;
  }
  if ((((size_t)dst | (size_t)bytes) & 3) == 0) {
    for (int i = t; i < (bytes >> 2); i += NT)
      reinterpret_cast<uint32_t*>(dst)[i] = reinterpret_cast<const uint32_t*>(src)[i];
  } else {
    for (int i = t; i < bytes; i += NT) dst[i] = src[i];
  }
  return false;
}

// Development switches (scripts/exp_rollout_code.sh): how the cold code of the step loop is emitted.
#ifndef MAVA_ROLL_GEN
#define MAVA_ROLL_GEN 3   // 0: generator inlined at both call sites; 1: one out-of-line copy, rolled
#endif                    // top-k loops, out-of-line Threefry blocks; 3: the same with unrolled top-k
                          // loops (default: +0.4 % over 1, within 0.2 % of 0); 2: one out-of-line
                          // copy, everything inlined into it (-6 %)
#ifndef MAVA_ROLL_COLD
#define MAVA_ROLL_COLD 1  // 1: delivery draws through out-of-line Threefry, whole-row build out of line
#endif
#if MAVA_ROLL_COLD
using ColdPrng = PrngCall;
#define MAVA_COLD_FN __noinline__
#else
using ColdPrng = PrngInline;
#define MAVA_COLD_FN __forceinline__
#endif

// A fresh episode into `rec`, by a whole warp (background warps, and the regeneration on the spot).
// (KA: the agent count is the kernel's template parameter, so the top-k lists are as short as needed)
#if MAVA_ROLL_GEN == 0
template <int KA>
__device__ __forceinline__ void generate_warp(const RwareConst& c, uint8_t* rec, Key key, int lane) {
  rware::generate<32>(c, rec, key, lane, 0xffffffffu);
}
#elif MAVA_ROLL_GEN == 1
template <int KA>
__device__ __noinline__ void generate_warp(const RwareConst& c, uint8_t* rec, Key key, int lane) {
  rware::generate<32, PrngCall, true, KA>(c, rec, key, lane, 0xffffffffu);
}
#elif MAVA_ROLL_GEN == 3
template <int KA>
__device__ __noinline__ void generate_warp(const RwareConst& c, uint8_t* rec, Key key, int lane) {
  rware::generate<32, PrngCall, false, KA>(c, rec, key, lane, 0xffffffffu);
}
#else
template <int KA>
__device__ __noinline__ void generate_warp(const RwareConst& c, uint8_t* rec, Key key, int lane) {
  rware::generate<32>(c, rec, key, lane, 0xffffffffu);
}
#endif

// The words of an observation row (rware::build_row) straight into tile row r of the bf16 X tile:
// [onehot(agent g) if add_id | row | 1 | 0...], the image expand_x_row builds from the int8 row.
template <int G, int R>
__device__ __forceinline__ void store_x_row(const NetDesc& d, const Tile& xt, int r, int g,
                                            const uint32_t (&w)[rware::ObsDims<R>::NW]) {
  using O = rware::ObsDims<R>;
  constexpr int KW = pad16(G + O::FR + 1) / 4;  // words of the longest int8 image
  static_assert(KW >= O::NW + (G + 3) / 4 && KW % 2 == 0, "image size");
  static_assert(G % 4 == 0 || G == 2, "agent-id columns");
  uint32_t img[KW];
#pragma unroll
  for (int j = 0; j < KW; ++j) img[j] = 0u;
  if (d.add_id) {
    if constexpr (G % 4 == 0) {
#pragma unroll
      for (int i = 0; i < O::NW; ++i) img[G / 4 + i] = w[i];
    } else {
      img[0] = w[0] << 16;
#pragma unroll
      for (int i = 1; i < O::NW; ++i) img[i] = __funnelshift_r(w[i - 1], w[i], 16);
    }
#pragma unroll
    for (int j = 0; j < (G + 3) / 4; ++j)
      if (j == (g >> 2)) img[j] |= 1u << (8 * (g & 3));
    img[(G + O::FR) >> 2] |= 1u << (8 * ((G + O::FR) & 3));
  } else {
#pragma unroll
    for (int i = 0; i < O::NW; ++i) img[i] = w[i];
    img[O::FR >> 2] |= 1u << (8 * (O::FR & 3));
  }
#pragma unroll
  for (int cg = 0; cg < KW / 2; ++cg)
    if (cg < d.k1p / 8)
      st_shared_v4(xt.base + chunk_off(xt, r, cg), s8x2_bf16x2(img[2 * cg]),
                   s8x2_bf16x2(img[2 * cg] >> 16), s8x2_bf16x2(img[2 * cg + 1]),
                   s8x2_bf16x2(img[2 * cg + 1] >> 16));
}

// The whole observation row by one thread (agent-id columns that leave the observation words
// unaligned in the image, or the exact grid replay of a terminal collision): out of line, it is not
// on the headline path and the step loop's instruction footprint matters.
template <int G>
struct OpkWords {
  uint32_t v[G];
};
template <int G, int R>
__device__ MAVA_COLD_FN uint32_t row_alone(const RwareConst& c, const NetDesc& d, const uint8_t* rec,
                                           int g, bool replay, OpkWords<G> ow, uint8_t* row,
                                           const Tile& xt, int r) {
  uint32_t w[rware::ObsDims<R>::NW];
  const uint32_t mk = rware::build_row<G, R>(c, rec, g, replay, ow.v, w);
  rware::store_row<R>(w, row, r & 1);
  store_x_row<G, R>(d, xt, r, g, w);
  return mk;
}

// Head row: masked logits -> sampled (Gumbel arg-max on the prepared noise) or greedy action and its
// log-prob.  N > 0: exactly N actions (the loops fold); N == 0: n <= NHEAD actions.
template <int N>
__device__ __forceinline__ void head_row(float (&out)[NHEAD], uint32_t mk, const float* nz,
                                         bool greedy, int n, int& act, float& la) {
  constexpr int NN = N ? N : NHEAD;
  float mx = kF32Min;
#pragma unroll
  for (int j = 0; j < NN; ++j) {
    if (N || j < n) {
      out[j] = ((mk >> j) & 1u) ? out[j] : kF32Min;
      mx = fmaxf(mx, out[j]);
    }
  }
  float se = 0.0f;
#pragma unroll
  for (int j = 0; j < NN; ++j)
    if (N || j < n) se += expf(out[j] - mx);
  const float lse = mx + logf(se);
  float best = 0.0f;
  act = 0;
#pragma unroll
  for (int j = 0; j < NN; ++j) {
    if (N || j < n) {
      const float z = (greedy ? 0.0f : nz[j]) + out[j];
      if (j == 0 || z > best) { best = z; act = j; }
    }
  }
  la = 0.0f;
#pragma unroll
  for (int j = 0; j < NN; ++j)
    if (j == act) la = out[j] - lse;
}

template <int G, int R>
__global__ void __launch_bounds__(NTR, 1)
rware_rollout_kernel(const __grid_constant__ RolloutArgs p) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ RCtrl ctrl;
  constexpr int EPC = TM / G;  // max envs per CTA; agents == G, so tile row r = el * G + g = thread r
  const RwareConst& c = p.c;
  const NetDesc& d = p.actor;
  const Lane L;
  const int t = L.t, warp = L.warp;

  const WImage wi{d.k1p};
  const uint32_t s_w = smem_u32(smem);
  const Tile xt{s_w + wi.total(), 128u, (uint32_t)(TM / 8) * 128u};
  const Tile ht{xt.base + tile_bytes(TM, d.k1p), 128u, (uint32_t)(TM / 8) * 128u};
  uint8_t* srec = smem + wi.total() + tile_bytes(TM, d.k1p) + tile_bytes(TM, HCOLS);
  uint8_t* sobs = srec + EPC * c.stride;
  float* snoise = reinterpret_cast<float*>(sobs + round_up(EPC * G * c.FR, 16));  // [2][TM][NHEAD]
  uint8_t* sspare = reinterpret_cast<uint8_t*>(snoise + 2 * TM * NHEAD);           // [EPC][stride]
  // padded int8 image of every tile row: [onehot(agent) | observation row | 1 | 0...]; the constant
  // parts are written once, the observation bytes by the four threads of the row every step
  uint8_t* simg = sspare + round_up(EPC * c.stride, 16);                           // [TM][IMG_S]
  const int IMG_S = img_stride(d.k1p);
  const int id_ofs = d.add_id ? G : 0;
  // four threads per row need the observation words 4-byte aligned in the image
  const bool split = (id_ofs & 3) == 0;

  const int env0 = blockIdx.x * p.epc;
  const int nenv = min(p.epc, p.num_envs - env0);
  const int rows_valid = nenv * G;
  const uint32_t rec_bytes = (uint32_t)nenv * (uint32_t)c.stride;
  uint8_t* gstate = p.state + (size_t)env0 * c.stride;
  const int obs_bytes = nenv * G * c.FR;
  const size_t obs_slot = (size_t)p.num_envs * G * c.FR;   // bytes of one time slot of `view`
  const size_t ea_slot = (size_t)p.num_envs * G;           // elements of one [NE][A] slot

  if (warp == 0) tmem_alloc<256>(&ctrl.tmem);
  if (t == 0) {
    mbar_init(&ctrl.wbar, 1);
    mbar_init(&ctrl.mbar, 1);
    mbar_init(&ctrl.rbar, 1);
    fence_mbar_init();
    ctrl.rcount = 0;
    ctrl.q_head = 0;
    ctrl.q_tail = 0;
    ctrl.quit = 0;
  }
  if (t < TM) ctrl.sstate[t] = 0;
  if (t < QN) ctrl.q_seq[t] = 0;
  fence_before_sync();
  __syncthreads();
  fence_after_sync();

  // ---- background warps: spare records on request, until the foreground says quit
  if (t >= NT) {
    for (;;) {
      uint32_t ticket = 0, stop = 0;
      if (L.lane == 0) {
        ticket = atomicAdd(&ctrl.q_tail, 1u);
        while (ld_vol(&ctrl.q_seq[ticket % QN]) != ticket + 1u) {
          if (ld_vol(&ctrl.quit)) {
            stop = 1;
            break;
          }
          __nanosleep(100);
        }
      }
      stop = __shfl_sync(0xffffffffu, stop, 0);
      if (stop) return;
      ticket = __shfl_sync(0xffffffffu, ticket, 0);
      __threadfence_block();
      const int slot = (int)(ticket % QN);
      const int bel = ctrl.q_el[slot];
      const Key bk{ctrl.q_key[slot][0], ctrl.q_key[slot][1]};
      Key nk, unused;
      split2(bk, nk, unused);
      generate_warp<G>(c, sspare + bel * c.stride, nk, L.lane);
      if (L.lane == 0) {
        ctrl.skey[bel][0] = bk.k0;
        ctrl.skey[bel][1] = bk.k1;
        __threadfence_block();
        st_vol(&ctrl.sstate[bel], 2u);
      }
    }
  }

  const uint32_t tmem = ctrl.tmem;
  if (t == 0) {
    load_weights(s_w, p.actor_img, wi.total(), &ctrl.wbar);
    mbar_expect_tx(&ctrl.rbar, rec_bytes);
    bulk_g2s(smem_u32(srec), gstate, rec_bytes, &ctrl.rbar);
  }
  // the observation rows step 0 acts on (slot 0 of `view`)
  {
    const uint8_t* src = reinterpret_cast<const uint8_t*>(p.view) + (size_t)env0 * G * c.FR;
    for (int i = t; i < obs_bytes; i += NT) sobs[i] = src[i];
    for (int i = t; i < TM * IMG_S / 4; i += NT) reinterpret_cast<uint32_t*>(simg)[i] = 0u;
  }
  // env-step role of the first TM threads: thread r is lane g of env el, and finishes tile row r
  const bool stepper = L.q == 0;
  const int el = L.r / G, g = L.r % G;
  const int env = env0 + el;
  const bool agent = stepper && el < nenv;
  const unsigned gmask = rware::group_mask<G>();
  uint8_t* rec = srec + el * c.stride;
  // the action mask of a row travels through the staging block (its builder is another thread)
  if (agent) ctrl.o_mask[L.r] = p.mask[(size_t)env * G + g];
  // row-building role: warp w builds quarter (w & 3) of the rows of tile-row quadrant (w >> 2), so
  // that the four warps of an SM sub-partition (w & 3) walk the same code
  const int bq = warp >> 2, bp = warp & 3;
  Lane Lb = L;
  Lb.r = bq * 32 + L.lane;
  const int bel = Lb.r / G, bg = Lb.r % G;
  uint8_t* brec = srec + bel * c.stride;
  const bool blive = bq * 32 < rows_valid;
  // Gumbel noise of a step: depends only on the step's policy key and (env, agent, action), so it
  // is produced one step ahead by the warps that idle during the env step (noise laid out
  // (envs_per_replica, A, N) as tfd.Categorical.sample draws it)
  auto make_noise = [&](int step, int first, int nthreads) {
    const Key key{__ldg(p.policy_keys + 2 * step), __ldg(p.policy_keys + 2 * step + 1)};
    const uint32_t size = (uint32_t)p.envs_per_replica * G * d.out;
    float* dst = snoise + (step & 1) * TM * NHEAD;
    for (int idx = first; idx < rows_valid * d.out; idx += nthreads) {
      const int row = idx / d.out, j = idx - row * d.out;
      const int e = (env0 + row / G) % p.envs_per_replica;
      const uint32_t base = (uint32_t)((e * G + row % G) * d.out);
      dst[row * NHEAD + j] = bits_to_gumbel(random_bits_at(key, base + j, size));
    }
  };
  if (!p.greedy) make_noise(0, t, NT);
  mbar_wait(&ctrl.wbar, 0);
  mbar_wait(&ctrl.rbar, 0);
  // lane 0 of an env: ask for a spare unless one for the current key is there or on its way (an
  // env whose key changed while its spare was being generated asks again on a later step)
  auto want_spare = [&](int el, const uint8_t* rec) {
    const uint32_t* k = reinterpret_cast<const uint32_t*>(rec + c.off_key);
    const uint32_t st = ld_vol(&ctrl.sstate[el]);
    if (st == 1u) return;
    if (st == 2u && ctrl.skey[el][0] == k[0] && ctrl.skey[el][1] == k[1]) return;
    st_vol(&ctrl.sstate[el], 1u);
    const uint32_t ticket = atomicAdd(&ctrl.q_head, 1u);
    const int slot = (int)(ticket % QN);
    ctrl.q_el[slot] = (uint16_t)el;
    ctrl.q_key[slot][0] = k[0];
    ctrl.q_key[slot][1] = k[1];
    __threadfence_block();
    st_vol(&ctrl.q_seq[slot], ticket + 1u);
  };
  if (agent && g == 0 && p.auto_reset) want_spare(el, rec);
  // X tile of step 0 from the observation rows in shared memory (later steps: written by the row's
  // own thread while it builds the row)
  fg_sync();
  if (t < rows_valid) {
    if (d.add_id) simg[t * IMG_S + g] = 1;
    simg[t * IMG_S + id_ofs + c.FR] = 1;
  }
  expand_x_row(d, xt, L, reinterpret_cast<const signed char*>(sobs) + L.r * c.FR, L.r < rows_valid,
               g);
  fence_proxy_async();
  fence_before_sync();
  fg_sync();

  uint32_t phase = 0;
  // warps whose 32 tile rows are all beyond this CTA's envs skip the hidden-layer epilogues (TMEM
  // reads are 64 B/clk per SM: half a tile of dead rows would cost as much as the live half)
  const bool live = (warp & 3) * 32 < rows_valid;
  for (int step = 0; step < p.T; ++step) {
    const int so = p.record ? step : 0, so1 = p.record ? step + 1 : 0;  // output slots of this step
    MAVA_RSTAMP(0);
    // ---- 1. actor MLP on the X tile
    if (mma_issuer()) {
      fence_after_sync();
      issue_gemm(tmem, xt, false, w1_tile(s_w, d.k1p), true, HID, d.k1p, false, &ctrl.mbar);
    }
    if (live) fg_wait_mma(&ctrl.mbar, phase);
    MAVA_RSTAMP(1);
    phase ^= 1;
    if (live) hidden_epilogue(L, tmem, ht);
    fence_proxy_async();
    fence_before_sync();
    fg_sync();
    MAVA_RSTAMP(2);
    if (t == 0) ctrl.rcount = 0;  // everybody has read the previous step's count
    if (mma_issuer()) {
      fence_after_sync();
      issue_gemm(tmem, ht, false, w2_tile(s_w, d.k1p), true, HID, HCOLS, false, &ctrl.mbar);
    }
    if (live) fg_wait_mma(&ctrl.mbar, phase);
    MAVA_RSTAMP(3);
    phase ^= 1;
    if (live) hidden_epilogue(L, tmem, ht);
    fence_proxy_async();
    fence_before_sync();
    fg_sync();
    MAVA_RSTAMP(4);
    if (mma_issuer()) {
      fence_after_sync();
      issue_gemm(tmem + HID, ht, false, w3_tile(s_w, d.k1p), true, NHEAD, HCOLS, false, &ctrl.mbar);
    }
    if (stepper && live) fg_wait_mma(&ctrl.mbar, phase);  // the other warps do not read the head
    MAVA_RSTAMP(5);
    phase ^= 1;
    // ---- 2. head epilogue, env step, reset from the spare, next observation row + X row
    //         (threads 0 .. TM-1; everything of an env happens on its own G lanes)
    bool needs_reset = false, replay = false, row_due = false, listed = false;
    uint32_t opk[G];
#pragma unroll
    for (int j = 0; j < G; ++j) opk[j] = 0u;
    if (stepper && live) {
      float out[NHEAD];
      ld16(tmem + L.tmem_lane() + (uint32_t)HID, out);
      int act = 0;
      if (agent) {
        const float* nz = snoise + (step & 1) * TM * NHEAD + L.r * NHEAD;
        float la;
        const uint32_t mk = ctrl.o_mask[L.r];
        if (d.out == 5) head_row<5>(out, mk, nz, p.greedy, d.out, act, la);        // RobotWarehouse
        else head_row<0>(out, mk, nz, p.greedy, d.out, act, la);
        ctrl.o_action[L.r] = (int8_t)act;
        ctrl.o_logp[L.r] = la;
      }
      MAVA_RSTAMP(6);
      // sampled from masked logits: the action respects the mask of the current state
      rware::step_group<G, true, ColdPrng>(c, rec, g, gmask, el < nenv, agent, act, el, p.auto_reset,
                                 ctrl.o_reward, ctrl.o_done, ctrl.o_ret, ctrl.o_len, needs_reset,
                                 replay, opk);
      MAVA_RSTAMP(7);
      __syncwarp();  // lane 0's State.key
      row_due = agent;
      bool hit = false;
      if (needs_reset) {
        const uint32_t* k = reinterpret_cast<const uint32_t*>(rec + c.off_key);
        hit = ld_vol(&ctrl.sstate[el]) == 2u && ctrl.skey[el][0] == k[0] &&
              ctrl.skey[el][1] == k[1];
        if (hit) {
          // the inner-env part of the record: [agents | queue | request bits | step | key] and
          // the shelf grid; the episode-metrics words in between stay
          __threadfence_block();
          const uint32_t* sp = reinterpret_cast<const uint32_t*>(sspare + el * c.stride);
          uint32_t* dst = reinterpret_cast<uint32_t*>(rec);
          for (int w = g; w < (c.off_mkey >> 2); w += G) dst[w] = sp[w];
          for (int w = g; w < c.cells_words; w += G)
            dst[(c.off_cells >> 2) + w] = sp[(c.off_cells >> 2) + w];
        } else {
          if (g == 0) ctrl.rlist[atomicAdd(&ctrl.rcount, 1)] = (uint16_t)el;
          listed = agent;
          row_due = false;  // after the regeneration below
        }
      }
      __syncwarp();  // the record is whole again; everybody has looked at the spare's state
      if (hit && g == 0) st_vol(&ctrl.sstate[el], 0u);
    } else {
      if (step + 1 < p.T && !p.greedy)
        make_noise(step + 1, t - TM, NT - TM);  // warps 4..15: next step's noise
    }
    const int xchunks = d.in_dim / 8 + 1;  // X chunks that hold observation bytes or the ones column
    for (int pass = 0;; ++pass) {
      // four threads per row, each a quarter of it into the int8 block and the padded image; then
      // every fourth chunk of the image into the bf16 X tile
      if (stepper && live) {
        const int mode = !row_due ? 0 : (split && !replay ? 1 : 2);
        ctrl.rmode[L.r] = (uint8_t)mode;
        if (mode == 2) {  // the whole row here (it needs this thread's registers)
          OpkWords<G> ow;
#pragma unroll
          for (int j = 0; j < G; ++j) ow.v[j] = opk[j];
          ctrl.o_mask[L.r] = (uint8_t)row_alone<G, R>(c, d, rec, g, replay, ow, sobs + L.r * c.FR, xt, L.r);
          if (g == 0 && p.auto_reset) want_spare(el, rec);
        }
        if (pass == 0) MAVA_RSTAMP(13);
        // "the records of quadrant `warp` are final for this pass" to its builders (warps 4 warp ..
        // 4 warp + 3; quadrant 0's are the stepper warps themselves: plain barrier below)
        if (warp != 0) {
          __threadfence_block();
          asm volatile("bar.arrive %0, 160;" ::"r"(2 + warp) : "memory");
        }
      }
      if (blive) {
        if (bq == 0) asm volatile("bar.sync 2, 128;" ::: "memory");
        else asm volatile("bar.sync %0, 160;" ::"r"(2 + bq) : "memory");
        if (pass == 0) MAVA_RSTAMP(14);
        const int mode = ctrl.rmode[Lb.r];
        uint8_t* img = simg + Lb.r * IMG_S;
        if (mode == 1) {
          const uint8_t m = rware::build_row_part<G, R>(c, brec, bg, bp, sobs + Lb.r * c.FR,
                                                        img + id_ofs);
          if (bp == 0) ctrl.o_mask[Lb.r] = m;
          // (the quarter with the least to do asks for the env's next spare)
          if (bp == 3 && bg == 0 && p.auto_reset) want_spare(bel, brec);
        }
        if (pass == 0) MAVA_RSTAMP(15);
        asm volatile("bar.sync %0, 128;" ::"r"(6 + bq) : "memory");  // the images are complete
        if (pass == 0) MAVA_RSTAMP(11);
        // (RobotWarehouse observation bytes are coordinates < 128 and flags: the cheaper conversion;
        //  up to twelve chunks are three per thread)
        if (mode == 1) {
          if (xchunks <= 12) expand_padded_row<3, true>(xt, Lb, smem_u32(img), true, xchunks, bp, 4);
          else expand_padded_row<4, true>(xt, Lb, smem_u32(img), true, xchunks, bp, 4);
        }
      }
      if (pass == 0) MAVA_RSTAMP(8);
      fence_proxy_async();
      fence_before_sync();
      fg_sync();
      if (pass == 0) MAVA_RSTAMP(9);
      const int nreset = ctrl.rcount;
      if (pass == 1 || nreset == 0) break;
      // ---- finished envs without a usable spare: one regeneration per warp at a time
      for (int i = warp; i < nreset; i += NWARPS) {
        const int rel = (int)ctrl.rlist[i];
        uint8_t* rrec = srec + rel * c.stride;
        for (uint32_t spins = 0; ld_vol(&ctrl.sstate[rel]) == 1u; ++spins) {  // on its way
          __nanosleep(100);
          if (spins > (1u << 22)) __trap();
        }
        __threadfence_block();
        const uint32_t* k = reinterpret_cast<const uint32_t*>(rrec + c.off_key);
        const uint32_t k0 = k[0], k1 = k[1];
        const bool hit = ld_vol(&ctrl.sstate[rel]) == 2u && ctrl.skey[rel][0] == k0 &&
                         ctrl.skey[rel][1] == k1;
        __syncwarp();
        if (hit) {
          const uint32_t* sp = reinterpret_cast<const uint32_t*>(sspare + rel * c.stride);
          uint32_t* dst = reinterpret_cast<uint32_t*>(rrec);
          for (int w = L.lane; w < (c.off_mkey >> 2); w += 32) dst[w] = sp[w];
          for (int w = L.lane; w < c.cells_words; w += 32)
            dst[(c.off_cells >> 2) + w] = sp[(c.off_cells >> 2) + w];
          __syncwarp();
          if (L.lane == 0) st_vol(&ctrl.sstate[rel], 0u);
        } else {
          Key nk, unused;
          split2(Key{k0, k1}, nk, unused);
          generate_warp<G>(c, rrec, nk, L.lane);
        }
      }
      fg_sync();
      row_due = listed;
    }
    // the step's outputs: staged by the env lanes, written by the other warps
    if (t >= TM) {
      const size_t o = (size_t)so * ea_slot + (size_t)env0 * G;
      for (int i = t - TM; i < rows_valid; i += NT - TM) {
        p.action[o + i] = ctrl.o_action[i];
        p.logp[o + i] = ctrl.o_logp[i];
        p.reward[o + i] = ctrl.o_reward[i];
        p.mask[(size_t)so1 * ea_slot + (size_t)env0 * G + i] = ctrl.o_mask[i];
      }
      // ... and the observation block (plain stores: a bulk store issued next to an MMA chain costs
      // the chain ~500 cycles)
      if ((obs_bytes & 3) == 0) {
        uint32_t* dst = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(p.view) +
                                                    (size_t)so1 * obs_slot + (size_t)env0 * G * c.FR);
        const uint32_t* src = reinterpret_cast<const uint32_t*>(sobs);
        for (int i = t - TM; i < (obs_bytes >> 2); i += NT - TM) dst[i] = src[i];
      } else {
        uint8_t* dst = reinterpret_cast<uint8_t*>(p.view) + (size_t)so1 * obs_slot +
                       (size_t)env0 * G * c.FR;
        for (int i = t - TM; i < obs_bytes; i += NT - TM) dst[i] = sobs[i];
      }
      const size_t oe = (size_t)step * p.num_envs + env0;
      for (int i = t - TM; i < nenv; i += NT - TM) {
        p.done[oe + i] = ctrl.o_done[i];
        p.ep_return[oe + i] = ctrl.o_ret[i];
        p.ep_length[oe + i] = ctrl.o_len[i];
      }
    }
    MAVA_RSTAMP(12);
  }
  // ---- records back to HBM
  if (t == 0) st_vol(&ctrl.quit, 1u);
  fence_proxy_async();
  fg_sync();
  const bool bulk = store_block(srec, gstate, (int)rec_bytes, t);
  if (t == kStoreThread && bulk) rware::bulk_commit_wait_read();
  fence_before_sync();
  fg_sync();
  if (warp == 0) tmem_dealloc<256>(tmem);
}

template <int G>
int launch_rollout(const RolloutArgs& a, cudaStream_t s) {
  constexpr int EPC = TM / G;
  const size_t smem = (size_t)WImage{a.actor.k1p}.total() + tile_bytes(TM, a.actor.k1p) +
                      tile_bytes(TM, HCOLS) + (size_t)EPC * a.c.stride +
                      (size_t)round_up(EPC * G * a.c.FR, 16) + 2 * TM * NHEAD * 4 +
                      (size_t)round_up(EPC * a.c.stride, 16) +          // ... noise, spare records
                      (size_t)TM * img_stride(a.actor.k1p) + 128;        // padded row images
  static size_t configured[kMaxDevices] = {};
  if (int rc = ensure_dyn_smem(rware_rollout_kernel<G, 1>, smem, configured)) return rc;
  // spread the envs over the SMs: a CTA takes whole warps of envs (32 / G each), at most a tile
  RolloutArgs b = a;
  const int per_warp = 32 / G;
  int epc = round_up(ceil_div(a.num_envs, sm_count()), per_warp);
  b.epc = epc < per_warp ? per_warp : (epc > EPC ? EPC : epc);
  rware_rollout_kernel<G, 1><<<ceil_div(a.num_envs, b.epc), NTR, smem, s>>>(b);
  return launch_status();
}

}  // namespace
}  // namespace tcmlp
}  // namespace mava

using namespace mava;
using namespace mava::tcmlp;

extern "C" {

#ifdef MAVA_PROFILE_PHASES
int mava_debug_rollout_phases(long long* out_host) {
  return (int)cudaMemcpyFromSymbol(out_host, g_rollout_clock, sizeof(long long) * 256);
}
#endif

int mava_ff_rollout_bf16_ex(mava_env_t env, const mava_mlp_desc* actor, const float* actor_params,
                            const void* actor_image, uint8_t* state, int8_t* view, uint8_t* mask,
                            const uint32_t* policy_keys, int envs_per_replica, int num_envs,
                            int rollout_length, int auto_reset, int greedy, int record,
                            int8_t* action, float* logp, float* reward, uint8_t* done,
                            float* ep_return, int32_t* ep_length, mava_stream_t s) {
  MAVA_CHECK_PTR(env);
  RolloutArgs a{};
  int rc = make_net(actor, actor_params, &a.actor);
  if (rc) return rc;
  MAVA_CHECK_PTR(actor_image);
  MAVA_CHECK_PTR(state);
  MAVA_CHECK_PTR(view);
  MAVA_CHECK_PTR(mask);
  MAVA_CHECK_PTR(policy_keys);
  MAVA_CHECK_PTR(action);
  MAVA_CHECK_PTR(logp);
  MAVA_CHECK_PTR(reward);
  MAVA_CHECK_PTR(done);
  MAVA_CHECK_PTR(ep_return);
  MAVA_CHECK_PTR(ep_length);
  MAVA_CHECK_ARG(num_envs > 0 && envs_per_replica > 0 && rollout_length > 0);
  MAVA_CHECK_ARG(actor->input_mode == MAVA_IN_AGENT_VIEW);
  if (env->kind != MAVA_ENV_RWARE) return MAVA_E_UNSUPPORTED;
  const RwareConst& c = env->rw;
  if (c.R != 1 || (c.A != 2 && c.A != 4 && c.A != 8)) return MAVA_E_UNSUPPORTED;
  MAVA_CHECK_ARG(actor->num_agents == c.A && actor->view_dim == c.FR);
  a.c = c;
  a.actor_img = static_cast<const unsigned char*>(actor_image);
  a.state = state;
  a.view = view;
  a.mask = mask;
  a.policy_keys = policy_keys;
  a.action = action;
  a.logp = logp;
  a.reward = reward;
  a.done = done;
  a.ep_return = ep_return;
  a.ep_length = ep_length;
  a.num_envs = num_envs;
  a.envs_per_replica = envs_per_replica;
  a.T = rollout_length;
  a.auto_reset = auto_reset != 0;
  a.greedy = greedy != 0;
  a.record = record != 0;
  if (c.A == 2) return launch_rollout<2>(a, as_stream(s));
  if (c.A == 4) return launch_rollout<4>(a, as_stream(s));
  return launch_rollout<8>(a, as_stream(s));
}

int mava_ff_rollout_bf16(mava_env_t env, const mava_mlp_desc* actor, const float* actor_params,
                         const void* actor_image, uint8_t* state, int8_t* view, uint8_t* mask,
                         const uint32_t* policy_keys, int envs_per_replica, int num_envs,
                         int rollout_length, int8_t* action, float* logp, float* reward,
                         uint8_t* done, float* ep_return, int32_t* ep_length, mava_stream_t s) {
  return mava_ff_rollout_bf16_ex(env, actor, actor_params, actor_image, state, view, mask,
                                 policy_keys, envs_per_replica, num_envs, rollout_length,
                                 /*auto_reset=*/1, /*greedy=*/0, /*record=*/1, action, logp, reward,
                                 done, ep_return, ep_length, s);
}

}  // extern "C"
