// Fused persistent rollout for ff_ippo / ff_mappo on RobotWarehouse: the whole scan over
// `rollout_length` env steps of mava/systems/ppo/ff_mappo.py:76-106 in ONE kernel launch.
//
// A CTA owns TM / A environments (32 for 4 agents = one 128-row tile of the actor) for the whole
// rollout.  Their packed records, their current observation rows and the actor's bf16 weight image
// stay in shared memory from the first step to the last; per step the CTA
//   1. expands the int8 observation rows to the bf16 X tile,
//   2. runs the actor MLP on the tcgen05 tensor cores (three MMA chains into TMEM, as act_kernel),
//   3. finishes each row in the head epilogue: action mask, Gumbel arg-max on the threefry bits of
//      this step's policy key, log-prob -- the sampled action stays in the register of the thread
//      that is also lane g of its env,
//   4. advances the envs (env_rware.cuh: sequential agent turns, collisions, deliveries, episode
//      metrics, CTA-queued regeneration of finished envs) and builds the next observation rows,
//   5. streams observation rows, mask, action, log-prob, reward, done and episode metrics of the
//      step to the rollout buffers (bulk store for the observation block).
// No state leaves the SM between steps and there is no launch per step: the rollout is latency
// bound at 2048 envs per GPU, and this removes 2 launches, a weight reload and an HBM round trip
// of the env records from every step.
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

struct RCtrl {
  uint64_t wbar, mbar, rbar;
  uint32_t tmem;
  int rcount;
  uint16_t rlist[TM];
  // Speculative regeneration.  What a finished env is regenerated from is its State.key, which only
  // deliveries change, and a regeneration is a long dependent threefry chain -- so the warps that
  // idle during the env step generate every env's NEXT initial state ahead of time into a spare
  // record (tagged with the key it was generated from).  A reset whose key still matches copies the
  // spare in; otherwise (a delivery since) it regenerates on the spot as before.  Same bits either
  // way.
  int bgcount;              // envs whose spare is missing or stale (found while emitting rows)
  uint16_t bglist[TM];
  uint32_t bgkey[TM][2];    // State.key snapshot of the listed env
  uint32_t skey[TM][2];     // key the spare of env el was generated from
  uint8_t sok[TM];          // spare of env el is valid
};

// Staged block -> HBM: one bulk store when size and address allow it, else a CTA-wide copy.
// Returns true when a bulk store was issued (the caller commits / waits).
__device__ __forceinline__ bool store_block(const uint8_t* src, uint8_t* dst, int bytes) {
  if ((((size_t)dst | (size_t)bytes) & 15) == 0) {
    if (threadIdx.x == 0) rware::bulk_s2g(dst, src, (uint32_t)bytes);
    return true;
  }
  if ((((size_t)dst | (size_t)bytes) & 3) == 0) {
    for (int i = threadIdx.x; i < (bytes >> 2); i += blockDim.x)
      reinterpret_cast<uint32_t*>(dst)[i] = reinterpret_cast<const uint32_t*>(src)[i];
  } else {
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) dst[i] = src[i];
  }
  return false;
}

template <int G, int R>
__global__ void __launch_bounds__(NT, 1)
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
    ctrl.bgcount = 0;
  }
  if (t < TM) ctrl.sok[t] = 0;
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
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
  }
  // env-step role of the first TM threads: thread r is lane g of env el, and finishes tile row r
  const bool stepper = L.q == 0;
  const int el = L.r / G, g = L.r % G;
  const int env = env0 + el;
  const bool agent = stepper && el < nenv;
  const unsigned gmask = rware::group_mask<G>();
  uint8_t* rec = srec + el * c.stride;
  uint32_t mk = agent ? p.mask[(size_t)env * G + g] : 0u;
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
  // every env starts without a spare
  auto want_spare = [&]() {
    const uint32_t* k = reinterpret_cast<const uint32_t*>(rec + c.off_key);
    if (!ctrl.sok[el] || ctrl.skey[el][0] != k[0] || ctrl.skey[el][1] != k[1]) {
      const int i = atomicAdd(&ctrl.bgcount, 1);
      ctrl.bglist[i] = (uint16_t)el;
      ctrl.bgkey[i][0] = k[0];
      ctrl.bgkey[i][1] = k[1];
    }
  };
  if (agent && g == 0 && p.auto_reset) want_spare();
  __syncthreads();

  uint32_t phase = 0;
  bool pending_store = false;
  // warps whose 32 tile rows are all beyond this CTA's envs skip the hidden-layer epilogues (TMEM
  // reads are 64 B/clk per SM: half a tile of dead rows would cost as much as the live half)
  const bool live = (warp & 3) * 32 < rows_valid;
  for (int step = 0; step < p.T; ++step) {
    const int so = p.record ? step : 0, so1 = p.record ? step + 1 : 0;  // output slots of this step
    MAVA_RSTAMP(0);
    // ---- 1. X tile from the observation rows in shared memory
    expand_x_row(d, xt, L, reinterpret_cast<const signed char*>(sobs) + L.r * c.FR,
                 L.r < rows_valid, g);
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    MAVA_RSTAMP(1);
    // ---- 2. actor MLP
    if (mma_issuer()) {
      fence_after_sync();
      issue_gemm(tmem, xt, false, w1_tile(s_w, d.k1p), true, HID, d.k1p, false, &ctrl.mbar);
    }
    wait_mma(&ctrl.mbar, phase);
    MAVA_RSTAMP(2);
    phase ^= 1;
    if (live) hidden_epilogue(L, tmem, ht);
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    MAVA_RSTAMP(3);
    if (mma_issuer()) {
      fence_after_sync();
      issue_gemm(tmem, ht, false, w2_tile(s_w, d.k1p), true, HID, HCOLS, false, &ctrl.mbar);
    }
    wait_mma(&ctrl.mbar, phase);
    MAVA_RSTAMP(4);
    phase ^= 1;
    if (live) hidden_epilogue(L, tmem, ht);
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    MAVA_RSTAMP(5);
    if (mma_issuer()) {
      fence_after_sync();
      issue_gemm(tmem + HID, ht, false, w3_tile(s_w, d.k1p), true, NHEAD, HCOLS, false, &ctrl.mbar);
    }
    wait_mma(&ctrl.mbar, phase);
    MAVA_RSTAMP(6);
    phase ^= 1;
    // ---- 3 + 4. head epilogue and env step (threads 0 .. TM-1)
    bool needs_reset = false, replay = false;
    uint32_t opk[G];
#pragma unroll
    for (int j = 0; j < G; ++j) opk[j] = 0u;
    if (stepper) {
      float out[NHEAD];
      ld16(tmem + L.tmem_lane() + (uint32_t)HID, out);
      int act = 0;
      if (agent) {
        float mx = kF32Min;
#pragma unroll
        for (int j = 0; j < NHEAD; ++j) {
          if (j < d.out) {
            out[j] = ((mk >> j) & 1u) ? out[j] : kF32Min;
            mx = fmaxf(mx, out[j]);
          }
        }
        float se = 0.0f;
#pragma unroll
        for (int j = 0; j < NHEAD; ++j)
          if (j < d.out) se += expf(out[j] - mx);
        const float lse = mx + logf(se);
        // Gumbel arg-max on the noise prepared during the previous step
        const float* nz = snoise + (step & 1) * TM * NHEAD + L.r * NHEAD;
        float best = 0.0f;
#pragma unroll
        for (int j = 0; j < NHEAD; ++j) {
          if (j < d.out) {
            const float z = (p.greedy ? 0.0f : nz[j]) + out[j];
            if (j == 0 || z > best) { best = z; act = j; }
          }
        }
        float la = 0.0f;
#pragma unroll
        for (int j = 0; j < NHEAD; ++j)
          if (j == act) la = out[j] - lse;
        const size_t o = (size_t)so * ea_slot + (size_t)env * G + g;
        p.action[o] = (int8_t)act;
        p.logp[o] = la;
      }
      if (el < nenv)
        rware::step_group<G>(c, rec, g, gmask, agent, act, env, p.auto_reset,
                             p.reward + (size_t)so * ea_slot,
                             p.done + (size_t)step * p.num_envs,
                             p.ep_return + (size_t)step * p.num_envs,
                             p.ep_length + (size_t)step * p.num_envs, needs_reset, replay, opk);
      if (needs_reset && g == 0) ctrl.rlist[atomicAdd(&ctrl.rcount, 1)] = (uint16_t)el;
    } else {
      if (step + 1 < p.T && !p.greedy)
        make_noise(step + 1, t - TM, NT - TM);  // warps 4..15: next step's noise
      // ... and one spare record per warp (the list is stable during this phase; half-warp
      // generators were tried and are slower)
      const int i = warp - TM / 32;
      if (i < ctrl.bgcount) {
        const int bel = ctrl.bglist[i];
        const Key bk{ctrl.bgkey[i][0], ctrl.bgkey[i][1]};
        Key nk, unused;
        split2(bk, nk, unused);
        rware::generate<32>(c, sspare + bel * c.stride, nk, L.lane, 0xffffffffu);
        if (L.lane == 0) {
          ctrl.skey[bel][0] = bk.k0;
          ctrl.skey[bel][1] = bk.k1;
          ctrl.sok[bel] = 1;
        }
      }
    }
    // the previous step's observation block must have left shared memory before it is rewritten
    if (t == 0 && pending_store) rware::bulk_commit_wait_read();
    fence_before_sync();
    __syncthreads();
    MAVA_RSTAMP(7);
    // ---- finished envs: one regeneration per warp at a time, all 16 warps take part
    {
      const int nreset = ctrl.rcount;
      if (t == 0) ctrl.bgcount = 0;  // consumed in the phase above; refilled after the barrier below
      for (int i = warp; i < nreset; i += NWARPS) {
        const int rel = (int)ctrl.rlist[i];
        uint8_t* rrec = srec + rel * c.stride;
        const uint32_t* k = reinterpret_cast<const uint32_t*>(rrec + c.off_key);
        const uint32_t k0 = k[0], k1 = k[1];
        const bool hit = ctrl.sok[rel] && ctrl.skey[rel][0] == k0 && ctrl.skey[rel][1] == k1;
        __syncwarp();
        if (hit) {
          // the inner-env part of the record: [agents | queue | request bits | step | key] and the
          // shelf grid; the episode-metrics words in between stay
          const uint32_t* sp = reinterpret_cast<const uint32_t*>(sspare + rel * c.stride);
          uint32_t* dst = reinterpret_cast<uint32_t*>(rrec);
          for (int w = L.lane; w < (c.off_mkey >> 2); w += 32) dst[w] = sp[w];
          for (int w = L.lane; w < c.cells_words; w += 32)
            dst[(c.off_cells >> 2) + w] = sp[(c.off_cells >> 2) + w];
        } else {
          Key nk, unused;
          split2(Key{k0, k1}, nk, unused);
          rware::generate<32>(c, rrec, nk, L.lane, 0xffffffffu);
        }
        __syncwarp();
        if (L.lane == 0) ctrl.sok[rel] = 0;
      }
    }
    __syncthreads();
    MAVA_RSTAMP(8);
    if (t == 0) ctrl.rcount = 0;
    // ---- next observation rows and masks
    if (agent) {
      mk = rware::emit_row<G, R>(c, rec, g, sobs + L.r * c.FR, L.r & 1, replay, opk);
      p.mask[(size_t)so1 * ea_slot + (size_t)env * G + g] = (uint8_t)mk;
      if (g == 0 && p.auto_reset) want_spare();
    }
    fence_proxy_async();
    __syncthreads();
    MAVA_RSTAMP(9);
    pending_store = store_block(
        sobs, reinterpret_cast<uint8_t*>(p.view) + (size_t)so1 * obs_slot +
                  (size_t)env0 * G * c.FR,
        obs_bytes);
    if (t == 0 && pending_store) asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    MAVA_RSTAMP(10);
  }
  // ---- records back to HBM
  fence_proxy_async();
  __syncthreads();
  const bool bulk = store_block(srec, gstate, (int)rec_bytes);
  if (t == 0 && (bulk || pending_store)) rware::bulk_commit_wait_read();
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem);
}

template <int G>
int launch_rollout(const RolloutArgs& a, cudaStream_t s) {
  constexpr int EPC = TM / G;
  const size_t smem = (size_t)WImage{a.actor.k1p}.total() + tile_bytes(TM, a.actor.k1p) +
                      tile_bytes(TM, HCOLS) + (size_t)EPC * a.c.stride +
                      (size_t)round_up(EPC * G * a.c.FR, 16) + 2 * TM * NHEAD * 4 +
                      (size_t)EPC * a.c.stride + 128;  // ... noise, spare records
  static size_t configured[kMaxDevices] = {};
  if (int rc = ensure_dyn_smem(rware_rollout_kernel<G, 1>, smem, configured)) return rc;
  // spread the envs over the SMs: a CTA takes whole warps of envs (32 / G each), at most a tile
  RolloutArgs b = a;
  const int per_warp = 32 / G;
  int epc = round_up(ceil_div(a.num_envs, sm_count()), per_warp);
  b.epc = epc < per_warp ? per_warp : (epc > EPC ? EPC : epc);
  rware_rollout_kernel<G, 1><<<ceil_div(a.num_envs, b.epc), NT, smem, s>>>(b);
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
