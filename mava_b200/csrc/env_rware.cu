// Fused RobotWarehouse env-step for sm_100a: one kernel does what the reference runs as
// jax.vmap(env.step) through RecordEpisodeMetrics(AutoResetWrapper(AgentIDWrapper(RwareWrapper(
// RobotWarehouse)))) -- mava/systems/ppo/ff_mappo.py:88, mava/utils/make_env.py:69-83,
// mava/wrappers/{episode_metrics.py:78-111, auto_reset_wrapper.py:60-101, jumanji.py:135-144}.
// The inner dynamics follow the published Jumanji RobotWarehouse algorithm (third party, absent
// from the reference tree).
//
// Mapping: G lanes cooperate on one env (G = 2, 4 or 8; lane g < A is agent g), 256/G envs per CTA.
//   * The packed records of a CTA are contiguous in HBM: ONE bulk (TMA) copy stages them into
//     shared memory, they are updated in place there, and one bulk copy writes them back.
//   * The record holds the SHELVES grid itself (one byte per cell) instead of shelf coordinates,
//     so nothing is rebuilt per step; the AGENTS grid is never materialised: with agents on
//     distinct cells (always true while an episode is alive) the value the reference's sequential
//     grid writes leave in a cell is a closed-form function of (old cell, new cell, moved) of the
//     A agents, exchanged with sub-warp shuffles.
//   * Agents still take their turns one after the other on the shelf grid (lane i in turn i),
//     exactly like the reference's scan over agents.
//   * Each lane assembles its whole observation row in registers as packed 32-bit words (the
//     5-byte "other agent" blocks come from a bit mask expanded with one multiply per word) and
//     writes it to a staging block that leaves the CTA as one contiguous bulk store.
//   * Rare paths (delivery -> new request, episode end -> in-kernel reset) run threefry on all
//     lanes of the group / warp.
#include "env_rware.cuh"

namespace mava {
namespace {

using namespace rware;

constexpr int kThreads = 256;
#ifndef MAVA_RWARE_MINB
#define MAVA_RWARE_MINB 6  // resident CTAs per SM the step kernel's register budget is sized for
#endif

__host__ inline size_t smem_bytes(const RwareConst& c, int envs_per_cta) {
  return (size_t)envs_per_cta * c.stride + (size_t)round_up(envs_per_cta * c.A * c.FR, 16) + 16 +
         2 * (size_t)envs_per_cta;  // mbarrier + reset counter, reset queue
}

// Staged observation block / records -> HBM.  Bulk stores when size and address allow it.
__device__ __forceinline__ void store_block(const uint8_t* src, uint8_t* dst, int bytes) {
  // caller has synchronised the CTA and fenced the async proxy
  if ((((size_t)dst | (size_t)bytes) & 15) == 0) {
    if (threadIdx.x == 0) bulk_s2g(dst, src, (uint32_t)bytes);
  } else if ((((size_t)dst | (size_t)bytes) & 3) == 0) {
    for (int i = threadIdx.x; i < (bytes >> 2); i += blockDim.x)
      reinterpret_cast<uint32_t*>(dst)[i] = reinterpret_cast<const uint32_t*>(src)[i];
  } else {
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) dst[i] = src[i];
  }
}

template <int G, int R>
__global__ void __launch_bounds__(kThreads, MAVA_RWARE_MINB)
rware_step_kernel(const __grid_constant__ RwareConst c, uint8_t* __restrict__ state,
                  const int8_t* __restrict__ action, int8_t* __restrict__ view,
                  uint8_t* __restrict__ mask, float* __restrict__ reward,
                  uint8_t* __restrict__ done, float* __restrict__ ep_return,
                  int32_t* __restrict__ ep_length, int num_envs, int auto_reset) {
  extern __shared__ uint4 smem_raw[];
  constexpr int EPC = kThreads / G;
  uint8_t* srec = reinterpret_cast<uint8_t*>(smem_raw);
  uint8_t* sobs = srec + EPC * c.stride;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sobs + round_up(EPC * c.A * c.FR, 16));
  int* rcount = reinterpret_cast<int*>(bar + 1);
  uint16_t* rlist = reinterpret_cast<uint16_t*>(bar + 2);

  const int env0 = blockIdx.x * EPC;
  const int nenv = min(EPC, num_envs - env0);
  const uint32_t rec_bytes = (uint32_t)nenv * (uint32_t)c.stride;
  uint8_t* gstate = state + (size_t)env0 * c.stride;
  if (threadIdx.x == 0) {
    *rcount = 0;
    mbar_init(bar, 1);
    mbar_expect_tx(bar, rec_bytes);
    bulk_g2s(srec, gstate, rec_bytes, bar);
  }
  const int el = threadIdx.x / G, g = threadIdx.x % G;
  const int env = env0 + el;
  const unsigned gmask = group_mask<G>();
  const bool active = el < nenv;
  const bool agent = active && g < c.A;
  int act = agent ? (int)action[(size_t)env * c.A + g] : 0;
  __syncthreads();  // the barrier is initialised for everyone
  mbar_wait(bar, 0);

  uint8_t* rec = srec + el * c.stride;
  uint8_t* cells = rec + c.off_cells;
  bool needs_reset = false, replay = false;
  Key key{0u, 0u};
  uint32_t opk[G];
#pragma unroll
  for (int j = 0; j < G; ++j) opk[j] = 0u;
  rware::step_group<G>(c, rec, g, gmask, active, agent, act, env, auto_reset, reward, done,
                       ep_return, ep_length, needs_reset, replay, opk);
  // --- AutoResetWrapper: on the last step the state and observation are those of a fresh episode
  //     seeded with split(state.key)[0] (auto_reset_wrapper.py:74-75).  Episode ends are rare and a
  //     regeneration is a long dependent chain of threefry calls, so finished envs go into a CTA
  //     queue and every warp takes one at a time with all 32 lanes: the tail is one regeneration
  //     per warp instead of all of a warp's finished envs back to back.
  if (needs_reset && g == 0) rlist[atomicAdd(rcount, 1)] = (uint16_t)el;
  __syncthreads();
  {
    const int nreset = *rcount;
    const unsigned lane = threadIdx.x & 31u;
    for (int i = threadIdx.x >> 5; i < nreset; i += kThreads / 32) {
      uint8_t* rrec = srec + (int)rlist[i] * c.stride;
      const uint32_t* k = reinterpret_cast<const uint32_t*>(rrec + c.off_key);
      Key nk, unused;
      split2(Key{k[0], k[1]}, nk, unused);
      __syncwarp();
      generate<32>(c, rrec, nk, (int)lane, 0xffffffffu);
    }
  }
  __syncthreads();
  // --- next observation and action mask
  if (agent) {
    const int r = el * c.A + g;
    mask[(size_t)env * c.A + g] = emit_row<G, R>(c, rec, g, sobs + r * c.FR, r & 1, replay, opk);
  }
  fence_proxy_async();
  __syncthreads();
  store_block(sobs, reinterpret_cast<uint8_t*>(view) + (size_t)env0 * c.A * c.FR,
              nenv * c.A * c.FR);
  store_block(srec, gstate, (int)rec_bytes);
  if (threadIdx.x == 0) bulk_commit_wait_read();
}

// vmap(env.reset)(keys): RecordEpisodeMetrics.reset splits the key, the inner generator builds the
// state, metrics start at zero (episode_metrics.py:59-76).
template <int G, int R>
__global__ void __launch_bounds__(kThreads)
rware_reset_kernel(const __grid_constant__ RwareConst c, const uint32_t* __restrict__ keys,
                   uint8_t* __restrict__ state, int8_t* __restrict__ view,
                   uint8_t* __restrict__ mask, int num_envs) {
  extern __shared__ uint4 smem_raw[];
  constexpr int EPC = kThreads / G;
  uint8_t* srec = reinterpret_cast<uint8_t*>(smem_raw);
  uint8_t* sobs = srec + EPC * c.stride;
  const int env0 = blockIdx.x * EPC;
  const int nenv = min(EPC, num_envs - env0);
  const int el = threadIdx.x / G, g = threadIdx.x % G;
  const int env = env0 + el;
  const unsigned gmask = group_mask<G>();
  uint32_t opk[G];
#pragma unroll
  for (int j = 0; j < G; ++j) opk[j] = 0u;
  if (el < nenv) {
    uint8_t* rec = srec + el * c.stride;
    for (int i = g; i < (c.stride >> 2); i += G) reinterpret_cast<uint32_t*>(rec)[i] = 0u;
    __syncwarp(gmask);
    Key key{keys[2 * (size_t)env], keys[2 * (size_t)env + 1]}, reset_key;
    split2(key, key, reset_key);
    generate<G>(c, rec, reset_key, g, gmask);
    if (g == 0) {
      uint32_t* mk = reinterpret_cast<uint32_t*>(rec + c.off_mkey);
      mk[0] = key.k0;
      mk[1] = key.k1;
    }
    __syncwarp(gmask);
    if (g < c.A) {
      const int r = el * c.A + g;
      mask[(size_t)env * c.A + g] =
          emit_row<G, R>(c, rec, g, sobs + r * c.FR, r & 1, false, opk);
    }
  }
  fence_proxy_async();
  __syncthreads();
  store_block(sobs, reinterpret_cast<uint8_t*>(view) + (size_t)env0 * c.A * c.FR,
              nenv * c.A * c.FR);
  store_block(srec, state + (size_t)env0 * c.stride, nenv * c.stride);
  if (threadIdx.x == 0) bulk_commit_wait_read();
}

__global__ void rware_peek_kernel(const __grid_constant__ RwareConst c,
                                  const uint8_t* __restrict__ state, int field,
                                  int32_t* __restrict__ out, int num_envs) {
  const int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= num_envs) return;
  const uint8_t* r = state + (size_t)env * c.stride;
  if (field == 0) {
    out[env] = (int32_t)(*reinterpret_cast<const uint32_t*>(r + c.off_step));
  } else if (field == 1) {
    const uint32_t* k = reinterpret_cast<const uint32_t*>(r + c.off_key);
    out[2 * env] = (int32_t)k[0];
    out[2 * env + 1] = (int32_t)k[1];
  } else if (field == 2) {
    const uint32_t* agents = reinterpret_cast<const uint32_t*>(r + c.off_agents);
    for (int i = 0; i < c.A; ++i) {
      int32_t* o = out + ((size_t)env * c.A + i) * 4;
      o[0] = agents[i] & 0xff;
      o[1] = (agents[i] >> 8) & 0xff;
      o[2] = (agents[i] >> 16) & 0xff;
      o[3] = agents[i] >> 24;
    }
  } else if (field == 3) {  // shelves: x, y, requested (positions read back from the grid)
    const uint8_t* cells = r + c.off_cells;
    for (int cell = 0; cell < c.HW; ++cell) {
      const int sid = cells[cell];
      if (sid != 0) {
        int32_t* o = out + ((size_t)env * c.n + sid - 1) * 3;
        o[0] = cell / c.W;
        o[1] = cell % c.W;
        o[2] = (reinterpret_cast<const uint32_t*>(r + c.off_reqbits)[(sid - 1) >> 5] >>
                ((sid - 1) & 31)) & 1u;
      }
    }
  } else if (field == 4) {  // request queue
    for (int q = 0; q < c.Q; ++q) out[(size_t)env * c.Q + q] = r[c.off_queue + q];
  }
}

template <typename K>
int prepare(K kernel, size_t smem) {
  static size_t configured[kMaxDevices] = {};  // one instance per kernel type
  return ensure_dyn_smem(kernel, smem, configured);
}

}  // namespace

int rware_create(const mava_rware_config* cfg, mava_env_s* env) {
  RwareConst& c = env->rw;
  MAVA_CHECK_ARG(cfg->column_height >= 1 && cfg->shelf_rows >= 1 && cfg->shelf_columns >= 1);
  MAVA_CHECK_ARG(cfg->shelf_columns % 2 == 1);
  MAVA_CHECK_ARG(cfg->num_agents >= 1 && cfg->num_agents <= kMaxAgents);
  MAVA_CHECK_ARG(cfg->request_queue_size >= 1 && cfg->request_queue_size <= kMaxQueue);
  MAVA_CHECK_ARG(cfg->sensor_range >= 1 && cfg->sensor_range <= 2);
  MAVA_CHECK_ARG(cfg->time_limit >= 1);
  c.H = (cfg->column_height + 1) * cfg->shelf_rows + 2;
  c.W = 3 * cfg->shelf_columns + 1;
  c.HW = c.H * c.W;
  if (c.HW > kMaxCells || c.H > 127 || c.W > 127) return MAVA_E_UNSUPPORTED;
  c.A = cfg->num_agents;
  c.Q = cfg->request_queue_size;
  c.R = cfg->sensor_range;
  c.time_limit = cfg->time_limit;
  const int loc = (2 * c.R + 1) * (2 * c.R + 1);
  c.FR = 8 + (loc - 1) * 5 + loc * 2;
  for (int i = 0; i < kMaxCells / 32; ++i) c.highway[i] = 0u;
  c.n = 0;
  for (int r = 0; r < c.H; ++r) {
    for (int col = 0; col < c.W; ++col) {
      const bool hw = (col % 3 == 0) || (r % (cfg->column_height + 1) == 0) || (r == c.H - 1) ||
                      ((r > c.H - (cfg->column_height + 3)) &&
                       (col == c.W / 2 - 1 || col == c.W / 2));
      const int cell = r * c.W + col;
      if (hw) {
        c.highway[cell >> 5] |= 1u << (cell & 31);
      } else {
        if (c.n >= kMaxShelves) return MAVA_E_UNSUPPORTED;
        c.shelf_home[c.n++] = (uint16_t)cell;
      }
    }
  }
  MAVA_CHECK_ARG(c.Q <= c.n && c.A <= c.HW);
  c.goal[0] = (c.H - 1) * c.W + c.W / 2 - 1;
  c.goal[1] = (c.H - 1) * c.W + c.W / 2;
  c.cells_words = (c.HW + 3) / 4;
  c.req_words = (c.n + 31) / 32;
  int o = 0;
  c.off_agents = o; o += 4 * c.A;
  c.off_queue = o; o += round_up(c.Q, 4);
  c.off_reqbits = o; o += 4 * c.req_words;
  c.off_step = o; o += 4;
  c.off_key = o; o += 8;
  c.off_mkey = o; o += 8;
  c.off_run_ret = o; o += 4;
  c.off_run_len = o; o += 4;
  c.off_ep_ret = o; o += 4;
  c.off_ep_len = o; o += 4;
  c.off_cells = o; o += 4 * c.cells_words;
  c.stride = round_up(o, 16);
  if ((c.stride / 16) % 2 == 0) c.stride += 16;  // odd multiple of 16: bank-conflict-free records

  mava_env_dims& d = env->dims;
  d.kind = MAVA_ENV_RWARE;
  d.num_agents = c.A;
  d.view_dim = c.FR;
  d.num_actions = 5;
  d.state_stride = c.stride;
  d.time_limit = c.time_limit;
  d.grid_h = c.H;
  d.grid_w = c.W;
  d.aux0 = c.n;
  d.aux1 = c.Q;
  // 2*S_state + A (action) + A*FR (obs i8) + A (mask) + 4 (reward) + 1 + 9
  const int s_state = 4 * c.A + 3 * c.n + c.Q + 2 + 8 + 24;
  d.algo_bytes_per_step = 2 * s_state + c.A + c.A * c.FR + c.A + 4 + 1 + 9;
  return 0;
}

#define MAVA_RWARE_LAUNCH(KERNEL, G_, R_, ...)                                              \
  do {                                                                                      \
    const size_t smem = smem_bytes(c, kThreads / G_);                                       \
    int rc = prepare(KERNEL<G_, R_>, smem);                                                 \
    if (rc) return rc;                                                                      \
    KERNEL<G_, R_><<<ceil_div(num_envs, kThreads / G_), kThreads, smem, s>>>(__VA_ARGS__);  \
  } while (0)

#define MAVA_RWARE_DISPATCH(KERNEL, ...)                                   \
  do {                                                                     \
    if (c.R == 1) {                                                        \
      if (c.A <= 2) MAVA_RWARE_LAUNCH(KERNEL, 2, 1, __VA_ARGS__);          \
      else if (c.A <= 4) MAVA_RWARE_LAUNCH(KERNEL, 4, 1, __VA_ARGS__);     \
      else MAVA_RWARE_LAUNCH(KERNEL, 8, 1, __VA_ARGS__);                   \
    } else {                                                               \
      if (c.A <= 2) MAVA_RWARE_LAUNCH(KERNEL, 2, 2, __VA_ARGS__);          \
      else if (c.A <= 4) MAVA_RWARE_LAUNCH(KERNEL, 4, 2, __VA_ARGS__);     \
      else MAVA_RWARE_LAUNCH(KERNEL, 8, 2, __VA_ARGS__);                   \
    }                                                                      \
  } while (0)

int rware_reset(const mava_env_s* env, const uint32_t* keys, uint8_t* state, int8_t* view,
                uint8_t* mask, int num_envs, cudaStream_t s) {
  const RwareConst& c = env->rw;
  MAVA_RWARE_DISPATCH(rware_reset_kernel, c, keys, state, view, mask, num_envs);
  return launch_status();
}

int rware_step(const mava_env_s* env, uint8_t* state, const int8_t* action, int8_t* view,
               uint8_t* mask, float* reward, uint8_t* done, float* ep_return, int32_t* ep_length,
               int num_envs, int auto_reset, cudaStream_t s) {
  const RwareConst& c = env->rw;
  MAVA_RWARE_DISPATCH(rware_step_kernel, c, state, action, view, mask, reward, done, ep_return,
                      ep_length, num_envs, auto_reset);
  return launch_status();
}

int rware_peek(const mava_env_s* env, const uint8_t* state, int field, int32_t* out, int num_envs,
               cudaStream_t s) {
  MAVA_CHECK_ARG(field >= 0 && field <= 4);
  rware_peek_kernel<<<ceil_div(num_envs, 128), 128, 0, s>>>(env->rw, state, field, out, num_envs);
  return launch_status();
}

}  // namespace mava
