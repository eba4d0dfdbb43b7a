// Fused RobotWarehouse env-step for sm_100a: one kernel does what the reference runs as
// jax.vmap(env.step) through RecordEpisodeMetrics(AutoResetWrapper(AgentIDWrapper(RwareWrapper(
// RobotWarehouse)))) -- mava/systems/ppo/ff_mappo.py:88, mava/utils/make_env.py:69-83,
// mava/wrappers/{episode_metrics.py:78-111, auto_reset_wrapper.py:60-101, jumanji.py:135-144}.
// The inner dynamics follow the published Jumanji RobotWarehouse algorithm (third party, absent
// from the reference tree).
//
// Mapping: G lanes cooperate on one env (G = 4 or 8, lane g < A is agent g), 256/G envs per CTA.
// The packed per-env records of a CTA are contiguous in HBM: they are staged into shared memory
// with 16-byte vector loads, worked on there (occupancy grids are rebuilt in shared memory, never
// stored), and written back the same way; observations are assembled in shared memory and stored
// as one contiguous int8 block per CTA.  Agents are advanced one after the other (lane i acts in
// turn i) exactly like the reference's scan over agents; collisions use a sub-warp ballot; the
// rare paths (delivery -> new request, episode end -> in-kernel reset) run threefry on all G lanes.
#include "env.cuh"
#include "prng.cuh"

namespace mava {
namespace {

constexpr int kThreads = 256;

template <int G>
__device__ __forceinline__ unsigned group_mask() {
  if (G == 32) return 0xffffffffu;
  const unsigned lane = threadIdx.x & 31u;
  return ((1u << G) - 1u) << ((lane / G) * G);
}

template <int G>
__device__ __forceinline__ unsigned long long group_min(unsigned long long v, unsigned gmask) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) {
    unsigned long long w = __shfl_xor_sync(gmask, v, o, G);
    v = w < v ? w : v;
  }
  return v;
}

// Per-env shared-memory working set.
struct EnvSmem {
  uint8_t* rec;   // packed record
  uint8_t* gsh;   // shelf id (+1) per cell
  uint8_t* gag;   // agent id (+1) per cell
};

__device__ __forceinline__ bool is_highway(const RwareConst& c, int cell) {
  return (c.highway[cell >> 5] >> (cell & 31)) & 1u;
}

__device__ __forceinline__ void forward_cell(const RwareConst& c, int x, int y, int d, int& nx,
                                             int& ny) {
  nx = x;
  ny = y;
  if (d == 0) nx = max(0, x - 1);
  else if (d == 1) ny = min(c.W - 1, y + 1);
  else if (d == 2) nx = min(c.H - 1, x + 1);
  else ny = max(0, y - 1);
}

// Rebuild both occupancy grids from the record.
template <int G>
__device__ __forceinline__ void build_grids(const RwareConst& c, const EnvSmem& m, int g,
                                            unsigned gmask) {
  const int words = (c.HW + 3) >> 2;
  uint32_t* s32 = reinterpret_cast<uint32_t*>(m.gsh);
  uint32_t* a32 = reinterpret_cast<uint32_t*>(m.gag);
  for (int i = g; i < words; i += G) {
    s32[i] = 0u;
    a32[i] = 0u;
  }
  __syncwarp(gmask);
  const uint8_t* sx = m.rec + c.off_sx;
  const uint8_t* sy = m.rec + c.off_sy;
  for (int s = g; s < c.n; s += G) m.gsh[sx[s] * c.W + sy[s]] = (uint8_t)(s + 1);
  if (g < c.A) m.gag[m.rec[c.off_ax + g] * c.W + m.rec[c.off_ay + g]] = (uint8_t)(g + 1);
  __syncwarp(gmask);
}

// K smallest of the composites (random_bits(sub, size)[i] << 32 | i), i.e. the first K entries of
// jax.random.permutation-by-stable-sort.  Every lane returns the same out[].
template <int G, int KMAX>
__device__ __forceinline__ void smallest_k(Key sub, int size, int K, int g, unsigned gmask,
                                           unsigned long long (&out)[KMAX]) {
  unsigned long long top[KMAX];
#pragma unroll
  for (int j = 0; j < KMAX; ++j) top[j] = ~0ull;
  const int half = (size + 1) >> 1;
  for (int p = g; p < half; p += G) {
    uint32_t lo, hi;
    random_bits_pair(sub, (uint32_t)p, (uint32_t)size, lo, hi);
    unsigned long long v = ((unsigned long long)lo << 32) | (unsigned)p;
#pragma unroll
    for (int j = 0; j < KMAX; ++j) {
      if (v < top[j]) {
        unsigned long long t = top[j];
        top[j] = v;
        v = t;
      }
    }
    if (p + half < size) {
      v = ((unsigned long long)hi << 32) | (unsigned)(p + half);
#pragma unroll
      for (int j = 0; j < KMAX; ++j) {
        if (v < top[j]) {
          unsigned long long t = top[j];
          top[j] = v;
          v = t;
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < KMAX; ++r) {
    out[r] = ~0ull;
    if (r < K) {
      unsigned long long mn = group_min<G>(top[0], gmask);
      if (top[0] == mn) {
#pragma unroll
        for (int j = 0; j + 1 < KMAX; ++j) top[j] = top[j + 1];
        top[KMAX - 1] = ~0ull;
      }
      out[r] = mn;
    }
  }
}

// jumanji RandomGenerator.__call__: agents on distinct random cells, random directions, shelves on
// their home cells, Q distinct requested shelves.  Writes the inner-env part of the record and
// returns the key left over (State.key).
template <int G>
__device__ __forceinline__ void generate(const RwareConst& c, const EnvSmem& m, Key key, int g,
                                         unsigned gmask) {
  Key pos_key, dir_key, q_key, unused, sub;
  unsigned long long pick[kMaxAgents];
  split2(key, key, pos_key);
  split2(pos_key, unused, sub);
  smallest_k<G, kMaxAgents>(sub, c.HW, c.A, g, gmask, pick);
  split2(key, key, dir_key);
  Key d_hi, d_lo;
  split2(dir_key, d_hi, d_lo);  // randint: span 4 -> only the low-bits draw matters
  if (g == 0) {
#pragma unroll
    for (int i = 0; i < kMaxAgents; ++i) {
      if (i < c.A) {
        int cell = (int)(pick[i] & 0xffffffffull);
        m.rec[c.off_ax + i] = (uint8_t)(cell / c.W);
        m.rec[c.off_ay + i] = (uint8_t)(cell % c.W);
        m.rec[c.off_dir + i] = (uint8_t)(random_bits_at(d_lo, (uint32_t)i, (uint32_t)c.A) & 3u);
        m.rec[c.off_carry + i] = 0;
      }
    }
  }
  split2(key, key, q_key);
  split2(q_key, unused, sub);
  unsigned long long qpick[kMaxQueue];
  smallest_k<G, kMaxQueue>(sub, c.n, c.Q, g, gmask, qpick);
  for (int s = g; s < c.n; s += G) {
    int cell = c.shelf_home[s];
    m.rec[c.off_sx + s] = (uint8_t)(cell / c.W);
    m.rec[c.off_sy + s] = (uint8_t)(cell % c.W);
    m.rec[c.off_req + s] = 0;
  }
  __syncwarp(gmask);
  if (g == 0) {
#pragma unroll
    for (int i = 0; i < kMaxQueue; ++i) {
      if (i < c.Q) {
        int s = (int)(qpick[i] & 0xffffffffull);
        m.rec[c.off_queue + i] = (uint8_t)s;
        m.rec[c.off_req + s] = 1;
      }
    }
    *reinterpret_cast<uint32_t*>(m.rec + c.off_step) = 0u;
    uint32_t* k = reinterpret_cast<uint32_t*>(m.rec + c.off_key);
    k[0] = key.k0;
    k[1] = key.k1;
  }
  __syncwarp(gmask);
}

// Action mask bits + int8 observation row of agent g (jumanji utils.make_agent_observation and
// compute_action_mask); grids must be current.
__device__ __forceinline__ uint8_t emit_obs_and_mask(const RwareConst& c, const EnvSmem& m, int g,
                                                     int8_t* row) {
  const int x = m.rec[c.off_ax + g], y = m.rec[c.off_ay + g];
  const int d = m.rec[c.off_dir + g], carry = m.rec[c.off_carry + g];
  row[0] = (int8_t)x;
  row[1] = (int8_t)y;
  row[2] = (int8_t)carry;
  row[3] = d == 0;
  row[4] = d == 1;
  row[5] = d == 2;
  row[6] = d == 3;
  row[7] = is_highway(c, x * c.W + y);
  int ia = 8;
  const int loc = (2 * c.R + 1) * (2 * c.R + 1);
  int is = 8 + (loc - 1) * 5;
  for (int dx = -c.R; dx <= c.R; ++dx) {
    for (int dy = -c.R; dy <= c.R; ++dy) {
      const int cx = x + dx, cy = y + dy;
      const bool inside = cx >= 0 && cx < c.H && cy >= 0 && cy < c.W;
      const int cell = cx * c.W + cy;
      const int aid = inside ? m.gag[cell] : 0;
      const int sid = inside ? m.gsh[cell] : 0;
      if (dx != 0 || dy != 0) {
        const int od = aid ? m.rec[c.off_dir + aid - 1] : -1;
        row[ia + 0] = aid != 0;
        row[ia + 1] = od == 0;
        row[ia + 2] = od == 1;
        row[ia + 3] = od == 2;
        row[ia + 4] = od == 3;
        ia += 5;
      }
      row[is + 0] = sid != 0;
      row[is + 1] = sid ? (int8_t)m.rec[c.off_req + sid - 1] : 0;
      is += 2;
    }
  }
  int nx, ny;
  forward_cell(c, x, y, d, nx, ny);
  const bool stuck = nx == x && ny == y;
  const bool blocked = carry && m.gsh[nx * c.W + ny] != 0;
  return (uint8_t)(0x1Du | ((stuck || blocked) ? 0u : 0x2u));
}

struct SmemLayout {
  int rec_stride, grid_stride, obs_stride, per_cta_rec, per_cta_grid;
};

__host__ __device__ inline SmemLayout smem_layout(const RwareConst& c, int envs_per_cta) {
  SmemLayout L;
  L.rec_stride = c.stride + 16;
  L.grid_stride = round_up(c.HW, 4) + 4;
  L.obs_stride = c.A * c.FR;
  L.per_cta_rec = envs_per_cta * L.rec_stride;
  L.per_cta_grid = envs_per_cta * L.grid_stride;
  return L;
}

__host__ inline size_t smem_bytes(const RwareConst& c, int envs_per_cta) {
  SmemLayout L = smem_layout(c, envs_per_cta);
  return (size_t)L.per_cta_rec + 2 * (size_t)L.per_cta_grid +
         (size_t)round_up(envs_per_cta * L.obs_stride, 16);
}

// Coalesced CTA-wide copies between HBM and the staged records / observation block.
__device__ __forceinline__ void load_records(const RwareConst& c, const SmemLayout& L,
                                             uint8_t* srec, const uint8_t* state, int env0,
                                             int nenv) {
  const int v = c.stride >> 4;
  const uint4* src = reinterpret_cast<const uint4*>(state + (size_t)env0 * c.stride);
  for (int i = threadIdx.x; i < nenv * v; i += blockDim.x) {
    const int e = i / v, w = i - e * v;
    reinterpret_cast<uint4*>(srec + e * L.rec_stride)[w] = src[i];
  }
}

__device__ __forceinline__ void store_records(const RwareConst& c, const SmemLayout& L,
                                              const uint8_t* srec, uint8_t* state, int env0,
                                              int nenv) {
  const int v = c.stride >> 4;
  uint4* dst = reinterpret_cast<uint4*>(state + (size_t)env0 * c.stride);
  for (int i = threadIdx.x; i < nenv * v; i += blockDim.x) {
    const int e = i / v, w = i - e * v;
    dst[i] = reinterpret_cast<const uint4*>(srec + e * L.rec_stride)[w];
  }
}

__device__ __forceinline__ void store_obs(const RwareConst& c, const uint8_t* sobs, int8_t* view,
                                          int env0, int nenv) {
  const size_t base = (size_t)env0 * c.A * c.FR;
  const int bytes = nenv * c.A * c.FR;
  if (((base | (size_t)bytes) & 15) == 0) {
    uint4* dst = reinterpret_cast<uint4*>(view + base);
    for (int i = threadIdx.x; i < (bytes >> 4); i += blockDim.x)
      dst[i] = reinterpret_cast<const uint4*>(sobs)[i];
  } else if (((base | (size_t)bytes) & 3) == 0) {
    uint32_t* dst = reinterpret_cast<uint32_t*>(view + base);
    for (int i = threadIdx.x; i < (bytes >> 2); i += blockDim.x)
      dst[i] = reinterpret_cast<const uint32_t*>(sobs)[i];
  } else {
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) view[base + i] = (int8_t)sobs[i];
  }
}

template <int G>
__global__ void __launch_bounds__(kThreads)
rware_step_kernel(const __grid_constant__ RwareConst c, uint8_t* __restrict__ state,
                  const int8_t* __restrict__ action, int8_t* __restrict__ view,
                  uint8_t* __restrict__ mask, float* __restrict__ reward,
                  uint8_t* __restrict__ done, float* __restrict__ ep_return,
                  int32_t* __restrict__ ep_length, int num_envs, int auto_reset) {
  extern __shared__ uint4 smem_raw[];
  constexpr int EPC = kThreads / G;
  const SmemLayout L = smem_layout(c, EPC);
  uint8_t* srec = reinterpret_cast<uint8_t*>(smem_raw);
  uint8_t* sgsh = srec + L.per_cta_rec;
  uint8_t* sgag = sgsh + L.per_cta_grid;
  uint8_t* sobs = sgag + L.per_cta_grid;

  const int env0 = blockIdx.x * EPC;
  const int nenv = min(EPC, num_envs - env0);
  load_records(c, L, srec, state, env0, nenv);
  __syncthreads();

  const int el = threadIdx.x / G, g = threadIdx.x % G;
  const int env = env0 + el;
  const unsigned gmask = group_mask<G>();
  const bool active = el < nenv;
  const EnvSmem m{srec + el * L.rec_stride, sgsh + el * L.grid_stride, sgag + el * L.grid_stride};
  bool needs_reset = false;
  Key key{0u, 0u};
  if (active) {
    build_grids<G>(c, m, g, gmask);

    // --- validate the action against the mask of the current state (utils.get_valid_actions)
    int x = 0, y = 0, d = 0, carry = 0, act = 0;
    if (g < c.A) {
      x = m.rec[c.off_ax + g];
      y = m.rec[c.off_ay + g];
      d = m.rec[c.off_dir + g];
      carry = m.rec[c.off_carry + g];
      act = action[(size_t)env * c.A + g];
      if (act == 1) {
        int nx, ny;
        forward_cell(c, x, y, d, nx, ny);
        const bool stuck = nx == x && ny == y;
        const bool blocked = carry && m.gsh[nx * c.W + ny] != 0;
        if (stuck || blocked) act = 0;
      }
    }
    // --- agents act one after the other on the shared grids (scan over agents in env.step)
    for (int i = 0; i < c.A; ++i) {
      if (g == i) {
        const int cell = x * c.W + y;
        if (act == 2) {
          d = (d + 3) & 3;
        } else if (act == 3) {
          d = (d + 1) & 3;
        } else if (act == 1) {
          int nx, ny;
          forward_cell(c, x, y, d, nx, ny);
          const int ncell = nx * c.W + ny;
          m.gag[cell] = 0;
          m.gag[ncell] = (uint8_t)(g + 1);
          if (carry) {
            const int sid = m.gsh[cell];
            const int s = sid ? sid - 1 : c.n - 1;  // jax .at[-1] wraps to the last shelf
            m.rec[c.off_sx + s] = (uint8_t)nx;
            m.rec[c.off_sy + s] = (uint8_t)ny;
            m.gsh[cell] = 0;
            m.gsh[ncell] = (uint8_t)sid;
          }
          x = nx;
          y = ny;
        } else if (act == 4) {
          const int sid = m.gsh[cell];
          if (!carry) {
            if (sid != 0) carry = 1;
          } else if (!is_highway(c, cell)) {
            carry = 0;
          }
        }
        m.rec[c.off_ax + g] = (uint8_t)x;
        m.rec[c.off_ay + g] = (uint8_t)y;
        m.rec[c.off_dir + g] = (uint8_t)d;
        m.rec[c.off_carry + g] = (uint8_t)carry;
      }
      __syncwarp(gmask);
    }
    // --- collision: the id left on my cell is not mine (utils.is_collision)
    const bool my_col = g < c.A && m.gag[x * c.W + y] != (uint8_t)(g + 1);
    const bool collision = (__ballot_sync(gmask, my_col) & gmask) != 0u;

    // --- deliveries at the goal cells; a delivered request is replaced by a uniformly drawn
    //     shelf that is not in the queue (env._update_reward_and_request_queue)
    float rew = 0.0f;
    {
      const uint32_t* k = reinterpret_cast<const uint32_t*>(m.rec + c.off_key);
      key = Key{k[0], k[1]};
    }
    for (int gi = 0; gi < 2; ++gi) {
      const int sid = m.gsh[c.goal[gi]];
      if (sid != 0 && m.rec[c.off_req + sid - 1] == 1) {
        Key rkey, unused, sub;
        split2(key, key, rkey);
        split2(rkey, unused, sub);
        const int msize = c.n - c.Q;
        unsigned long long best = ~0ull;
        for (int s = g; s < c.n; s += G) {
          int below = 0;
          bool inq = false;
          for (int q = 0; q < c.Q; ++q) {
            const int qs = m.rec[c.off_queue + q];
            inq |= qs == s;
            below += qs < s;
          }
          if (!inq) {
            const int p = s - below;  // position in the sorted not-in-queue list
            const uint32_t b = random_bits_at(sub, (uint32_t)p, (uint32_t)msize);
            const unsigned long long v =
                ((unsigned long long)b << 32) | ((unsigned long long)p << 16) | (unsigned)s;
            best = v < best ? v : best;
          }
        }
        best = group_min<G>(best, gmask);
        const int new_req = (int)(best & 0xffffull);
        __syncwarp(gmask);
        if (g == 0) {
          for (int q = 0; q < c.Q; ++q) {
            if (m.rec[c.off_queue + q] == sid - 1) {
              m.rec[c.off_queue + q] = (uint8_t)new_req;
              break;
            }
          }
          m.rec[c.off_req + sid - 1] = 0;
          m.rec[c.off_req + new_req] = 1;
        }
        rew += 1.0f;
        __syncwarp(gmask);
      }
    }
    // --- step count, termination
    uint32_t* pstep = reinterpret_cast<uint32_t*>(m.rec + c.off_step);
    const int step = (int)(*pstep) + 1;
    const bool is_done = collision || step >= c.time_limit;
    __syncwarp(gmask);
    if (g == 0) {
      *pstep = (uint32_t)step;
      uint32_t* k = reinterpret_cast<uint32_t*>(m.rec + c.off_key);
      k[0] = key.k0;
      k[1] = key.k1;
      // RecordEpisodeMetrics.step (episode_metrics.py:83-110)
      float* run_ret = reinterpret_cast<float*>(m.rec + c.off_run_ret);
      int32_t* run_len = reinterpret_cast<int32_t*>(m.rec + c.off_run_len);
      float* e_ret = reinterpret_cast<float*>(m.rec + c.off_ep_ret);
      int32_t* e_len = reinterpret_cast<int32_t*>(m.rec + c.off_ep_len);
      const float new_ret = *run_ret + rew;  // mean over agents of a shared reward
      const int32_t new_len = *run_len + 1;
      const float nd = is_done ? 0.0f : 1.0f, dd = is_done ? 1.0f : 0.0f;
      const float ret_info = *e_ret * nd + new_ret * dd;
      const int32_t len_info = is_done ? new_len : *e_len;
      *run_ret = new_ret * nd;
      *run_len = is_done ? 0 : new_len;
      *e_ret = ret_info;
      *e_len = len_info;
      done[env] = is_done ? 1 : 0;
      ep_return[env] = ret_info;
      ep_length[env] = len_info;
    }
    if (g < c.A) reward[(size_t)env * c.A + g] = rew;
    needs_reset = is_done && auto_reset != 0;
  }
  // --- AutoResetWrapper: on the last step the state and observation are those of a fresh episode
  //     seeded with split(state.key)[0] (auto_reset_wrapper.py:74-75).  Episode ends are rare, so
  //     the whole warp regenerates one finished env at a time: 32 lanes share the threefry draws
  //     and the top-k selections instead of leaving G lanes with the long tail.
  __syncwarp();
  {
    const unsigned lane = threadIdx.x & 31u;
    unsigned pending = __ballot_sync(0xffffffffu, needs_reset && g == 0);
    while (pending) {
      const int leader = __ffs(pending) - 1;
      pending &= pending - 1;
      const int rel = __shfl_sync(0xffffffffu, el, leader);
      const uint32_t k0 = __shfl_sync(0xffffffffu, key.k0, leader);
      const uint32_t k1 = __shfl_sync(0xffffffffu, key.k1, leader);
      const EnvSmem mr{srec + rel * L.rec_stride, sgsh + rel * L.grid_stride,
                       sgag + rel * L.grid_stride};
      Key nk, unused;
      split2(Key{k0, k1}, nk, unused);
      generate<32>(c, mr, nk, (int)lane, 0xffffffffu);
      build_grids<32>(c, mr, (int)lane, 0xffffffffu);
    }
  }
  __syncwarp();
  // --- next observation and action mask
  if (active && g < c.A) {
    int8_t* row = reinterpret_cast<int8_t*>(sobs) + (el * c.A + g) * c.FR;
    const uint8_t mk = emit_obs_and_mask(c, m, g, row);
    mask[(size_t)env * c.A + g] = mk;
  }
  __syncthreads();
  store_obs(c, sobs, view, env0, nenv);
  store_records(c, L, srec, state, env0, nenv);
}

// vmap(env.reset)(keys): RecordEpisodeMetrics.reset splits the key, the inner generator builds the
// state, metrics start at zero (episode_metrics.py:59-76).
template <int G>
__global__ void __launch_bounds__(kThreads)
rware_reset_kernel(const __grid_constant__ RwareConst c, const uint32_t* __restrict__ keys,
                   uint8_t* __restrict__ state, int8_t* __restrict__ view,
                   uint8_t* __restrict__ mask, int num_envs) {
  extern __shared__ uint4 smem_raw[];
  constexpr int EPC = kThreads / G;
  const SmemLayout L = smem_layout(c, EPC);
  uint8_t* srec = reinterpret_cast<uint8_t*>(smem_raw);
  uint8_t* sgsh = srec + L.per_cta_rec;
  uint8_t* sgag = sgsh + L.per_cta_grid;
  uint8_t* sobs = sgag + L.per_cta_grid;
  const int env0 = blockIdx.x * EPC;
  const int nenv = min(EPC, num_envs - env0);
  const int el = threadIdx.x / G, g = threadIdx.x % G;
  const int env = env0 + el;
  const unsigned gmask = group_mask<G>();
  if (el < nenv) {
    EnvSmem m{srec + el * L.rec_stride, sgsh + el * L.grid_stride, sgag + el * L.grid_stride};
    for (int i = g; i < (c.stride >> 2); i += G) reinterpret_cast<uint32_t*>(m.rec)[i] = 0u;
    __syncwarp(gmask);
    Key key{keys[2 * (size_t)env], keys[2 * (size_t)env + 1]}, reset_key;
    split2(key, key, reset_key);
    generate<G>(c, m, reset_key, g, gmask);
    if (g == 0) {
      uint32_t* mk = reinterpret_cast<uint32_t*>(m.rec + c.off_mkey);
      mk[0] = key.k0;
      mk[1] = key.k1;
    }
    build_grids<G>(c, m, g, gmask);
    if (g < c.A) {
      int8_t* row = reinterpret_cast<int8_t*>(sobs) + (el * c.A + g) * c.FR;
      mask[(size_t)env * c.A + g] = emit_obs_and_mask(c, m, g, row);
    }
  }
  __syncthreads();
  store_obs(c, sobs, view, env0, nenv);
  store_records(c, L, srec, state, env0, nenv);
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
    for (int i = 0; i < c.A; ++i) {
      int32_t* o = out + ((size_t)env * c.A + i) * 4;
      o[0] = r[c.off_ax + i];
      o[1] = r[c.off_ay + i];
      o[2] = r[c.off_dir + i];
      o[3] = r[c.off_carry + i];
    }
  } else if (field == 3) {  // shelves: x, y, requested
    for (int s = 0; s < c.n; ++s) {
      int32_t* o = out + ((size_t)env * c.n + s) * 3;
      o[0] = r[c.off_sx + s];
      o[1] = r[c.off_sy + s];
      o[2] = r[c.off_req + s];
    }
  } else if (field == 4) {  // request queue
    for (int q = 0; q < c.Q; ++q) out[(size_t)env * c.Q + q] = r[c.off_queue + q];
  }
}

template <typename K>
int prepare(K kernel, size_t smem) {
  static size_t configured = 0;  // one instance per kernel type
  if (smem > 48 * 1024 && smem > configured) {
    configured = smem;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  return 0;
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
  int o = 0;
  c.off_ax = o; o += c.A;
  c.off_ay = o; o += c.A;
  c.off_dir = o; o += c.A;
  c.off_carry = o; o += c.A;
  c.off_sx = o; o += c.n;
  c.off_sy = o; o += c.n;
  c.off_req = o; o += c.n;
  c.off_queue = o; o += c.Q;
  o = round_up(o, 4);
  c.off_step = o; o += 4;
  c.off_key = o; o += 8;
  c.off_mkey = o; o += 8;
  c.off_run_ret = o; o += 4;
  c.off_run_len = o; o += 4;
  c.off_ep_ret = o; o += 4;
  c.off_ep_len = o; o += 4;
  c.stride = round_up(o, 16);

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

#define MAVA_RWARE_DISPATCH(KERNEL, ...)                                              \
  do {                                                                                \
    if (c.A <= 4) {                                                                   \
      constexpr int G = 4;                                                            \
      const size_t smem = smem_bytes(c, kThreads / G);                                \
      int rc = prepare(KERNEL<G>, smem);                                              \
      if (rc) return rc;                                                              \
      KERNEL<G><<<ceil_div(num_envs, kThreads / G), kThreads, smem, s>>>(__VA_ARGS__); \
    } else {                                                                          \
      constexpr int G = 8;                                                            \
      const size_t smem = smem_bytes(c, kThreads / G);                                \
      int rc = prepare(KERNEL<G>, smem);                                              \
      if (rc) return rc;                                                              \
      KERNEL<G><<<ceil_div(num_envs, kThreads / G), kThreads, smem, s>>>(__VA_ARGS__); \
    }                                                                                 \
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
