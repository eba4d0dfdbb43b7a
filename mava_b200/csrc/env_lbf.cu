// Fused Level-Based Foraging env-step for sm_100a: jax.vmap(env.step) through
// RecordEpisodeMetrics(AutoResetWrapper(AgentIDWrapper(LbfWrapper(LevelBasedForaging)))) in one
// kernel -- mava/systems/ppo/ff_ippo.py:88, mava/utils/make_env.py:69-83,
// mava/wrappers/{jumanji.py:158-215, auto_reset_wrapper.py:60-101, episode_metrics.py:78-111}.
// The inner dynamics follow the published Jumanji LevelBasedForaging algorithm (third party, absent
// from the reference tree).
//
// Mapping: G lanes per env (lane g < A is agent g), 256/G envs per CTA; packed per-env records are
// staged through shared memory with 16-byte loads.  Agents move simultaneously (against the old
// positions), clashes are found with sub-warp shuffles, food levels are summed with a sub-warp
// reduction, and the rare episode end regenerates the env in-kernel (threefry, inverse-CDF food
// placement, Gumbel top-k agent placement ordered on the raw random bits).
#include "env.cuh"
#include "prng.cuh"

namespace mava {
namespace {

constexpr int kThreads = 256;
constexpr int kMaxFood = 8;
constexpr int kMaxCellsLbf = 256;  // grid up to 16 x 16

template <int G>
__device__ __forceinline__ unsigned group_mask() {
  if (G == 32) return 0xffffffffu;
  const unsigned lane = threadIdx.x & 31u;
  return ((1u << G) - 1u) << ((lane / G) * G);
}

template <int G>
__device__ __forceinline__ int group_sum(int v, unsigned gmask) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(gmask, v, o, G);
  return v;
}

// Minimum of a 64-bit composite over the lanes of a group: two hardware warp reductions
// (redux.sync), high word first, then the low word among the ties.
template <int G>
__device__ __forceinline__ unsigned long long group_min(unsigned long long v, unsigned gmask) {
  const uint32_t hi = (uint32_t)(v >> 32);
  const uint32_t m = __reduce_min_sync(gmask, hi);
  const uint32_t lo = hi == m ? (uint32_t)v : 0xffffffffu;
  const uint32_t l = __reduce_min_sync(gmask, lo);
  return ((unsigned long long)m << 32) | l;
}

__device__ __forceinline__ void move_of(int a, int& dx, int& dy) {
  dx = a == 1 ? -1 : (a == 2 ? 1 : 0);
  dy = a == 3 ? -1 : (a == 4 ? 1 : 0);
}

// jax.random.randint(key, (n,), 1, span + 1)[i]
__device__ __forceinline__ int randint_1(Key key, int i, int n, int span) {
  Key k_hi, k_lo;
  split2(key, k_hi, k_lo);
  const uint32_t hi = random_bits_at(k_hi, (uint32_t)i, (uint32_t)n);
  const uint32_t lo = random_bits_at(k_lo, (uint32_t)i, (uint32_t)n);
  const uint32_t s = (uint32_t)span;
  uint32_t mult = 65536u % s;
  mult = (mult * mult) % s;
  return 1 + (int)((((hi % s) * mult) + (lo % s)) % s);
}

// RandomGenerator.__call__ (jumanji lbf/generator.py) for one env, cooperatively on G lanes.
template <int G>
__device__ __forceinline__ void generate(const LbfConst& c, uint8_t* rec, uint32_t* cellmask,
                                         Key key, int g, unsigned gmask) {
  const int S = c.S, flat = S * S;
  Key k_food, k_agents, k_flevel, k_alevel, k_state;
  uint2 y3 = make_uint2(0u, 0u);  // G == 32: level-3 blocks (food draws, agent-level draws)
  if constexpr (G == 32) {
    // Warp-cooperative form (in-step regeneration).  Its ~26 Threefry blocks are a tree of depth
    // three (+ the agent-position draw); a lane computes ONE block per level and the results travel
    // by shuffles, instead of every lane computing every block:
    //   level 1: split(key, 5): blocks p = 0..4 (counters p, p + 5) on lanes 0..4
    //   level 2: split(k_food, NF): blocks 0..NF-1 on lanes 0..7; split(k_alevel): lanes 8, 9
    //   level 3: the foods' uniform draws on lanes 0..7; randint's high / low draws for the agent
    //            levels on lanes 8..11 / 12..15
    auto blk = [](Key k, uint32_t a, uint32_t b) {
      uint32_t o0, o1;
      threefry2x32(k, a, b, o0, o1);
      return make_uint2(o0, o1);
    };
    auto sh = [](uint32_t v, int src) { return __shfl_sync(0xffffffffu, v, src); };
    const uint32_t p1 = g < 5 ? (uint32_t)g : 0u;
    const uint2 y1 = blk(key, p1, p1 + 5u);
    // flat output f of split(key, 5): f < 5 ? x of block f : y of block f - 5; key j = (2j, 2j + 1)
    k_food = Key{sh(y1.x, 0), sh(y1.x, 1)};
    k_agents = Key{sh(y1.x, 2), sh(y1.x, 3)};
    k_flevel = Key{sh(y1.x, 4), sh(y1.y, 0)};
    k_alevel = Key{sh(y1.y, 1), sh(y1.y, 2)};
    k_state = Key{sh(y1.y, 3), sh(y1.y, 4)};
    const uint32_t NF = (uint32_t)c.NF, A = (uint32_t)c.A, ahalf = (A + 1u) >> 1;
    const uint32_t a2 = g < 8 ? ((uint32_t)g < NF ? (uint32_t)g : 0u) : (uint32_t)(g & 1);
    const uint2 y2 = blk(g < 8 ? k_food : k_alevel, a2, g < 8 ? a2 + NF : a2 + 2u);
    const Key ka_hi{sh(y2.x, 8), sh(y2.x, 9)}, ka_lo{sh(y2.y, 8), sh(y2.y, 9)};  // split(k_alevel)
    // lane f < NF: the key of food f = flat outputs (2f, 2f + 1) of split(k_food, NF)
    const uint32_t fl = (uint32_t)g < NF ? (uint32_t)g : 0u, i0 = 2u * fl, i1 = i0 + 1u;
    const int s0 = (int)(i0 < NF ? i0 : i0 - NF), s1 = (int)(i1 < NF ? i1 : i1 - NF);
    const uint32_t x0 = sh(y2.x, s0), w0 = sh(y2.y, s0), x1 = sh(y2.x, s1), w1 = sh(y2.y, s1);
    const Key kf{i0 < NF ? x0 : w0, i1 < NF ? x1 : w1};
    const uint32_t q3 = (uint32_t)(g & 3);
    y3 = blk(g < 8 ? kf : (g < 12 ? ka_hi : ka_lo), g < 8 ? 0u : q3,
             g < 8 ? 0u : (q3 + ahalf < A ? q3 + ahalf : 0u));
  } else {
    k_food = split_n(key, 5, 0);
    k_agents = split_n(key, 5, 1);
    k_flevel = split_n(key, 5, 2);
    k_alevel = split_n(key, 5, 3);
    k_state = split_n(key, 5, 4);
  }
  // ---- food: inverse CDF over the cells that are not on the border / next to an earlier food
  if constexpr (G == 32) {
    // lane w < 8 owns mask word w; the cumulative count is a prefix sum over the word popcounts
    // and the hit is the k-th set bit of one word
    uint32_t word = g < kMaxCellsLbf / 32 ? c.interior[g] : 0u;
    for (int f = 0; f < c.NF; ++f) {
      const uint32_t bits = __shfl_sync(0xffffffffu, y3.x, f);  // random_bits(key of food f, (1,))
      const float u = __uint_as_float((bits >> 9) | 0x3F800000u) - 1.0f;
      const int pc = __popc(word);
      int incl = pc;
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (g >= o) incl += v;
      }
      const int count = __shfl_sync(0xffffffffu, incl, 7);
      const float r = (float)count * (1.0f - u);
      const int k = r <= 0.0f ? 0 : (int)ceilf(r);  // first i with (float)cumsum[i] >= r
      const int excl = incl - pc;
      const bool hit = g < 8 && k > excl && k <= incl;
      const unsigned b = __ballot_sync(0xffffffffu, hit);
      int pos = flat;
      if (k == 0) {
        pos = 0;
      } else if (b != 0u) {
        const int src = __ffs(b) - 1;
        const int local = hit ? (int)__fns(word, 0u, k - excl) : 0;
        pos = 32 * src + __shfl_sync(0xffffffffu, local, src);
      }
      const int adj[5] = {pos, pos + 1, pos - 1, pos + S, pos - S};
#pragma unroll
      for (int q = 0; q < 5; ++q)
        if (adj[q] >= 0 && adj[q] < flat && (adj[q] >> 5) == g) word &= ~(1u << (adj[q] & 31));
      if (g == 0) {
        rec[c.off_fx + f] = (uint8_t)(pos / S);
        rec[c.off_fy + f] = (uint8_t)(pos % S);
        rec[c.off_featen + f] = 0;
      }
    }
    __syncwarp(gmask);
    // ---- the agents may stand anywhere but on food
    if (g < kMaxCellsLbf / 32) {
      uint32_t w = c.allcells[g];
      for (int f = 0; f < c.NF; ++f) {
        const int cell = rec[c.off_fx + f] * S + rec[c.off_fy + f];
        if ((cell >> 5) == g) w &= ~(1u << (cell & 31));
      }
      cellmask[g] = w;
    }
  } else if (g == 0) {
    for (int w = 0; w < kMaxCellsLbf / 32; ++w) cellmask[w] = c.interior[w];
    for (int f = 0; f < c.NF; ++f) {
      const Key kf = split_n(k_food, (uint32_t)c.NF, (uint32_t)f);
      const uint32_t bits = random_bits_at(kf, 0u, 1u);
      const float u = __uint_as_float((bits >> 9) | 0x3F800000u) - 1.0f;
      int count = 0;
      for (int w = 0; w < kMaxCellsLbf / 32; ++w) count += __popc(cellmask[w]);
      const float r = (float)count * (1.0f - u);
      int cum = 0, pos = flat;  // searchsorted(cumsum(mask), r, side="left")
      for (int i = 0; i < flat; ++i) {
        cum += (cellmask[i >> 5] >> (i & 31)) & 1u;
        if ((float)cum >= r) { pos = i; break; }
      }
      const int adj[5] = {pos, pos + 1, pos - 1, pos + S, pos - S};
      for (int q = 0; q < 5; ++q)
        if (adj[q] >= 0 && adj[q] < flat) cellmask[adj[q] >> 5] &= ~(1u << (adj[q] & 31));
      rec[c.off_fx + f] = (uint8_t)(pos / S);
      rec[c.off_fy + f] = (uint8_t)(pos % S);
      rec[c.off_featen + f] = 0;
    }
    // ---- the agents may stand anywhere but on food
    for (int w = 0; w < kMaxCellsLbf / 32; ++w) cellmask[w] = c.allcells[w];
    for (int f = 0; f < c.NF; ++f) {
      const int cell = rec[c.off_fx + f] * S + rec[c.off_fy + f];
      cellmask[cell >> 5] &= ~(1u << (cell & 31));
    }
  }
  __syncwarp(gmask);
  // ---- Gumbel top-k without replacement: -gumbel(u) is decreasing in u, so the A smallest keys
  //      are the A largest uniform mantissas (ties by cell index, as a stable argsort would)
  unsigned long long top[kMaxAgents];
#pragma unroll
  for (int j = 0; j < kMaxAgents; ++j) top[j] = ~0ull;
  auto offer = [&](uint32_t bits, int i) {
    if ((cellmask[i >> 5] >> (i & 31)) & 1u) {
      const uint32_t m = bits >> 9;
      unsigned long long v = ((unsigned long long)(0x7FFFFFu - m) << 32) | (unsigned)i;
#pragma unroll
      for (int j = 0; j < kMaxAgents; ++j) {
        if (v < top[j]) {
          const unsigned long long t = top[j];
          top[j] = v;
          v = t;
        }
      }
    }
  };
  {
    // one block per PAIR of cells (i, i + half): both halves of every Threefry block are used
    const int half = (flat + 1) >> 1;
    for (int p = g; p < half; p += G) {
      uint32_t lo, hi;
      random_bits_pair(k_agents, (uint32_t)p, (uint32_t)flat, lo, hi);
      offer(lo, p);
      if (p + half < flat) offer(hi, p + half);
    }
  }
  int lv_sorted[3] = {1 << 20, 1 << 20, 1 << 20};
#pragma unroll
  for (int r = 0; r < kMaxAgents; ++r) {
    if (r < c.A) {
      const unsigned long long mn = group_min<G>(top[0], gmask);
      if (top[0] == mn) {
#pragma unroll
        for (int j = 0; j + 1 < kMaxAgents; ++j) top[j] = top[j + 1];
        top[kMaxAgents - 1] = ~0ull;
      }
      const int cell = (int)(mn & 0xffffffffull);
      int lvl;
      if constexpr (G == 32) {
        // randint(1, max_level + 1)[r] from the level-3 draws: element r of random_bits(k, (A,)) is
        // half r / ahalf of the block of pair r mod ahalf
        const int ahalf = (c.A + 1) >> 1, q = r < ahalf ? r : r - ahalf;
        const uint32_t hx = __shfl_sync(0xffffffffu, y3.x, 8 + q), hy = __shfl_sync(0xffffffffu, y3.y, 8 + q);
        const uint32_t lx = __shfl_sync(0xffffffffu, y3.x, 12 + q), ly = __shfl_sync(0xffffffffu, y3.y, 12 + q);
        const uint32_t hi = r < ahalf ? hx : hy, lo = r < ahalf ? lx : ly;
        const uint32_t sp = (uint32_t)c.max_level;
        uint32_t mult = 65536u % sp;
        mult = (mult * mult) % sp;
        lvl = 1 + (int)((((hi % sp) * mult) + (lo % sp)) % sp);
      } else {
        lvl = randint_1(k_alevel, r, c.A, c.max_level);
      }
      if (g == 0) {
        rec[c.off_ax + r] = (uint8_t)(cell / S);
        rec[c.off_ay + r] = (uint8_t)(cell % S);
        rec[c.off_alvl + r] = (uint8_t)lvl;
      }
      // keep the three smallest agent levels (max_food_level = their sum)
      int v = lvl;
#pragma unroll
      for (int j = 0; j < 3; ++j)
        if (v < lv_sorted[j]) { const int t = lv_sorted[j]; lv_sorted[j] = v; v = t; }
    }
  }
  int max_food_level = 0;
#pragma unroll
  for (int j = 0; j < 3; ++j)
    if (j < c.A) max_food_level += lv_sorted[j];
  if (g == 0) {
    for (int f = 0; f < c.NF; ++f)
      rec[c.off_flvl + f] = (uint8_t)(c.force_coop ? max_food_level
                                                   : randint_1(k_flevel, f, c.NF, max_food_level));
    *reinterpret_cast<uint32_t*>(rec + c.off_step) = 0u;
    uint32_t* k = reinterpret_cast<uint32_t*>(rec + c.off_key);
    k[0] = k_state.k0;
    k[1] = k_state.k1;
  }
  __syncwarp(gmask);
}

// VectorObserver.make_observation + compute_action_mask for agent g.
// TA / TNF > 0: the agent / food counts are compile-time constants (the loops unroll fully).
template <int TA = 0, int TNF = 0>
__device__ __forceinline__ uint8_t emit_obs_and_mask(const LbfConst& c, const uint8_t* rec, int g,
                                                     int8_t* row) {
  const int A_ = TA ? TA : c.A, NF_ = TNF ? TNF : c.NF;
  const int px = rec[c.off_ax + g], py = rec[c.off_ay + g];
  const int ox = min(c.fov, px), oy = min(c.fov, py);
  int o = 0;
  bool adj_food = false;
  // cells a move cannot enter (another agent, food that has not been eaten): one bit per cell,
  // filled while the observation entries are written, so that the action mask below is four bit
  // tests instead of 6 x (A + NF) position compares on bytes re-read from shared memory
  unsigned long long occ[kMaxCellsLbf / 64];
#pragma unroll
  for (int i = 0; i < kMaxCellsLbf / 64; ++i) occ[i] = 0ull;
  auto occupy = [&](int cell) {
    const unsigned long long bit = 1ull << (cell & 63);
#pragma unroll
    for (int i = 0; i < kMaxCellsLbf / 64; ++i)
      if (i == (cell >> 6)) occ[i] |= bit;
  };
#pragma unroll
  for (int f = 0; f < NF_; ++f) {
    const int fx = rec[c.off_fx + f], fy = rec[c.off_fy + f];
    const bool alive = !rec[c.off_featen + f];
    const bool vis = abs(px - fx) <= c.fov && abs(py - fy) <= c.fov && alive;
    row[o++] = vis ? (int8_t)(fx - px + ox) : -1;
    row[o++] = vis ? (int8_t)(fy - py + oy) : -1;
    row[o++] = vis ? (int8_t)rec[c.off_flvl + f] : 0;
    adj_food |= alive && (abs(px - fx) + abs(py - fy) == 1);
    if (alive) occupy(fx * c.S + fy);
  }
#pragma unroll
  for (int q = 0; q < A_; ++q) {  // own entry first, then the others in index order
    const int j = q == 0 ? g : (q <= g ? q - 1 : q);
    const int ax = rec[c.off_ax + j], ay = rec[c.off_ay + j];
    const bool vis = abs(px - ax) <= c.fov && abs(py - ay) <= c.fov;
    row[o++] = vis ? (int8_t)(ax - px + ox) : -1;
    row[o++] = vis ? (int8_t)(ay - py + oy) : -1;
    row[o++] = vis ? (int8_t)rec[c.off_alvl + j] : 0;
    if (q != 0) occupy(ax * c.S + ay);
  }
  // NOOP is always legal, LOAD next to food; a move into the grid and onto a free cell
  uint8_t mk = (uint8_t)(1u | (adj_food ? 1u << 5 : 0u));
#pragma unroll
  for (int a = 1; a < 5; ++a) {
    int dx, dy;
    move_of(a, dx, dy);
    const int nx = px + dx, ny = py + dy;
    const bool inside = nx >= 0 && ny >= 0 && nx < c.S && ny < c.S;
    const int cell = inside ? nx * c.S + ny : 0;
    unsigned long long w = occ[0];
#pragma unroll
    for (int i = 1; i < kMaxCellsLbf / 64; ++i)
      if (i == (cell >> 6)) w = occ[i];
    const bool free_cell = !((w >> (cell & 63)) & 1ull);
    mk |= (uint8_t)((inside && free_cell ? 1u : 0u) << a);
  }
  return mk;
}

struct SmemLayout {
  int rec_stride, per_cta_rec, per_cta_mask, obs_stride;
};

__host__ __device__ inline SmemLayout smem_layout(const LbfConst& c, int envs_per_cta) {
  SmemLayout L;
  L.rec_stride = c.stride + 16;
  L.per_cta_rec = envs_per_cta * L.rec_stride;
  L.per_cta_mask = envs_per_cta * (kMaxCellsLbf / 32) * 4;
  L.obs_stride = c.A * c.FR;
  return L;
}

__host__ inline size_t smem_bytes(const LbfConst& c, int envs_per_cta) {
  const SmemLayout L = smem_layout(c, envs_per_cta);
  return (size_t)L.per_cta_rec + L.per_cta_mask + (size_t)round_up(envs_per_cta * L.obs_stride, 16);
}

__device__ __forceinline__ void load_records(const LbfConst& c, const SmemLayout& L, uint8_t* srec,
                                             const uint8_t* state, int env0, int nenv) {
  const int v = c.stride >> 4;
  const uint4* src = reinterpret_cast<const uint4*>(state + (size_t)env0 * c.stride);
  for (int i = threadIdx.x; i < nenv * v; i += blockDim.x) {
    const int e = i / v, w = i - e * v;
    reinterpret_cast<uint4*>(srec + e * L.rec_stride)[w] = src[i];
  }
}

__device__ __forceinline__ void store_records(const LbfConst& c, const SmemLayout& L,
                                              const uint8_t* srec, uint8_t* state, int env0,
                                              int nenv) {
  const int v = c.stride >> 4;
  uint4* dst = reinterpret_cast<uint4*>(state + (size_t)env0 * c.stride);
  for (int i = threadIdx.x; i < nenv * v; i += blockDim.x) {
    const int e = i / v, w = i - e * v;
    dst[i] = reinterpret_cast<const uint4*>(srec + e * L.rec_stride)[w];
  }
}

__device__ __forceinline__ void store_obs(const LbfConst& c, const uint8_t* sobs, int8_t* view,
                                          int env0, int nenv) {
  const size_t base = (size_t)env0 * c.A * c.FR;
  const int bytes = nenv * c.A * c.FR;
  if (((base | (size_t)bytes) & 3) == 0) {
    uint32_t* dst = reinterpret_cast<uint32_t*>(view + base);
    for (int i = threadIdx.x; i < (bytes >> 2); i += blockDim.x)
      dst[i] = reinterpret_cast<const uint32_t*>(sobs)[i];
  } else {
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) view[base + i] = (int8_t)sobs[i];
  }
}

template <int G, int TA = 0, int TNF = 0>
__global__ void __launch_bounds__(kThreads)
lbf_step_kernel(const __grid_constant__ LbfConst c, uint8_t* __restrict__ state,
                const int8_t* __restrict__ action, int8_t* __restrict__ view,
                uint8_t* __restrict__ mask, float* __restrict__ reward, uint8_t* __restrict__ done,
                float* __restrict__ ep_return, int32_t* __restrict__ ep_length, int num_envs,
                int auto_reset) {
  extern __shared__ uint4 smem_raw[];
  constexpr int EPC = kThreads / G;
  const SmemLayout L = smem_layout(c, EPC);
  uint8_t* srec = reinterpret_cast<uint8_t*>(smem_raw);
  uint32_t* smask = reinterpret_cast<uint32_t*>(srec + L.per_cta_rec);
  uint8_t* sobs = srec + L.per_cta_rec + L.per_cta_mask;
  const int env0 = blockIdx.x * EPC;
  const int nenv = min(EPC, num_envs - env0);
  __shared__ int rcount;
  __shared__ uint16_t rlist[kThreads / 2];
  if (threadIdx.x == 0) rcount = 0;
  load_records(c, L, srec, state, env0, nenv);
  __syncthreads();
  const int el = threadIdx.x / G, g = threadIdx.x % G;
  const int env = env0 + el;
  bool needs_reset = false;
  // Every lane of a warp runs the step (lanes without an env compute on whatever their record slot
  // holds and store nothing): the exchanges between an env's lanes are then whole-warp shuffles /
  // barriers, one instruction each, instead of collectives on a run-time lane mask (a MATCH.ANY +
  // vote + divergence check in front of each, see env_rware.cuh::step_group)
  constexpr unsigned kAll = 0xffffffffu;
  const int A_ = TA ? TA : c.A, NF_ = TNF ? TNF : c.NF;  // compile-time in the specialised kernel
  const bool active = el < nenv;
  {
    uint8_t* rec = srec + el * L.rec_stride;
    const bool is_agent = active && g < A_;
    // ---- simultaneous moves against the old positions (utils.simulate_agent_movement)
    const int act = is_agent ? action[(size_t)env * A_ + g] : 0;
    const int ox = is_agent ? rec[c.off_ax + g] : -100 - g, oy = is_agent ? rec[c.off_ay + g] : -100;
    int nx = ox, ny = oy;
    if (is_agent) {
      int dx, dy;
      move_of(act, dx, dy);
      const int tx = ox + dx, ty = oy + dy;
      bool bad = tx < 0 || ty < 0 || tx >= c.S || ty >= c.S;
      _Pragma("unroll") for (int j = 0; j < A_; ++j)
        bad |= j != g && rec[c.off_ax + j] == tx && rec[c.off_ay + j] == ty;
      _Pragma("unroll") for (int f = 0; f < NF_; ++f)
        bad |= !rec[c.off_featen + f] && rec[c.off_fx + f] == tx && rec[c.off_fy + f] == ty;
      if (!bad) { nx = tx; ny = ty; }
    }
    // ---- fix_collisions: everybody whose target is shared stays where they were
    bool dup = false;
    for (int j = 0; j < G; ++j) {
      const int jx = __shfl_sync(kAll, nx, j, G), jy = __shfl_sync(kAll, ny, j, G);
      dup |= j != g && j < A_ && jx == nx && jy == ny;
    }
    if (dup) { nx = ox; ny = oy; }
    const bool loading = is_agent && act == 5;
    const int lvl = is_agent ? rec[c.off_alvl + g] : 0;
    // ---- eat_food + get_reward (normalised, no penalty)
    int total_food_level = 0;
    _Pragma("unroll") for (int f = 0; f < NF_; ++f) total_food_level += rec[c.off_flvl + f];
    float rew = 0.0f;
    bool all_eaten = true;
    uint32_t eaten_bits = 0;
    _Pragma("unroll") for (int f = 0; f < NF_; ++f) {
      const int fx = rec[c.off_fx + f], fy = rec[c.off_fy + f], fl = rec[c.off_flvl + f];
      const bool was = rec[c.off_featen + f];
      const int lv = (is_agent && loading && !was && (abs(nx - fx) + abs(ny - fy) == 1)) ? lvl : 0;
      const int sum = group_sum<G>(lv, kAll);
      const bool now = sum >= fl;
      if (sum != 0) rew += (float)(lv * (now ? 1 : 0) * fl) / (float)(sum * total_food_level);
      eaten_bits |= (uint32_t)(now || was) << f;
      all_eaten &= now || was;
    }
    __syncwarp();
    if (is_agent) {
      rec[c.off_ax + g] = (uint8_t)nx;
      rec[c.off_ay + g] = (uint8_t)ny;
    }
    if (g == 0 && active)
      _Pragma("unroll") for (int f = 0; f < NF_; ++f) rec[c.off_featen + f] = (eaten_bits >> f) & 1u;
    // ---- LbfWrapper.aggregate_rewards (sum over agents in index order) / individual rewards
    float team = 0.0f, mean = 0.0f;
    _Pragma("unroll") for (int j = 0; j < A_; ++j) team += __shfl_sync(kAll, rew, j, G);
    const float my_reward = c.individual_rewards ? rew : team;
    _Pragma("unroll") for (int j = 0; j < A_; ++j) mean += __shfl_sync(kAll, my_reward, j, G);
    mean = mean / (float)A_;
    uint32_t* pstep = reinterpret_cast<uint32_t*>(rec + c.off_step);
    const int step = (int)(*pstep) + 1;
    const bool is_done = all_eaten || step >= c.time_limit;
    Key key;
    {
      const uint32_t* k = reinterpret_cast<const uint32_t*>(rec + c.off_key);
      key = Key{k[0], k[1]};
    }
    __syncwarp();
    if (g == 0 && active) {
      *pstep = (uint32_t)step;
      float* run_ret = reinterpret_cast<float*>(rec + c.off_run_ret);
      int32_t* run_len = reinterpret_cast<int32_t*>(rec + c.off_run_len);
      float* e_ret = reinterpret_cast<float*>(rec + c.off_ep_ret);
      int32_t* e_len = reinterpret_cast<int32_t*>(rec + c.off_ep_len);
      const float new_ret = *run_ret + mean;
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
    if (is_agent) reward[(size_t)env * A_ + g] = my_reward;
    __syncwarp();
    needs_reset = active && is_done && auto_reset != 0;
  }
  // ---- AutoResetWrapper (auto_reset_wrapper.py:74-75): finished envs go into a CTA queue and every
  //      warp regenerates one at a time with all 32 lanes (warp-cooperative generator), instead of
  //      the G lanes of the env holding up the other envs of their warp
  if (needs_reset && g == 0) rlist[atomicAdd(&rcount, 1)] = (uint16_t)el;
  __syncthreads();
  {
    const int nreset = rcount;
    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x >> 5; i < nreset; i += kThreads / 32) {
      const int rel = rlist[i];
      uint8_t* rrec = srec + rel * L.rec_stride;
      const uint32_t* k = reinterpret_cast<const uint32_t*>(rrec + c.off_key);
      Key nk, unused;
      split2(Key{k[0], k[1]}, nk, unused);
      __syncwarp();
      generate<32>(c, rrec, smask + rel * (kMaxCellsLbf / 32), nk, lane, 0xffffffffu);
    }
  }
  __syncthreads();
  if (el < nenv && g < A_) {
    const uint8_t* rec = srec + el * L.rec_stride;
    int8_t* row = reinterpret_cast<int8_t*>(sobs) + (el * A_ + g) * c.FR;
    mask[(size_t)env * A_ + g] = emit_obs_and_mask<TA, TNF>(c, rec, g, row);
  }
  __syncthreads();
  store_obs(c, sobs, view, env0, nenv);
  store_records(c, L, srec, state, env0, nenv);
}

template <int G>
__global__ void __launch_bounds__(kThreads)
lbf_reset_kernel(const __grid_constant__ LbfConst c, const uint32_t* __restrict__ keys,
                 uint8_t* __restrict__ state, int8_t* __restrict__ view, uint8_t* __restrict__ mask,
                 int num_envs) {
  extern __shared__ uint4 smem_raw[];
  constexpr int EPC = kThreads / G;
  const SmemLayout L = smem_layout(c, EPC);
  uint8_t* srec = reinterpret_cast<uint8_t*>(smem_raw);
  uint32_t* smask = reinterpret_cast<uint32_t*>(srec + L.per_cta_rec);
  uint8_t* sobs = srec + L.per_cta_rec + L.per_cta_mask;
  const int env0 = blockIdx.x * EPC;
  const int nenv = min(EPC, num_envs - env0);
  const int el = threadIdx.x / G, g = threadIdx.x % G;
  const int env = env0 + el;
  const unsigned gmask = group_mask<G>();
  if (el < nenv) {
    uint8_t* rec = srec + el * L.rec_stride;
    for (int i = g; i < (c.stride >> 2); i += G) reinterpret_cast<uint32_t*>(rec)[i] = 0u;
    __syncwarp(gmask);
    Key key{keys[2 * (size_t)env], keys[2 * (size_t)env + 1]}, reset_key;
    split2(key, key, reset_key);  // RecordEpisodeMetrics.reset, episode_metrics.py:61
    generate<G>(c, rec, smask + el * (kMaxCellsLbf / 32), reset_key, g, gmask);
    if (g == 0) {
      uint32_t* mk = reinterpret_cast<uint32_t*>(rec + c.off_mkey);
      mk[0] = key.k0;
      mk[1] = key.k1;
    }
    __syncwarp(gmask);
    if (g < c.A) {
      int8_t* row = reinterpret_cast<int8_t*>(sobs) + (el * c.A + g) * c.FR;
      mask[(size_t)env * c.A + g] = emit_obs_and_mask(c, rec, g, row);
    }
  }
  __syncthreads();
  store_obs(c, sobs, view, env0, nenv);
  store_records(c, L, srec, state, env0, nenv);
}

__global__ void lbf_peek_kernel(const __grid_constant__ LbfConst c,
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
  } else if (field == 2) {  // agents: x, y, level, 0
    for (int i = 0; i < c.A; ++i) {
      int32_t* o = out + ((size_t)env * c.A + i) * 4;
      o[0] = r[c.off_ax + i];
      o[1] = r[c.off_ay + i];
      o[2] = r[c.off_alvl + i];
      o[3] = 0;
    }
  } else if (field == 3) {  // food: x, y, level
    for (int f = 0; f < c.NF; ++f) {
      int32_t* o = out + ((size_t)env * c.NF + f) * 3;
      o[0] = r[c.off_fx + f];
      o[1] = r[c.off_fy + f];
      o[2] = r[c.off_flvl + f];
    }
  } else if (field == 4) {  // eaten flags
    for (int f = 0; f < c.NF; ++f) out[(size_t)env * c.NF + f] = r[c.off_featen + f];
  }
}

}  // namespace

int lbf_create(const mava_lbf_config* cfg, mava_env_s* env) {
  LbfConst& c = env->lbf;
  MAVA_CHECK_ARG(cfg->grid_size >= 3 && cfg->grid_size * cfg->grid_size <= kMaxCellsLbf);
  MAVA_CHECK_ARG(cfg->num_agents >= 1 && cfg->num_agents <= kMaxAgents);
  MAVA_CHECK_ARG(cfg->num_food >= 1 && cfg->num_food <= kMaxFood);
  MAVA_CHECK_ARG(cfg->max_agent_level >= 1 && cfg->max_agent_level <= 8 && cfg->fov >= 1);
  MAVA_CHECK_ARG(cfg->time_limit >= 1);
  c.S = cfg->grid_size;
  c.fov = cfg->fov;
  c.A = cfg->num_agents;
  c.NF = cfg->num_food;
  c.max_level = cfg->max_agent_level;
  c.force_coop = cfg->force_coop;
  c.time_limit = cfg->time_limit;
  c.individual_rewards = cfg->use_individual_rewards;
  c.FR = 3 * (c.NF + c.A);
  for (int w = 0; w < kMaxCellsLbf / 32; ++w) c.interior[w] = c.allcells[w] = 0u;
  for (int i = 0; i < c.S * c.S; ++i) {
    const int x = i / c.S, y = i % c.S;
    c.allcells[i >> 5] |= 1u << (i & 31);
    if (x > 0 && x < c.S - 1 && y > 0 && y < c.S - 1) c.interior[i >> 5] |= 1u << (i & 31);
  }
  int o = 0;
  c.off_ax = o; o += c.A;
  c.off_ay = o; o += c.A;
  c.off_alvl = o; o += c.A;
  c.off_fx = o; o += c.NF;
  c.off_fy = o; o += c.NF;
  c.off_flvl = o; o += c.NF;
  c.off_featen = o; o += c.NF;
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
  d.kind = MAVA_ENV_LBF;
  d.num_agents = c.A;
  d.view_dim = c.FR;
  d.num_actions = 6;
  d.state_stride = c.stride;
  d.time_limit = c.time_limit;
  d.grid_h = d.grid_w = c.S;
  d.aux0 = c.NF;
  d.aux1 = c.max_level;
  // 2*S_state + A (action) + A*FR (obs) + A (mask) + 4A (reward) + 1 + 9
  const int s_state = 4 * c.A + 4 * c.NF + 2 + 8 + 24;
  d.algo_bytes_per_step = 2 * s_state + c.A + c.A * c.FR + c.A + 4 * c.A + 1 + 9;
  return 0;
}

#define MAVA_LBF_DISPATCH(KERNEL, ...)                                                \
  do {                                                                                \
    if (c.A <= 2) {                                                                   \
      constexpr int G = 2;                                                            \
      const size_t smem = smem_bytes(c, kThreads / G);                                \
      KERNEL<G><<<ceil_div(num_envs, kThreads / G), kThreads, smem, s>>>(__VA_ARGS__); \
    } else if (c.A <= 4) {                                                            \
      constexpr int G = 4;                                                            \
      const size_t smem = smem_bytes(c, kThreads / G);                                \
      KERNEL<G><<<ceil_div(num_envs, kThreads / G), kThreads, smem, s>>>(__VA_ARGS__); \
    } else {                                                                          \
      constexpr int G = 8;                                                            \
      const size_t smem = smem_bytes(c, kThreads / G);                                \
      KERNEL<G><<<ceil_div(num_envs, kThreads / G), kThreads, smem, s>>>(__VA_ARGS__); \
    }                                                                                 \
  } while (0)

int lbf_reset(const mava_env_s* env, const uint32_t* keys, uint8_t* state, int8_t* view,
              uint8_t* mask, int num_envs, cudaStream_t s) {
  const LbfConst& c = env->lbf;
  MAVA_LBF_DISPATCH(lbf_reset_kernel, c, keys, state, view, mask, num_envs);
  return launch_status();
}

int lbf_step(const mava_env_s* env, uint8_t* state, const int8_t* action, int8_t* view,
             uint8_t* mask, float* reward, uint8_t* done, float* ep_return, int32_t* ep_length,
             int num_envs, int auto_reset, cudaStream_t s) {
  const LbfConst& c = env->lbf;
  if (c.A == 2 && c.NF == 2) {  // the benchmark scenarios: agent / food loops unrolled at compile time
    const size_t smem = smem_bytes(c, kThreads / 2);
    lbf_step_kernel<2, 2, 2><<<ceil_div(num_envs, kThreads / 2), kThreads, smem, s>>>(
        c, state, action, view, mask, reward, done, ep_return, ep_length, num_envs, auto_reset);
    return launch_status();
  }
  MAVA_LBF_DISPATCH(lbf_step_kernel, c, state, action, view, mask, reward, done, ep_return,
                    ep_length, num_envs, auto_reset);
  return launch_status();
}

int lbf_peek(const mava_env_s* env, const uint8_t* state, int field, int32_t* out, int num_envs,
             cudaStream_t s) {
  MAVA_CHECK_ARG(field >= 0 && field <= 4);
  lbf_peek_kernel<<<ceil_div(num_envs, 128), 128, 0, s>>>(env->lbf, state, field, out, num_envs);
  return launch_status();
}

}  // namespace mava
