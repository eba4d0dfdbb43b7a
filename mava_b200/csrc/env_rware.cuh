// Device-side RobotWarehouse: the env-step of one env on its shared-memory record, the generator,
// and the observation-row builder.  Shared by the stand-alone env kernels (env_rware.cu) and the
// fused rollout kernel (rollout_tc.cu).  See env_rware.cu for the design notes.
#pragma once
#include "env.cuh"
#include "prng.cuh"

namespace mava {
namespace rware {

// ---- small PTX helpers (mbarrier + bulk copies) ------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t ok = 0;
  for (uint32_t spins = 0; !ok; ++spins) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (spins > (1u << 24)) __trap();  // a bug surfaces as an error instead of a hang
  }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst),
               "r"(smem_u32(src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_wait_read() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

template <int G>
__device__ __forceinline__ unsigned group_mask() {
  if (G == 32) return 0xffffffffu;
  const unsigned lane = threadIdx.x & 31u;
  return ((1u << G) - 1u) << ((lane / G) * G);
}

// Minimum of a 64-bit composite over the lanes of a group: two hardware warp reductions
// (redux.sync) instead of a shuffle tree -- high word first, then the low word among the ties.
template <int G>
__device__ __forceinline__ unsigned long long group_min(unsigned long long v, unsigned gmask) {
  const uint32_t hi = (uint32_t)(v >> 32);
  const uint32_t m = __reduce_min_sync(gmask, hi);
  const uint32_t lo = hi == m ? (uint32_t)v : 0xffffffffu;
  const uint32_t l = __reduce_min_sync(gmask, lo);
  return ((unsigned long long)m << 32) | l;
}

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

__device__ __forceinline__ uint32_t pack_agent(int x, int y, int d, int carry) {
  return (uint32_t)x | ((uint32_t)y << 8) | ((uint32_t)d << 16) | ((uint32_t)carry << 24);
}

__device__ __forceinline__ bool requested(const RwareConst& c, const uint8_t* rec, int s) {
  return (reinterpret_cast<const uint32_t*>(rec + c.off_reqbits)[s >> 5] >> (s & 31)) & 1u;
}

// K smallest of the composites (random_bits(sub, size)[i] << 32 | i), i.e. the first K entries of
// jax.random.permutation-by-stable-sort.  Every lane returns the same out[].
// COMPACT: the two insertion passes and the K extraction rounds as rolled loops (a third of the
// instructions; for kernels whose instruction footprint matters more than this function's speed).
// pre[0..4): composites of pairs [0, p_start) this lane has computed already (~0 = none); the other
// pairs are computed here.
template <int G, int KMAX, class P = PrngInline, bool COMPACT = false>
__device__ __forceinline__ void smallest_k(Key sub, int size, int K, int g, unsigned gmask,
                                           unsigned long long (&out)[KMAX],
                                           const unsigned long long (&pre)[4], int p_start) {
  unsigned long long top[KMAX];
#pragma unroll
  for (int j = 0; j < KMAX; ++j) top[j] = ~0ull;
  const int half = (size + 1) >> 1;
  auto insert = [&](unsigned long long v) {
#pragma unroll
    for (int j = 0; j < KMAX; ++j) {
      if (v < top[j]) {
        unsigned long long t = top[j];
        top[j] = v;
        v = t;
      }
    }
  };
  if (p_start > 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) insert(pre[i]);  // (~0 never displaces anything)
  }
  for (int p = p_start + g; p < half; p += G) {
    uint32_t lo, hi;
    P::bits_pair(sub, (uint32_t)p, (uint32_t)size, lo, hi);
    if (COMPACT) {
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        const int idx = p + h * half;
        if (idx < size) insert(((unsigned long long)(h ? hi : lo) << 32) | (unsigned)idx);
      }
    } else {
      insert(((unsigned long long)lo << 32) | (unsigned)p);
      if (p + half < size) insert(((unsigned long long)hi << 32) | (unsigned)(p + half));
    }
  }
  auto extract = [&](int r) {
    unsigned long long mn = ~0ull;
    if (r < K) {
      mn = group_min<G>(top[0], gmask);
      if (top[0] == mn) {
#pragma unroll
        for (int j = 0; j + 1 < KMAX; ++j) top[j] = top[j + 1];
        top[KMAX - 1] = ~0ull;
      }
    }
#pragma unroll
    for (int j = 0; j < KMAX; ++j)
      if (j == r) out[j] = mn;
  };
  if (COMPACT) {
#pragma unroll 1
    for (int r = 0; r < KMAX; ++r) extract(r);
  } else {
#pragma unroll
    for (int r = 0; r < KMAX; ++r) extract(r);
  }
}

template <int G, int KMAX, class P = PrngInline, bool COMPACT = false>
__device__ __forceinline__ void smallest_k(Key sub, int size, int K, int g, unsigned gmask,
                                           unsigned long long (&out)[KMAX]) {
  const unsigned long long none[4] = {~0ull, ~0ull, ~0ull, ~0ull};
  smallest_k<G, KMAX, P, COMPACT>(sub, size, K, g, gmask, out, none, 0);
}

// Agents on the KA-list's cells, requested shelves from the KQ-list, shelf grid, step, key: the part
// of generate() behind the key derivation (KA >= c.A, KQ >= c.Q: the top-k lists are as short as
// the caller can promise).
template <int GG, class P, bool COMPACT, int KA, int KQ>
__device__ __forceinline__ void place(const RwareConst& c, uint8_t* rec, Key sub_pos, Key sub_q,
                                      Key k3, int my_dir, int g, unsigned gmask,
                                      const unsigned long long (&pre)[4], int p_start) {
  unsigned long long pick[KA];
  smallest_k<GG, KA, P, COMPACT>(sub_pos, c.HW, c.A, g, gmask, pick, pre, p_start);
  if (g < c.A) {
    unsigned long long mine = 0ull;
#pragma unroll
    for (int i = 0; i < KA; ++i)
      if (i == g) mine = pick[i];
    const int cell = (int)(mine & 0xffffffffull);
    reinterpret_cast<uint32_t*>(rec + c.off_agents)[g] = pack_agent(cell / c.W, cell % c.W, my_dir, 0);
  }
  unsigned long long qpick[KQ];
  smallest_k<GG, KQ, P, COMPACT>(sub_q, c.n, c.Q, g, gmask, qpick);
  uint32_t* cw = reinterpret_cast<uint32_t*>(rec + c.off_cells);
  for (int i = g; i < c.cells_words; i += GG) cw[i] = 0u;
  __syncwarp(gmask);
  uint8_t* cells = rec + c.off_cells;
  for (int s = g; s < c.n; s += GG) cells[c.shelf_home[s]] = (uint8_t)(s + 1);
  if (g == 0) {
    uint32_t* rq = reinterpret_cast<uint32_t*>(rec + c.off_reqbits);
    for (int i = 0; i < c.req_words; ++i) rq[i] = 0u;
#pragma unroll
    for (int i = 0; i < KQ; ++i) {
      if (i < c.Q) {
        const int s = (int)(qpick[i] & 0xffffffffull);
        rec[c.off_queue + i] = (uint8_t)s;
        rq[s >> 5] |= 1u << (s & 31);
      }
    }
    *reinterpret_cast<uint32_t*>(rec + c.off_step) = 0u;
    uint32_t* k = reinterpret_cast<uint32_t*>(rec + c.off_key);
    k[0] = k3.k0;
    k[1] = k3.k1;
  }
  __syncwarp(gmask);
}

// jumanji RandomGenerator.__call__: agents on distinct random cells, random directions, shelves on
// their home cells, Q distinct requested shelves.  Writes the inner-env part of the record (GG
// lanes cooperate; State.key is what is left of `key`).
template <int GG, class P = PrngInline, bool COMPACT = false, int KA = kMaxAgents>
__device__ __forceinline__ void generate(const RwareConst& c, uint8_t* rec, Key key, int g,
                                         unsigned gmask) {
  // Regenerations are on the hot path of an untrained policy (in tiny-4ag an episode lasts a few
  // steps under random actions: they are most of the step kernel's instructions there), and a
  // regeneration is a chain of dependent Threefry blocks.  The six key splits are twelve blocks in
  // a tree of depth four: every lane would compute all of them for itself; instead a lane computes
  // ONE block per level (lanes 0..3: role r = g & 3, r & 1 = the block with counters (0, 2) or (1, 3)
  // of a split, r >> 1 = which of the level's two keys) and the halves are exchanged with shuffles.
  // A whole warp has 28 lanes to spare on the last two levels: they draw the first 54 pairs of the
  // position permutation (its key is known after level two) and the agents' directions, so a
  // regeneration is five block times (four levels + the request draw) instead of fifteen.
  Key k1, pos_key, k2, dir_key, k3, q_key, unused, sub_pos, sub_q, d_hi, d_lo;
  unsigned long long pre[4] = {~0ull, ~0ull, ~0ull, ~0ull};
  int p_start = 0, my_dir = 0;
  if constexpr (GG >= 4) {
    const int r = g & 3;
    const uint32_t c0 = (uint32_t)(r & 1), c1 = c0 + 2u;
    auto sh = [&](uint32_t v, int src) { return __shfl_sync(gmask, v, src, GG); };
    uint2 y = P::block(key.k0, key.k1, c0, c1);                       // split(key)
    k1 = Key{sh(y.x, 0), sh(y.x, 1)};
    pos_key = Key{sh(y.y, 0), sh(y.y, 1)};
    Key ks = (r >> 1) ? pos_key : k1;                                 // split(k1) | split(pos_key)
    y = P::block(ks.k0, ks.k1, c0, c1);
    k2 = Key{sh(y.x, 0), sh(y.x, 1)};
    dir_key = Key{sh(y.y, 0), sh(y.y, 1)};
    sub_pos = Key{sh(y.y, 2), sh(y.y, 3)};
    if constexpr (GG == 32) {
      const int half = (c.HW + 1) >> 1, dhalf = (c.A + 1) >> 1;
      auto pair_c1 = [](int p, int hf, int size) { return (uint32_t)(p + hf < size ? p + hf : 0); };
      auto keep = [&](int slot, uint2 v, int p) {  // both halves of position pair p
        if (p < half) {
          pre[slot] = ((unsigned long long)v.x << 32) | (unsigned)p;
          if (p + half < c.HW) pre[slot + 1] = ((unsigned long long)v.y << 32) | (unsigned)(p + half);
        }
      };
      // level 3: split(k2) | split(dir_key) | position pairs 0..27
      const int p3 = g - 4;
      ks = g < 2 ? k2 : (g < 4 ? dir_key : sub_pos);
      y = P::block(ks.k0, ks.k1, g < 4 ? c0 : (uint32_t)p3, g < 4 ? c1 : pair_c1(p3, half, c.HW));
      k3 = Key{sh(y.x, 0), sh(y.x, 1)};
      q_key = Key{sh(y.y, 0), sh(y.y, 1)};
      d_lo = Key{sh(y.y, 2), sh(y.y, 3)};  // randint: span 4 -> only the low-bits draw matters
      if (g >= 4) keep(0, y, p3);
      // level 4: split(q_key) | position pairs 28..53 | direction pairs
      const int p4 = 26 + g, dp = g - 28;
      ks = g < 2 ? q_key : (g < 28 ? sub_pos : d_lo);
      y = P::block(ks.k0, ks.k1, g < 2 ? c0 : (g < 28 ? (uint32_t)p4 : (uint32_t)dp),
                   g < 2 ? c1 : (g < 28 ? pair_c1(p4, half, c.HW) : pair_c1(dp, dhalf, c.A)));
      sub_q = Key{sh(y.y, 0), sh(y.y, 1)};
      if (g >= 2 && g < 28) keep(2, y, p4);
      p_start = 54;
      // randint(0, 4) of agent g = element g of random_bits(d_lo, (A,)): pair g mod dhalf
      const int dpl = g < c.A ? (g < dhalf ? g : g - dhalf) : 0;
      const uint32_t dx = sh(y.x, 28 + dpl), dy = sh(y.y, 28 + dpl);
      my_dir = g < c.A ? (int)((g < dhalf ? dx : dy) & 3u) : 0;
    } else {
      ks = (r >> 1) ? dir_key : k2;                                   // split(k2) | split(dir_key)
      y = P::block(ks.k0, ks.k1, c0, c1);
      k3 = Key{sh(y.x, 0), sh(y.x, 1)};
      q_key = Key{sh(y.y, 0), sh(y.y, 1)};
      d_lo = Key{sh(y.y, 2), sh(y.y, 3)};
      y = P::block(q_key.k0, q_key.k1, c0, c1);                       // split(q_key)
      sub_q = Key{sh(y.y, 0), sh(y.y, 1)};
      my_dir = g < c.A ? (int)(P::bits_at(d_lo, (uint32_t)g, (uint32_t)c.A) & 3u) : 0;
    }
  } else {
    P::split(key, k1, pos_key);
    P::split(k1, k2, dir_key);
    P::split(pos_key, unused, sub_pos);
    P::split(k2, k3, q_key);
    P::split(dir_key, d_hi, d_lo);  // randint: span 4 -> only the low-bits draw matters
    P::split(q_key, unused, sub_q);
    my_dir = g < c.A ? (int)(P::bits_at(d_lo, (uint32_t)g, (uint32_t)c.A) & 3u) : 0;
  }
  // (a kernel that minds its instruction footprint gets one instantiation; the others the short
  //  lists whenever the configuration allows them)
  if (COMPACT) {
    place<GG, P, COMPACT, KA, kMaxQueue>(c, rec, sub_pos, sub_q, k3, my_dir, g, gmask, pre, p_start);
  } else if (c.A <= 4 && c.Q <= 4) {
    place<GG, P, COMPACT, 4, 4>(c, rec, sub_pos, sub_q, k3, my_dir, g, gmask, pre, p_start);
  } else {
    place<GG, P, COMPACT, kMaxAgents, kMaxQueue>(c, rec, sub_pos, sub_q, k3, my_dir, g, gmask, pre,
                                                 p_start);
  }
}

// ---- observation rows --------------------------------------------------------------------------
// jumanji utils.make_agent_observation: [x, y, carrying, onehot(dir, 4), on_highway] then, for the
// (2R+1)^2 - 1 cells around the agent, [other agent present, onehot(its dir, 4)], then for all
// (2R+1)^2 cells [shelf present, shelf requested]; int8, FR = 8 + 5 (L - 1) + 2 L bytes.
template <int R>
struct ObsDims {
  static constexpr int SIDE = 2 * R + 1;
  static constexpr int LOC = SIDE * SIDE;
  static constexpr int CENTER = LOC / 2;
  static constexpr int NAG = (LOC - 1) * 5;   // bytes of "other agent" features
  static constexpr int FR = 8 + NAG + 2 * LOC;
  static constexpr int NH = FR / 2;           // halfwords per row (odd)
  static constexpr int NW = NH / 2 + 1;       // 32-bit words, the last one half used
  static constexpr int MW = (NAG + 63) / 64;  // 64-bit words of the agent-feature bit mask
  static constexpr int SH0 = (8 + NAG) / 4;   // first word of the shelf features
  static_assert(FR % 4 == 2 && NAG % 4 == 0, "row layout assumptions");
};

// OR the 5-bit pattern p into bit position s of a multi-word mask.
template <int MW>
__device__ __forceinline__ void set5(unsigned long long (&M)[MW], int s, unsigned long long p) {
  if (MW == 1) {
    M[0] |= p << s;
  } else {
    const int w = s >> 6, b = s & 63;
#pragma unroll
    for (int i = 0; i < MW; ++i) {
      if (i == w) M[i] |= p << b;
      if (i == w + 1 && b > 59) M[i] |= p >> (64 - b);
    }
  }
}

// Builds the row of agent g from the (post-step) record as O::NW little-endian words (the upper half
// of the last one is zero), with the action mask bits returned.  opk[j] = old cell | new cell << 10 |
// moved << 20 of agent j for this step; when `replay` is set the AGENTS grid is evaluated exactly as
// the reference's sequence of grid writes leaves it (needed only for the terminal observation of a
// collision without auto-reset), else agents are on distinct cells and the grid is simply
// {cell of j -> j}.
template <int G, int R>
__device__ __forceinline__ uint8_t build_row(const RwareConst& c, const uint8_t* rec, int g,
                                             bool replay, const uint32_t (&opk)[G],
                                             uint32_t (&w)[ObsDims<R>::NW]) {
  using O = ObsDims<R>;
  const uint32_t* agents = reinterpret_cast<const uint32_t*>(rec + c.off_agents);
  const uint8_t* cells = rec + c.off_cells;
  const uint32_t me = agents[g];
  const int x = me & 0xff, y = (me >> 8) & 0xff, d = (me >> 16) & 0xff, carry = me >> 24;
#pragma unroll
  for (int i = 0; i < O::NW; ++i) w[i] = 0u;
  w[0] = (uint32_t)x | ((uint32_t)y << 8) | ((uint32_t)carry << 16) | ((uint32_t)(d == 0) << 24);
  w[1] = (uint32_t)(d == 1) | ((uint32_t)(d == 2) << 8) | ((uint32_t)(d == 3) << 16) |
         ((uint32_t)is_highway(c, x * c.W + y) << 24);
  unsigned long long M[O::MW];
#pragma unroll
  for (int i = 0; i < O::MW; ++i) M[i] = 0ull;
  uint32_t ag[G];
#pragma unroll
  for (int j = 0; j < G; ++j) ag[j] = j < c.A ? agents[j] : 0u;
  if (!replay) {
#pragma unroll
    for (int j = 0; j < G; ++j) {
      if (j < c.A && j != g) {
        const int dx = (int)(ag[j] & 0xff) - x + R, dy = (int)((ag[j] >> 8) & 0xff) - y + R;
        if ((unsigned)dx <= 2u * R && (unsigned)dy <= 2u * R) {
          int k = dx * O::SIDE + dy;
          if (k != O::CENTER) {
            k -= k > O::CENTER;
            set5<O::MW>(M, 5 * k, 1ull | (2ull << ((ag[j] >> 16) & 3u)));
          }
        }
      }
    }
  } else {
    int k = 0;
    for (int dx = -R; dx <= R; ++dx) {
      for (int dy = -R; dy <= R; ++dy) {
        if (dx == 0 && dy == 0) continue;
        const int cx = x + dx, cy = y + dy;
        if (cx >= 0 && cx < c.H && cy >= 0 && cy < c.W) {
          const uint32_t q = (uint32_t)(cx * c.W + cy);
          int v = 0;
#pragma unroll
          for (int m = 0; m < G; ++m)
            if (m < c.A && (opk[m] & 1023u) == q) v = m + 1;
#pragma unroll
          for (int m = 0; m < G; ++m) {
            if (m < c.A && ((opk[m] >> 20) & 1u)) {
              if ((opk[m] & 1023u) == q) v = 0;
              if (((opk[m] >> 10) & 1023u) == q) v = m + 1;
            }
          }
          if (v != 0) {
            uint32_t dj = 0;
#pragma unroll
            for (int m = 0; m < G; ++m)
              if (m + 1 == v) dj = (ag[m] >> 16) & 3u;
            set5<O::MW>(M, 5 * k, 1ull | (2ull << dj));
          }
        }
        ++k;
      }
    }
  }
  // bit b of the mask -> byte b of the feature block: 4 bits per word, spread with one multiply
#pragma unroll
  for (int i = 0; i < O::NAG / 4; ++i) {
    const uint32_t nib = (uint32_t)(M[i >> 4] >> ((i & 15) * 4)) & 0xFu;
    w[2 + i] = (nib * 0x00204081u) & 0x01010101u;
  }
  const uint32_t* rq = reinterpret_cast<const uint32_t*>(rec + c.off_reqbits);
#pragma unroll
  for (int ci = 0; ci < O::LOC; ++ci) {
    const int cx = x + ci / O::SIDE - R, cy = y + ci % O::SIDE - R;
    const bool inside = (unsigned)cx < (unsigned)c.H && (unsigned)cy < (unsigned)c.W;
    const int sid = inside ? (int)cells[cx * c.W + cy] : 0;
    // s = sid - 1 is -1 for an empty cell: read word 0 and mask the result instead of branching
    const int s1 = sid - 1;
    const uint32_t word = rq[sid ? s1 >> 5 : 0];
    const uint32_t present = sid != 0;
    const uint32_t h = present | (((word >> (s1 & 31)) & present) << 8);
    w[O::SH0 + ci / 2] |= h << (16 * (ci & 1));
  }
  // utils.compute_action_mask: only FORWARD can be illegal
  int nx, ny;
  forward_cell(c, x, y, d, nx, ny);
  const bool stuck = nx == x && ny == y;
  const bool blocked = carry && cells[nx * c.W + ny] != 0;
  return (uint8_t)(0x1Du | ((stuck || blocked) ? 0u : 0x2u));
}

// A quarter of build_row + store_row, for callers that put four threads on a row (the agents are on
// distinct cells: not the replay case).  part 0: the agent's own features -- returns the action
// mask; part 1: the other agents around it; parts 2, 3: the shelf cells.  Every word / half word
// goes to the int8 row (`row`, 2-byte aligned) and to a second, 4-byte aligned copy `img`.
template <int G, int R>
__device__ __forceinline__ uint8_t build_row_part(const RwareConst& c, const uint8_t* rec, int g,
                                                  int part, uint8_t* row, uint8_t* img) {
  using O = ObsDims<R>;
  const uint32_t* agents = reinterpret_cast<const uint32_t*>(rec + c.off_agents);
  const uint8_t* cells = rec + c.off_cells;
  const uint32_t me = agents[g];
  const int x = me & 0xff, y = (me >> 8) & 0xff, d = (me >> 16) & 0xff, carry = me >> 24;
  auto put32 = [&](int off, uint32_t v) {
    *reinterpret_cast<uint32_t*>(img + off) = v;
    reinterpret_cast<uint16_t*>(row + off)[0] = (uint16_t)v;
    reinterpret_cast<uint16_t*>(row + off)[1] = (uint16_t)(v >> 16);
  };
  if (part == 0) {
    put32(0, (uint32_t)x | ((uint32_t)y << 8) | ((uint32_t)carry << 16) | ((uint32_t)(d == 0) << 24));
    put32(4, (uint32_t)(d == 1) | ((uint32_t)(d == 2) << 8) | ((uint32_t)(d == 3) << 16) |
                 ((uint32_t)is_highway(c, x * c.W + y) << 24));
    // utils.compute_action_mask: only FORWARD can be illegal
    int nx, ny;
    forward_cell(c, x, y, d, nx, ny);
    const bool stuck = nx == x && ny == y;
    const bool blocked = carry && cells[nx * c.W + ny] != 0;
    return (uint8_t)(0x1Du | ((stuck || blocked) ? 0u : 0x2u));
  }
  if (part == 1) {
    unsigned long long M[O::MW];
#pragma unroll
    for (int i = 0; i < O::MW; ++i) M[i] = 0ull;
#pragma unroll
    for (int j = 0; j < G; ++j) {
      if (j < c.A && j != g) {
        const uint32_t aj = agents[j];
        const int dx = (int)(aj & 0xff) - x + R, dy = (int)((aj >> 8) & 0xff) - y + R;
        if ((unsigned)dx <= 2u * R && (unsigned)dy <= 2u * R) {
          int k = dx * O::SIDE + dy;
          if (k != O::CENTER) {
            k -= k > O::CENTER;
            set5<O::MW>(M, 5 * k, 1ull | (2ull << ((aj >> 16) & 3u)));
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < O::NAG / 4; ++i) {
      const uint32_t nib = (uint32_t)(M[i >> 4] >> ((i & 15) * 4)) & 0xFu;
      put32(8 + 4 * i, (nib * 0x00204081u) & 0x01010101u);
    }
    return 0;
  }
  const uint32_t* rq = reinterpret_cast<const uint32_t*>(rec + c.off_reqbits);
  constexpr int HALF = (O::LOC + 1) / 2;
#pragma unroll
  for (int i = 0; i < HALF; ++i) {
    const int ci = part == 2 ? i : HALF + i;
    if (ci < O::LOC) {
      const int cx = x + ci / O::SIDE - R, cy = y + ci % O::SIDE - R;
      const bool inside = (unsigned)cx < (unsigned)c.H && (unsigned)cy < (unsigned)c.W;
      const int sid = inside ? (int)cells[cx * c.W + cy] : 0;
      const int s1 = sid - 1;
      const uint32_t word = rq[sid ? s1 >> 5 : 0];
      const uint32_t present = sid != 0;
      const uint16_t h = (uint16_t)(present | (((word >> (s1 & 31)) & present) << 8));
      *reinterpret_cast<uint16_t*>(img + 4 * O::SH0 + 2 * ci) = h;
      *reinterpret_cast<uint16_t*>(row + 4 * O::SH0 + 2 * ci) = h;
    }
  }
  return 0;
}

// The words of build_row to the int8 row at `row` (rows are FR = 2 (mod 4) bytes long: odd rows
// start on a half word).
template <int R>
__device__ __forceinline__ void store_row(const uint32_t (&w)[ObsDims<R>::NW], uint8_t* row,
                                          int row_parity) {
  using O = ObsDims<R>;
  uint32_t* wp = reinterpret_cast<uint32_t*>(row + 2 * row_parity);
  const int sh = 16 * row_parity;
#pragma unroll
  for (int k = 0; k + 1 < O::NW; ++k) wp[k] = __funnelshift_r(w[k], w[k + 1], sh);
  uint16_t* hp = reinterpret_cast<uint16_t*>(row_parity ? row : row + 4 * (O::NW - 1));
  *hp = (uint16_t)(row_parity ? w[0] : w[O::NW - 1]);
}

template <int G, int R>
__device__ __forceinline__ uint8_t emit_row(const RwareConst& c, const uint8_t* rec, int g,
                                            uint8_t* row, int row_parity, bool replay,
                                            const uint32_t (&opk)[G]) {
  uint32_t w[ObsDims<R>::NW];
  const uint8_t mk = build_row<G, R>(c, rec, g, replay, opk, w);
  store_row<R>(w, row, row_parity);
  return mk;
}

// One env-step of the G lanes of an env on its shared-memory record (everything of env.step and the
// wrapper stack up to, not including, the auto-reset and the next observation).  Called by ALL 32
// lanes of a warp together -- lanes whose env does not exist pass active = false (and any readable
// record) -- so that the exchanges between an env's lanes are whole-warp shuffles / votes / barriers
// (one instruction each) instead of collectives on a run-time lane mask (a MATCH.ANY + vote +
// divergence check each).  `act` is lane g's action; VALIDATED: it is known to respect the action
// mask of the current state (sampled from masked logits).  Outputs go to the per-env / per-agent
// slots of this step (index env, env * A + g).  Returns through the references whether the env must
// be regenerated, whether the terminal observation needs the exact grid replay, and the (old cell,
// new cell, moved) words of all agents.
template <int G, bool VALIDATED = false, class P = PrngInline>
__device__ __forceinline__ void step_group(const RwareConst& c, uint8_t* rec, int g, unsigned gmask,
                                           bool active, bool agent, int act, int env,
                                           int auto_reset, float* __restrict__ reward,
                                           uint8_t* __restrict__ done,
                                           float* __restrict__ ep_return,
                                           int32_t* __restrict__ ep_length, bool& needs_reset,
                                           bool& replay, uint32_t (&opk)[G]) {
  constexpr unsigned kAll = 0xffffffffu;
  uint8_t* cells = rec + c.off_cells;
  Key key{0u, 0u};
  // --- validate the action against the mask of the current state (utils.get_valid_actions)
  int x = 0, y = 0, d = 0, carry = 0, nx = 0, ny = 0;
  if (agent) {
    const uint32_t me = reinterpret_cast<const uint32_t*>(rec + c.off_agents)[g];
    x = me & 0xff;
    y = (me >> 8) & 0xff;
    d = (me >> 16) & 0xff;
    carry = me >> 24;
    forward_cell(c, x, y, d, nx, ny);
    if (!VALIDATED && act == 1) {
      const bool stuck = nx == x && ny == y;
      const bool blocked = carry && cells[nx * c.W + ny] != 0;
      if (stuck || blocked) act = 0;
    }
  }
  const int oldcell = x * c.W + y;
  const bool moved = act == 1;
  // where everybody is and goes is known before anybody acts: the forward cell depends only on the
  // agent's own state.  (old cell, new cell, moved) of all agents, exchanged with shuffles:
  const int newcell = moved ? nx * c.W + ny : oldcell;
  const uint32_t pk = (uint32_t)oldcell | ((uint32_t)newcell << 10) | ((uint32_t)moved << 20);
#pragma unroll
  for (int j = 0; j < G; ++j) opk[j] = __shfl_sync(kAll, pk, j, G);
  // --- collision (utils.is_collision): grid[AGENTS, pos_i] != i + 1 after the sequential writes
  //     "old cell <- 0, new cell <- j + 1" of every agent j that moved.  The last write to my
  //     cell is not mine iff some mover j entered or left it after my own write.
  //     `touch`: my move lands on a cell another agent occupies or enters -- only then can the order
  //     of the agents' turns matter for the shelf grid.
  bool my_col = false, touch = false;
#pragma unroll
  for (int j = 0; j < G; ++j) {
    if (j < c.A && j != g) {
      const uint32_t oj = opk[j] & 1023u, nj = (opk[j] >> 10) & 1023u;
      const bool mj = (opk[j] >> 20) & 1u;
      if (mj && (nj == (uint32_t)newcell || oj == (uint32_t)newcell) && (!moved || j > g))
        my_col = true;
      if (moved && (nj == (uint32_t)newcell || oj == (uint32_t)newcell)) touch = true;
    }
  }
  const unsigned touch_b = __ballot_sync(kAll, touch && agent);
  const unsigned col_b = __ballot_sync(kAll, my_col && agent);
  const bool collision = (col_b & gmask) != 0u;
  // rotations and the position update do not depend on the other agents
  if (act == 2) d = (d + 3) & 3;
  else if (act == 3) d = (d + 1) & 3;
  if (moved) {
    x = nx;
    y = ny;
  }
  if (touch_b == 0u) {
    // every cell this step reads or writes belongs to exactly one agent: the turns commute
    if (moved && carry) {
      const uint8_t sid = cells[oldcell];
      cells[oldcell] = 0;
      cells[newcell] = sid;
    } else if (act == 4 && agent) {
      if (!carry) carry = cells[oldcell] != 0;
      else if (!is_highway(c, oldcell)) carry = 0;
    }
  } else {
    // --- some env of this warp has agents meeting on a cell: agents act one after the other on
    //     the shelf grid (scan over agents in env.step); for the other envs the order is immaterial
    for (int i = 0; i < c.A; ++i) {
      if (g == i && agent) {
        if (moved && carry) {
          const uint8_t sid = cells[oldcell];
          cells[oldcell] = 0;
          cells[newcell] = sid;
        } else if (act == 4) {
          if (!carry) carry = cells[oldcell] != 0;
          else if (!is_highway(c, oldcell)) carry = 0;
        }
      }
      __syncwarp();
    }
  }
  if (agent) reinterpret_cast<uint32_t*>(rec + c.off_agents)[g] = pack_agent(x, y, d, carry);
  __syncwarp();

  // --- deliveries at the goal cells; a delivered request is replaced by a uniformly drawn
  //     shelf that is not in the queue (env._update_reward_and_request_queue)
  float rew = 0.0f;
  {
    const uint32_t* k = reinterpret_cast<const uint32_t*>(rec + c.off_key);
    key = Key{k[0], k[1]};
  }
#pragma unroll 1
  for (int gi = 0; gi < 2; ++gi) {
    const int sid = cells[c.goal[gi]];
    if (active && sid != 0 && requested(c, rec, sid - 1)) {
      Key rkey, unused, sub;
      P::split(key, key, rkey);
      P::split(rkey, unused, sub);
      const int msize = c.n - c.Q;
      unsigned long long best = ~0ull;
      for (int s = g; s < c.n; s += G) {
        int below = 0;
        bool inq = false;
        for (int q = 0; q < c.Q; ++q) {
          const int qs = rec[c.off_queue + q];
          inq |= qs == s;
          below += qs < s;
        }
        if (!inq) {
          const int p = s - below;  // position in the sorted not-in-queue list
          const uint32_t b = P::bits_at(sub, (uint32_t)p, (uint32_t)msize);
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
          if (rec[c.off_queue + q] == sid - 1) {
            rec[c.off_queue + q] = (uint8_t)new_req;
            break;
          }
        }
        uint32_t* rq = reinterpret_cast<uint32_t*>(rec + c.off_reqbits);
        rq[(sid - 1) >> 5] &= ~(1u << ((sid - 1) & 31));
        rq[new_req >> 5] |= 1u << (new_req & 31);
      }
      rew += 1.0f;
      __syncwarp(gmask);
    }
  }
  // --- step count, termination
  uint32_t* pstep = reinterpret_cast<uint32_t*>(rec + c.off_step);
  const int step = (int)(*pstep) + 1;
  const bool is_done = collision || step >= c.time_limit;
  __syncwarp();
  if (g == 0 && active) {
    *pstep = (uint32_t)step;
    uint32_t* k = reinterpret_cast<uint32_t*>(rec + c.off_key);
    k[0] = key.k0;
    k[1] = key.k1;
    // RecordEpisodeMetrics.step (episode_metrics.py:83-110)
    float* run_ret = reinterpret_cast<float*>(rec + c.off_run_ret);
    int32_t* run_len = reinterpret_cast<int32_t*>(rec + c.off_run_len);
    float* e_ret = reinterpret_cast<float*>(rec + c.off_ep_ret);
    int32_t* e_len = reinterpret_cast<int32_t*>(rec + c.off_ep_len);
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
  if (agent) reward[(size_t)env * c.A + g] = rew;
  needs_reset = active && is_done && auto_reset != 0;
  replay = active && collision && !needs_reset;
}

}  // namespace rware
}  // namespace mava
