// fp32 actor / critic path: fused 3-layer MLP forward with acting / loss epilogues, backward and
// weight-gradient kernels.  This is the full-precision path (rtol 1e-5 against the oracle); the
// bf16 tcgen05 path in mlp_tc.cu is the fast one and is checked against this one as well.
//
// Reference regions replaced: FeedForwardActor / FeedForwardValueNet / MLPTorso /
// DiscreteActionHead (mava/networks.py:39-58,88-124,172-207), pi.sample / log_prob / entropy
// (mava/distributions.py:146-165 over tfd.Categorical), _actor_loss_fn / _critic_loss_fn and their
// value_and_grad (mava/systems/ppo/ff_mappo.py:150-218), pmean over "batch" (:224-234).
#include "common.cuh"
#include "prng.cuh"

namespace mava {
namespace {

constexpr int BM = 64;        // rows per CTA tile
constexpr int NT = 256;       // threads
constexpr int HMAX = 128;     // max hidden width
constexpr int KC = 16;        // K chunk streamed through shared memory
constexpr int OMAX = 16;      // max head width (actions)
constexpr int HS = HMAX + 4;  // row stride of activation tiles in shared memory
constexpr float kF32Min = -3.402823466e38f;

struct InputDesc {
  const int8_t* view;   // [S][A][FR]
  const int32_t* rows;  // env-step index per tile row group, or nullptr for identity
  int mode, add_id, A, FR, in_dim;
  int rows_per_step;    // A for MAVA_IN_AGENT_VIEW, 1 for MAVA_IN_GLOBAL
};

struct NetPtrs {
  const float *w1, *b1, *w2, *b2, *w3, *b3;
  int in_dim, h1, h2, out;
};

__host__ __device__ inline NetPtrs net_ptrs(const float* p, int in_dim, int h1, int h2, int out) {
  NetPtrs n;
  n.in_dim = in_dim; n.h1 = h1; n.h2 = h2; n.out = out;
  n.w1 = p; p += (size_t)in_dim * h1;
  n.b1 = p; p += h1;
  n.w2 = p; p += (size_t)h1 * h2;
  n.b2 = p; p += h2;
  n.w3 = p; p += (size_t)h2 * out;
  n.b3 = p;
  return n;
}

// Per-tile row bookkeeping: where a tile row reads its observation and where it writes.
struct RowMap {
  int64_t obs_off[BM];  // byte offset of the row's first view feature
  int32_t agent[BM];    // agent index of the row (one-hot id / per-agent arrays)
  int64_t flat[BM];     // index into [S][A] arrays (AGENT_VIEW) or [S] (GLOBAL: env-step index)
};

__device__ __forceinline__ void fill_row_map(const InputDesc& d, int64_t row0, int64_t M,
                                             RowMap& rm) {
  for (int r = threadIdx.x; r < BM; r += NT) {
    const int64_t row = row0 + r;
    if (row < M) {
      if (d.mode == MAVA_IN_GLOBAL) {
        const int64_t s = d.rows ? d.rows[row] : row;
        rm.obs_off[r] = s * d.A * d.FR;
        rm.agent[r] = 0;
        rm.flat[r] = s;
      } else {
        const int64_t j = row / d.A;
        const int a = (int)(row - j * d.A);
        const int64_t s = d.rows ? d.rows[j] : j;
        rm.obs_off[r] = (s * d.A + a) * d.FR;
        rm.agent[r] = a;
        rm.flat[r] = s * d.A + a;
      }
    } else {
      rm.obs_off[r] = -1;
      rm.agent[r] = 0;
      rm.flat[r] = -1;
    }
  }
}

__device__ __forceinline__ float input_at(const InputDesc& d, const RowMap& rm, int r, int k) {
  if (rm.obs_off[r] < 0 || k >= d.in_dim) return 0.0f;
  if (d.mode == MAVA_IN_AGENT_VIEW && d.add_id) {
    if (k < d.A) return k == rm.agent[r] ? 1.0f : 0.0f;
    k -= d.A;
  }
  return (float)d.view[rm.obs_off[r] + k];
}

// acc[4][8] += A[rows ty*4..+3][k0..k0+kn) * B chunk.  A row-major in smem with stride lda,
// B chunk Bs[kk][HMAX].  Thread columns: tx*4..+3 and 64+tx*4..+3.
__device__ __forceinline__ void mma_chunk(const float* __restrict__ As, int lda, int acol0,
                                          const float (*Bs)[HMAX], int kn, int ty, int tx,
                                          float (&acc)[4][8]) {
#pragma unroll 4
  for (int kk = 0; kk < kn; ++kk) {
    float a[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = As[(ty * 4 + i) * lda + acol0 + kk];
    const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
    const float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][64 + tx * 4]);
    const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
  }
}

__device__ __forceinline__ int thread_col(int tx, int j) { return j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4); }

// Stream W[k0..k0+KC)[0..width) (row-major, row length width) into Bs, zero padded.
__device__ __forceinline__ void load_w_chunk(const float* __restrict__ W, int K, int width, int k0,
                                             float (*Bs)[HMAX]) {
  for (int i = threadIdx.x; i < KC * HMAX; i += NT) {
    const int kk = i / HMAX, c = i - kk * HMAX;
    const int k = k0 + kk;
    Bs[kk][c] = (k < K && c < width) ? __ldg(W + (size_t)k * width + c) : 0.0f;
  }
}

enum Mode { kAct = 0, kValue = 1, kTrainActor = 2, kTrainCritic = 3 };

struct FwdArgs {
  InputDesc in;
  NetPtrs net;
  int64_t M;
  // activation stash for training (nullptr when acting)
  float* H1;
  float* H2;
  // acting
  const uint8_t* mask;       // [S][A]
  const uint32_t* policy_key;
  int envs_per_replica;
  int greedy;
  const int8_t* actions_in;  // replay
  int8_t* action;            // [S][A]
  float* logp;               // [S][A]
  float* value;              // [S][A]
  // training
  const int8_t* action_old;
  const float* old_logp;
  const float* old_value;
  const float* adv;
  const float* targets;
  const double* adv_stats;   // [U][2] sum, sumsq
  int mb_size;               // env-steps per replica
  int num_replicas;
  float clip_eps, ent_coef, vf_coef;
  float* dOut;               // [M][OMAX] (actor) or [M] (critic)
  double* loss_acc;          // [5]
};

struct FwdSmem {
  float Ha[BM][HS];
  float Hb[BM][HS];
  float Xs[BM][KC + 1];
  float Bs[KC][HMAX];
  float red[4][BM][OMAX];
  RowMap rm;
  double lsum[8];
};

template <int MODE>
__global__ void __launch_bounds__(NT) mlp_fwd_kernel(const FwdArgs p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FwdSmem& sm = *reinterpret_cast<FwdSmem*>(smem_raw);
  const int tid = threadIdx.x, ty = tid / 16, tx = tid % 16;
  const int64_t row0 = (int64_t)blockIdx.x * BM;
  const NetPtrs& n = p.net;
  fill_row_map(p.in, row0, p.M, sm.rm);
  if (tid < 8) sm.lsum[tid] = 0.0;
  __syncthreads();

  float acc[4][8];
  // ---------------- layer 1: X (built from the int8 view) x W1
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
  for (int k0 = 0; k0 < n.in_dim; k0 += KC) {
    for (int i = tid; i < BM * KC; i += NT) {
      const int r = i / KC, kk = i - r * KC;
      sm.Xs[r][kk] = input_at(p.in, sm.rm, r, k0 + kk);
    }
    load_w_chunk(n.w1, n.in_dim, n.h1, k0, sm.Bs);
    __syncthreads();
    mma_chunk(&sm.Xs[0][0], KC + 1, 0, sm.Bs, KC, ty, tx, acc);
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = ty * 4 + i;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = thread_col(tx, j);
      const float v = c < n.h1 ? fmaxf(acc[i][j] + __ldg(n.b1 + c), 0.0f) : 0.0f;
      sm.Ha[r][c] = v;
      if (p.H1 && c < n.h1 && row0 + r < p.M) p.H1[(row0 + r) * n.h1 + c] = v;
    }
  }
  __syncthreads();
  // ---------------- layer 2
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
  for (int k0 = 0; k0 < n.h1; k0 += KC) {
    load_w_chunk(n.w2, n.h1, n.h2, k0, sm.Bs);
    __syncthreads();
    mma_chunk(&sm.Ha[0][0], HS, k0, sm.Bs, min(KC, n.h1 - k0), ty, tx, acc);
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = ty * 4 + i;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = thread_col(tx, j);
      const float v = c < n.h2 ? fmaxf(acc[i][j] + __ldg(n.b2 + c), 0.0f) : 0.0f;
      sm.Hb[r][c] = v;
      if (p.H2 && c < n.h2 && row0 + r < p.M) p.H2[(row0 + r) * n.h2 + c] = v;
    }
  }
  __syncthreads();
  // ---------------- head: out[r][o] = sum_k Hb[r][k] W3[k][o]; 4 k-slices per row
  {
    const int r = tid % BM, part = tid / BM;
    float o[OMAX];
#pragma unroll
    for (int j = 0; j < OMAX; ++j) o[j] = 0.0f;
    const int kslice = (n.h2 + 3) / 4;
    for (int k = part * kslice; k < min(n.h2, (part + 1) * kslice); ++k) {
      const float h = sm.Hb[r][k];
#pragma unroll
      for (int j = 0; j < OMAX; ++j)
        if (j < n.out) o[j] = fmaf(h, __ldg(n.w3 + (size_t)k * n.out + j), o[j]);
    }
#pragma unroll
    for (int j = 0; j < OMAX; ++j) sm.red[part][r][j] = o[j];
  }
  __syncthreads();
  // ---------------- epilogue: one thread per row
  double l0 = 0.0, l1 = 0.0, l2 = 0.0;
  if (tid < BM && row0 + tid < p.M) {
    const int r = tid;
    const int64_t row = row0 + r;
    float out[OMAX];
#pragma unroll
    for (int j = 0; j < OMAX; ++j)
      out[j] = j < n.out ? sm.red[0][r][j] + sm.red[1][r][j] + sm.red[2][r][j] + sm.red[3][r][j] +
                               __ldg(n.b3 + j)
                         : 0.0f;
    const int64_t flat = sm.rm.flat[r];
    if (MODE == kValue || MODE == kTrainCritic) {
      const float v = out[0];
      const int reps = p.in.mode == MAVA_IN_GLOBAL ? p.in.A : 1;
      const int64_t base = p.in.mode == MAVA_IN_GLOBAL ? flat * p.in.A : flat;
      if (MODE == kValue) {
        for (int a = 0; a < reps; ++a) p.value[base + a] = v;
      } else {
        // _critic_loss_fn, ff_mappo.py:190-201, and its gradient w.r.t. the value
        const float w = 1.0f / ((float)p.num_replicas * (float)p.mb_size * (float)p.in.A);
        float dv = 0.0f;
        for (int a = 0; a < reps; ++a) {
          const float vo = p.old_value[base + a], tg = p.targets[base + a];
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
          l0 += 0.5 * (double)fmaxf(a1, a2);
        }
        p.dOut[row] = dv * w * p.vf_coef;
      }
    } else {
      // masked categorical (networks.py:116-124)
      const uint8_t mk = p.mask[flat];
      float mx = kF32Min;
#pragma unroll
      for (int j = 0; j < OMAX; ++j) {
        if (j < n.out) {
          out[j] = ((mk >> j) & 1) ? out[j] : kF32Min;
          mx = fmaxf(mx, out[j]);
        }
      }
      float se = 0.0f;
#pragma unroll
      for (int j = 0; j < OMAX; ++j)
        if (j < n.out) se += expf(out[j] - mx);
      const float lse = mx + logf(se);
      if (MODE == kAct) {
        int a = 0;
        if (p.actions_in) {
          a = p.actions_in[flat];
        } else if (p.greedy) {
          float best = out[0];
#pragma unroll
          for (int j = 1; j < OMAX; ++j)
            if (j < n.out && out[j] > best) { best = out[j]; a = j; }
        } else {
          // Gumbel arg-max; noise laid out (envs_per_replica, A, N) as tfd.Categorical.sample
          const Key key{p.policy_key[0], p.policy_key[1]};
          const int64_t s = flat / p.in.A;
          const int ag = (int)(flat - s * p.in.A);
          const int64_t e = s % p.envs_per_replica;
          const uint32_t size = (uint32_t)p.envs_per_replica * p.in.A * n.out;
          const uint32_t base = (uint32_t)((e * p.in.A + ag) * n.out);
          float best = 0.0f;
#pragma unroll
          for (int j = 0; j < OMAX; ++j) {
            if (j < n.out) {
              const float z = bits_to_gumbel(random_bits_at(key, base + j, size)) + out[j];
              if (j == 0 || z > best) { best = z; a = j; }
            }
          }
        }
        float la = 0.0f;
#pragma unroll
        for (int j = 0; j < OMAX; ++j)
          if (j == a) la = out[j] - lse;
        p.action[flat] = (int8_t)a;
        p.logp[flat] = la;
      } else {
        // _actor_loss_fn, ff_mappo.py:159-180, and d(total_loss)/d(logits)
        const int a = p.action_old[flat];
        const int64_t j_step = (flat / p.in.A);
        (void)j_step;
        const int u = (int)((row / p.in.A) / p.mb_size);
        const double cnt = (double)p.mb_size * p.in.A;
        const double mean_d = p.adv_stats[2 * u] / cnt;
        const double var_d = fmax(p.adv_stats[2 * u + 1] / cnt - mean_d * mean_d, 0.0);
        const float mean = (float)mean_d, sd = (float)sqrt(var_d);
        float logp[OMAX], pr[OMAX];
        float la = 0.0f, ent = 0.0f;
#pragma unroll
        for (int j = 0; j < OMAX; ++j) {
          logp[j] = 0.0f;
          pr[j] = 0.0f;
          if (j < n.out) {
            logp[j] = out[j] - lse;
            pr[j] = expf(logp[j]);
            if (pr[j] != 0.0f) ent -= pr[j] * logp[j];
            if (j == a) la = logp[j];
          }
        }
        const float ratio = expf(la - p.old_logp[flat]);
        const float g = (p.adv[flat] - mean) / (sd + 1e-8f);
        const float lo = 1.0f - p.clip_eps, hi = 1.0f + p.clip_eps;
        const float t1 = ratio * g, t2 = fminf(fmaxf(ratio, lo), hi) * g;
        const bool inside = ratio > lo && ratio < hi;
        float dr;
        if (t1 < t2) dr = -g;
        else if (t1 > t2) dr = inside ? -g : 0.0f;
        else dr = -g * (0.5f + (inside ? 0.5f : 0.0f));
        const float dla = dr * ratio;
        const float w = 1.0f / ((float)p.num_replicas * (float)p.mb_size * (float)p.in.A);
#pragma unroll
        for (int j = 0; j < OMAX; ++j) {
          float dl = 0.0f;
          if (j < n.out && ((mk >> j) & 1)) {
            dl = dla * ((j == a ? 1.0f : 0.0f) - pr[j]);
            if (pr[j] != 0.0f) dl += p.ent_coef * pr[j] * (logp[j] + ent);
          }
          p.dOut[row * OMAX + j] = dl * w;
        }
        l0 = (double)(-fminf(t1, t2));
        l1 = (double)ent;
      }
    }
  }
  if (MODE == kTrainActor || MODE == kTrainCritic) {
    // block-reduce the loss sums, then one atomic per CTA
    for (int o = 16; o > 0; o >>= 1) {
      l0 += __shfl_xor_sync(0xffffffffu, l0, o);
      l1 += __shfl_xor_sync(0xffffffffu, l1, o);
    }
    if ((tid & 31) == 0 && tid < BM) {
      atomicAdd(&sm.lsum[0], l0);
      atomicAdd(&sm.lsum[1], l1);
    }
    __syncthreads();
    if (tid == 0) {
      if (MODE == kTrainActor) {
        atomicAdd(p.loss_acc + 0, sm.lsum[0]);
        atomicAdd(p.loss_acc + 1, sm.lsum[1]);
      } else {
        atomicAdd(p.loss_acc + 2, sm.lsum[0]);
      }
    }
  }
  (void)l2;
}

// out[M][K] = (D[M][Nd] x W[K][Nd]^T) * (H > 0), written over H (dZ = dH * relu').
struct BwdSmem {
  float Ds[BM][KC + 1];
  float Bs[KC][HMAX];
};

__global__ void __launch_bounds__(NT)
mlp_bwd_kernel(const float* __restrict__ D, int ldd, int Nd, const float* __restrict__ W, int K,
               float* __restrict__ H, int64_t M) {
  __shared__ BwdSmem sm;
  const int tid = threadIdx.x, ty = tid / 16, tx = tid % 16;
  const int64_t row0 = (int64_t)blockIdx.x * BM;
  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
  for (int n0 = 0; n0 < Nd; n0 += KC) {
    for (int i = tid; i < BM * KC; i += NT) {
      const int r = i / KC, nn = i - r * KC;
      const int64_t row = row0 + r;
      sm.Ds[r][nn] = (row < M && n0 + nn < Nd) ? D[row * ldd + n0 + nn] : 0.0f;
    }
    for (int i = tid; i < KC * HMAX; i += NT) {
      const int nn = i / HMAX, k = i - nn * HMAX;
      sm.Bs[nn][k] = (k < K && n0 + nn < Nd) ? __ldg(W + (size_t)k * Nd + n0 + nn) : 0.0f;
    }
    __syncthreads();
    mma_chunk(&sm.Ds[0][0], KC + 1, 0, sm.Bs, KC, ty, tx, acc);
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t row = row0 + ty * 4 + i;
    if (row >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = thread_col(tx, j);
      if (c < K) {
        const float h = H[row * K + c];
        H[row * K + c] = h > 0.0f ? acc[i][j] : 0.0f;
      }
    }
  }
}

// dW[K][N] += sum_rows A[row][k] * D[row][n], db[n] += sum_rows D[row][n].
// A is either a dense f32 matrix (lda = K) or built from the int8 view (layer 1).
constexpr int WG_ROWS = 1024;  // rows reduced per CTA before the atomics
struct WgSmem {
  float As[KC][BM];
  float Ds[KC][HMAX];
  RowMap rm;
};

template <bool FROM_VIEW>
__global__ void __launch_bounds__(NT)
mlp_wgrad_kernel(const float* __restrict__ A, InputDesc in, int K, const float* __restrict__ D,
                 int ldd, int N, int64_t M, float* __restrict__ dW, float* __restrict__ db) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  WgSmem& sm = *reinterpret_cast<WgSmem*>(smem_raw);
  const int tid = threadIdx.x, ty = tid / 16, tx = tid % 16;
  const int k0 = blockIdx.y * BM;  // this CTA's 64 input features
  const int64_t rbeg = (int64_t)blockIdx.x * WG_ROWS;
  const int64_t rend = rbeg + WG_ROWS < M ? rbeg + WG_ROWS : M;
  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
  float bsum = 0.0f;
  for (int64_t r0 = rbeg; r0 < rend; r0 += BM) {
    if (FROM_VIEW) {
      __syncthreads();
      fill_row_map(in, r0, rend, sm.rm);
      __syncthreads();
    }
    for (int rc = 0; rc < BM; rc += KC) {
      for (int i = tid; i < KC * BM; i += NT) {
        const int rr = i / BM, k = i - rr * BM;
        const int64_t row = r0 + rc + rr;
        float v = 0.0f;
        if (row < rend && k0 + k < K) {
          if (FROM_VIEW) v = input_at(in, sm.rm, rc + rr, k0 + k);
          else v = A[row * K + k0 + k];
        }
        sm.As[rr][k] = v;
      }
      for (int i = tid; i < KC * HMAX; i += NT) {
        const int rr = i / HMAX, c = i - rr * HMAX;
        const int64_t row = r0 + rc + rr;
        sm.Ds[rr][c] = (row < rend && c < N) ? D[row * ldd + c] : 0.0f;
      }
      __syncthreads();
#pragma unroll 4
      for (int rr = 0; rr < KC; ++rr) {
        const float4 a4 = *reinterpret_cast<const float4*>(&sm.As[rr][ty * 4]);
        const float a[4] = {a4.x, a4.y, a4.z, a4.w};
        const float4 b0 = *reinterpret_cast<const float4*>(&sm.Ds[rr][tx * 4]);
        const float4 b1 = *reinterpret_cast<const float4*>(&sm.Ds[rr][64 + tx * 4]);
        const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      if (blockIdx.y == 0 && tid < HMAX) {
#pragma unroll
        for (int rr = 0; rr < KC; ++rr) bsum += sm.Ds[rr][tid];
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k = k0 + ty * 4 + i;
    if (k >= K) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = thread_col(tx, j);
      if (c < N) atomicAdd(dW + (size_t)k * N + c, acc[i][j]);
    }
  }
  if (blockIdx.y == 0 && tid < N) atomicAdd(db + tid, bsum);
}

// Per-replica sum and sum of squares of the advantages of a minibatch (gae.mean(), gae.std()).
__global__ void __launch_bounds__(256)
adv_stats_kernel(const float* __restrict__ adv, const int32_t* __restrict__ rows, int mb_size,
                 int A, double* __restrict__ stats) {
  const int u = blockIdx.y;
  double s = 0.0, ss = 0.0;
  const int64_t n = (int64_t)mb_size * A;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = i / A;
    const int a = (int)(i - j * A);
    const float v = adv[(int64_t)rows[(int64_t)u * mb_size + j] * A + a];
    s += v;
    ss += (double)v * v;
  }
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
  }
  __shared__ double rs[8], rss[8];
  if ((threadIdx.x & 31) == 0) {
    rs[threadIdx.x >> 5] = s;
    rss[threadIdx.x >> 5] = ss;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) {
      s += rs[w];
      ss += rss[w];
    }
    atomicAdd(stats + 2 * u, s);
    atomicAdd(stats + 2 * u + 1, ss);
  }
}

__global__ void finalize_loss_kernel(const double* __restrict__ acc, double denom, float ent_coef,
                                     float vf_coef, float* __restrict__ out5) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double actor_loss = acc[0] / denom, entropy = acc[1] / denom, value_loss = acc[2] / denom;
  out5[0] = (float)(actor_loss - (double)ent_coef * entropy);
  out5[1] = (float)actor_loss;
  out5[2] = (float)entropy;
  out5[3] = (float)((double)vf_coef * value_loss);
  out5[4] = (float)value_loss;
}

int check_desc(const mava_mlp_desc* d) {
  if (!d) return MAVA_E_NULL;
  if (d->h1 < 1 || d->h1 > HMAX || d->h2 < 1 || d->h2 > HMAX) return MAVA_E_UNSUPPORTED;
  if (d->out_dim < 1 || d->out_dim > OMAX) return MAVA_E_UNSUPPORTED;
  if (d->num_agents < 1 || d->view_dim < 1) return MAVA_E_BADARG;
  int in_dim = d->input_mode == MAVA_IN_GLOBAL ? d->num_agents * d->view_dim
                                               : d->view_dim + (d->add_agent_id ? d->num_agents : 0);
  if (d->in_dim != in_dim) return MAVA_E_BADARG;
  if (d->input_mode != MAVA_IN_GLOBAL && d->input_mode != MAVA_IN_AGENT_VIEW) return MAVA_E_BADARG;
  return 0;
}

InputDesc make_input(const mava_mlp_desc* d, const int8_t* view, const int32_t* rows) {
  InputDesc in;
  in.view = view;
  in.rows = rows;
  in.mode = d->input_mode;
  in.add_id = d->add_agent_id;
  in.A = d->num_agents;
  in.FR = d->view_dim;
  in.in_dim = d->in_dim;
  in.rows_per_step = d->input_mode == MAVA_IN_GLOBAL ? 1 : d->num_agents;
  return in;
}

template <int MODE>
int launch_fwd(const FwdArgs& a, cudaStream_t s) {
  static size_t configured[kMaxDevices] = {};
  if (int rc = ensure_dyn_smem(mlp_fwd_kernel<MODE>, sizeof(FwdSmem), configured)) return rc;
  const int blocks = (int)ceil_div64(a.M, BM);
  mlp_fwd_kernel<MODE><<<blocks, NT, sizeof(FwdSmem), s>>>(a);
  return launch_status();
}

}  // namespace

// shared with the bf16 path (ppo_tc.cu)
int launch_adv_stats(const float* adv, const int32_t* rows, int mb_size, int A, int num_replicas,
                     double* stats, cudaStream_t s) {
  dim3 grid((unsigned)min((int64_t)sm_count() * 2, ceil_div64((int64_t)mb_size * A, 256)),
            (unsigned)num_replicas);
  adv_stats_kernel<<<grid, 256, 0, s>>>(adv, rows, mb_size, A, stats);
  return launch_status();
}

int launch_finalize_loss(const double* acc, double denom, float ent_coef, float vf_coef, float* out5,
                         cudaStream_t s) {
  finalize_loss_kernel<<<1, 32, 0, s>>>(acc, denom, ent_coef, vf_coef, out5);
  return launch_status();
}

}  // namespace mava

using namespace mava;

extern "C" {

int64_t mava_mlp_param_count(const mava_mlp_desc* d) {
  if (!d) return -1;
  return (int64_t)d->in_dim * d->h1 + d->h1 + (int64_t)d->h1 * d->h2 + d->h2 +
         (int64_t)d->h2 * d->out_dim + d->out_dim;
}

int mava_ff_act(const mava_mlp_desc* actor, const float* actor_params, const mava_mlp_desc* critic,
                const float* critic_params, const int8_t* view, const uint8_t* mask,
                const uint32_t* policy_key, int envs_per_replica, int num_envs, int greedy,
                const int8_t* actions_in, int8_t* action, float* logp, float* value,
                mava_stream_t s) {
  int rc = check_desc(actor);
  if (rc) return rc;
  MAVA_CHECK_PTR(actor_params);
  MAVA_CHECK_PTR(view);
  MAVA_CHECK_PTR(mask);
  MAVA_CHECK_PTR(action);
  MAVA_CHECK_PTR(logp);
  MAVA_CHECK_ARG(num_envs > 0 && envs_per_replica > 0);
  MAVA_CHECK_ARG(actor->input_mode == MAVA_IN_AGENT_VIEW);
  if (!greedy && !actions_in) MAVA_CHECK_PTR(policy_key);
  FwdArgs a{};
  a.in = make_input(actor, view, nullptr);
  a.net = net_ptrs(actor_params, actor->in_dim, actor->h1, actor->h2, actor->out_dim);
  a.M = (int64_t)num_envs * actor->num_agents;
  a.mask = mask;
  a.policy_key = policy_key;
  a.envs_per_replica = envs_per_replica;
  a.greedy = greedy;
  a.actions_in = actions_in;
  a.action = action;
  a.logp = logp;
  rc = launch_fwd<kAct>(a, as_stream(s));
  if (rc) return rc;
  if (value) return mava_ff_value(critic, critic_params, view, num_envs, value, s);
  return 0;
}

int mava_ff_value(const mava_mlp_desc* critic, const float* critic_params, const int8_t* view,
                  int num_envs, float* value, mava_stream_t s) {
  int rc = check_desc(critic);
  if (rc) return rc;
  MAVA_CHECK_PTR(critic_params);
  MAVA_CHECK_PTR(view);
  MAVA_CHECK_PTR(value);
  MAVA_CHECK_ARG(num_envs > 0 && critic->out_dim == 1);
  FwdArgs a{};
  a.in = make_input(critic, view, nullptr);
  a.net = net_ptrs(critic_params, critic->in_dim, critic->h1, critic->h2, 1);
  a.M = (int64_t)num_envs * a.in.rows_per_step;
  a.value = value;
  return launch_fwd<kValue>(a, as_stream(s));
}

static int64_t align_up64(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

int64_t mava_ppo_workspace_bytes(const mava_mlp_desc* actor, const mava_mlp_desc* critic,
                                 int rows_total) {
  if (!actor || !critic || rows_total <= 0) return -1;
  const int64_t Ma = (int64_t)rows_total * actor->num_agents;
  const int64_t Mc = (int64_t)rows_total *
                     (critic->input_mode == MAVA_IN_GLOBAL ? 1 : critic->num_agents);
  int64_t b = 256;  // stats + loss accumulators
  b += align_up64(Ma * (actor->h1 + actor->h2 + OMAX) * 4, 256);
  b += align_up64(Mc * (critic->h1 + critic->h2 + 1) * 4, 256);
  return b + 1024;
}

int mava_ppo_loss_grad(const mava_mlp_desc* actor, const float* actor_params,
                       const mava_mlp_desc* critic, const float* critic_params,
                       const mava_ppo_hyper* hyper, const int8_t* view, const uint8_t* mask,
                       const int8_t* action, const float* old_logp, const float* old_value,
                       const float* adv, const float* targets, const int32_t* rows,
                       int num_replicas, int mb_size, float* grad_out, void* workspace,
                       mava_stream_t stream) {
  int rc = check_desc(actor);
  if (rc) return rc;
  rc = check_desc(critic);
  if (rc) return rc;
  MAVA_CHECK_PTR(hyper);
  MAVA_CHECK_PTR(actor_params);
  MAVA_CHECK_PTR(critic_params);
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
  MAVA_CHECK_ARG(num_replicas > 0 && mb_size > 0 && critic->out_dim == 1);
  MAVA_CHECK_ARG(actor->input_mode == MAVA_IN_AGENT_VIEW);
  cudaStream_t s = as_stream(stream);
  const int A = actor->num_agents;
  const int R = num_replicas * mb_size;
  const int64_t Ma = (int64_t)R * A;
  const int64_t Mc = (int64_t)R * (critic->input_mode == MAVA_IN_GLOBAL ? 1 : A);
  const int64_t na = mava_mlp_param_count(actor), nc = mava_mlp_param_count(critic);

  // carve the workspace
  unsigned char* w = static_cast<unsigned char*>(workspace);
  double* stats = reinterpret_cast<double*>(w);           // [U][2], U <= 8
  double* loss_acc = reinterpret_cast<double*>(w + 128);  // [5]
  if (num_replicas > 8) return MAVA_E_UNSUPPORTED;
  w += 256;
  float* aH1 = reinterpret_cast<float*>(w);
  float* aH2 = aH1 + Ma * actor->h1;
  float* aD = aH2 + Ma * actor->h2;
  w += align_up64(Ma * (actor->h1 + actor->h2 + OMAX) * 4, 256);
  float* cH1 = reinterpret_cast<float*>(w);
  float* cH2 = cH1 + Mc * critic->h1;
  float* cD = cH2 + Mc * critic->h2;

  cudaError_t e = cudaMemsetAsync(workspace, 0, 256, s);
  if (e != cudaSuccess) return (int)e;
  e = cudaMemsetAsync(grad_out, 0, (size_t)(na + nc + 8) * sizeof(float), s);
  if (e != cudaSuccess) return (int)e;

  {
    dim3 grid((unsigned)min((int64_t)sm_count() * 2, ceil_div64((int64_t)mb_size * A, 256)),
              (unsigned)num_replicas);
    adv_stats_kernel<<<grid, 256, 0, s>>>(adv, rows, mb_size, A, stats);
  }
  // ---- actor
  FwdArgs a{};
  a.in = make_input(actor, view, rows);
  a.net = net_ptrs(actor_params, actor->in_dim, actor->h1, actor->h2, actor->out_dim);
  a.M = Ma;
  a.H1 = aH1;
  a.H2 = aH2;
  a.mask = mask;
  a.action_old = action;
  a.old_logp = old_logp;
  a.adv = adv;
  a.adv_stats = stats;
  a.mb_size = mb_size;
  a.num_replicas = num_replicas;
  a.clip_eps = hyper->clip_eps;
  a.ent_coef = hyper->ent_coef;
  a.vf_coef = hyper->vf_coef;
  a.dOut = aD;
  a.loss_acc = loss_acc;
  rc = launch_fwd<kTrainActor>(a, s);
  if (rc) return rc;
  // ---- critic
  FwdArgs c{};
  c.in = make_input(critic, view, rows);
  c.net = net_ptrs(critic_params, critic->in_dim, critic->h1, critic->h2, 1);
  c.M = Mc;
  c.H1 = cH1;
  c.H2 = cH2;
  c.old_value = old_value;
  c.targets = targets;
  c.mb_size = mb_size;
  c.num_replicas = num_replicas;
  c.clip_eps = hyper->clip_eps;
  c.ent_coef = hyper->ent_coef;
  c.vf_coef = hyper->vf_coef;
  c.dOut = cD;
  c.loss_acc = loss_acc;
  rc = launch_fwd<kTrainCritic>(c, s);
  if (rc) return rc;

  static bool wg_configured = false;
  if (!wg_configured) {
    cudaFuncSetAttribute(mlp_wgrad_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)sizeof(WgSmem));
    cudaFuncSetAttribute(mlp_wgrad_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)sizeof(WgSmem));
    wg_configured = true;
  }
  auto backward = [&](const FwdArgs& f, const mava_mlp_desc* d, float* H1, float* H2, float* D,
                      int ldd, int64_t M, float* g) -> int {
    NetPtrs gp = net_ptrs(g, d->in_dim, d->h1, d->h2, d->out_dim);
    float* gw1 = const_cast<float*>(gp.w1); float* gb1 = const_cast<float*>(gp.b1);
    float* gw2 = const_cast<float*>(gp.w2); float* gb2 = const_cast<float*>(gp.b2);
    float* gw3 = const_cast<float*>(gp.w3); float* gb3 = const_cast<float*>(gp.b3);
    const unsigned rb = (unsigned)ceil_div64(M, WG_ROWS);
    const int row_blocks = (int)ceil_div64(M, BM);
    InputDesc none{};
    // head: dW3 = H2^T D, then dZ2 = (D W3^T) * relu'(H2) in place
    mlp_wgrad_kernel<false><<<dim3(rb, ceil_div(d->h2, BM)), NT, sizeof(WgSmem), s>>>(
        H2, none, d->h2, D, ldd, d->out_dim, M, gw3, gb3);
    mlp_bwd_kernel<<<row_blocks, NT, 0, s>>>(D, ldd, d->out_dim, f.net.w3, d->h2, H2, M);
    mlp_wgrad_kernel<false><<<dim3(rb, ceil_div(d->h1, BM)), NT, sizeof(WgSmem), s>>>(
        H1, none, d->h1, H2, d->h2, d->h2, M, gw2, gb2);
    mlp_bwd_kernel<<<row_blocks, NT, 0, s>>>(H2, d->h2, d->h2, f.net.w2, d->h1, H1, M);
    mlp_wgrad_kernel<true><<<dim3(rb, ceil_div(d->in_dim, BM)), NT, sizeof(WgSmem), s>>>(
        nullptr, f.in, d->in_dim, H1, d->h1, d->h1, M, gw1, gb1);
    return launch_status();
  };
  rc = backward(a, actor, aH1, aH2, aD, OMAX, Ma, grad_out);
  if (rc) return rc;
  rc = backward(c, critic, cH1, cH2, cD, 1, Mc, grad_out + na);
  if (rc) return rc;
  finalize_loss_kernel<<<1, 32, 0, s>>>(loss_acc, (double)R * A, hyper->ent_coef, hyper->vf_coef,
                                        grad_out + na + nc);
  return launch_status();
}

}  // extern "C"
