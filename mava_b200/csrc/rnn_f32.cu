// Recurrent (GRU) actor / critic for rec_ippo / rec_mappo, full precision.
//
// Reference regions replaced:
//   ScannedRNN / RecurrentActor / RecurrentValueNet      mava/networks.py:238-331
//   _env_step of the recurrent systems                   mava/systems/ppo/rec_mappo.py:91-149
//   _actor_loss_fn / _critic_loss_fn + value_and_grad    mava/systems/ppo/rec_mappo.py:208-293
//   the chunk reshape + shuffle of the batch             mava/systems/ppo/rec_mappo.py:334-360
// flax.linen.GRUCell (third party, published algorithm):
//   r = sigmoid(x Wir + bir + h Whr)        z = sigmoid(x Wiz + biz + h Whz)
//   n = tanh(x Win + bin + r * (h Whn + bhn))        h' = (1 - z) * n + z * h
//
// Structure: everything that is not sequential is ONE dense contraction over all (time, sequence)
// rows (pre-torso, the x-side of the gates, post-torso, head, and every weight gradient); only
// h W_h (forward) and dG W_h^T (backward) run once per time step.  All contractions go through one
// tiled SGEMM (128 x 128 x 16 tiles, 8 x 8 register micro-tiles) with transposed-operand and
// split-K modes; the gate non-linearities, their backward, the categorical head and the PPO losses
// are element-wise kernels.  A tensor-core (tcgen05) version of the scan is the next step; this
// path is the rtol-1e-5 one the bf16 kernels will be checked against.
#include "common.cuh"
#include <cstdlib>

#include "gemm.cuh"
#include "gru_scan.cuh"
#include "prng.cuh"

namespace mava {
int launch_finalize_loss(const double* acc, double denom, float ent_coef, float vf_coef, float* out5,
                         cudaStream_t s);
namespace {

constexpr float kF32Min = -3.402823466e38f;
constexpr int OMAX = 16;  // max head width, also the row stride of logits buffers

// ------------------------------------------------------------------------------------------------
// SGEMM: C[M][N] (op)= alpha * opA(A)[M][K] * opB(B)[K][N] (+ bias) (relu) (* relu_ref > 0)
// ------------------------------------------------------------------------------------------------
constexpr int GBM = 128, GBK = 16, GT = 256;

template <int TN>
__global__ void __launch_bounds__(GT, 2) sgemm_kernel(const GemmArgs p) {
  constexpr int BN = TN * 16;
  constexpr int ALD = GBM + 4, BLD = BN + 4;
  constexpr int A_PER = GBM * GBK / GT;
  constexpr int B_PER = (BN * GBK + GT - 1) / GT;
  __shared__ __align__(16) float As[GBK][ALD];
  __shared__ __align__(16) float Bs[GBK][BLD];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * GBM, n0 = blockIdx.y * BN;
  const int kbeg = blockIdx.z * p.kchunk;
  const int kend = min(p.K, kbeg + p.kchunk);
  float acc[8][TN];
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int c = 0; c < TN; ++c) acc[r][c] = 0.0f;
  float ra[A_PER], rb[B_PER];

  auto fetch = [&](int k0) {
#pragma unroll
    for (int l = 0; l < A_PER; ++l) {
      const int idx = l * GT + tid;
      int i, k;
      if (p.ta == 0) { k = idx & (GBK - 1); i = idx / GBK; } else { i = idx & (GBM - 1); k = idx / GBM; }
      const int gi = m0 + i, gk = k0 + k;
      float v = 0.0f;
      if (gi < p.M && gk < kend)
        v = p.ta == 0 ? __ldg(p.A + (int64_t)gi * p.lda + gk) : __ldg(p.A + (int64_t)gk * p.lda + gi);
      ra[l] = v;
    }
#pragma unroll
    for (int l = 0; l < B_PER; ++l) {
      const int idx = l * GT + tid;
      float v = 0.0f;
      if (idx < BN * GBK) {
        int j, k;
        if (p.tb == 0) { j = idx % BN; k = idx / BN; } else { k = idx & (GBK - 1); j = idx / GBK; }
        const int gj = n0 + j, gk = k0 + k;
        if (gj < p.N && gk < kend)
          v = p.tb == 0 ? __ldg(p.B + (int64_t)gk * p.ldb + gj) : __ldg(p.B + (int64_t)gj * p.ldb + gk);
      }
      rb[l] = v;
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int l = 0; l < A_PER; ++l) {
      const int idx = l * GT + tid;
      int i, k;
      if (p.ta == 0) { k = idx & (GBK - 1); i = idx / GBK; } else { i = idx & (GBM - 1); k = idx / GBM; }
      As[k][i] = ra[l];
    }
#pragma unroll
    for (int l = 0; l < B_PER; ++l) {
      const int idx = l * GT + tid;
      if (idx < BN * GBK) {
        int j, k;
        if (p.tb == 0) { j = idx % BN; k = idx / BN; } else { k = idx & (GBK - 1); j = idx / GBK; }
        Bs[k][j] = rb[l];
      }
    }
  };

  if (kbeg < kend) fetch(kbeg);
  for (int k0 = kbeg; k0 < kend; k0 += GBK) {
    stash();
    __syncthreads();
    if (k0 + GBK < kend) fetch(k0 + GBK);  // next tile's loads fly while this one is multiplied
#pragma unroll
    for (int kk = 0; kk < GBK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][64 + ty * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[TN];
      if constexpr (TN == 8) {
        const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
        const float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][64 + tx * 4]);
        b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
        b[TN - 4] = b1.x; b[TN - 3] = b1.y; b[TN - 2] = b1.z; b[TN - 1] = b1.w;
      } else {
        const float2 b0 = *reinterpret_cast<const float2*>(&Bs[kk][tx * 2]);
        b[0] = b0.x; b[1] = b0.y;
      }
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < TN; ++c) acc[r][c] = fmaf(a[r], b[c], acc[r][c]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int i = m0 + (r < 4 ? ty * 4 + r : 64 + ty * 4 + (r - 4));
    if (i >= p.M) continue;
#pragma unroll
    for (int c = 0; c < TN; ++c) {
      const int j = n0 + (TN == 8 ? (c < 4 ? tx * 4 + c : 64 + tx * 4 + (c - 4)) : tx * 2 + c);
      if (j >= p.N) continue;
      float v = p.alpha * acc[r][c];
      if (p.bias != nullptr && blockIdx.z == 0) v += __ldg(p.bias + j);
      if (p.relu) v = fmaxf(v, 0.0f);
      if (p.relu_ref != nullptr && !(p.relu_ref[(int64_t)i * p.ldr + j] > 0.0f)) v = 0.0f;
      float* dst = p.C + (int64_t)i * p.ldc + j;
      if (p.mode == 0) *dst = v;
      else if (p.mode == 1) *dst += v;
      else atomicAdd(dst, v);
    }
  }
}

struct Gemm {
  GemmArgs a{};
  bool use_tc = false;
  Gemm(const float* A, int ta, int64_t lda, const float* B, int tb, int64_t ldb, float* C,
       int64_t ldc, int M, int N, int K) {
    a.A = A; a.ta = ta; a.lda = lda; a.B = B; a.tb = tb; a.ldb = ldb; a.C = C; a.ldc = ldc;
    a.M = M; a.N = N; a.K = K; a.alpha = 1.0f; a.kchunk = round_up(K, 64);
    a.bias = nullptr; a.relu_ref = nullptr; a.ldr = 0; a.relu = 0; a.mode = 0;
  }
  Gemm& bias(const float* b) { a.bias = b; return *this; }
  Gemm& relu() { a.relu = 1; return *this; }
  Gemm& relu_ref(const float* r, int64_t ldr) { a.relu_ref = r; a.ldr = ldr; return *this; }
  Gemm& accumulate() { a.mode = 1; return *this; }
  Gemm& tc(bool on) { use_tc = on; return *this; }
  // reduction over a very long K (weight gradients): split K over CTAs, combine with atomics
  Gemm& split_k_atomic() {
    a.mode = 2;
    const int tiles = ceil_div(a.M, GBM) * ceil_div(a.N, a.N > 32 ? 128 : 32);
    int splits = max(1, min(ceil_div(4 * sm_count(), tiles), ceil_div(a.K, 8 * GBK)));
    a.kchunk = round_up(ceil_div(a.K, splits), 64);
    return *this;
  }
  int run(cudaStream_t s) const { return use_tc ? launch_tc_gemm(a, s) : launch_sgemm(a, s); }
};

}  // namespace

int launch_sgemm(const GemmArgs& a, cudaStream_t s) {
  if (a.M <= 0 || a.N <= 0 || a.K <= 0) return 0;
  const int splits = ceil_div(a.K, a.kchunk);
  if (a.N > 32) {
    dim3 grid(ceil_div(a.M, GBM), ceil_div(a.N, 128), splits);
    sgemm_kernel<8><<<grid, GT, 0, s>>>(a);
  } else {
    dim3 grid(ceil_div(a.M, GBM), ceil_div(a.N, 32), splits);
    sgemm_kernel<2><<<grid, GT, 0, s>>>(a);
  }
  return launch_status();
}

namespace {

// db[j] += sum_i D[i][j]
__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ D, int64_t ldd, int64_t M, int N, int64_t rows_per_block,
              float* __restrict__ db) {
  const int j = blockIdx.x * 32 + (threadIdx.x & 31);
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  const int64_t r1 = min(M, r0 + rows_per_block);
  float s = 0.0f;
  if (j < N)
    for (int64_t i = r0 + (threadIdx.x >> 5); i < r1; i += 8) s += D[i * ldd + j];
  __shared__ float red[8][33];
  red[threadIdx.x >> 5][threadIdx.x & 31] = s;
  __syncthreads();
  if (threadIdx.x < 32 && j < N) {
    float t = 0.0f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
    atomicAdd(db + j, t);
  }
}

int launch_colsum(const float* D, int64_t ldd, int64_t M, int N, float* db, cudaStream_t s) {
  if (M <= 0 || N <= 0) return 0;
  const int64_t rpb = max((int64_t)256, ceil_div64(M, (int64_t)sm_count() * 2));
  dim3 grid(ceil_div(N, 32), (unsigned)ceil_div64(M, rpb));
  colsum_kernel<<<grid, 256, 0, s>>>(D, ldd, M, N, rpb, db);
  return launch_status();
}

// ------------------------------------------------------------------------------------------------
// network description
// ------------------------------------------------------------------------------------------------
struct Net {
  int in_dim, H, Q, out, rows_per_env, mode, add_id, A, FR;
  bool tc;  // bf16 tensor-core contractions (mava_rnn_desc.precision == 1)
  const float *w_pre, *b_pre, *w_i, *b_i, *w_h, *b_hn, *w_post, *b_post, *w_head, *b_head;
};

int check_desc(const mava_rnn_desc* d) {
  if (!d) return MAVA_E_NULL;
  if (d->hidden < 1 || d->hidden > 1024 || d->post < 1 || d->post > 1024) return MAVA_E_UNSUPPORTED;
  if (d->out_dim < 1 || d->out_dim > OMAX) return MAVA_E_UNSUPPORTED;
  if (d->num_agents < 1) return MAVA_E_BADARG;
  if (d->precision != 0 && d->precision != 1) return MAVA_E_BADARG;
  if (d->input_mode == MAVA_IN_DENSE) {
    if (d->in_dim < 1) return MAVA_E_BADARG;
    if (d->rows_per_env != 1 && d->rows_per_env != d->num_agents) return MAVA_E_BADARG;
  } else if (d->input_mode == MAVA_IN_GLOBAL) {
    if (d->view_dim < 1 || d->in_dim != d->num_agents * d->view_dim || d->rows_per_env != 1)
      return MAVA_E_BADARG;
  } else if (d->input_mode == MAVA_IN_AGENT_VIEW) {
    if (d->view_dim < 1 || d->rows_per_env != d->num_agents ||
        d->in_dim != d->view_dim + (d->add_agent_id ? d->num_agents : 0))
      return MAVA_E_BADARG;
  } else {
    return MAVA_E_BADARG;
  }
  return 0;
}

template <typename T>
Net make_net(const mava_rnn_desc* d, T* p) {
  Net n;
  n.in_dim = d->in_dim; n.H = d->hidden; n.Q = d->post; n.out = d->out_dim;
  n.rows_per_env = d->rows_per_env; n.mode = d->input_mode; n.add_id = d->add_agent_id;
  n.A = d->num_agents; n.FR = d->view_dim;
  n.tc = d->precision == 1;
  const int H = n.H;
  n.w_pre = p; p += (int64_t)n.in_dim * H;
  n.b_pre = p; p += H;
  n.w_i = p; p += (int64_t)H * 3 * H;
  n.b_i = p; p += 3 * H;
  n.w_h = p; p += (int64_t)H * 3 * H;
  n.b_hn = p; p += H;
  n.w_post = p; p += (int64_t)H * n.Q;
  n.b_post = p; p += n.Q;
  n.w_head = p; p += (int64_t)n.Q * n.out;
  n.b_head = p;
  return n;
}

// ------------------------------------------------------------------------------------------------
// element-wise kernels
// ------------------------------------------------------------------------------------------------
// X[(l*Senv + q)*rpe + a][:] = network input of env-step steps[l*Senv + q] (identity when steps is
// null): [onehot(a) | view] (AGENT_VIEW), concat_a view (GLOBAL), or the dense f32 row.
// Rows of X are `ldx` floats apart (in_dim rounded up to 8): 32-byte aligned rows let the contractions
// read them with 256-bit loads; the padding columns are never read.
__global__ void __launch_bounds__(256)
expand_obs_kernel(const int8_t* __restrict__ view, const float* __restrict__ dense,
                  const int32_t* __restrict__ steps, int64_t rows, int rpe, int mode, int add_id,
                  int A, int FR, int in_dim, int ldx, float* __restrict__ X) {
  const int64_t total = rows * in_dim;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = idx / in_dim;
    const int k = (int)(idx - row * in_dim);
    const int64_t q = row / rpe;
    const int a = (int)(row - q * rpe);
    const int64_t st = steps ? steps[q] : q;
    float v;
    if (mode == MAVA_IN_DENSE) {
      v = dense[(st * rpe + a) * in_dim + k];
    } else if (mode == MAVA_IN_GLOBAL) {
      v = (float)view[st * A * FR + k];
    } else if (add_id && k < A) {
      v = k == a ? 1.0f : 0.0f;
    } else {
      v = (float)view[(st * A + a) * FR + (add_id ? k - A : k)];
    }
    X[row * ldx + k] = v;
  }
}

// steps[l][u*mbc + j] = (l*nc + c)*NE + u*E + e with (c, e) = divmod(cols[j], E): the env-step a
// sequence position reads after the reference's reshape (T, E) -> (chunk, E*nc) + take(cols).
__global__ void rec_steps_kernel(const int32_t* __restrict__ cols, int mbc, int U, int E, int nc,
                                 int NE, int L, int32_t* __restrict__ steps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int Senv = U * mbc;
  if (i >= L * Senv) return;
  const int l = i / Senv, q = i - l * Senv;
  const int u = q / mbc, j = q - u * mbc;
  const int col = cols[j];
  const int c = col / E, e = col - c * E;
  steps[i] = (l * nc + c) * NE + u * E + e;
}

// Hin[row][:] = done_in[step(row)] ? 0 : hsrc[srow][:], where for the first position of a chunk
// hsrc is the stored chunk-start state of that env (hs[c][env*rpe + a]) -- ScannedRNN's reset.
__global__ void __launch_bounds__(256)
mask_hidden_kernel(const float* __restrict__ hsrc, const int32_t* __restrict__ steps,
                   const uint8_t* __restrict__ done_in, int64_t rows, int rpe, int H, int NE,
                   int gather, float* __restrict__ Hin) {
  const int64_t total = rows * H;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = idx / H;
    const int j = (int)(idx - row * H);
    const int64_t q = row / rpe;
    const int a = (int)(row - q * rpe);
    const int64_t st = steps ? steps[q] : q;
    float v = 0.0f;
    if (!done_in[st]) {
      // gather: hsrc is [T'][NE*rpe][H] indexed by the env-step itself (chunk starts: st < nc*NE)
      const int64_t src = gather ? (st * rpe + a) : row;
      v = hsrc[src * H + j];
    }
    Hin[idx] = v;
  }
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// Gates of one time step.  Gx = x Wi + bi, Gh = hin Wh (no bias).  Writes y = h' and, when given,
// the stash (r, z, n, q = hin Whn + bhn) for the backward pass and next_hin = done_next ? 0 : h'.
__global__ void __launch_bounds__(256)
gru_fwd_kernel(const float* __restrict__ Gx, const float* __restrict__ Gh,
               const float* __restrict__ b_hn, const float* __restrict__ Hin, int64_t rows, int H,
               float* __restrict__ Hout, float* __restrict__ gates,
               const int32_t* __restrict__ next_steps, const uint8_t* __restrict__ done_in, int rpe,
               float* __restrict__ next_hin) {
  const int64_t total = rows * H;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = idx / H;
    const int j = (int)(idx - row * H);
    const float* gx = Gx + row * 3 * H;
    const float* gh = Gh + row * 3 * H;
    const float r = sigmoidf_(gx[j] + gh[j]);
    const float z = sigmoidf_(gx[H + j] + gh[H + j]);
    const float q = gh[2 * H + j] + b_hn[j];
    const float n = tanhf(gx[2 * H + j] + r * q);
    const float hin = Hin[idx];
    const float h = (1.0f - z) * n + z * hin;
    Hout[idx] = h;
    if (gates) {
      float* g = gates + row * 4 * H;
      g[j] = r;
      g[H + j] = z;
      g[2 * H + j] = n;
      g[3 * H + j] = q;
    }
    if (next_hin) next_hin[idx] = done_in[next_steps[row / rpe]] ? 0.0f : h;
  }
}

// Backward of the gates of one time step.  dh = dHout (+ carry where the next step did not reset).
// Overwrites the stash row with dGh = [da_r | da_z | dq | .], writes dGx = [da_r | da_z | da_n]
// and dHin = dh * z; the W_h term dGh W_h^T is produced by the GEMM that follows into its own buffer
// (carry2 of the next call), so that GEMM is a plain store instead of a read-modify-write.
__global__ void __launch_bounds__(256)
gru_bwd_kernel(const float* __restrict__ dHout, const float* __restrict__ carry,
               const float* __restrict__ carry2, const int32_t* __restrict__ next_steps, const uint8_t* __restrict__ done_in, int rpe,
               const float* __restrict__ Hin, float* __restrict__ gates, float* __restrict__ dGx,
               float* __restrict__ dHin, int64_t rows, int H) {
  const int64_t total = rows * H;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = idx / H;
    const int j = (int)(idx - row * H);
    float dh = dHout[idx];
    if (carry != nullptr && !done_in[next_steps[row / rpe]]) dh += carry[idx] + carry2[idx];
    float* g = gates + row * 4 * H;
    const float r = g[j], z = g[H + j], n = g[2 * H + j], q = g[3 * H + j];
    const float hin = Hin[idx];
    const float dn = dh * (1.0f - z);
    const float dz = dh * (hin - n);
    const float da_n = dn * (1.0f - n * n);
    const float da_r = da_n * q * r * (1.0f - r);
    const float da_z = dz * z * (1.0f - z);
    const float dq = da_n * r;
    g[j] = da_r;
    g[H + j] = da_z;
    g[2 * H + j] = dq;
    float* gx = dGx + row * 3 * H;
    gx[j] = da_r;
    gx[H + j] = da_z;
    gx[2 * H + j] = da_n;
    dHin[idx] = dh * z;
  }
}

struct HeadArgs {
  const float* logits;  // [rows][OMAX] (actor) or values [rows] (critic)
  int64_t rows;
  int rpe, A, N;
  const int32_t* steps;  // env-step per row / rpe (null: identity)
  // acting
  const void* mask;      // [steps][A]: uint8 entries, uint16 when N > 8 (bit k = action k legal)
  const uint32_t* policy_key;
  int envs_per_replica, greedy;
  const int8_t* actions_in;
  int8_t* action;
  float* logp;
  float* value;  // [steps][A]
  // training
  const int8_t* action_old;
  const float* old_logp;
  const float* old_value;
  const float* adv;
  const float* targets;
  const double* adv_stats;  // [U][2]
  int64_t rows_per_replica;  // rows of one replica per time position
  int64_t rows_per_pos;      // rows of all replicas per time position
  double count_per_replica;  // elements the per-replica means run over
  int num_replicas;
  float clip_eps, ent_coef, vf_coef;
  float* dOut;       // dlogits [rows][OMAX] or dvalue [rows]
  double* loss_acc;  // [5]
};

__device__ __forceinline__ uint32_t mask_at(const HeadArgs& p, int64_t flat) {
  return p.N > 8 ? (uint32_t)static_cast<const uint16_t*>(p.mask)[flat]
                 : (uint32_t)static_cast<const uint8_t*>(p.mask)[flat];
}

// masked categorical: sample / mode / replay, log-prob (acting)
__global__ void __launch_bounds__(256) rec_sample_kernel(const HeadArgs p) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= p.rows) return;
  const int64_t q = row / p.A;
  const int a = (int)(row - q * p.A);
  const int64_t flat = (p.steps ? p.steps[q] : q) * p.A + a;
  const uint32_t mk = mask_at(p, flat);
  float out[OMAX];
  float mx = kF32Min;
#pragma unroll
  for (int j = 0; j < OMAX; ++j) {
    out[j] = kF32Min;
    if (j < p.N) {
      out[j] = ((mk >> j) & 1) ? p.logits[row * OMAX + j] : kF32Min;
      mx = fmaxf(mx, out[j]);
    }
  }
  float se = 0.0f;
#pragma unroll
  for (int j = 0; j < OMAX; ++j)
    if (j < p.N) se += expf(out[j] - mx);
  const float lse = mx + logf(se);
  int act = 0;
  if (p.actions_in) {
    act = p.actions_in[flat];
  } else if (p.greedy) {
    float best = out[0];
#pragma unroll
    for (int j = 1; j < OMAX; ++j)
      if (j < p.N && out[j] > best) { best = out[j]; act = j; }
  } else {
    const Key key{p.policy_key[0], p.policy_key[1]};
    const int64_t e = q % p.envs_per_replica;
    const uint32_t size = (uint32_t)p.envs_per_replica * p.A * p.N;
    const uint32_t base = (uint32_t)((e * p.A + a) * p.N);
    float best = 0.0f;
#pragma unroll
    for (int j = 0; j < OMAX; ++j) {
      if (j < p.N) {
        const float zz = bits_to_gumbel(random_bits_at(key, base + j, size)) + out[j];
        if (j == 0 || zz > best) { best = zz; act = j; }
      }
    }
  }
  float la = 0.0f;
#pragma unroll
  for (int j = 0; j < OMAX; ++j)
    if (j == act) la = out[j] - lse;
  p.action[flat] = (int8_t)act;
  p.logp[flat] = la;
}

// value[step][a] = v[row] (rpe == A) or v[row] for every a (centralised critic evaluated per env)
__global__ void __launch_bounds__(256) rec_value_kernel(const HeadArgs p) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= p.rows) return;
  const float v = p.logits[row];
  if (p.rpe == 1) {
    for (int a = 0; a < p.A; ++a) p.value[row * p.A + a] = v;
  } else {
    p.value[row] = v;
  }
}

__device__ __forceinline__ void block_add2(double l0, double l1, double* dst0, double* dst1) {
  for (int o = 16; o > 0; o >>= 1) {
    l0 += __shfl_xor_sync(0xffffffffu, l0, o);
    l1 += __shfl_xor_sync(0xffffffffu, l1, o);
  }
  __shared__ double s0[8], s1[8];
  if ((threadIdx.x & 31) == 0) {
    s0[threadIdx.x >> 5] = l0;
    s1[threadIdx.x >> 5] = l1;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
      l0 += s0[w];
      l1 += s1[w];
    }
    atomicAdd(dst0, l0);
    if (dst1) atomicAdd(dst1, l1);
  }
}

// per-replica sum / sum of squares of the minibatch advantages (gae.mean(), gae.std())
__global__ void __launch_bounds__(256)
rec_adv_stats_kernel(const float* __restrict__ adv, const int32_t* __restrict__ steps, int L,
                     int Senv, int mbc, int A, double* __restrict__ stats) {
  const int u = blockIdx.y;
  double s = 0.0, ss = 0.0;
  const int64_t n = (int64_t)L * mbc * A;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int a = (int)(i % A);
    const int64_t lj = i / A;
    const int l = (int)(lj / mbc), j = (int)(lj - (int64_t)l * mbc);
    const float v = adv[(int64_t)steps[(int64_t)l * Senv + u * mbc + j] * A + a];
    s += v;
    ss += (double)v * v;
  }
  block_add2(s, ss, stats + 2 * u, stats + 2 * u + 1);
}

// _actor_loss_fn (rec_mappo.py:208-243) and d(total_loss)/d(logits)
__global__ void __launch_bounds__(256) rec_actor_loss_kernel(const HeadArgs p) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double l0 = 0.0, l1 = 0.0;
  if (row < p.rows) {
    const int64_t q = row / p.A;
    const int a = (int)(row - q * p.A);
    const int64_t flat = (int64_t)p.steps[q] * p.A + a;
    const int u = (int)((row % p.rows_per_pos) / p.rows_per_replica);
    const uint32_t mk = mask_at(p, flat);
    float out[OMAX], logp[OMAX], pr[OMAX];
    float mx = kF32Min;
#pragma unroll
    for (int j = 0; j < OMAX; ++j) {
      out[j] = kF32Min;
      if (j < p.N) {
        out[j] = ((mk >> j) & 1) ? p.logits[row * OMAX + j] : kF32Min;
        mx = fmaxf(mx, out[j]);
      }
    }
    float se = 0.0f;
#pragma unroll
    for (int j = 0; j < OMAX; ++j)
      if (j < p.N) se += expf(out[j] - mx);
    const float lse = mx + logf(se);
    const int act = p.action_old[flat];
    float la = 0.0f, ent = 0.0f;
#pragma unroll
    for (int j = 0; j < OMAX; ++j) {
      logp[j] = 0.0f;
      pr[j] = 0.0f;
      if (j < p.N) {
        logp[j] = out[j] - lse;
        pr[j] = expf(logp[j]);
        if (pr[j] != 0.0f) ent -= pr[j] * logp[j];
        if (j == act) la = logp[j];
      }
    }
    const double cnt = p.count_per_replica;
    const double mean_d = p.adv_stats[2 * u] / cnt;
    const double var_d = fmax(p.adv_stats[2 * u + 1] / cnt - mean_d * mean_d, 0.0);
    const float mean = (float)mean_d, sd = (float)sqrt(var_d);
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
    const float w = (float)(1.0 / ((double)p.num_replicas * cnt));
#pragma unroll
    for (int j = 0; j < OMAX; ++j) {
      float dl = 0.0f;
      if (j < p.N && ((mk >> j) & 1)) {
        dl = dla * ((j == act ? 1.0f : 0.0f) - pr[j]);
        if (pr[j] != 0.0f) dl += p.ent_coef * pr[j] * (logp[j] + ent);
      }
      p.dOut[row * OMAX + j] = dl * w;
    }
    l0 = (double)(-fminf(t1, t2));
    l1 = (double)ent;
  }
  block_add2(l0, l1, p.loss_acc + 0, p.loss_acc + 1);
}

// _critic_loss_fn (rec_mappo.py:245-268) and d(total_loss)/d(value)
__global__ void __launch_bounds__(256) rec_critic_loss_kernel(const HeadArgs p) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double l0 = 0.0;
  if (row < p.rows) {
    const int64_t q = row / p.rpe;
    const int a0 = (int)(row - q * p.rpe);
    const int reps = p.rpe == 1 ? p.A : 1;
    const int64_t base = (int64_t)p.steps[q] * p.A + (p.rpe == 1 ? 0 : a0);
    const float v = p.logits[row];
    const float w = (float)(1.0 / ((double)p.num_replicas * p.count_per_replica));
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
  block_add2(l0, 0.0, p.loss_acc + 2, nullptr);
}

inline unsigned ew_blocks(int64_t n) {
  return (unsigned)max((int64_t)1, min(ceil_div64(n, 256), (int64_t)sm_count() * 8));
}

int64_t align256(int64_t v) { return (v + 255) / 256 * 256; }

// ------------------------------------------------------------------------------------------------
// forward of one network over `L` time positions of `S` sequences (rows = S per position)
// ------------------------------------------------------------------------------------------------
struct Work {
  float *X, *E1, *Gx, *gates, *Hin, *Hout, *P, *out, *Gh, *dH0, *dH1, *dT0, *dT1;
};

int64_t work_floats(const mava_rnn_desc* d, int64_t R, int64_t S, bool train) {
  const int64_t H = d->hidden, Q = d->post;
  int64_t f = 0;
  f += align256(R * round_up(d->in_dim, 8)) + align256(R * H) + align256(R * 3 * H);  // X, E1, Gx
  f += train ? align256(R * 4 * H) : 0;                                   // gates
  f += align256(R * H) * 2 + align256(R * Q) + align256(R * OMAX);        // Hin, Hout, P, out
  f += align256(S * 3 * H) + 4 * align256(S * H);                         // Gh, dH0/1, dT0/1
  return f;
}

Work carve(float* w, const mava_rnn_desc* d, int64_t R, int64_t S, bool train) {
  const int64_t H = d->hidden, Q = d->post;
  Work k;
  k.X = w; w += align256(R * round_up(d->in_dim, 8));
  k.E1 = w; w += align256(R * H);
  k.Gx = w; w += align256(R * 3 * H);
  k.gates = train ? w : nullptr; w += train ? align256(R * 4 * H) : 0;
  k.Hin = w; w += align256(R * H);
  k.Hout = w; w += align256(R * H);
  k.P = w; w += align256(R * Q);
  k.out = w; w += align256(R * OMAX);
  k.Gh = w; w += align256(S * 3 * H);
  k.dH0 = w; w += align256(S * H);
  k.dH1 = w; w += align256(S * H);
  k.dT0 = w; w += align256(S * H);
  k.dT1 = w;
  return k;
}

struct SeqInput {
  const int8_t* view;
  const float* dense;
  const int32_t* steps;    // [L][Senv] or null (acting: identity over NE envs)
  const uint8_t* done_in;  // indexed by env-step
  const float* h0;         // acting: [S][H] current state; training: chunk-start stash
  int h0_gather;           // training: h0 indexed by env-step
  int L;
  int64_t Senv;  // env-sequences per position
};

// bf16 path, training: the whole scan in one persistent launch (W_h resident in shared memory, the
// hidden state in registers; gru_scan.cu); the per-step schedule stays for fp32 and for acting.
bool scan_capable(const Net& n, const SeqInput& in, const Work& k) {
  const bool off = getenv("MAVA_NO_GRU_SCAN") != nullptr;  // development / A-B test switch
  return !off && n.tc && n.H == 128 && in.L > 1 && in.steps != nullptr && k.gates != nullptr;
}

GruScanNet scan_net(const Net& n, const SeqInput& in, const Work& k) {
  GruScanNet g;
  g.Wh = n.w_h; g.b_hn = n.b_hn; g.Gx = k.Gx; g.gates = k.gates; g.Hin = k.Hin; g.Hout = k.Hout;
  g.steps = in.steps; g.done_in = in.done_in; g.Senv = in.Senv;
  g.S = in.Senv * n.rows_per_env; g.rpe = n.rows_per_env; g.L = in.L;
  return g;
}

// forward, part 1: observation rows, pre-torso, x-side of the gates, reset-masked chunk-start state
int forward_pre(const Net& n, const SeqInput& in, const Work& k, cudaStream_t s) {
  const int H = n.H, L = in.L;
  const int64_t S = in.Senv * n.rows_per_env, R = S * L;
  expand_obs_kernel<<<ew_blocks(R * n.in_dim), 256, 0, s>>>(
      in.view, in.dense, in.steps, R, n.rows_per_env, n.mode, n.add_id, n.A, n.FR, n.in_dim,
      round_up(n.in_dim, 8), k.X);
  int rc = Gemm(k.X, 0, round_up(n.in_dim, 8), n.w_pre, 0, H, k.E1, H, (int)R, H, n.in_dim).bias(n.b_pre).relu().tc(n.tc).run(s);
  if (rc) return rc;
  rc = Gemm(k.E1, 0, H, n.w_i, 0, 3 * H, k.Gx, 3 * H, (int)R, 3 * H, H).bias(n.b_i).tc(n.tc).run(s);
  if (rc) return rc;
  mask_hidden_kernel<<<ew_blocks(S * H), 256, 0, s>>>(in.h0, in.steps, in.done_in, S,
                                                       n.rows_per_env, H, 0, in.h0_gather, k.Hin);
  return launch_status();
}

// forward, part 2: the scan over the L positions, one GEMM + one gate kernel per position
int forward_scan_steps(const Net& n, const SeqInput& in, const Work& k, cudaStream_t s) {
  const int H = n.H, L = in.L;
  const int64_t S = in.Senv * n.rows_per_env;
  for (int l = 0; l < L; ++l) {
    const float* hin = k.Hin + (int64_t)l * S * H;
    int rc = Gemm(hin, 0, H, n.w_h, 0, 3 * H, k.Gh, 3 * H, (int)S, 3 * H, H).tc(n.tc).run(s);
    if (rc) return rc;
    const bool more = l + 1 < L;
    gru_fwd_kernel<<<ew_blocks(S * H), 256, 0, s>>>(
        k.Gx + (int64_t)l * S * 3 * H, k.Gh, n.b_hn, hin, S, H, k.Hout + (int64_t)l * S * H,
        k.gates ? k.gates + (int64_t)l * S * 4 * H : nullptr,
        more ? in.steps + (int64_t)(l + 1) * in.Senv : nullptr, in.done_in, n.rows_per_env,
        more ? k.Hin + (int64_t)(l + 1) * S * H : nullptr);
  }
  return launch_status();
}

// forward, part 3: post-torso and head.  k.out holds logits [R][OMAX] or values [R].
int forward_post(const Net& n, const SeqInput& in, const Work& k, cudaStream_t s) {
  const int H = n.H;
  const int64_t S = in.Senv * n.rows_per_env, R = S * in.L;
  int rc = Gemm(k.Hout, 0, H, n.w_post, 0, n.Q, k.P, n.Q, (int)R, n.Q, H).bias(n.b_post).relu().tc(n.tc).run(s);
  if (rc) return rc;
  const int ldo = n.out == 1 ? 1 : OMAX;
  rc = Gemm(k.P, 0, n.Q, n.w_head, 0, n.out, k.out, ldo, (int)R, n.out, n.Q).bias(n.b_head).tc(n.tc).run(s);
  if (rc) return rc;
  return launch_status();
}

// Runs pre-torso, GRU scan, post-torso and head of one network.
int forward(const Net& n, const SeqInput& in, const Work& k, cudaStream_t s) {
  int rc = forward_pre(n, in, k, s);
  if (rc) return rc;
  if (scan_capable(n, in, k)) {
    const GruScanNet g = scan_net(n, in, k);
    rc = launch_gru_scan_fwd(&g, 1, s);
  } else {
    rc = forward_scan_steps(n, in, k, s);
  }
  if (rc) return rc;
  return forward_post(n, in, k, s);
}

// Backward from dOut (in k.out: dlogits [R][OMAX] or dvalue [R]) into the flat gradient g.
// Part 1: head and post-torso, down to dHout (in place over Hout).
int backward_pre(const Net& n, const Net& g, const SeqInput& in, const Work& k, cudaStream_t s) {
  const int H = n.H, L = in.L, Q = n.Q;
  const int64_t S = in.Senv * n.rows_per_env, R = S * L;
  const int ldo = n.out == 1 ? 1 : OMAX;
  auto G = [](const float* p) { return const_cast<float*>(p); };
  int rc;
  // head
  rc = Gemm(k.P, 1, Q, k.out, 0, ldo, G(g.w_head), n.out, Q, n.out, (int)R).split_k_atomic().tc(n.tc).run(s);
  if (rc) return rc;
  rc = launch_colsum(k.out, ldo, R, n.out, G(g.b_head), s);
  if (rc) return rc;
  // dP = (dOut W_head^T) * relu'(P), in place over P
  rc = Gemm(k.out, 0, ldo, n.w_head, 1, n.out, k.P, Q, (int)R, Q, n.out).relu_ref(k.P, Q).tc(n.tc).run(s);
  if (rc) return rc;
  rc = Gemm(k.Hout, 1, H, k.P, 0, Q, G(g.w_post), Q, H, Q, (int)R).split_k_atomic().tc(n.tc).run(s);
  if (rc) return rc;
  rc = launch_colsum(k.P, Q, R, Q, G(g.b_post), s);
  if (rc) return rc;
  // dHout = dP W_post^T, in place over Hout
  return Gemm(k.P, 0, Q, n.w_post, 1, Q, k.Hout, H, (int)R, H, Q).tc(n.tc).run(s);
}

// Part 2: the reverse scan, one gate kernel + one GEMM per position.
int backward_scan_steps(const Net& n, const SeqInput& in, const Work& k, cudaStream_t s) {
  const int H = n.H, L = in.L;
  const int64_t S = in.Senv * n.rows_per_env;
  float* dH[2] = {k.dH0, k.dH1};
  float* dT[2] = {k.dT0, k.dT1};
  for (int l = L - 1; l >= 0; --l) {
    const bool has_next = l + 1 < L;
    gru_bwd_kernel<<<ew_blocks(S * H), 256, 0, s>>>(
        k.Hout + (int64_t)l * S * H, has_next ? dH[(l + 1) & 1] : nullptr,
        has_next ? dT[(l + 1) & 1] : nullptr,
        has_next ? in.steps + (int64_t)(l + 1) * in.Senv : nullptr, in.done_in, n.rows_per_env,
        k.Hin + (int64_t)l * S * H, k.gates + (int64_t)l * S * 4 * H,
        k.Gx + (int64_t)l * S * 3 * H, dH[l & 1], S, H);
    if (l > 0) {  // dGh W_h^T (the chunk-start state carries no gradient)
      int rc = Gemm(k.gates + (int64_t)l * S * 4 * H, 0, 4 * H, n.w_h, 1, 3 * H, dT[l & 1], H, (int)S,
                    H, 3 * H).tc(n.tc).run(s);
      if (rc) return rc;
    }
  }
  return launch_status();
}

// Part 3: recurrent and input weights of the cell and the pre-torso -- contractions over all
// (position, sequence) rows.
int backward_post(const Net& n, const Net& g, const SeqInput& in, const Work& k, cudaStream_t s) {
  const int H = n.H, L = in.L;
  const int64_t S = in.Senv * n.rows_per_env, R = S * L;
  auto G = [](const float* p) { return const_cast<float*>(p); };
  int rc = Gemm(k.Hin, 1, H, k.gates, 0, 4 * H, G(g.w_h), 3 * H, H, 3 * H, (int)R).split_k_atomic().tc(n.tc).run(s);
  if (rc) return rc;
  rc = launch_colsum(k.gates + 2 * H, 4 * H, R, H, G(g.b_hn), s);
  if (rc) return rc;
  rc = Gemm(k.E1, 1, H, k.Gx, 0, 3 * H, G(g.w_i), 3 * H, H, 3 * H, (int)R).split_k_atomic().tc(n.tc).run(s);
  if (rc) return rc;
  rc = launch_colsum(k.Gx, 3 * H, R, 3 * H, G(g.b_i), s);
  if (rc) return rc;
  // dE1 = (dGx W_i^T) * relu'(E1), in place over E1
  rc = Gemm(k.Gx, 0, 3 * H, n.w_i, 1, 3 * H, k.E1, H, (int)R, H, 3 * H).relu_ref(k.E1, H).tc(n.tc).run(s);
  if (rc) return rc;
  rc = Gemm(k.X, 1, round_up(n.in_dim, 8), k.E1, 0, H, G(g.w_pre), H, n.in_dim, H, (int)R).split_k_atomic().tc(n.tc).run(s);
  if (rc) return rc;
  return launch_colsum(k.E1, H, R, H, G(g.b_pre), s);
}

}  // namespace
}  // namespace mava

using namespace mava;

extern "C" {

int mava_gemm(int use_tc, const float* A, int ta, int64_t lda, const float* B, int tb, int64_t ldb,
              float* C, int64_t ldc, int M, int N, int K, const float* bias, int relu,
              const float* relu_ref, int64_t ldr, int mode, int k_splits, mava_stream_t s) {
  MAVA_CHECK_PTR(A);
  MAVA_CHECK_PTR(B);
  MAVA_CHECK_PTR(C);
  MAVA_CHECK_ARG(M > 0 && N > 0 && K > 0 && mode >= 0 && mode <= 2 && k_splits >= 1);
  MAVA_CHECK_ARG(k_splits == 1 || mode == 2);
  GemmArgs a{};
  a.A = A; a.ta = ta; a.lda = lda; a.B = B; a.tb = tb; a.ldb = ldb; a.C = C; a.ldc = ldc;
  a.M = M; a.N = N; a.K = K; a.bias = bias; a.relu = relu; a.relu_ref = relu_ref; a.ldr = ldr;
  a.mode = mode; a.alpha = 1.0f;
  a.kchunk = round_up(ceil_div(K, k_splits), 64);
  return use_tc ? launch_tc_gemm(a, as_stream(s)) : launch_sgemm(a, as_stream(s));
}

int64_t mava_rnn_param_count(const mava_rnn_desc* d) {
  if (!d) return -1;
  const int64_t H = d->hidden, Q = d->post;
  return (int64_t)d->in_dim * H + H + 2 * (H * 3 * H) + 3 * H + H + H * Q + Q + Q * d->out_dim +
         d->out_dim;
}

int64_t mava_rec_act_workspace_bytes(const mava_rnn_desc* actor, const mava_rnn_desc* critic,
                                     int num_envs) {
  if (!critic || num_envs <= 0) return -1;
  int64_t f = work_floats(critic, (int64_t)num_envs * critic->rows_per_env,
                          (int64_t)num_envs * critic->rows_per_env, false);
  if (actor)
    f = max(f, work_floats(actor, (int64_t)num_envs * actor->rows_per_env,
                           (int64_t)num_envs * actor->rows_per_env, false));
  return f * 4 + 1024;
}

int mava_rec_act(const mava_rnn_desc* actor, const float* actor_params,
                 const mava_rnn_desc* critic, const float* critic_params, const int8_t* view,
                 const float* obs_actor, const float* obs_critic, const void* mask,
                 const uint8_t* done_in, const float* h_actor_in, float* h_actor_out,
                 const float* h_critic_in, float* h_critic_out, const uint32_t* policy_key,
                 int envs_per_replica, int num_envs, int greedy, const int8_t* actions_in,
                 int8_t* action, float* logp, float* value, void* workspace,
                 mava_stream_t stream) {
  MAVA_CHECK_PTR(workspace);
  MAVA_CHECK_PTR(done_in);
  MAVA_CHECK_ARG(num_envs > 0 && envs_per_replica > 0);
  cudaStream_t s = as_stream(stream);
  float* w = static_cast<float*>(workspace);
  int rc;
  if (actor != nullptr) {
    rc = check_desc(actor);
    if (rc) return rc;
    MAVA_CHECK_PTR(actor_params);
    MAVA_CHECK_PTR(mask);
    MAVA_CHECK_PTR(h_actor_in);
    MAVA_CHECK_PTR(action);
    MAVA_CHECK_PTR(logp);
    MAVA_CHECK_ARG(actor->rows_per_env == actor->num_agents);
    MAVA_CHECK_ARG(actor->input_mode == MAVA_IN_DENSE ? obs_actor != nullptr : view != nullptr);
    MAVA_CHECK_ARG(greedy || actions_in != nullptr || policy_key != nullptr);
    const Net n = make_net(actor, actor_params);
    const int64_t S = (int64_t)num_envs * actor->rows_per_env;
    Work k = carve(w, actor, S, S, false);
    SeqInput in{view, obs_actor, nullptr, done_in, h_actor_in, 0, 1, num_envs};
    rc = forward(n, in, k, s);
    if (rc) return rc;
    HeadArgs h{};
    h.logits = k.out; h.rows = S; h.rpe = actor->rows_per_env; h.A = actor->num_agents;
    h.N = actor->out_dim; h.mask = mask; h.policy_key = policy_key;
    h.envs_per_replica = envs_per_replica; h.greedy = greedy; h.actions_in = actions_in;
    h.action = action; h.logp = logp;
    rec_sample_kernel<<<(unsigned)ceil_div64(S, 256), 256, 0, s>>>(h);
    if (h_actor_out) {
      cudaError_t e = cudaMemcpyAsync(h_actor_out, k.Hout, (size_t)S * actor->hidden * 4,
                                      cudaMemcpyDeviceToDevice, s);
      if (e != cudaSuccess) return (int)e;
    }
  }
  if (critic != nullptr && value != nullptr) {
    rc = check_desc(critic);
    if (rc) return rc;
    MAVA_CHECK_PTR(critic_params);
    MAVA_CHECK_PTR(h_critic_in);
    MAVA_CHECK_ARG(critic->out_dim == 1);
    MAVA_CHECK_ARG(critic->input_mode == MAVA_IN_DENSE ? obs_critic != nullptr : view != nullptr);
    const Net n = make_net(critic, critic_params);
    const int64_t S = (int64_t)num_envs * critic->rows_per_env;
    Work k = carve(w, critic, S, S, false);
    SeqInput in{view, obs_critic, nullptr, done_in, h_critic_in, 0, 1, num_envs};
    rc = forward(n, in, k, s);
    if (rc) return rc;
    HeadArgs h{};
    h.logits = k.out; h.rows = S; h.rpe = critic->rows_per_env; h.A = critic->num_agents;
    h.value = value;
    rec_value_kernel<<<(unsigned)ceil_div64(S, 256), 256, 0, s>>>(h);
    if (h_critic_out) {
      cudaError_t e = cudaMemcpyAsync(h_critic_out, k.Hout, (size_t)S * critic->hidden * 4,
                                      cudaMemcpyDeviceToDevice, s);
      if (e != cudaSuccess) return (int)e;
    }
  }
  return launch_status();
}

int64_t mava_rec_ppo_workspace_bytes(const mava_rnn_desc* actor, const mava_rnn_desc* critic,
                                     int seq_envs_total, int chunk) {
  if (!actor || !critic || seq_envs_total <= 0 || chunk <= 0) return -1;
  const int64_t Sa = (int64_t)seq_envs_total * actor->rows_per_env;
  const int64_t Sc = (int64_t)seq_envs_total * critic->rows_per_env;
  // both networks' work areas side by side: their scans run in one launch (gru_scan.cu)
  const int64_t f = work_floats(actor, Sa * chunk, Sa, true) +
                    work_floats(critic, Sc * chunk, Sc, true);
  return 1024 + align256((int64_t)seq_envs_total * chunk * 4) + f * 4 + 1024;
}

int mava_rec_ppo_loss_grad(const mava_rnn_desc* actor, const float* actor_params,
                           const mava_rnn_desc* critic, const float* critic_params,
                           const mava_ppo_hyper* hyper, const int8_t* view,
                           const float* obs_actor, const float* obs_critic, const void* mask,
                           const int8_t* action, const float* old_logp, const float* old_value,
                           const float* adv, const float* targets, const uint8_t* done_in,
                           const float* hs_actor, const float* hs_critic, const int32_t* cols,
                           int num_replicas, int envs_per_replica, int mb_cols, int chunk,
                           int num_chunks, float* grad_out, void* workspace,
                           mava_stream_t stream) {
  int rc = check_desc(actor);
  if (rc) return rc;
  rc = check_desc(critic);
  if (rc) return rc;
  MAVA_CHECK_PTR(hyper);
  MAVA_CHECK_PTR(actor_params);
  MAVA_CHECK_PTR(critic_params);
  MAVA_CHECK_PTR(mask);
  MAVA_CHECK_PTR(action);
  MAVA_CHECK_PTR(old_logp);
  MAVA_CHECK_PTR(old_value);
  MAVA_CHECK_PTR(adv);
  MAVA_CHECK_PTR(targets);
  MAVA_CHECK_PTR(done_in);
  MAVA_CHECK_PTR(hs_actor);
  MAVA_CHECK_PTR(hs_critic);
  MAVA_CHECK_PTR(cols);
  MAVA_CHECK_PTR(grad_out);
  MAVA_CHECK_PTR(workspace);
  MAVA_CHECK_ARG(num_replicas > 0 && num_replicas <= 8 && envs_per_replica > 0 && mb_cols > 0);
  MAVA_CHECK_ARG(chunk > 0 && num_chunks > 0 && critic->out_dim == 1);
  MAVA_CHECK_ARG(actor->rows_per_env == actor->num_agents);
  MAVA_CHECK_ARG(actor->input_mode == MAVA_IN_DENSE ? obs_actor != nullptr : view != nullptr);
  MAVA_CHECK_ARG(critic->input_mode == MAVA_IN_DENSE ? obs_critic != nullptr : view != nullptr);
  cudaStream_t s = as_stream(stream);
  const int A = actor->num_agents, L = chunk, U = num_replicas;
  const int NE = U * envs_per_replica;
  const int Senv = U * mb_cols;
  const int64_t na = mava_rnn_param_count(actor), nc = mava_rnn_param_count(critic);

  unsigned char* wb = static_cast<unsigned char*>(workspace);
  double* stats = reinterpret_cast<double*>(wb);           // [U][2]
  double* loss_acc = reinterpret_cast<double*>(wb + 512);  // [5]
  int32_t* steps = reinterpret_cast<int32_t*>(wb + 1024);
  float* wf = reinterpret_cast<float*>(wb + 1024 + align256((int64_t)Senv * L * 4));

  cudaError_t e = cudaMemsetAsync(workspace, 0, 1024, s);
  if (e != cudaSuccess) return (int)e;
  e = cudaMemsetAsync(grad_out, 0, (size_t)(na + nc + 8) * sizeof(float), s);
  if (e != cudaSuccess) return (int)e;
  rec_steps_kernel<<<ceil_div(L * Senv, 256), 256, 0, s>>>(cols, mb_cols, U, envs_per_replica,
                                                           num_chunks, NE, L, steps);
  {
    dim3 grid((unsigned)min((int64_t)sm_count() * 2, ceil_div64((int64_t)L * mb_cols * A, 256)),
              (unsigned)U);
    rec_adv_stats_kernel<<<grid, 256, 0, s>>>(adv, steps, L, Senv, mb_cols, A, stats);
  }
  const double cnt = (double)L * mb_cols * A;  // elements per replica the means run over
  // Both networks advance phase by phase, so that their scans -- 128 CTAs of actor sequences and 16
  // of critic sequences at the SMAX shapes, each bound by the latency of a time step -- share one
  // launch per direction.
  const Net an = make_net(actor, actor_params), ag = make_net(actor, grad_out);
  const Net cn = make_net(critic, critic_params), cg = make_net(critic, grad_out + na);
  const int64_t Sa = (int64_t)Senv * actor->rows_per_env, Ra = Sa * L;
  const int64_t Sc = (int64_t)Senv * critic->rows_per_env, Rc = Sc * L;
  const Work ka = carve(wf, actor, Ra, Sa, true);
  const Work kc = carve(wf + work_floats(actor, Ra, Sa, true), critic, Rc, Sc, true);
  const SeqInput ia{view, obs_actor, steps, done_in, hs_actor, 1, L, Senv};
  const SeqInput ic{view, obs_critic, steps, done_in, hs_critic, 1, L, Senv};
  const bool scan = scan_capable(an, ia, ka) && scan_capable(cn, ic, kc);
  const GruScanNet sn[2] = {scan_net(an, ia, ka), scan_net(cn, ic, kc)};
  // ---- forward
  rc = forward_pre(an, ia, ka, s);
  if (rc) return rc;
  rc = forward_pre(cn, ic, kc, s);
  if (rc) return rc;
  if (scan) {
    rc = launch_gru_scan_fwd(sn, 2, s);
  } else {
    rc = forward_scan_steps(an, ia, ka, s);
    if (rc) return rc;
    rc = forward_scan_steps(cn, ic, kc, s);
  }
  if (rc) return rc;
  rc = forward_post(an, ia, ka, s);
  if (rc) return rc;
  rc = forward_post(cn, ic, kc, s);
  if (rc) return rc;
  // ---- losses: d(loss)/d(head output) in place over the head outputs
  {
    HeadArgs h{};
    h.logits = ka.out; h.rows = Ra; h.rpe = actor->rows_per_env; h.A = A; h.N = actor->out_dim;
    h.steps = steps; h.mask = mask; h.action_old = action; h.old_logp = old_logp; h.adv = adv;
    h.adv_stats = stats; h.rows_per_replica = (int64_t)mb_cols * A; h.rows_per_pos = Sa;
    h.count_per_replica = cnt; h.num_replicas = U; h.clip_eps = hyper->clip_eps;
    h.ent_coef = hyper->ent_coef; h.vf_coef = hyper->vf_coef; h.dOut = ka.out;
    h.loss_acc = loss_acc;
    rec_actor_loss_kernel<<<(unsigned)ceil_div64(Ra, 256), 256, 0, s>>>(h);
  }
  {
    HeadArgs h{};
    h.logits = kc.out; h.rows = Rc; h.rpe = critic->rows_per_env; h.A = A; h.steps = steps;
    h.old_value = old_value; h.targets = targets;
    h.rows_per_replica = (int64_t)mb_cols * critic->rows_per_env; h.rows_per_pos = Sc;
    h.count_per_replica = cnt; h.num_replicas = U; h.clip_eps = hyper->clip_eps;
    h.ent_coef = hyper->ent_coef; h.vf_coef = hyper->vf_coef; h.dOut = kc.out;
    h.loss_acc = loss_acc;
    rec_critic_loss_kernel<<<(unsigned)ceil_div64(Rc, 256), 256, 0, s>>>(h);
  }
  // ---- backward
  rc = backward_pre(an, ag, ia, ka, s);
  if (rc) return rc;
  rc = backward_pre(cn, cg, ic, kc, s);
  if (rc) return rc;
  if (scan) {
    rc = launch_gru_scan_bwd(sn, 2, s);
  } else {
    rc = backward_scan_steps(an, ia, ka, s);
    if (rc) return rc;
    rc = backward_scan_steps(cn, ic, kc, s);
  }
  if (rc) return rc;
  rc = backward_post(an, ag, ia, ka, s);
  if (rc) return rc;
  rc = backward_post(cn, cg, ic, kc, s);
  if (rc) return rc;
  return launch_finalize_loss(loss_acc, cnt * U, hyper->ent_coef, hyper->vf_coef,
                              grad_out + na + nc, s);
}

}  // extern "C"
