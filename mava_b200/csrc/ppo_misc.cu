// GAE reverse scan, minibatch row lists and the fused clip-by-global-norm + Adam step.
#include <cstdlib>

#include "common.cuh"
#include "mlp_tc.cuh"  // layout of the packed bf16 weight images (clip_adam_pair_pack)

using namespace mava;

namespace {

// One thread per env-agent, walking time backwards (ff_mappo.py:117-137).  Loads of a block of
// kU steps are issued together so the HBM latency of the sequential scan overlaps.
template <bool REC>
__global__ void __launch_bounds__(256)
gae_kernel(const float* __restrict__ reward, const float* __restrict__ value,
           const uint8_t* __restrict__ done, const float* __restrict__ last_val,
           const uint8_t* __restrict__ last_done, float gamma, float gamma_lambda, int T,
           int num_envs, int A, float* __restrict__ adv, float* __restrict__ targets) {
  constexpr int kU = 8;
  const int64_t n = (int64_t)num_envs * A;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int e = (int)(i / A);
  float gae = 0.0f;
  float next_value = last_val[i];
  float next_nd = REC ? (last_done[e] ? 0.0f : 1.0f) : 1.0f;
  for (int t1 = T; t1 > 0; t1 -= kU) {
    float r[kU], v[kU], nd[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int t = t1 - 1 - u;
      if (t >= 0) {
        r[u] = __ldg(reward + (int64_t)t * n + i);
        v[u] = __ldg(value + (int64_t)t * n + i);
        nd[u] = __ldg(done + (int64_t)t * num_envs + e) ? 0.0f : 1.0f;
      }
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int t = t1 - 1 - u;
      if (t >= 0) {
        const float m = REC ? next_nd : nd[u];
        const float delta = r[u] + gamma * next_value * m - v[u];
        gae = delta + gamma_lambda * m * gae;
        adv[(int64_t)t * n + i] = gae;
        targets[(int64_t)t * n + i] = gae + v[u];
        next_value = v[u];
        if (REC) next_nd = nd[u];
      }
    }
  }
}

// The same scan for small problems (config 2: 8192 env-agents x 128 steps = 18 MB, where one thread per
// env-agent leaves most SMs without a warp and every block of kU steps pays a DRAM round trip): a CTA
// owns 32 env-agents; its four warps bring kTS time steps of reward / value / done into shared
// memory with every load in flight at once, warp 0 walks them backwards (same arithmetic, same
// order: bit-identical to gae_kernel), and the four warps write adv / targets back.
constexpr int kGaeCols = 32, kGaeTS = 128, kGaeThreads = 128;
template <bool REC>
__global__ void __launch_bounds__(kGaeThreads)
gae_tile_kernel(const float* __restrict__ reward, const float* __restrict__ value,
                const uint8_t* __restrict__ done, const float* __restrict__ last_val,
                const uint8_t* __restrict__ last_done, float gamma, float gamma_lambda, int T,
                int num_envs, int A, float* __restrict__ adv, float* __restrict__ targets) {
  __shared__ float s_r[kGaeTS][kGaeCols], s_v[kGaeTS][kGaeCols];  // in: reward, value; out: adv, targets
  __shared__ float s_nd[kGaeTS][kGaeCols];
  const int64_t n = (int64_t)num_envs * A;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t i = (int64_t)blockIdx.x * kGaeCols + lane;
  const bool ok = i < n;
  const int e = ok ? (int)(i / A) : 0;
  float gae = 0.0f, next_value = 0.0f, next_nd = 1.0f;
  if (warp == 0 && ok) {
    next_value = last_val[i];
    if (REC) next_nd = last_done[e] ? 0.0f : 1.0f;
  }
  for (int t1 = T; t1 > 0; t1 -= kGaeTS) {
    const int t0 = t1 > kGaeTS ? t1 - kGaeTS : 0, nt = t1 - t0;
    for (int k = warp; k < nt; k += kGaeThreads / 32) {
      const int t = t0 + k;
      s_r[k][lane] = ok ? __ldg(reward + (int64_t)t * n + i) : 0.0f;
      s_v[k][lane] = ok ? __ldg(value + (int64_t)t * n + i) : 0.0f;
      s_nd[k][lane] = (ok && __ldg(done + (int64_t)t * num_envs + e)) ? 0.0f : 1.0f;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll 4
      for (int k = nt - 1; k >= 0; --k) {
        const float r = s_r[k][lane], v = s_v[k][lane], nd = s_nd[k][lane];
        const float m = REC ? next_nd : nd;
        const float delta = r + gamma * next_value * m - v;
        gae = delta + gamma_lambda * m * gae;
        s_r[k][lane] = gae;
        s_v[k][lane] = gae + v;
        next_value = v;
        if (REC) next_nd = nd;
      }
    }
    __syncthreads();
    if (ok) {
      for (int k = warp; k < nt; k += kGaeThreads / 32) {
        const int t = t0 + k;
        adv[(int64_t)t * n + i] = s_r[k][lane];
        targets[(int64_t)t * n + i] = s_v[k][lane];
      }
    }
    __syncthreads();
  }
}

__global__ void minibatch_rows_kernel(const int32_t* __restrict__ perm, int mb_index, int mb_size,
                                      int num_replicas, int E, int32_t* __restrict__ rows) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= num_replicas * mb_size) return;
  const int u = i / mb_size, j = i - u * mb_size;
  const int p = perm[(int64_t)mb_index * mb_size + j];
  const int t = p / E, e = p - t * E;
  rows[i] = t * (num_replicas * E) + u * E + e;
}

// optax.clip_by_global_norm -> optax.adam(eps=1e-5) -> apply_updates, one CTA (n is ~5e4).
__global__ void __launch_bounds__(1024)
clip_adam_kernel(float* __restrict__ params, float* __restrict__ mu, float* __restrict__ nu,
                 int32_t* __restrict__ count, const float* __restrict__ grad, int64_t n,
                 float grad_scale, float lr, float max_norm, int lr_decay_num_updates,
                 int steps_per_update) {
  __shared__ double red[32];
  __shared__ float s_norm;
  double ss = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const float g = grad[i] * grad_scale;
    ss += (double)g * (double)g;
  }
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) s_norm = (float)sqrt(v);
  }
  __syncthreads();
  const float g_norm = s_norm;
  const bool keep = g_norm < max_norm;
  const int c0 = *count;
  const int c = c0 + 1;
  const float b1 = 0.9f, b2 = 0.999f, eps = 1e-5f;
  const float bc1 = 1.0f - powf(b1, (float)c), bc2 = 1.0f - powf(b2, (float)c);
  float step_lr = lr;
  if (lr_decay_num_updates > 0)
    step_lr = lr * (1.0f - (float)(c0 / steps_per_update) / (float)lr_decay_num_updates);
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    float g = grad[i] * grad_scale;
    if (!keep) g = (g / g_norm) * max_norm;
    const float m = (1.0f - b1) * g + b1 * mu[i];
    const float v = (1.0f - b2) * g * g + b2 * nu[i];
    mu[i] = m;
    nu[i] = v;
    const float upd = (m / bc1) / (sqrtf(v / bc2) + eps);
    params[i] += -step_lr * upd;
  }
  __syncthreads();
  if (threadIdx.x == 0) *count = c;
}

// clip_by_global_norm -> adam -> apply_updates for both networks: a first launch accumulates the
// squared global norms (one slice per CTA, fp32 partial sums -- B200's fp64 pipe is far too slow for
// an element-wise pass -- combined in fp64), a second launch applies the update slice by slice.
// The last CTA of the second launch to have read norm and step count (ticket) resets the norm
// accumulator and advances the count, so late CTAs never see the new values.
__device__ unsigned int g_adam_ticket[2];
__device__ double g_adam_norm2[2];

struct AdamPairArgs {
  float *params, *mu, *nu;
  int32_t* counts;
  const float* grad;
  int64_t n[2];
  float lr[2];
  float grad_scale, max_norm;
  int lr_decay_num_updates, steps_per_update;
  // optional: the bf16 operand images of the tensor-core kernels (mlp_tc.cuh), refreshed by the
  // thread that updates a parameter instead of by separate packing launches
  unsigned char* image[2];
  int in_dim[2], k1p[2], out[2];
};

__global__ void __launch_bounds__(256) grad_sqnorm_kernel(const AdamPairArgs a) {
  const int net = blockIdx.y;
  const float* grad = a.grad + (net == 0 ? 0 : a.n[0]);
  const int64_t n = a.n[net];
  float p4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n; i += 4 * stride) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float g = __ldg(grad + i + u * stride) * a.grad_scale;
      p4[u] = fmaf(g, g, p4[u]);
    }
  }
  for (; i < n; i += stride) {
    const float g = __ldg(grad + i) * a.grad_scale;
    p4[0] = fmaf(g, g, p4[0]);
  }
  double ss = ((double)p4[0] + (double)p4[1]) + ((double)p4[2] + (double)p4[3]);
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  __shared__ double red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  if (threadIdx.x == 0) {
    double v = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += red[w];
    atomicAdd(&g_adam_norm2[net], v);
  }
}

__global__ void __launch_bounds__(256) clip_adam_pair_kernel(const AdamPairArgs a) {
  const int net = blockIdx.y;
  const int64_t off = net == 0 ? 0 : a.n[0];
  const int64_t n = a.n[net];
  const float* grad = a.grad + off;
  float* params = a.params + off;
  float* mu = a.mu + off;
  float* nu = a.nu + off;
  __shared__ float s_norm;
  __shared__ int s_count;
  if (threadIdx.x == 0) {
    s_count = a.counts[net];
    s_norm = (float)sqrt(*reinterpret_cast<volatile double*>(&g_adam_norm2[net]));
    __threadfence();
    const unsigned int ticket = atomicAdd(&g_adam_ticket[net], 1u);
    if (ticket == gridDim.x - 1) {  // every CTA of this network has read norm and count
      g_adam_ticket[net] = 0u;
      g_adam_norm2[net] = 0.0;
      a.counts[net] = s_count + 1;
    }
  }
  __syncthreads();
  const float g_norm = s_norm;
  const bool keep = g_norm < a.max_norm;
  const int c0 = s_count, c = c0 + 1;
  const float b1 = 0.9f, b2 = 0.999f, eps = 1e-5f;
  const float bc1 = 1.0f - powf(b1, (float)c), bc2 = 1.0f - powf(b2, (float)c);
  float step_lr = a.lr[net];
  if (a.lr_decay_num_updates > 0)
    step_lr *= 1.0f - (float)(c0 / a.steps_per_update) / (float)a.lr_decay_num_updates;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    float g = grad[i] * a.grad_scale;
    if (!keep) g = (g / g_norm) * a.max_norm;
    const float m = (1.0f - b1) * g + b1 * mu[i];
    const float v = (1.0f - b2) * g * g + b2 * nu[i];
    mu[i] = m;
    nu[i] = v;
    const float pnew = params[i] - step_lr * ((m / bc1) / (sqrtf(v / bc2) + eps));
    params[i] = pnew;
    if (a.image[net])
      *reinterpret_cast<__nv_bfloat16*>(a.image[net] +
                                        tcmlp::image_offset((int)i, a.in_dim[net], a.k1p[net], a.out[net])) =
          __float2bfloat16_rn(pnew);
  }
}


// get_final_step_metrics (mava/wrappers/episode_metrics.py:114-132) + describe()
// (mava/utils/logger.py:44-58) on the device: count, sum, sum of squares, min and max of the return
// and the length of the episodes that END in this block of env-steps, accumulated into
// stats[10] = {count, sum_r, sumsq_r, min_r, max_r, sum_l, sumsq_l, min_l, max_l, unused}.
// The run loop then moves 80 bytes to the host instead of three [T][NE] arrays.
__device__ __forceinline__ void atomic_min_f64(double* addr, double v) {
  unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
  unsigned long long old = *a;
  while (__longlong_as_double((long long)old) > v) {
    const unsigned long long seen = atomicCAS(a, old, (unsigned long long)__double_as_longlong(v));
    if (seen == old) break;
    old = seen;
  }
}
__device__ __forceinline__ void atomic_max_f64(double* addr, double v) {
  unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
  unsigned long long old = *a;
  while (__longlong_as_double((long long)old) < v) {
    const unsigned long long seen = atomicCAS(a, old, (unsigned long long)__double_as_longlong(v));
    if (seen == old) break;
    old = seen;
  }
}

__global__ void __launch_bounds__(256)
episode_stats_kernel(const uint8_t* __restrict__ done, const float* __restrict__ ep_return,
                     const int32_t* __restrict__ ep_length, int64_t n, double* __restrict__ stats) {
  // per-thread partial sums in fp32 would lose the small returns of long runs: the terminal steps
  // are few (one per episode), so fp64 only runs on them
  double cnt = 0.0, sr = 0.0, qr = 0.0, sl = 0.0, ql = 0.0;
  float mnr = INFINITY, mxr = -INFINITY, mnl = INFINITY, mxl = -INFINITY;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    if (__ldg(done + i)) {
      const float r = __ldg(ep_return + i), l = (float)__ldg(ep_length + i);
      cnt += 1.0;
      sr += (double)r;
      qr += (double)r * (double)r;
      sl += (double)l;
      ql += (double)l * (double)l;
      mnr = fminf(mnr, r);
      mxr = fmaxf(mxr, r);
      mnl = fminf(mnl, l);
      mxl = fmaxf(mxl, l);
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    sr += __shfl_xor_sync(0xffffffffu, sr, o);
    qr += __shfl_xor_sync(0xffffffffu, qr, o);
    sl += __shfl_xor_sync(0xffffffffu, sl, o);
    ql += __shfl_xor_sync(0xffffffffu, ql, o);
    mnr = fminf(mnr, __shfl_xor_sync(0xffffffffu, mnr, o));
    mxr = fmaxf(mxr, __shfl_xor_sync(0xffffffffu, mxr, o));
    mnl = fminf(mnl, __shfl_xor_sync(0xffffffffu, mnl, o));
    mxl = fmaxf(mxl, __shfl_xor_sync(0xffffffffu, mxl, o));
  }
  if ((threadIdx.x & 31) == 0 && cnt > 0.0) {
    atomicAdd(stats + 0, cnt);
    atomicAdd(stats + 1, sr);
    atomicAdd(stats + 2, qr);
    atomic_min_f64(stats + 3, (double)mnr);
    atomic_max_f64(stats + 4, (double)mxr);
    atomicAdd(stats + 5, sl);
    atomicAdd(stats + 6, ql);
    atomic_min_f64(stats + 7, (double)mnl);
    atomic_max_f64(stats + 8, (double)mxl);
  }
}

// evaluator.py:143-150: metrics at the first terminal step of every env (argmax of the done flag)
__global__ void first_terminal_kernel(const uint8_t* __restrict__ done,
                                      const float* __restrict__ ep_return,
                                      const int32_t* __restrict__ ep_length, int T, int num_envs,
                                      float* __restrict__ out_return, int32_t* __restrict__ out_length) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= num_envs) return;
  int first = 0;
  for (int t = 0; t < T; ++t) {
    if (__ldg(done + (size_t)t * num_envs + e)) {
      first = t;
      break;
    }
  }
  out_return[e] = ep_return[(size_t)first * num_envs + e];
  out_length[e] = ep_length[(size_t)first * num_envs + e];
}

__global__ void episode_stats_init_kernel(double* stats) {
  const int i = threadIdx.x;
  if (i < 10) stats[i] = (i == 3 || i == 7) ? INFINITY : (i == 4 || i == 8) ? -INFINITY : 0.0;
}

}  // namespace

extern "C" {

int mava_clip_adam_pair(float* params, float* mu, float* nu, int32_t* counts, const float* grad,
                        int64_t n_actor, int64_t n_critic, float grad_scale, float lr_actor,
                        float lr_critic, float max_norm, int lr_decay_num_updates,
                        int steps_per_update, mava_stream_t s) {
  MAVA_CHECK_PTR(params);
  MAVA_CHECK_PTR(mu);
  MAVA_CHECK_PTR(nu);
  MAVA_CHECK_PTR(counts);
  MAVA_CHECK_PTR(grad);
  MAVA_CHECK_ARG(n_actor > 0 && n_critic > 0 && steps_per_update > 0);
  AdamPairArgs a;
  a.params = params; a.mu = mu; a.nu = nu; a.counts = counts; a.grad = grad;
  a.n[0] = n_actor; a.n[1] = n_critic;
  a.lr[0] = lr_actor; a.lr[1] = lr_critic;
  a.grad_scale = grad_scale; a.max_norm = max_norm;
  a.lr_decay_num_updates = lr_decay_num_updates; a.steps_per_update = steps_per_update;
  a.image[0] = a.image[1] = nullptr;
  const unsigned ctas = (unsigned)max((int64_t)1, min((int64_t)64, ceil_div64(max(n_actor, n_critic), 1024)));
  grad_sqnorm_kernel<<<dim3(ctas, 2), 256, 0, as_stream(s)>>>(a);
  clip_adam_pair_kernel<<<dim3(ctas, 2), 256, 0, as_stream(s)>>>(a);
  return launch_status();
}

int mava_clip_adam_pair_pack(float* params, float* mu, float* nu, int32_t* counts, const float* grad,
                             const mava_mlp_desc* actor, void* actor_image,
                             const mava_mlp_desc* critic, void* critic_image, float grad_scale,
                             float lr_actor, float lr_critic, float max_norm,
                             int lr_decay_num_updates, int steps_per_update, mava_stream_t s) {
  MAVA_CHECK_PTR(params);
  MAVA_CHECK_PTR(mu);
  MAVA_CHECK_PTR(nu);
  MAVA_CHECK_PTR(counts);
  MAVA_CHECK_PTR(grad);
  MAVA_CHECK_PTR(actor);
  MAVA_CHECK_PTR(critic);
  MAVA_CHECK_PTR(actor_image);
  MAVA_CHECK_PTR(critic_image);
  MAVA_CHECK_ARG(steps_per_update > 0);
  const mava_mlp_desc* nets[2] = {actor, critic};
  AdamPairArgs a;
  a.params = params; a.mu = mu; a.nu = nu; a.counts = counts; a.grad = grad;
  a.lr[0] = lr_actor; a.lr[1] = lr_critic;
  a.grad_scale = grad_scale; a.max_norm = max_norm;
  a.lr_decay_num_updates = lr_decay_num_updates; a.steps_per_update = steps_per_update;
  a.image[0] = static_cast<unsigned char*>(actor_image);
  a.image[1] = static_cast<unsigned char*>(critic_image);
  for (int k = 0; k < 2; ++k) {
    const mava_mlp_desc* d = nets[k];
    MAVA_CHECK_ARG(d->h1 == tcmlp::HID && d->h2 == tcmlp::HID && d->out_dim <= tcmlp::NHEAD);
    a.n[k] = mava_mlp_param_count(d);
    a.in_dim[k] = d->in_dim;
    a.k1p[k] = tcmlp::pad16(d->in_dim + 1);
    a.out[k] = d->out_dim;
  }
  const unsigned ctas = (unsigned)max((int64_t)1, min((int64_t)64, ceil_div64(max(a.n[0], a.n[1]), 1024)));
  grad_sqnorm_kernel<<<dim3(ctas, 2), 256, 0, as_stream(s)>>>(a);
  clip_adam_pair_kernel<<<dim3(ctas, 2), 256, 0, as_stream(s)>>>(a);
  return launch_status();
}

int mava_gae(const float* reward, const float* value, const uint8_t* done, const float* last_val,
             const uint8_t* last_done, float gamma, float gae_lambda, int T, int num_envs,
             int num_agents, int rec, float* adv, float* targets, mava_stream_t s) {
  MAVA_CHECK_PTR(reward);
  MAVA_CHECK_PTR(value);
  MAVA_CHECK_PTR(done);
  MAVA_CHECK_PTR(last_val);
  MAVA_CHECK_PTR(adv);
  MAVA_CHECK_PTR(targets);
  MAVA_CHECK_ARG(T > 0 && num_envs > 0 && num_agents > 0);
  if (rec) MAVA_CHECK_PTR(last_done);
  const int64_t n = (int64_t)num_envs * num_agents;
  const int blocks = (int)ceil_div64(n, 256);
  // gamma * gae_lambda is folded in double like the reference's Python floats (ff_mappo.py:127)
  const float gl = (float)((double)gamma * (double)gae_lambda);
  const char* no_tile = getenv("MAVA_GAE_NO_TILE");  // development switch (sweep_gae.py compares)
  if (n <= (int64_t)1 << 16 && !(no_tile && no_tile[0] == '1')) {
    // small: spread over the SMs, stage through shared memory
    const int tb = (int)ceil_div64(n, kGaeCols);
    if (rec)
      gae_tile_kernel<true><<<tb, kGaeThreads, 0, as_stream(s)>>>(
          reward, value, done, last_val, last_done, gamma, gl, T, num_envs, num_agents, adv, targets);
    else
      gae_tile_kernel<false><<<tb, kGaeThreads, 0, as_stream(s)>>>(
          reward, value, done, last_val, last_done, gamma, gl, T, num_envs, num_agents, adv, targets);
    return launch_status();
  }
  if (rec)
    gae_kernel<true><<<blocks, 256, 0, as_stream(s)>>>(reward, value, done, last_val, last_done,
                                                       gamma, gl, T, num_envs, num_agents, adv,
                                                       targets);
  else
    gae_kernel<false><<<blocks, 256, 0, as_stream(s)>>>(reward, value, done, last_val, last_done,
                                                        gamma, gl, T, num_envs, num_agents, adv,
                                                        targets);
  return launch_status();
}

int mava_ppo_minibatch_rows(const int32_t* perm, int mb_index, int mb_size, int num_replicas,
                            int envs_per_replica, int32_t* rows, mava_stream_t s) {
  MAVA_CHECK_PTR(perm);
  MAVA_CHECK_PTR(rows);
  MAVA_CHECK_ARG(mb_index >= 0 && mb_size > 0 && num_replicas > 0 && envs_per_replica > 0);
  const int n = num_replicas * mb_size;
  minibatch_rows_kernel<<<ceil_div(n, 256), 256, 0, as_stream(s)>>>(
      perm, mb_index, mb_size, num_replicas, envs_per_replica, rows);
  return launch_status();
}

int mava_clip_adam(float* params, float* mu, float* nu, int32_t* count, const float* grad,
                   int64_t n, float grad_scale, float lr, float max_norm, int lr_decay_num_updates,
                   int steps_per_update, mava_stream_t s) {
  MAVA_CHECK_PTR(params);
  MAVA_CHECK_PTR(mu);
  MAVA_CHECK_PTR(nu);
  MAVA_CHECK_PTR(count);
  MAVA_CHECK_PTR(grad);
  MAVA_CHECK_ARG(n > 0 && steps_per_update > 0);
  clip_adam_kernel<<<1, 1024, 0, as_stream(s)>>>(params, mu, nu, count, grad, n, grad_scale, lr,
                                                 max_norm, lr_decay_num_updates, steps_per_update);
  return launch_status();
}

int mava_episode_first_terminal(const uint8_t* done, const float* ep_return,
                                const int32_t* ep_length, int T, int num_envs, float* out_return,
                                int32_t* out_length, mava_stream_t s) {
  MAVA_CHECK_PTR(done);
  MAVA_CHECK_PTR(ep_return);
  MAVA_CHECK_PTR(ep_length);
  MAVA_CHECK_PTR(out_return);
  MAVA_CHECK_PTR(out_length);
  MAVA_CHECK_ARG(T > 0 && num_envs > 0);
  first_terminal_kernel<<<ceil_div(num_envs, 128), 128, 0, as_stream(s)>>>(
      done, ep_return, ep_length, T, num_envs, out_return, out_length);
  return launch_status();
}

int mava_episode_stats(const uint8_t* done, const float* ep_return, const int32_t* ep_length,
                       int64_t n, int reset, double* stats, mava_stream_t s) {
  MAVA_CHECK_PTR(stats);
  MAVA_CHECK_ARG(n >= 0);
  if (reset) episode_stats_init_kernel<<<1, 32, 0, as_stream(s)>>>(stats);
  if (n > 0) {
    MAVA_CHECK_PTR(done);
    MAVA_CHECK_PTR(ep_return);
    MAVA_CHECK_PTR(ep_length);
    const int blocks = (int)min((int64_t)4 * sm_count(), ceil_div64(n, 256));
    episode_stats_kernel<<<blocks, 256, 0, as_stream(s)>>>(done, ep_return, ep_length, n, stats);
  }
  return launch_status();
}

}  // extern "C"
