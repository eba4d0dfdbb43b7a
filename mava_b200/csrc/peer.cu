// pmean("device") fused with the optimiser: ONE launch per minibatch does
//   all-reduce(sum) of [actor grads | critic grads | 5 loss scalars] over the ranks' peer-mapped
//   gradient buffers  ->  optax.clip_by_global_norm  ->  optax.adam(eps=1e-5)  ->  apply_updates
//   ->  refresh of the packed bf16 operand images  ->  loss metrics of the minibatch
// (ff_mappo.py:224-250,359-366; rec_mappo.py:283-305).  No NCCL call sits between the loss kernel
// and the next minibatch any more.
//
// Every rank owns one exchange buffer (cudaMalloc + cudaIpc, mapped by all ranks of the node over
// NVLink / NVSwitch): its gradient vector followed by a flag block.  One-shot all-reduce: after a
// flag handshake ("my gradients are complete") every rank READS the vectors of all ranks and adds
// them in rank order 0..W-1, so all ranks hold bit-identical sums (the replicated parameters stay
// replicated), 7 x 307 KB per rank and minibatch for the headline networks.  A second handshake at
// the end of the kernel ("I have read yours") lets a rank overwrite its buffer as soon as the
// kernel is done: one buffer, no parity games with the host-side schedule.
//
// The kernel is one thread-block cluster (8 CTAs x 1024 threads): the squared global norms are
// combined through distributed shared memory and a cluster barrier -- the hardware co-schedules a
// cluster, so no grid-wide barrier or device-global scratch is needed and the kernel is re-entrant
// across learners and streams.  Spin loops give up after 2 s and raise the buffer's error word
// instead of hanging the device (a rank that died, a mismatched call sequence).
#include <cooperative_groups.h>

#include <cstring>

#include "common.cuh"
#include "mlp_tc.cuh"

namespace cg = cooperative_groups;
using namespace mava;

namespace {

constexpr int kMaxRanks = MAVA_PEER_MAX_RANKS;
constexpr int kCluster = 8;
constexpr int kThreads = 1024;
// flag block (uint32 words) behind the gradient vector of an exchange buffer
constexpr int F_READY = 0;            // [kMaxRanks] written by rank p: "p's gradients of call #seq are complete"
constexpr int F_DONE = kMaxRanks;     // [kMaxRanks] written by rank p: "p has read this buffer in call #seq"
constexpr int F_SEQ = 2 * kMaxRanks;  // calls completed on this buffer (local)
constexpr int F_ERR = 2 * kMaxRanks + 1;  // != 0: a handshake timed out (local)
constexpr int F_WORDS = 32;

__host__ __device__ inline int64_t grad_bytes_padded(int64_t n_grad) {
  return (n_grad * 4 + 255) / 256 * 256;
}

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_sys_f4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p)
               : "memory");
  return v;
}
__device__ __forceinline__ float ld_sys_f(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// spin until *flag has reached seq (wrap-safe); false on timeout
__device__ __forceinline__ bool wait_flag(const uint32_t* flag, uint32_t seq) {
  const unsigned long long t0 = global_ns();
  while ((int32_t)(ld_acquire_sys(flag) - seq) < 0) {
    if (global_ns() - t0 > 2000000000ull) return false;
    __nanosleep(64);
  }
  return true;
}

struct ReduceAdamArgs {
  float *params, *mu, *nu;
  int32_t* counts;
  const float* grad[kMaxRanks];  // this rank's mapping of every rank's gradient vector
  uint32_t* flags[kMaxRanks];    // ... and of every rank's flag block
  int rank, world;
  float* gsum;  // local scratch, n[0] + n[1] floats: the reduced, scaled gradients
  int64_t n[2];
  float lr[2];
  float grad_scale, max_norm;
  int lr_decay_num_updates, steps_per_update;
  unsigned char* image[2];
  int in_dim[2], k1p[2], out[2];
  float* loss_out;  // [5] or null: mean over ranks of the loss scalars behind the gradients
};

__device__ __forceinline__ void adam_one(const ReduceAdamArgs& a, int64_t i, float g, float g_norm,
                                         bool keep, int net, int64_t off, float step_lr, float bc1,
                                         float bc2) {
  const float b1 = 0.9f, b2 = 0.999f, eps = 1e-5f;
  if (!keep) g = (g / g_norm) * a.max_norm;
  const float m = (1.0f - b1) * g + b1 * a.mu[i];
  const float v = (1.0f - b2) * g * g + b2 * a.nu[i];
  a.mu[i] = m;
  a.nu[i] = v;
  const float pnew = a.params[i] - step_lr * ((m / bc1) / (sqrtf(v / bc2) + eps));
  a.params[i] = pnew;
  if (a.image[net])
    *reinterpret_cast<__nv_bfloat16*>(
        a.image[net] + tcmlp::image_offset(i - off, a.in_dim[net], a.k1p[net], a.out[net])) =
        __float2bfloat16_rn(pnew);
}

__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(kThreads)
reduce_clip_adam_kernel(const ReduceAdamArgs a) {
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned crank = cluster.block_rank();
  const int t = threadIdx.x;
  const int64_t tid = (int64_t)crank * kThreads + t;
  constexpr int64_t nthreads = (int64_t)kCluster * kThreads;
  uint32_t* my_flags = a.flags[a.rank];
  __shared__ double s_part[kCluster][2];  // CTA 0's copy collects every CTA's partial norms
  __shared__ double s_red[kThreads / 32][2];
  __shared__ float s_norm[2];

  const int c0[2] = {a.counts[0], a.counts[1]};
  const uint32_t seq = (a.world > 1 ? my_flags[F_SEQ] : 0u) + 1u;

  // ---- handshake 1: every rank's gradients of this call are complete ---------------------------
  if (a.world > 1) {
    if (crank == 0 && t < a.world && t != a.rank) {
      __threadfence_system();
      st_release_sys(a.flags[t] + F_READY + a.rank, seq);
    }
    if (t < a.world && t != a.rank) {
      if (!wait_flag(my_flags + F_READY + t, seq)) my_flags[F_ERR] = 1u;
    }
    __syncthreads();
  }

  // ---- sum over ranks (rank order), scale, squared norms per network ----------------------------
  const int64_t n0 = a.n[0], n01 = a.n[0] + a.n[1], n_all = n01 + 8;
  const int64_t chunks = n_all >> 2;  // the buffers are padded: n_all floats are always readable
  float ss[2] = {0.0f, 0.0f};
  for (int64_t c = tid; c < chunks; c += nthreads) {
    float4 s;
    if (a.world > 1) {
      // all ranks' chunks in flight at once (an NVLink round trip is ~2 us), added in rank order
      float4 v[kMaxRanks];
#pragma unroll
      for (int r = 0; r < kMaxRanks; ++r)
        if (r < a.world) v[r] = ld_sys_f4(a.grad[r] + 4 * c);
      s = v[0];
#pragma unroll
      for (int r = 1; r < kMaxRanks; ++r)
        if (r < a.world) { s.x += v[r].x; s.y += v[r].y; s.z += v[r].z; s.w += v[r].w; }
    } else {
      s = *reinterpret_cast<const float4*>(a.grad[0] + 4 * c);
    }
    float g[4] = {s.x * a.grad_scale, s.y * a.grad_scale, s.z * a.grad_scale, s.w * a.grad_scale};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t i = 4 * c + j;
      if (i < n01) {
        a.gsum[i] = g[j];
        const int net = i < n0 ? 0 : 1;
        ss[net] = fmaf(g[j], g[j], ss[net]);
      } else if (i < n01 + 5 && a.loss_out != nullptr) {
        a.loss_out[i - n01] = g[j];
      }
    }
  }
  for (int64_t i = 4 * chunks + tid; i < n_all; i += nthreads) {  // tail (n_all not a multiple of 4)
    float s = a.world > 1 ? ld_sys_f(a.grad[0] + i) : a.grad[0][i];
    for (int r = 1; r < a.world; ++r) s += ld_sys_f(a.grad[r] + i);
    const float g = s * a.grad_scale;
    if (i < n01) {
      a.gsum[i] = g;
      const int net = i < n0 ? 0 : 1;
      ss[net] = fmaf(g, g, ss[net]);
    } else if (i < n01 + 5 && a.loss_out != nullptr) {
      a.loss_out[i - n01] = g;
    }
  }
  double d0 = (double)ss[0], d1 = (double)ss[1];
  for (int o = 16; o > 0; o >>= 1) {
    d0 += __shfl_xor_sync(0xffffffffu, d0, o);
    d1 += __shfl_xor_sync(0xffffffffu, d1, o);
  }
  if ((t & 31) == 0) {
    s_red[t >> 5][0] = d0;
    s_red[t >> 5][1] = d1;
  }
  __syncthreads();
  if (t == 0) {
    double v0 = 0.0, v1 = 0.0;
    for (int w = 0; w < kThreads / 32; ++w) {
      v0 += s_red[w][0];
      v1 += s_red[w][1];
    }
    double* dst = cluster.map_shared_rank(&s_part[0][0], 0);  // CTA 0 collects
    dst[2 * crank + 0] = v0;
    dst[2 * crank + 1] = v1;
  }
  cluster.sync();  // every CTA has consumed its share of the peers' buffers and posted its norms

  // ---- handshake 2, first half: tell the peers their buffers have been read ---------------------
  if (a.world > 1 && crank == 0 && t < a.world && t != a.rank)
    st_release_sys(a.flags[t] + F_DONE + a.rank, seq);

  if (t == 0) {
    const double* src = cluster.map_shared_rank(&s_part[0][0], 0);
    double v0 = 0.0, v1 = 0.0;
    for (int c = 0; c < kCluster; ++c) {  // fixed order: every CTA computes the same norms
      v0 += src[2 * c + 0];
      v1 += src[2 * c + 1];
    }
    s_norm[0] = (float)sqrt(v0);
    s_norm[1] = (float)sqrt(v1);
  }
  __syncthreads();

  // ---- clip + Adam + apply (+ bf16 image refresh); every thread re-reads what it wrote ----------
  float g_norm[2], step_lr[2], bc1[2], bc2[2];
  bool keep[2];
#pragma unroll
  for (int net = 0; net < 2; ++net) {
    g_norm[net] = s_norm[net];
    keep[net] = g_norm[net] < a.max_norm;
    const int c = c0[net] + 1;
    bc1[net] = 1.0f - powf(0.9f, (float)c);
    bc2[net] = 1.0f - powf(0.999f, (float)c);
    step_lr[net] = a.lr[net];
    if (a.lr_decay_num_updates > 0)
      step_lr[net] *= 1.0f - (float)(c0[net] / a.steps_per_update) / (float)a.lr_decay_num_updates;
  }
  for (int64_t c = tid; c < chunks; c += nthreads) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t i = 4 * c + j;
      if (i < n01) {
        const int net = i < n0 ? 0 : 1;
        adam_one(a, i, a.gsum[i], g_norm[net], keep[net], net, net ? n0 : 0, step_lr[net], bc1[net],
                 bc2[net]);
      }
    }
  }
  for (int64_t i = 4 * chunks + tid; i < n01; i += nthreads) {
    const int net = i < n0 ? 0 : 1;
    adam_one(a, i, a.gsum[i], g_norm[net], keep[net], net, net ? n0 : 0, step_lr[net], bc1[net],
             bc2[net]);
  }

  // ---- handshake 2, second half: nobody reads this rank's buffer any more -----------------------
  if (a.world > 1 && crank == 0 && t < a.world && t != a.rank) {
    if (!wait_flag(my_flags + F_DONE + t, seq)) my_flags[F_ERR] = 1u;
  }
  cluster.sync();  // CTA 0's shared memory stays alive until every CTA has read the norms
  if (crank == 0 && t == 0) {
    a.counts[0] = c0[0] + 1;
    a.counts[1] = c0[1] + 1;
    if (a.world > 1) my_flags[F_SEQ] = seq;
  }
}

}  // namespace

extern "C" {

int64_t mava_peer_buffer_bytes(int64_t n_grad) {
  if (n_grad <= 0) return -1;
  return grad_bytes_padded(n_grad) + F_WORDS * 4;
}

int mava_peer_alloc(int64_t bytes, void** buf_out, void* ipc_handle64_host) {
  MAVA_CHECK_PTR(buf_out);
  MAVA_CHECK_ARG(bytes > 0);
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, (size_t)bytes);
  if (e != cudaSuccess) return (int)e;
  e = cudaMemset(p, 0, (size_t)bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e == cudaSuccess && ipc_handle64_host != nullptr)
    e = cudaIpcGetMemHandle(static_cast<cudaIpcMemHandle_t*>(ipc_handle64_host), p);
  if (e != cudaSuccess) {
    cudaFree(p);
    return (int)e;
  }
  *buf_out = p;
  return 0;
}

int mava_peer_open(const void* ipc_handle64_host, void** buf_out) {
  MAVA_CHECK_PTR(ipc_handle64_host);
  MAVA_CHECK_PTR(buf_out);
  cudaIpcMemHandle_t h;
  memcpy(&h, ipc_handle64_host, sizeof(h));
  cudaError_t e = cudaIpcOpenMemHandle(buf_out, h, cudaIpcMemLazyEnablePeerAccess);
  return e == cudaSuccess ? 0 : (int)e;
}

int mava_peer_close(void* buf) {
  MAVA_CHECK_PTR(buf);
  cudaError_t e = cudaIpcCloseMemHandle(buf);
  return e == cudaSuccess ? 0 : (int)e;
}

int mava_peer_free(void* buf) {
  MAVA_CHECK_PTR(buf);
  cudaError_t e = cudaFree(buf);
  return e == cudaSuccess ? 0 : (int)e;
}

int mava_peer_status(const void* buf, int64_t n_grad, uint32_t* seq_out_host, uint32_t* err_out_host,
                     mava_stream_t s) {
  MAVA_CHECK_PTR(buf);
  MAVA_CHECK_ARG(n_grad > 0);
  uint32_t w[2] = {0, 0};
  const unsigned char* f = static_cast<const unsigned char*>(buf) + grad_bytes_padded(n_grad);
  cudaError_t e = cudaMemcpyAsync(w, f + F_SEQ * 4, 8, cudaMemcpyDeviceToHost, as_stream(s));
  if (e == cudaSuccess) e = cudaStreamSynchronize(as_stream(s));
  if (e != cudaSuccess) return (int)e;
  if (seq_out_host) *seq_out_host = w[0];
  if (err_out_host) *err_out_host = w[1];
  return 0;
}

int mava_reduce_clip_adam_pair(float* params, float* mu, float* nu, int32_t* counts,
                               const mava_peer_group* group_host, float* gsum, int64_t n_actor,
                               int64_t n_critic, const mava_mlp_desc* actor, void* actor_image,
                               const mava_mlp_desc* critic, void* critic_image, float grad_scale,
                               float lr_actor, float lr_critic, float max_norm,
                               int lr_decay_num_updates, int steps_per_update, float* loss_out5,
                               mava_stream_t s) {
  MAVA_CHECK_PTR(params);
  MAVA_CHECK_PTR(mu);
  MAVA_CHECK_PTR(nu);
  MAVA_CHECK_PTR(counts);
  MAVA_CHECK_PTR(group_host);
  MAVA_CHECK_PTR(gsum);
  MAVA_CHECK_ARG(n_actor > 0 && n_critic > 0 && steps_per_update > 0);
  MAVA_CHECK_ARG(group_host->world >= 1 && group_host->world <= kMaxRanks &&
                 group_host->rank >= 0 && group_host->rank < group_host->world);
  ReduceAdamArgs a{};
  a.params = params; a.mu = mu; a.nu = nu; a.counts = counts;
  a.rank = group_host->rank; a.world = group_host->world;
  const int64_t n_grad = n_actor + n_critic + 8;
  for (int r = 0; r < a.world; ++r) {
    MAVA_CHECK_PTR(group_host->buf[r]);
    MAVA_CHECK_ARG((reinterpret_cast<size_t>(group_host->buf[r]) & 15) == 0);
    a.grad[r] = static_cast<const float*>(group_host->buf[r]);
    a.flags[r] = reinterpret_cast<uint32_t*>(static_cast<unsigned char*>(group_host->buf[r]) +
                                             grad_bytes_padded(n_grad));
  }
  a.gsum = gsum;
  a.n[0] = n_actor; a.n[1] = n_critic;
  a.lr[0] = lr_actor; a.lr[1] = lr_critic;
  a.grad_scale = grad_scale; a.max_norm = max_norm;
  a.lr_decay_num_updates = lr_decay_num_updates; a.steps_per_update = steps_per_update;
  a.loss_out = loss_out5;
  const mava_mlp_desc* nets[2] = {actor, critic};
  void* images[2] = {actor_image, critic_image};
  for (int k = 0; k < 2; ++k) {
    a.image[k] = nullptr;
    if (nets[k] != nullptr && images[k] != nullptr) {
      const mava_mlp_desc* d = nets[k];
      MAVA_CHECK_ARG(d->h1 == tcmlp::HID && d->h2 == tcmlp::HID && d->out_dim <= tcmlp::NHEAD);
      MAVA_CHECK_ARG(mava_mlp_param_count(d) == a.n[k]);
      a.image[k] = static_cast<unsigned char*>(images[k]);
      a.in_dim[k] = d->in_dim;
      a.k1p[k] = tcmlp::pad16(d->in_dim + 1);
      a.out[k] = d->out_dim;
    }
  }
  reduce_clip_adam_kernel<<<kCluster, kThreads, 0, as_stream(s)>>>(a);
  return launch_status();
}

}  // extern "C"
