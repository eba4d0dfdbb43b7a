// pmean("device") fused with the optimiser: ONE launch per minibatch does
//   all-reduce(sum) of [actor grads | critic grads | 5 loss scalars] over the ranks' peer-mapped
//   gradient buffers  ->  optax.clip_by_global_norm  ->  optax.adam(eps=1e-5)  ->  apply_updates
//   ->  refresh of the packed bf16 operand images  ->  loss metrics of the minibatch
// (ff_mappo.py:224-250,359-366; rec_mappo.py:283-305).  No NCCL call sits between the loss kernel
// and the next minibatch any more.
//
// Every rank owns one exchange buffer (cudaMalloc + cudaIpc, mapped by all ranks of the node over
// NVLink / NVSwitch): its gradient vector, a flag block and a receive area.  One-shot all-reduce,
// PUSH protocol (default): every rank writes its vector into its slot of every peer's receive area
// as 128-byte lines of seven 16-byte data chunks + one chunk that carries the call number -- the
// eight lanes that own a line store it with one coalesced 128-byte write, which NVLink delivers
// whole, so a receiver that reads the line back (one coalesced 128-byte read) and finds the call
// number it expects has the data too: no separate flag, no fence, no round trip -- one one-way
// NVLink traversal between "my gradients are complete" and "I hold everybody's" (NCCL's LL128 idea).
// Every rank then adds the vectors in rank order 0..W-1, so all ranks hold bit-identical sums (the
// replicated parameters stay replicated).  The slots are double buffered by call parity: a rank
// pushes call k+2 only after it has completed call k+1, for which it needed every peer's push of
// k+1, which a peer issues after it has consumed call k -- so nothing is overwritten before it is
// read and there is no "I have read yours" handshake either.
// PULL protocol (MAVA_PEER_PULL=1, the checker): flag handshake ("my gradients are complete"), every
// rank READS the vectors of all ranks, second handshake ("I have read yours").
//
// The kernel is a cooperative launch of up to 128 small CTAs (all co-resident): the squared global
// norms are combined through two fp64 accumulators and one grid barrier, both kept in the rank's own
// flag block, so the kernel is re-entrant across learners and streams (no device-global scratch).
// Peer handshakes give up after 2 s and raise the buffer's error word instead of hanging the
// device (a rank that died, a mismatched call sequence).
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "mlp_tc.cuh"

using namespace mava;

namespace {

constexpr int kMaxRanks = MAVA_PEER_MAX_RANKS;
constexpr int kThreads = 256;
constexpr int kMaxCtas = 128;  // x 256 threads, one 16-byte chunk per thread for the headline networks
// flag block (uint32 words) behind the gradient vector of an exchange buffer
constexpr int F_READY = 0;            // [kMaxRanks] written by rank p: "p's gradients of call #seq are complete"
constexpr int F_DONE = kMaxRanks;     // [kMaxRanks] written by rank p: "p has read this buffer in call #seq"
constexpr int F_SEQ = 2 * kMaxRanks;  // calls completed on this buffer (local)
constexpr int F_ERR = 2 * kMaxRanks + 1;  // != 0: a handshake timed out (local)
constexpr int F_GRID = 2 * kMaxRanks + 2;  // arrivals at the kernel's grid barrier, ever growing (local)
constexpr int F_NORM = 2 * kMaxRanks + 4;  // 2 x 2 doubles: squared norms of call parity 0 / 1 (local)
constexpr int F_WORDS = 32;
constexpr int kLineData = 7;  // 16-byte data chunks per 128-byte line of the receive area

__host__ __device__ inline int64_t grad_bytes_padded(int64_t n_grad) {
  return (n_grad * 4 + 255) / 256 * 256;
}
// receive area: [parity 0 | 1][source rank 0 .. kMaxRanks) slots of `lines` 128-byte lines
__host__ __device__ inline int64_t recv_lines(int64_t n_grad) {
  return ((n_grad + 3) / 4 + kLineData - 1) / kLineData;
}
__host__ __device__ inline int64_t recv_offset(int64_t n_grad) {
  return grad_bytes_padded(n_grad) + 256;  // behind the flag block, 128-byte aligned
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
__device__ __forceinline__ uint4 ld_sys_u4(const void* p) {
  uint4 v;
  asm volatile("ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p)
               : "memory");
  return v;
}
__device__ __forceinline__ void st_sys_u4(void* p, uint4 v) {
  asm volatile("st.relaxed.sys.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w)
               : "memory");
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

// Phase timing (development aid, -DMAVA_PEER_STAMPS, scripts/exp_peer_stamps.sh): thread 0 of CTA 0
// accumulates globaltimer deltas of the kernel's phases over the calls.
#ifdef MAVA_PEER_STAMPS
__device__ unsigned long long g_peer_ns[8];
#define MAVA_PSTAMP(k)                                                  \
  do {                                                                  \
    if (blockIdx.x == 0 && threadIdx.x == 0) {                          \
      const unsigned long long now_ = global_ns();                      \
      g_peer_ns[k] += now_ - pst_;                                      \
      pst_ = now_;                                                      \
    }                                                                   \
  } while (0)
#else
#define MAVA_PSTAMP(k) do { } while (0)
#endif

struct ReduceAdamArgs {
  float *params, *mu, *nu;
  int32_t* counts;
  const float* grad[kMaxRanks];  // this rank's mapping of every rank's gradient vector
  uint32_t* flags[kMaxRanks];    // ... and of every rank's flag block (world > 1)
  uint32_t* local_flags;         // this rank's flag block (world == 1: a block owned by the library)
  unsigned char* recv[kMaxRanks];  // ... and of every rank's receive area (push protocol)
  int64_t lines;                 // 128-byte lines per slot
  int push;                      // 1: push protocol, 0: pull
  int rank, world;
  float* gsum;  // local scratch, n[0] + n[1] floats: the reduced, scaled gradients
  int64_t n[2];
  float lr[2];
  float grad_scale, max_norm;
  int lr_decay_num_updates, steps_per_update;
  unsigned char* image[2];
  int in_dim[2], k1p[2], out[2];
  float* loss_out;  // [5] or null: mean over ranks of the loss scalars behind the gradients
  // accumulating pairing with mava_ppo_loss_grad_bf16_acc: the five loss scalars behind the
  // gradients are computed here from the loss kernel's accumulators, and both the rank's gradient
  // vector and the accumulators are left zero for the next minibatch
  double* loss_acc;  // null: the scalars are in the buffer already, nothing is cleared
  double loss_denom;
  float ent_coef, vf_coef;
};

// element i of the rank's own [gradients | 5 loss scalars | pad] vector
__device__ __forceinline__ float loss_scalar(const ReduceAdamArgs& a, int k) {
  const double actor_loss = a.loss_acc[0] / a.loss_denom, entropy = a.loss_acc[1] / a.loss_denom,
               value_loss = a.loss_acc[2] / a.loss_denom;
  switch (k) {  // as finalize_loss_kernel (mlp_f32.cu)
    case 0: return (float)(actor_loss - (double)a.ent_coef * entropy);
    case 1: return (float)actor_loss;
    case 2: return (float)entropy;
    case 3: return (float)((double)a.vf_coef * value_loss);
    case 4: return (float)value_loss;
    default: return 0.0f;
  }
}
__device__ __forceinline__ float4 own_chunk(const ReduceAdamArgs& a, int64_t c, int64_t n01) {
  float4 v = *reinterpret_cast<const float4*>(a.grad[a.rank] + 4 * c);
  if (a.loss_acc != nullptr && 4 * c + 4 > n01) {
    float* f = reinterpret_cast<float*>(&v);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t i = 4 * c + j;
      if (i >= n01) f[j] = loss_scalar(a, (int)(i - n01));
    }
  }
  return v;
}

struct AdamStep {  // per network: what every element of one call shares
  float g_norm, step_lr, bc1, bc2;
  bool keep;
};

// optax.clip_by_global_norm -> optax.adam(eps=1e-5) -> apply_updates on one element (registers)
__device__ __forceinline__ float adam_math(float g, float& m, float& v, float p, const AdamStep& st,
                                           float max_norm) {
  const float b1 = 0.9f, b2 = 0.999f, eps = 1e-5f;
  if (!st.keep) g = (g / st.g_norm) * max_norm;
  m = (1.0f - b1) * g + b1 * m;
  v = (1.0f - b2) * g * g + b2 * v;
  return p - st.step_lr * ((m / st.bc1) / (sqrtf(v / st.bc2) + eps));
}

__device__ __forceinline__ void store_image(const ReduceAdamArgs& a, int net, int i_local, float p) {
  if (a.image[net])
    *reinterpret_cast<__nv_bfloat16*>(
        a.image[net] + tcmlp::image_offset(i_local, a.in_dim[net], a.k1p[net], a.out[net])) =
        __float2bfloat16_rn(p);
}

// Grid barrier of a cooperative launch (all CTAs co-resident): arrivals are counted in a word of the
// rank's flag block that only ever grows; generation k of a kernel that runs `per_call` barriers is
// complete at (calls_before * per_call + k) * gridDim.x arrivals.
__device__ __forceinline__ void grid_arrive_wait(uint32_t* counter, uint32_t target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    while ((int32_t)(ld_acquire_sys(counter) - target) < 0) __nanosleep(32);
    __threadfence();
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kThreads, 3) reduce_clip_adam_kernel(const ReduceAdamArgs a) {
  const int t = threadIdx.x;
  const int64_t tid = (int64_t)blockIdx.x * kThreads + t;
  const int64_t nthreads = (int64_t)gridDim.x * kThreads;
  uint32_t* my_flags = a.local_flags;
  __shared__ double s_red[kThreads / 32][2];

  const int c0[2] = {a.counts[0], a.counts[1]};
  const uint32_t calls = my_flags[F_SEQ];  // calls completed on this flag block
  const uint32_t seq = calls + 1u;
  double* norm2 = reinterpret_cast<double*>(my_flags + F_NORM) + 2 * (seq & 1u);
#ifdef MAVA_PEER_STAMPS
  unsigned long long pst_ = global_ns();
  if (blockIdx.x == 0 && t == 0) g_peer_ns[7] += 1;
#endif

  const int64_t n0 = a.n[0], n01 = a.n[0] + a.n[1], n_all = n01 + 8;
  const int64_t chunks = (n_all + 3) >> 2;  // the buffers are padded: whole chunks are always readable
  float ss[2] = {0.0f, 0.0f};
  // one reduced chunk: scale, keep, squared norms, loss metrics
  auto take = [&](int64_t c, float4 s) {
    float g[4] = {s.x * a.grad_scale, s.y * a.grad_scale, s.z * a.grad_scale, s.w * a.grad_scale};
    if (4 * c + 4 <= n01) *reinterpret_cast<float4*>(a.gsum + 4 * c) = make_float4(g[0], g[1], g[2], g[3]);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t i = 4 * c + j;
      if (i < n01) {
        if (4 * c + 4 > n01) a.gsum[i] = g[j];
        const int net = i < n0 ? 0 : 1;
        ss[net] = fmaf(g[j], g[j], ss[net]);
      } else if (i < n01 + 5 && a.loss_out != nullptr) {
        a.loss_out[i - n01] = g[j];
      }
    }
  };
  if (a.world > 1 && a.push) {
    // ---- push: my vector into my slot of every peer's receive area, 128-byte lines -------------
    const int sub = t & 7;                       // chunk of the line this lane owns (7: call number)
    const int64_t grp = tid >> 3, ngrp = nthreads >> 3;
    const int64_t slot_bytes = a.lines * 128;
    const int64_t my_slot = ((int64_t)(seq & 1u) * kMaxRanks + a.rank) * slot_bytes;
    for (int64_t line = grp; line < a.lines; line += ngrp) {
      const int64_t c = line * kLineData + sub;
      uint4 w = make_uint4(seq, seq, seq, seq);
      if (sub < kLineData) {
        w = make_uint4(0u, 0u, 0u, 0u);
        if (c < chunks) {
          const float4 f = own_chunk(a, c, n01);
          w = make_uint4(__float_as_uint(f.x), __float_as_uint(f.y), __float_as_uint(f.z),
                         __float_as_uint(f.w));
        }
      }
#pragma unroll
      for (int r = 0; r < kMaxRanks; ++r)
        if (r < a.world && r != a.rank) st_sys_u4(a.recv[r] + my_slot + line * 128 + sub * 16, w);
    }
    MAVA_PSTAMP(0);
    // ---- receive: every peer's line must carry this call's number; sum in rank order -------------
    const unsigned char* mine = a.recv[a.rank] + (int64_t)(seq & 1u) * kMaxRanks * slot_bytes;
    for (int64_t line0 = (tid >> 5) * 4; line0 < a.lines; line0 += (nthreads >> 5) * 4) {
      const int64_t line = line0 + ((t & 31) >> 3);  // a warp takes four lines at a time
      const bool live = line < a.lines;
      const int64_t c = line * kLineData + sub;
      uint4 v[kMaxRanks];
      unsigned pending = 0u;  // peers whose line has not arrived yet
#pragma unroll
      for (int r = 0; r < kMaxRanks; ++r)
        if (r < a.world && r != a.rank) pending |= 1u << r;
      const unsigned long long t0 = global_ns();
      for (uint32_t spins = 0;; ++spins) {
#pragma unroll
        for (int r = 0; r < kMaxRanks; ++r)
          if (live && ((pending >> r) & 1u)) v[r] = ld_sys_u4(mine + r * slot_bytes + line * 128 + sub * 16);
        unsigned still = 0u;
#pragma unroll
        for (int r = 0; r < kMaxRanks; ++r) {
          if (r < a.world && r != a.rank) {
            const uint32_t tag = __shfl_sync(0xffffffffu, v[r].x, 7, 8);  // the line's eighth chunk
            if (live && ((pending >> r) & 1u) && tag != seq) still |= 1u << r;
          }
        }
        pending = still;
        if (!__any_sync(0xffffffffu, pending != 0u)) break;
        if ((spins & 255u) == 255u && global_ns() - t0 > 2000000000ull) {
          my_flags[F_ERR] = 1u;
          break;
        }
      }
      if (live && sub < kLineData && c < chunks) {
        float4 s = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll
        for (int r = 0; r < kMaxRanks; ++r) {
          if (r < a.world) {
            float4 x;
            if (r == a.rank) {
              x = own_chunk(a, c, n01);
            } else {
              x = make_float4(__uint_as_float(v[r].x), __uint_as_float(v[r].y),
                              __uint_as_float(v[r].z), __uint_as_float(v[r].w));
            }
            if (r == 0) s = x;
            else { s.x += x.x; s.y += x.y; s.z += x.z; s.w += x.w; }
          }
        }
        take(c, s);
      }
    }
  } else {
    // ---- pull, handshake 1: every rank's gradients of this call are complete ---------------------
    if (a.world > 1) {
      if (blockIdx.x == 0 && t < a.world && t != a.rank)
        st_release_sys(a.flags[t] + F_READY + a.rank, seq);
      if (t < a.world && t != a.rank) {
        if (!wait_flag(my_flags + F_READY + t, seq)) my_flags[F_ERR] = 1u;
      }
      __syncthreads();
    }
    MAVA_PSTAMP(0);
    // ---- sum over ranks (rank order) --------------------------------------------------------------
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
        s = own_chunk(a, c, n01);
      }
      take(c, s);
    }
  }
  MAVA_PSTAMP(1);
  // fp32 partial sums per thread (a handful of elements), combined in fp64
  for (int o = 16; o > 0; o >>= 1) {
    ss[0] += __shfl_xor_sync(0xffffffffu, ss[0], o);
    ss[1] += __shfl_xor_sync(0xffffffffu, ss[1], o);
  }
  if ((t & 31) == 0) {
    s_red[t >> 5][0] = (double)ss[0];
    s_red[t >> 5][1] = (double)ss[1];
  }
  __syncthreads();
  if (t == 0) {
    double v0 = 0.0, v1 = 0.0;
    for (int w = 0; w < kThreads / 32; ++w) {
      v0 += s_red[w][0];
      v1 += s_red[w][1];
    }
    atomicAdd(norm2 + 0, v0);
    atomicAdd(norm2 + 1, v1);
  }
  // every CTA of this rank has consumed its share of the peers' buffers and posted its norms
  grid_arrive_wait(my_flags + F_GRID, seq * gridDim.x);
  MAVA_PSTAMP(2);

  // ---- handshake 2, first half: tell the peers their buffers have been read ---------------------
  if (a.world > 1 && !a.push && blockIdx.x == 0 && t < a.world && t != a.rank)
    st_release_sys(a.flags[t] + F_DONE + a.rank, seq);

  // ---- clip + Adam + apply (+ bf16 image refresh); every thread re-reads what it wrote ----------
  AdamStep st[2];
#pragma unroll
  for (int net = 0; net < 2; ++net) {
    st[net].g_norm = (float)sqrt(*reinterpret_cast<volatile double*>(norm2 + net));
    st[net].keep = st[net].g_norm < a.max_norm;
    const int c = c0[net] + 1;
    st[net].bc1 = 1.0f - powf(0.9f, (float)c);
    st[net].bc2 = 1.0f - powf(0.999f, (float)c);
    st[net].step_lr = a.lr[net];
    if (a.lr_decay_num_updates > 0)
      st[net].step_lr *=
          1.0f - (float)(c0[net] / a.steps_per_update) / (float)a.lr_decay_num_updates;
  }
  // Whole 16-byte chunks: the four moments / parameters of a chunk are loaded together (no store
  // sits between the loads of a thread's elements), updated in registers and stored together.
  const int n0i = (int)n0, n01i = (int)n01;
  float* own = const_cast<float*>(a.grad[a.rank]);
  for (int64_t c = tid; c < chunks; c += nthreads) {
    const int i0 = (int)(4 * c);
    // every CTA is past the grid barrier: the rank's own vector has been consumed
    if (a.loss_acc != nullptr) *reinterpret_cast<float4*>(own + i0) = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (i0 + 4 <= n01i) {
      const float4 g4 = *reinterpret_cast<const float4*>(a.gsum + i0);
      float4 m4 = *reinterpret_cast<const float4*>(a.mu + i0);
      float4 v4 = *reinterpret_cast<const float4*>(a.nu + i0);
      float4 p4 = *reinterpret_cast<const float4*>(a.params + i0);
      const float g[4] = {g4.x, g4.y, g4.z, g4.w};
      float m[4] = {m4.x, m4.y, m4.z, m4.w}, v[4] = {v4.x, v4.y, v4.z, v4.w};
      float pn[4] = {p4.x, p4.y, p4.z, p4.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int net = i0 + j < n0i ? 0 : 1;
        pn[j] = adam_math(g[j], m[j], v[j], pn[j], st[net], a.max_norm);
      }
      *reinterpret_cast<float4*>(a.mu + i0) = make_float4(m[0], m[1], m[2], m[3]);
      *reinterpret_cast<float4*>(a.nu + i0) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(a.params + i0) = make_float4(pn[0], pn[1], pn[2], pn[3]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int net = i0 + j < n0i ? 0 : 1;
        store_image(a, net, i0 + j - (net ? n0i : 0), pn[j]);
      }
    } else {
      for (int i = i0; i < n01i; ++i) {  // the chunk that holds the end of the vector
        const int net = i < n0i ? 0 : 1;
        float m = a.mu[i], v = a.nu[i];
        const float pn = adam_math(a.gsum[i], m, v, a.params[i], st[net], a.max_norm);
        a.mu[i] = m;
        a.nu[i] = v;
        a.params[i] = pn;
        store_image(a, net, i - (net ? n0i : 0), pn);
      }
    }
  }
  for (int64_t i64 = 4 * chunks + tid; i64 < n01; i64 += nthreads) {
    const int i = (int)i64;
    const int net = i < n0i ? 0 : 1;
    float m = a.mu[i], v = a.nu[i];
    const float pn = adam_math(a.gsum[i], m, v, a.params[i], st[net], a.max_norm);
    a.mu[i] = m;
    a.nu[i] = v;
    a.params[i] = pn;
    store_image(a, net, i - (net ? n0i : 0), pn);
  }

  MAVA_PSTAMP(3);
  // ---- handshake 2, second half: nobody reads this rank's buffer any more -----------------------
  if (blockIdx.x == 0) {
    if (a.world > 1 && !a.push && t < a.world && t != a.rank) {
      if (!wait_flag(my_flags + F_DONE + t, seq)) my_flags[F_ERR] = 1u;
    }
    __syncthreads();
    if (t == 0) {
      // every CTA has passed the grid barrier, i.e. has read counts, F_SEQ and the norms of this call
      a.counts[0] = c0[0] + 1;
      a.counts[1] = c0[1] + 1;
      if (a.loss_acc != nullptr) a.loss_acc[0] = a.loss_acc[1] = a.loss_acc[2] = 0.0;
      double* other = reinterpret_cast<double*>(my_flags + F_NORM) + 2 * ((seq + 1u) & 1u);
      other[0] = 0.0;  // the next call's accumulators (this call's are zeroed by the call after it)
      other[1] = 0.0;
      my_flags[F_SEQ] = seq;
    }
  }
  MAVA_PSTAMP(4);
}

}  // namespace

extern "C" {

#ifdef MAVA_PEER_STAMPS
int mava_debug_peer_stamps(unsigned long long* out8_host, int reset) {
  cudaError_t e = cudaMemcpyFromSymbol(out8_host, g_peer_ns, sizeof(unsigned long long) * 8);
  if (e == cudaSuccess && reset) {
    const unsigned long long z[8] = {};
    e = cudaMemcpyToSymbol(g_peer_ns, z, sizeof(z));
  }
  return (int)e;
}
#endif

int64_t mava_peer_buffer_bytes(int64_t n_grad) {
  if (n_grad <= 0) return -1;
  return recv_offset(n_grad) + 2 * kMaxRanks * recv_lines(n_grad) * 128;
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

static int reduce_clip_adam_impl(float* params, float* mu, float* nu, int32_t* counts,
                                 const mava_peer_group* group_host, float* gsum, int64_t n_actor,
                                 int64_t n_critic, const mava_mlp_desc* actor, void* actor_image,
                                 const mava_mlp_desc* critic, void* critic_image, float grad_scale,
                                 float lr_actor, float lr_critic, float max_norm,
                                 int lr_decay_num_updates, int steps_per_update, float* loss_out5,
                                 double* loss_acc, double loss_denom, float ent_coef, float vf_coef,
                                 mava_stream_t s) {
  MAVA_CHECK_PTR(params);
  MAVA_CHECK_PTR(mu);
  MAVA_CHECK_PTR(nu);
  MAVA_CHECK_PTR(counts);
  MAVA_CHECK_PTR(group_host);
  MAVA_CHECK_PTR(gsum);
  MAVA_CHECK_ARG(n_actor > 0 && n_critic > 0 && steps_per_update > 0 &&
                 n_actor + n_critic < ((int64_t)1 << 31));
  MAVA_CHECK_ARG(((reinterpret_cast<size_t>(params) | reinterpret_cast<size_t>(mu) |
                   reinterpret_cast<size_t>(nu) | reinterpret_cast<size_t>(gsum)) & 15) == 0);
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
  static_assert(F_WORDS * 4 <= 256, "flag block fits in front of the receive area");
  for (int r = 0; r < a.world; ++r)
    a.recv[r] = static_cast<unsigned char*>(group_host->buf[r]) + recv_offset(n_grad);
  a.lines = recv_lines(n_grad);
  {
    const char* pull = getenv("MAVA_PEER_PULL");  // read per call: the tests compare both protocols
    a.push = (pull != nullptr && pull[0] == '1') ? 0 : 1;
  }
  a.local_flags = a.flags[a.rank];
  // clearing the rank's vector is only legal when nobody else reads it (one rank, or push)
  if (loss_acc != nullptr && a.world > 1 && !a.push) return MAVA_E_UNSUPPORTED;
  MAVA_CHECK_ARG(loss_acc == nullptr || loss_denom > 0.0);
  a.loss_acc = loss_acc;
  a.loss_denom = loss_denom;
  a.ent_coef = ent_coef;
  a.vf_coef = vf_coef;
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
  const int64_t chunks = (n_grad + 3) / 4;
  int ctas = (int)ceil_div64(a.world > 1 && a.push ? a.lines * 8 : chunks, kThreads);
  ctas = ctas < 1 ? 1 : (ctas > kMaxCtas ? kMaxCtas : ctas);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)ctas);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = as_stream(s);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;  // the grid barrier needs every CTA resident
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, reduce_clip_adam_kernel, a);
  if (e != cudaSuccess) return (int)e;
  return launch_status();
}

int mava_reduce_clip_adam_pair(float* params, float* mu, float* nu, int32_t* counts,
                               const mava_peer_group* group_host, float* gsum, int64_t n_actor,
                               int64_t n_critic, const mava_mlp_desc* actor, void* actor_image,
                               const mava_mlp_desc* critic, void* critic_image, float grad_scale,
                               float lr_actor, float lr_critic, float max_norm,
                               int lr_decay_num_updates, int steps_per_update, float* loss_out5,
                               mava_stream_t s) {
  return reduce_clip_adam_impl(params, mu, nu, counts, group_host, gsum, n_actor, n_critic, actor,
                               actor_image, critic, critic_image, grad_scale, lr_actor, lr_critic,
                               max_norm, lr_decay_num_updates, steps_per_update, loss_out5, nullptr,
                               0.0, 0.0f, 0.0f, s);
}

int mava_reduce_clip_adam_pair_acc(float* params, float* mu, float* nu, int32_t* counts,
                                   const mava_peer_group* group_host, float* gsum, int64_t n_actor,
                                   int64_t n_critic, const mava_mlp_desc* actor, void* actor_image,
                                   const mava_mlp_desc* critic, void* critic_image,
                                   float grad_scale, float lr_actor, float lr_critic, float max_norm,
                                   int lr_decay_num_updates, int steps_per_update, float* loss_out5,
                                   void* loss_workspace, const mava_ppo_hyper* hyper,
                                   int64_t loss_rows, mava_stream_t s) {
  MAVA_CHECK_PTR(loss_workspace);
  MAVA_CHECK_PTR(hyper);
  MAVA_CHECK_ARG(loss_rows > 0);
  // the loss accumulators of mava_ppo_loss_grad_bf16_acc: three doubles at byte 128 of its workspace
  double* acc = reinterpret_cast<double*>(static_cast<unsigned char*>(loss_workspace) + 128);
  return reduce_clip_adam_impl(params, mu, nu, counts, group_host, gsum, n_actor, n_critic, actor,
                               actor_image, critic, critic_image, grad_scale, lr_actor, lr_critic,
                               max_norm, lr_decay_num_updates, steps_per_update, loss_out5, acc,
                               (double)loss_rows, hyper->ent_coef, hyper->vf_coef, s);
}

}  // extern "C"
