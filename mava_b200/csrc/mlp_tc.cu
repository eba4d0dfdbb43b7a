// bf16 tensor-core (tcgen05 + TMEM) actor / critic kernels for sm_100a.
//
//   mava_mlp_pack_bf16   fp32 parameters -> bf16 weight image in the shared-memory operand format
//   mava_ff_act_bf16     acting step: both networks' forward pass + masked categorical sampling
//
// One CTA = 128 threads = one 128-row tile; thread t owns tile row t and TMEM lane t.  Weight
// images are staged into shared memory by bulk (TMA) copies, the int8 observations are expanded to
// bf16 in shared memory, each layer is a chain of tcgen05.mma (M=128, N=128|16, K=16) issued by one
// thread into a TMEM accumulator, and the epilogues (bias, relu, softmax, Gumbel arg-max on threefry
// bits) read the accumulator back with tcgen05.ld.  Actor tiles and critic tiles run as different
// CTAs of the same launch.
//
// Replaces the same reference regions as mlp_f32.cu (mava/networks.py:39-58,88-124,172-207;
// mava/systems/ppo/ff_mappo.py:81-85) with bf16 operands / fp32 accumulation (tolerance 2e-2).
#include "mlp_tc.cuh"

namespace mava {
namespace tcmlp {
namespace {

// ------------------------------------------------------------------------------------------------
// weight packing
// ------------------------------------------------------------------------------------------------
__global__ void pack_kernel(const float* __restrict__ params, int in_dim, int k1p, int out,
                            unsigned char* __restrict__ image) {
  // flat flax order: W1 (in,H) | b1 | W2 (H,H) | b2 | W3 (H,out) | b3
  const float* w1 = params;
  const float* b1 = w1 + (size_t)in_dim * HID;
  const float* w2 = b1 + HID;
  const float* b2 = w2 + (size_t)HID * HID;
  const float* w3 = b2 + HID;
  const float* b3 = w3 + (size_t)HID * out;
  const int n1 = k1p * (HID / 8), n2 = HCOLS * (HID / 8), n3 = HCOLS * (NHEAD / 8);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n1 + n2 + n3;
       i += gridDim.x * blockDim.x) {
    const float *src, *bias;
    int rows, r, cg, ld, rows_valid, cols_valid;
    size_t base;
    if (i < n1) {
      rows = k1p; r = i % rows; cg = i / rows; src = w1; bias = b1; ld = HID; rows_valid = in_dim;
      cols_valid = HID; base = 0;
    } else if (i < n1 + n2) {
      const int j = i - n1;
      rows = HCOLS; r = j % rows; cg = j / rows; src = w2; bias = b2; ld = HID; rows_valid = HID;
      cols_valid = HID; base = (size_t)k1p * HID * 2;
    } else {
      const int j = i - n1 - n2;
      rows = HCOLS; r = j % rows; cg = j / rows; src = w3; bias = b3; ld = out; rows_valid = HID;
      cols_valid = out; base = (size_t)k1p * HID * 2 + (size_t)HCOLS * HID * 2;
    }
    uint32_t w[4];
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      float v[2];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int c = cg * 8 + h * 2 + q;
        float x = 0.0f;
        if (c < cols_valid) {
          if (r < rows_valid) x = src[(size_t)r * ld + c];
          else if (r == rows_valid) x = bias[c];  // the bias row, met by the tiles' ones column
        }
        v[q] = x;
      }
      w[h] = pack_bf16(v[0], v[1]);
    }
    const size_t off = base + (size_t)(r >> 3) * 128 + (size_t)cg * (rows / 8) * 128 + (r & 7) * 16;
    *reinterpret_cast<uint4*>(image + off) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// ------------------------------------------------------------------------------------------------
// acting kernel
// ------------------------------------------------------------------------------------------------
struct ActArgs {
  NetDesc actor, critic;
  const unsigned char *actor_img, *critic_img;
  const int8_t* view;
  const uint8_t* mask;
  const uint32_t* policy_key;
  const int8_t* actions_in;
  int8_t* action;
  float *logp, *value;
  int num_envs, envs_per_replica, greedy;
  int actor_ctas, critic_ctas;
};

struct Ctrl {
  uint64_t wbar;   // weights landed
  uint64_t mbar;   // MMA chain done
  uint32_t tmem;
};

// shared memory: [weights image][X tile][H tile]
__global__ void __launch_bounds__(NT, 1) act_kernel(const ActArgs p) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ Ctrl ctrl;
  const Lane L;
  const int t = L.t, warp = L.warp;
  const bool is_actor = (int)blockIdx.x < p.actor_ctas;
  const NetDesc& d = is_actor ? p.actor : p.critic;
  const unsigned char* img = is_actor ? p.actor_img : p.critic_img;
  const int tile = is_actor ? blockIdx.x : blockIdx.x - p.actor_ctas;
  const int64_t M = (int64_t)p.num_envs * (d.mode == MAVA_IN_GLOBAL ? 1 : d.A);
  const int64_t row0 = (int64_t)tile * TM;

  const WImage wi{d.k1p};
  const uint32_t s_w = smem_u32(smem);
  const Tile xt{s_w + wi.total(), 128u, (uint32_t)(TM / 8) * 128u};
  const Tile ht{xt.base + tile_bytes(TM, d.k1p), 128u, (uint32_t)(TM / 8) * 128u};

  if (warp == 0) tmem_alloc<256>(&ctrl.tmem);
  if (t == 0) {
    mbar_init(&ctrl.wbar, 1);
    mbar_init(&ctrl.mbar, 1);
    fence_mbar_init();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = ctrl.tmem;
  if (t == 0) load_weights(s_w, img, wi.total(), &ctrl.wbar);

  // while acting, rows are (env, agent) for AGENT_VIEW and env for GLOBAL, in buffer order
  // (the H tile doubles as the staging area: it is not live before layer 1)
  build_x_tile(d, p.view, xt, smem + (ht.base - s_w), (int)row0, (int)M, [](int64_t j) { return j; });
  fence_proxy_async();
  mbar_wait(&ctrl.wbar, 0);
  fence_before_sync();
  __syncthreads();
  uint32_t phase = 0;
  // ---- layer 1
  if (mma_issuer()) {
    fence_after_sync();
    issue_gemm(tmem, xt, false, w1_tile(s_w, d.k1p), true, HID, d.k1p, false, &ctrl.mbar);
  }
  wait_mma(&ctrl.mbar, phase);
  phase ^= 1;
  hidden_epilogue(L, tmem, ht);
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  // ---- layer 2 (accumulator reused: every thread has drained its columns)
  if (mma_issuer()) {
    fence_after_sync();
    issue_gemm(tmem, ht, false, w2_tile(s_w, d.k1p), true, HID, HCOLS, false, &ctrl.mbar);
  }
  wait_mma(&ctrl.mbar, phase);
  phase ^= 1;
  hidden_epilogue(L, tmem, ht);  // layer-2 MMAs have completed: H1 may be overwritten
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  // ---- head
  if (mma_issuer()) {
    fence_after_sync();
    issue_gemm(tmem + HID, ht, false, w3_tile(s_w, d.k1p), true, NHEAD, HCOLS, false, &ctrl.mbar);
  }
  wait_mma(&ctrl.mbar, phase);
  const int64_t row = row0 + L.r;
  if (L.q == 0) {  // the first four warps finish the rows (one thread per row)
    float out[NHEAD];
    ld16(tmem + L.tmem_lane() + (uint32_t)HID, out);
    if (row < M) {
      if (!is_actor) {
        const float v = out[0];
        if (d.mode == MAVA_IN_GLOBAL) {
          for (int a = 0; a < d.A; ++a) p.value[row * d.A + a] = v;
        } else {
          p.value[row] = v;
        }
      } else {
        const uint8_t mk = p.mask[row];
        float mx = kF32Min;
#pragma unroll
        for (int j = 0; j < NHEAD; ++j) {
          if (j < d.out) {
            out[j] = ((mk >> j) & 1) ? out[j] : kF32Min;
            mx = fmaxf(mx, out[j]);
          }
        }
        float se = 0.0f;
#pragma unroll
        for (int j = 0; j < NHEAD; ++j)
          if (j < d.out) se += expf(out[j] - mx);
        const float lse = mx + logf(se);
        int a = 0;
        if (p.actions_in) {
          a = p.actions_in[row];
        } else if (p.greedy) {
          float best = out[0];
#pragma unroll
          for (int j = 1; j < NHEAD; ++j)
            if (j < d.out && out[j] > best) { best = out[j]; a = j; }
        } else {
          const Key key{p.policy_key[0], p.policy_key[1]};
          const int64_t s = row / d.A;
          const int ag = (int)(row - s * d.A);
          const int64_t e = s % p.envs_per_replica;
          const uint32_t size = (uint32_t)p.envs_per_replica * d.A * d.out;
          const uint32_t base = (uint32_t)((e * d.A + ag) * d.out);
          float best = 0.0f;
#pragma unroll
          for (int j = 0; j < NHEAD; ++j) {
            if (j < d.out) {
              const float z = bits_to_gumbel(random_bits_at(key, base + j, size)) + out[j];
              if (j == 0 || z > best) { best = z; a = j; }
            }
          }
        }
        float la = 0.0f;
#pragma unroll
        for (int j = 0; j < NHEAD; ++j)
          if (j == a) la = out[j] - lse;
        p.action[row] = (int8_t)a;
        p.logp[row] = la;
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem);
}

// ------------------------------------------------------------------------------------------------
// batched critic pass on joint observations (centralised critic, all T + 1 slots of a rollout)
// ------------------------------------------------------------------------------------------------
// The rows of a tile are 128 consecutive env-steps of in_dim = A * FR bytes: ONE bulk (TMA) copy per
// tile.  Persistent CTAs: the weight image is loaded once per CTA (the one-tile-per-CTA acting
// kernel reloads 111 KB per 128 rows), and the next tile's bytes are requested as soon as layer 1
// has consumed X -- into the upper half of the X region, which is why the expansion goes through
// registers (every thread reads its chunks, then all of them write).
struct ValueCtrl {
  uint64_t wbar, mbar, xbar;  // weights landed | MMA chain | observation bytes landed
  uint32_t tmem;
};

__global__ void __launch_bounds__(NT, 1) value_batch_kernel(const ActArgs p) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ ValueCtrl ctrl;
  const Lane L;
  const int t = L.t, warp = L.warp;
  const NetDesc& d = p.critic;
  const int M = p.num_envs;  // one row per env-step
  const int n_tiles = ceil_div(M, TM);
  const WImage wi{d.k1p};
  const uint32_t s_w = smem_u32(smem);
  const uint32_t x_bytes = tile_bytes(TM, d.k1p);
  const Tile xt{s_w + wi.total(), 128u, (uint32_t)(TM / 8) * 128u};
  const Tile ht{xt.base + x_bytes, 128u, (uint32_t)(TM / 8) * 128u};
  const uint32_t row_bytes = (uint32_t)d.in_dim;
  const uint32_t stage_off = (x_bytes - TM * row_bytes) & ~15u;  // upper part of the X region
  unsigned char* stage = smem + wi.total() + stage_off;

  if (warp == 0) tmem_alloc<256>(&ctrl.tmem);
  if (t == 0) {
    mbar_init(&ctrl.wbar, 1);
    mbar_init(&ctrl.mbar, 1);
    mbar_init(&ctrl.xbar, 1);
    fence_mbar_init();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = ctrl.tmem;
  // a tile whose byte count is a multiple of 16 arrives by one bulk copy; the (partial) last tile
  // may not be: it is copied with plain loads
  auto tile_rows = [&](int tile) { return M - tile * TM < TM ? M - tile * TM : TM; };
  auto bulk_ok = [&](int tile) { return (((uint32_t)tile_rows(tile) * row_bytes) & 15u) == 0u; };
  auto request = [&](int tile) {  // one thread
    const uint32_t bytes = (uint32_t)tile_rows(tile) * row_bytes;
    if (bulk_ok(tile)) {
      mbar_expect_tx(&ctrl.xbar, bytes);
      bulk_g2s(smem_u32(stage), p.view + (size_t)tile * TM * row_bytes, bytes, &ctrl.xbar);
    } else {
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&ctrl.xbar)) : "memory");
    }
  };
  if (t == 0) {
    load_weights(s_w, p.critic_img, wi.total(), &ctrl.wbar);
    if ((int)blockIdx.x < n_tiles) request(blockIdx.x);
  }
  mbar_wait(&ctrl.wbar, 0);
  uint32_t phase = 0, xphase = 0;
  const int nchunks = d.in_dim >> 3;  // chunks of observation bytes; chunk nchunks is [1 | 0...]
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int row0 = tile * TM;
    mbar_wait(&ctrl.xbar, xphase);
    xphase ^= 1u;
    if (!bulk_ok(tile)) {  // plain copy of the last tile (8-byte units)
      const int units = tile_rows(tile) * (int)(row_bytes >> 3);
      const uint2* src = reinterpret_cast<const uint2*>(p.view + (size_t)tile * TM * row_bytes);
      for (int i = t; i < units; i += NT) reinterpret_cast<uint2*>(stage)[i] = __ldg(src + i);
      __syncthreads();
    }
    // ---- X tile: observation bytes -> registers -> bf16 chunks (the staging bytes live in the
    //      region the chunks are written to)
    constexpr int kMaxChunks = 9;  // per thread: in_dim <= 288
    uint32_t w[kMaxChunks][2];
    const bool valid = row0 + L.r < M;
    const uint32_t src = smem_u32(stage) + (uint32_t)L.r * row_bytes;
#pragma unroll
    for (int i = 0; i < kMaxChunks; ++i) {
      const int cg = L.q + 4 * i;
      w[i][0] = w[i][1] = 0u;
      if (valid && cg < nchunks)
        asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(w[i][0]), "=r"(w[i][1]) : "r"(src + 8u * cg));
    }
    __syncthreads();  // staging has been read: X may overwrite it
#pragma unroll
    for (int i = 0; i < kMaxChunks; ++i) {
      const int cg = L.q + 4 * i;
      if (cg < nchunks)
        st_shared_v4(xt.base + chunk_off(xt, L.r, cg), s8x2_bf16x2(w[i][0]), s8x2_bf16x2(w[i][0] >> 16),
                     s8x2_bf16x2(w[i][1]), s8x2_bf16x2(w[i][1] >> 16));
    }
    for (int cg = nchunks + L.q; cg < d.k1p / 8; cg += 4)  // ones column, padding
      st_shared_v4(xt.base + chunk_off(xt, L.r, cg), (valid && cg == nchunks) ? 0x00003F80u : 0u, 0u,
                   0u, 0u);
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    // ---- layer 1
    if (mma_issuer()) {
      fence_after_sync();
      issue_gemm(tmem, xt, false, w1_tile(s_w, d.k1p), true, HID, d.k1p, false, &ctrl.mbar);
    }
    wait_mma(&ctrl.mbar, phase);
    phase ^= 1;
    // X is dead: the next tile's bytes may land in its upper half
    if (t == 0 && tile + (int)gridDim.x < n_tiles) request(tile + gridDim.x);
    hidden_epilogue(L, tmem, ht);
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    // ---- layer 2
    if (mma_issuer()) {
      fence_after_sync();
      issue_gemm(tmem, ht, false, w2_tile(s_w, d.k1p), true, HID, HCOLS, false, &ctrl.mbar);
    }
    wait_mma(&ctrl.mbar, phase);
    phase ^= 1;
    hidden_epilogue(L, tmem, ht);
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    // ---- head
    if (mma_issuer()) {
      fence_after_sync();
      issue_gemm(tmem + HID, ht, false, w3_tile(s_w, d.k1p), true, NHEAD, HCOLS, false, &ctrl.mbar);
    }
    wait_mma(&ctrl.mbar, phase);
    phase ^= 1;
    if (L.q == 0) {
      float out[8];
      ld8(tmem + L.tmem_lane() + (uint32_t)HID, out);
      if (valid)
        for (int a = 0; a < d.A; ++a) p.value[(size_t)(row0 + L.r) * d.A + a] = out[0];
    }
    fence_before_sync();
    __syncthreads();  // the head accumulator has been read before the next tile's chain writes
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem);
}

}  // namespace

int make_net(const mava_mlp_desc* d, const float* params, NetDesc* n) {
  if (!d) return MAVA_E_NULL;
  if (d->h1 != HID || d->h2 != HID) return MAVA_E_UNSUPPORTED;
  if (d->out_dim < 1 || d->out_dim > NHEAD) return MAVA_E_UNSUPPORTED;
  n->mode = d->input_mode;
  n->add_id = d->add_agent_id;
  n->A = d->num_agents;
  n->FR = d->view_dim;
  n->in_dim = d->in_dim;
  n->k1p = pad16(d->in_dim + 1);
  n->out = d->out_dim;
  if (n->k1p > 304) return MAVA_E_UNSUPPORTED;  // shared-memory budget of the training kernel
  if (stage_bytes(n->A, n->FR, n->mode == MAVA_IN_GLOBAL ? 1 : n->A) > tile_bytes(TM, HCOLS))
    return MAVA_E_UNSUPPORTED;  // the observation staging area aliases an activation tile
  (void)params;
  return 0;
}

}  // namespace tcmlp
}  // namespace mava

using namespace mava;
using namespace mava::tcmlp;

extern "C" {

int64_t mava_mlp_pack_bytes(const mava_mlp_desc* d) {
  NetDesc n;
  if (make_net(d, nullptr, &n)) return -1;
  return (int64_t)WImage{n.k1p}.total();
}

int mava_mlp_pack_bf16(const mava_mlp_desc* d, const float* params, void* image,
                       mava_stream_t s) {
  NetDesc n;
  int rc = make_net(d, params, &n);
  if (rc) return rc;
  MAVA_CHECK_PTR(params);
  MAVA_CHECK_PTR(image);
  pack_kernel<<<32, 256, 0, as_stream(s)>>>(params, n.in_dim, n.k1p, n.out,
                                            static_cast<unsigned char*>(image));
  return launch_status();
}

int mava_ff_act_bf16(const mava_mlp_desc* actor, const float* actor_params, const void* actor_image,
                     const mava_mlp_desc* critic, const float* critic_params,
                     const void* critic_image, const int8_t* view, const uint8_t* mask,
                     const uint32_t* policy_key, int envs_per_replica, int num_envs, int greedy,
                     const int8_t* actions_in, int8_t* action, float* logp, float* value,
                     mava_stream_t s) {
  ActArgs a{};
  int rc = 0;
  MAVA_CHECK_PTR(view);
  MAVA_CHECK_ARG(num_envs > 0 && envs_per_replica > 0);
  MAVA_CHECK_ARG((int64_t)num_envs * 8 < ((int64_t)1 << 31));  // tile rows are indexed in 32 bits
  MAVA_CHECK_ARG(actor != nullptr || value != nullptr);
  if (actor) {  // actor == NULL: critic only (bootstrap value, ff_mappo.py:110)
    rc = make_net(actor, actor_params, &a.actor);
    if (rc) return rc;
    MAVA_CHECK_PTR(actor_params);
    MAVA_CHECK_PTR(actor_image);
    MAVA_CHECK_PTR(mask);
    MAVA_CHECK_PTR(action);
    MAVA_CHECK_PTR(logp);
    MAVA_CHECK_ARG(actor->input_mode == MAVA_IN_AGENT_VIEW);
    if (!greedy && !actions_in) MAVA_CHECK_PTR(policy_key);
    a.actor_img = static_cast<const unsigned char*>(actor_image);
    a.actor_ctas = (int)ceil_div64((int64_t)num_envs * actor->num_agents, TM);
  }
  int k1p_max = actor ? a.actor.k1p : 16;
  if (value) {
    rc = make_net(critic, critic_params, &a.critic);
    if (rc) return rc;
    MAVA_CHECK_PTR(critic_params);
    MAVA_CHECK_PTR(critic_image);
    MAVA_CHECK_ARG(critic->out_dim == 1);
    a.critic_img = static_cast<const unsigned char*>(critic_image);
    const int64_t rows = (int64_t)num_envs * (critic->input_mode == MAVA_IN_GLOBAL ? 1 : critic->num_agents);
    a.critic_ctas = (int)ceil_div64(rows, TM);
    if (a.critic.k1p > k1p_max) k1p_max = a.critic.k1p;
  }
  a.view = view;
  a.mask = mask;
  a.policy_key = policy_key;
  a.actions_in = actions_in;
  a.action = action;
  a.logp = logp;
  a.value = value;
  a.num_envs = num_envs;
  a.envs_per_replica = envs_per_replica;
  a.greedy = greedy;
  const size_t smem = (size_t)WImage{k1p_max}.total() + tile_bytes(TM, k1p_max) +
                      tile_bytes(TM, HCOLS) + 128;
  // critic only, joint observations: the persistent batched kernel (one bulk copy per tile)
  if (!actor && critic->input_mode == MAVA_IN_GLOBAL && (a.critic.in_dim & 7) == 0 &&
      a.critic.in_dim <= 288 && a.critic.k1p > a.critic.in_dim &&
      (size_t)TM * a.critic.in_dim + 16 <= tile_bytes(TM, a.critic.k1p) &&
      (reinterpret_cast<size_t>(view) & 15) == 0) {
    static size_t configured_v[kMaxDevices] = {};
    if (int rc2 = ensure_dyn_smem(value_batch_kernel, smem, configured_v)) return rc2;
    const int ctas = a.critic_ctas < sm_count() ? a.critic_ctas : sm_count();
    value_batch_kernel<<<ctas, NT, smem, as_stream(s)>>>(a);
    return launch_status();
  }
  static size_t configured[kMaxDevices] = {};
  if (int rc2 = ensure_dyn_smem(act_kernel, smem, configured)) return rc2;
  act_kernel<<<a.actor_ctas + a.critic_ctas, NT, smem, as_stream(s)>>>(a);
  return launch_status();
}

}  // extern "C"
