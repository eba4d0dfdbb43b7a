// Synthetic SMAX-shaped step source for the recurrent benchmark (BASELINE.json configs[3],
// SURVEY.md 8d config 4): the SMAX dynamics live in jaxmarl (third party, not under
// /root/reference, unpinned), so the rec_mappo workload is driven by tensors of the SMAX 3s5z
// SHAPE -- per-agent f32 observations, a world-state row per env, 13-way action masks with the
// first five actions always legal, team reward ~ N(0, reward_std), done ~ Bernoulli(done_prob) --
// produced on the device by a counter hash.  It exercises the GRU acting / GAE / sequence-minibatch
// PPO kernels at their real shapes; it is not an environment and makes no parity claim.
// RecordEpisodeMetrics (mava/wrappers/episode_metrics.py:78-111) is applied for real.
#include "common.cuh"

namespace mava {
namespace {

__device__ __forceinline__ uint32_t hash32(uint32_t x) {  // PCG output permutation
  x = x * 747796405u + 2891336453u;
  const uint32_t w = ((x >> ((x >> 28u) + 4u)) ^ x) * 277803737u;
  return (w >> 22u) ^ w;
}
__device__ __forceinline__ float u01(uint32_t h) { return (float)(h >> 8) * (1.0f / 16777216.0f); }

struct SynthArgs {
  mava_synth_config cfg;
  const uint32_t* key;  // 2 words, device
  uint8_t* state;       // [NE][16]: run_ret f32, run_len i32, ep_ret f32, ep_len i32
  float* obs_actor;     // [NE][A][obs_dim]
  float* obs_critic;    // [NE][state_dim]
  uint16_t* mask;       // [NE][A]
  float* reward;        // [NE][A]
  uint8_t* done;        // [NE]
  float* ep_return;
  int32_t* ep_length;
  int num_envs, reset;
};

__global__ void __launch_bounds__(256) synth_kernel(const SynthArgs p) {
  const mava_synth_config& c = p.cfg;
  const uint32_t seed = hash32(p.key[0] ^ 0x5EED5EEDu) ^ p.key[1];
  const int64_t n_obs = (int64_t)p.num_envs * c.num_agents * c.obs_dim;
  const int64_t n_state = (int64_t)p.num_envs * c.state_dim;
  const int64_t n_ea = (int64_t)p.num_envs * c.num_agents;
  const int64_t total = n_obs + n_state + n_ea + p.num_envs;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t h = hash32(seed + (uint32_t)i * 2654435761u);
    if (i < n_obs) {
      p.obs_actor[i] = u01(h);
    } else if (i < n_obs + n_state) {
      p.obs_critic[i - n_obs] = u01(h);
    } else if (i < n_obs + n_state + n_ea) {
      // 16 bits of h, each kept with probability ~0.7 by and-or of two more draws; 0..4 legal
      const uint32_t a = hash32(h), b = hash32(a);
      uint32_t m = (h | a) & (h | b | ~a);  // P(bit) = 1 - (1/2)(1/2) ... ~0.69
      m = (m | 0x1Fu) & ((1u << c.num_actions) - 1u);
      p.mask[i - n_obs - n_state] = (uint16_t)m;
    } else {
      const int e = (int)(i - n_obs - n_state - n_ea);
      float* st = reinterpret_cast<float*>(p.state + (size_t)e * 16);
      int32_t* sti = reinterpret_cast<int32_t*>(st);
      if (p.reset) {
        st[0] = 0.0f; sti[1] = 0; st[2] = 0.0f; sti[3] = 0;
        continue;
      }
      const uint32_t h2 = hash32(h);
      const float r = c.reward_std * sqrtf(-2.0f * logf(fmaxf(u01(h), 1e-7f))) *
                      cospif(2.0f * u01(h2));
      const bool is_done = u01(hash32(h2)) < c.done_prob;
      for (int a = 0; a < c.num_agents; ++a) p.reward[(size_t)e * c.num_agents + a] = r;
      const float new_ret = st[0] + r;
      const int32_t new_len = sti[1] + 1;
      const float ret_info = is_done ? new_ret : st[2];
      const int32_t len_info = is_done ? new_len : sti[3];
      st[0] = is_done ? 0.0f : new_ret;
      sti[1] = is_done ? 0 : new_len;
      st[2] = ret_info;
      sti[3] = len_info;
      p.done[e] = is_done ? 1 : 0;
      p.ep_return[e] = ret_info;
      p.ep_length[e] = len_info;
    }
  }
}

int launch(const SynthArgs& a, cudaStream_t s) {
  const mava_synth_config& c = a.cfg;
  const int64_t total = (int64_t)a.num_envs * (c.num_agents * c.obs_dim + c.state_dim + c.num_agents + 1);
  const unsigned blocks = (unsigned)min(ceil_div64(total, 256 * 4), (int64_t)sm_count() * 16);
  synth_kernel<<<blocks, 256, 0, s>>>(a);
  return launch_status();
}

int check(const mava_synth_config* c) {
  if (!c) return MAVA_E_NULL;
  if (c->num_agents < 1 || c->obs_dim < 1 || c->state_dim < 1) return MAVA_E_BADARG;
  if (c->num_actions < 5 || c->num_actions > 16) return MAVA_E_BADARG;
  return 0;
}

}  // namespace
}  // namespace mava

using namespace mava;

extern "C" {

int mava_synth_reset(const mava_synth_config* cfg, const uint32_t* key, uint8_t* state,
                     float* obs_actor, float* obs_critic, uint16_t* mask, int num_envs,
                     mava_stream_t s) {
  int rc = check(cfg);
  if (rc) return rc;
  MAVA_CHECK_PTR(key);
  MAVA_CHECK_PTR(state);
  MAVA_CHECK_PTR(obs_actor);
  MAVA_CHECK_PTR(obs_critic);
  MAVA_CHECK_PTR(mask);
  MAVA_CHECK_ARG(num_envs > 0);
  SynthArgs a{*cfg, key, state, obs_actor, obs_critic, mask, nullptr, nullptr, nullptr, nullptr,
              num_envs, 1};
  return launch(a, as_stream(s));
}

int mava_synth_step(const mava_synth_config* cfg, const uint32_t* key, uint8_t* state,
                    const int8_t* action, float* obs_actor, float* obs_critic, uint16_t* mask,
                    float* reward, uint8_t* done, float* ep_return, int32_t* ep_length,
                    int num_envs, mava_stream_t s) {
  int rc = check(cfg);
  if (rc) return rc;
  MAVA_CHECK_PTR(key);
  MAVA_CHECK_PTR(state);
  MAVA_CHECK_PTR(action);
  MAVA_CHECK_PTR(obs_actor);
  MAVA_CHECK_PTR(obs_critic);
  MAVA_CHECK_PTR(mask);
  MAVA_CHECK_PTR(reward);
  MAVA_CHECK_PTR(done);
  MAVA_CHECK_PTR(ep_return);
  MAVA_CHECK_PTR(ep_length);
  MAVA_CHECK_ARG(num_envs > 0);
  SynthArgs a{*cfg, key, state, obs_actor, obs_critic, mask, reward, done, ep_return, ep_length,
              num_envs, 0};
  return launch(a, as_stream(s));
}

}  // extern "C"
