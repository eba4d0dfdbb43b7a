/* mava_b200.h - C ABI of libmava_b200.so: the B200 (sm_100a) kernels behind Mava's Anakin PPO path.
 *
 * The reference (RuanJohn/Mava v0.2.0) has no FFI of its own: the seam is a set of Python
 * callables that XLA compiles.  Each entry point below replaces one XLA-compiled region and
 * cites the reference lines it stands in for.  This is what a jax.ffi / ctypes binding binds.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - every call is asynchronous on the caller-supplied CUDA stream (cudaStream_t passed as
 *     void*), never synchronises, never allocates on the hot path;
 *   - return value: 0 ok, <0 invalid argument (MAVA_E_*), >0 a cudaError_t;
 *   - buffers are caller-owned; the library keeps only immutable per-scenario constants inside
 *     the opaque env handle.
 *
 * Layouts (NE = envs resident on this GPU = update_batch_size * num_envs, A = agents,
 * FR = raw observation features per agent, T = rollout_length):
 *   env state      uint8 [NE][state_stride]           packed per-env record, see DESIGN.md
 *   view           int8  [NE][A][FR]                   raw integer observation (no ids)
 *   mask           uint8 [NE][A]                       bit k set = action k legal
 *   action         int8  [NE][A]
 *   logp, value    f32   [NE][A]
 *   reward         f32   [NE][A]
 *   done           uint8 [NE]
 *   ep_return f32 [NE], ep_length int32 [NE]           RecordEpisodeMetrics outputs
 * Rollout buffers stack these with a leading time axis.
 */
#ifndef MAVA_B200_H_
#define MAVA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MAVA_B200_ABI_VERSION 1

#define MAVA_E_BADARG (-1)
#define MAVA_E_UNSUPPORTED (-2)
#define MAVA_E_NULL (-3)

typedef void* mava_stream_t; /* cudaStream_t */
typedef struct mava_env_s* mava_env_t;

int mava_abi_version(void);
/* Human readable text for a return code of any function below (static storage). */
const char* mava_error_string(int code);
/* Device properties the host sizes grids with.  out[0]=SM count, out[1]=cc major, out[2]=cc minor. */
int mava_device_info(int* out3_host);

/* ------------------------------------------------------------------------------------------
 * PRNG - jax.random (threefry2x32) as the reference uses it.
 * ---------------------------------------------------------------------------------------- */
/* Sequential chain of n `key, sub = jax.random.split(key)` (ff_mappo.py:81).  subkeys[n][2]
 * receives the n sub keys, key_io[2] is advanced in place.  One thread; n is small (T). */
int mava_prng_split_chain(uint32_t* key_io, uint32_t* subkeys, int n, mava_stream_t s);
/* jax.random.split(key, num) -> out[num][2] (ff_mappo.py:392-394). */
int mava_prng_split(const uint32_t* key, uint32_t* out, int num, mava_stream_t s);
/* jax random_bits(key, (n,)) -> out[n] uint32; the sort keys of jax.random.permutation
 * (ff_mappo.py:273). */
int mava_prng_random_bits(const uint32_t* key, uint32_t* out, int64_t n, mava_stream_t s);

/* One round of jax.random.permutation's sort (ff_mappo.py:273; jax: sort_key_val(random_bits, x)):
 * val_out = val_in stably sorted by the uint32 keys (mava_prng_random_bits).  Keys are assumed to be
 * (close to) uniform: a two-level bucket + shared-memory bitonic sort; *overflow_flag (device int,
 * zero it once) is set if a bucket ever exceeds its capacity, in which case val_out is invalid.
 * workspace: mava_sort_workspace_bytes(n) bytes.  val_in and val_out must not alias. */
int64_t mava_sort_workspace_bytes(int64_t n);
int mava_sort_by_key(const uint32_t* keys, const int32_t* val_in, int32_t* val_out, int64_t n,
                     void* workspace, int* overflow_flag, mava_stream_t s);

/* ------------------------------------------------------------------------------------------
 * Environments - replaces jax.vmap(env.step) / jax.vmap(env.reset) through the whole wrapper
 * stack RecordEpisodeMetrics(AutoResetWrapper(AgentIDWrapper(RwareWrapper|LbfWrapper(env))))
 * (ff_mappo.py:88,395; mava/utils/make_env.py:69-83; mava/wrappers/ *.py).
 * ---------------------------------------------------------------------------------------- */
#define MAVA_ENV_RWARE 1
#define MAVA_ENV_LBF 2

typedef struct mava_rware_config {
  int32_t column_height, shelf_rows, shelf_columns;
  int32_t num_agents, sensor_range, request_queue_size;
  int32_t time_limit;
} mava_rware_config; /* configs/env/scenario/tiny-4ag.yaml task_config + env/rware.yaml kwargs */

typedef struct mava_lbf_config {
  int32_t grid_size, fov, num_agents, num_food, max_agent_level, force_coop;
  int32_t time_limit;
  int32_t use_individual_rewards; /* configs/env/lbf.yaml */
} mava_lbf_config;

typedef struct mava_env_dims {
  int32_t kind;
  int32_t num_agents;   /* A  */
  int32_t view_dim;     /* FR */
  int32_t num_actions;  /* N  */
  int32_t state_stride; /* bytes per env record, multiple of 16 */
  int32_t time_limit;
  int32_t grid_h, grid_w;
  int32_t aux0, aux1;   /* RWARE: n_shelves, queue size.  LBF: n_food, max level */
  int32_t algo_bytes_per_step; /* minimal lossless bytes one env-step moves */
} mava_env_dims;

int mava_env_create(int kind, const void* config_host, size_t config_size, mava_env_t* out);
int mava_env_destroy(mava_env_t env);
int mava_env_dims_of(mava_env_t env, mava_env_dims* out_host);

/* vmap(env.reset)(keys): keys[NE][2] -> state, first observation (view, mask). */
int mava_env_reset(mava_env_t env, const uint32_t* keys, uint8_t* state, int8_t* view,
                   uint8_t* mask, int num_envs, mava_stream_t s);
/* vmap(env.step)(state, action).  State is updated in place.  auto_reset != 0 applies
 * AutoResetWrapper (training env); 0 is the evaluation env (make_env.py:79-81).  reward is the
 * per-agent reward after the wrapper's aggregation.  ep_return/ep_length are the
 * `episode_metrics` extras; `is_terminal_step` equals done. */
int mava_env_step(mava_env_t env, uint8_t* state, const int8_t* action, int8_t* view,
                  uint8_t* mask, float* reward, uint8_t* done, float* ep_return,
                  int32_t* ep_length, int num_envs, int auto_reset, mava_stream_t s);
/* Decode fields of the packed state for inspection/tests: writes int32 per env.
 * field ids: 0 step_count, 1 env key (2 words), 2 agents (A*4: x,y,dir,carry ...) */
int mava_env_peek(mava_env_t env, const uint8_t* state, int field, int32_t* out, int num_envs,
                  mava_stream_t s);

/* ------------------------------------------------------------------------------------------
 * Networks - FeedForwardActor / FeedForwardValueNet (mava/networks.py:39-58,88-124,172-207).
 * Parameters are one flat f32 vector per network in flax order and layout:
 *   [Dense_0.kernel (in,h1) | Dense_0.bias | Dense_1.kernel (h1,h2) | Dense_1.bias |
 *    head.kernel (h2,out) | head.bias]
 * ---------------------------------------------------------------------------------------- */
#define MAVA_IN_AGENT_VIEW 0 /* x = [onehot(agent) if add_agent_id | view[e][a][:]]            */
#define MAVA_IN_GLOBAL 1     /* x = concat_a view[e][a][:]  (ObservationGlobalState.global_state) */

typedef struct mava_mlp_desc {
  int32_t input_mode;   /* MAVA_IN_*                                   */
  int32_t add_agent_id; /* AgentIDWrapper on (observation.py:41-53)     */
  int32_t num_agents, view_dim;
  int32_t in_dim;       /* derived: A*FR, or FR (+A)                    */
  int32_t h1, h2;       /* MLPTorso.layer_sizes (two hidden layers)     */
  int32_t out_dim;      /* action_dim for the actor, 1 for the critic   */
} mava_mlp_desc;

/* Number of f32 parameters of a network. */
int64_t mava_mlp_param_count(const mava_mlp_desc* d_host);

/* One acting step (ff_mappo.py:81-85): logits -> masked categorical sample (Gumbel arg-max on
 * threefry bits of policy_key, noise laid out (envs_per_replica, A, N)), log_prob, value.
 * greedy != 0 takes pi.mode() instead (evaluator.py:183).  value may be NULL (evaluator).
 * If actions_in != NULL the sample is replaced by the given action (replay for parity tests). */
int mava_ff_act(const mava_mlp_desc* actor_host, const float* actor_params,
                const mava_mlp_desc* critic_host, const float* critic_params,
                const int8_t* view, const uint8_t* mask, const uint32_t* policy_key,
                int envs_per_replica, int num_envs, int greedy, const int8_t* actions_in,
                int8_t* action, float* logp, float* value, mava_stream_t s);
/* Critic only (last_val, ff_mappo.py:110). */
int mava_ff_value(const mava_mlp_desc* critic_host, const float* critic_params, const int8_t* view,
                  int num_envs, float* value, mava_stream_t s);

/* ------------------------------------------------------------------------------------------
 * GAE (ff_mappo.py:112-139; rec_mappo.py:177-199).  One thread per env-agent, reverse scan.
 * reward/value/adv/targets: [T][NE][A] f32; done [T][NE] uint8; last_val [NE][A].
 * rec != 0: `done` holds the flag entering each step and last_done[NE] seeds the carry.
 * ---------------------------------------------------------------------------------------- */
int mava_gae(const float* reward, const float* value, const uint8_t* done, const float* last_val,
             const uint8_t* last_done, float gamma, float gae_lambda, int T, int num_envs,
             int num_agents, int rec, float* adv, float* targets, mava_stream_t s);

/* ------------------------------------------------------------------------------------------
 * PPO minibatch (ff_mappo.py:144-266).
 * ---------------------------------------------------------------------------------------- */
typedef struct mava_ppo_hyper {
  float clip_eps, ent_coef, vf_coef;
} mava_ppo_hyper;

/* Episode metrics of the finished episodes, reduced on the device.
 * Replaces get_final_step_metrics (mava/wrappers/episode_metrics.py:114-132) followed by the
 * logger's describe() (mava/utils/logger.py:44-58: mean / std / min / max) for the run loop
 * (mava/systems/ppo/ff_mappo.py:499-505): instead of moving the three [T][NE] arrays of
 * ExperimentOutput.episode_metrics to the host, the loop moves stats[10] (f64):
 *   {count, sum_return, sumsq_return, min_return, max_return,
 *    sum_length, sumsq_length, min_length, max_length, 0}
 * over the n env-steps with done != 0.  reset != 0 re-initialises stats first; otherwise the
 * call accumulates (one call per update of an evaluation interval). */
int mava_episode_stats(const uint8_t* done, const float* ep_return, const int32_t* ep_length,
                       int64_t n, int reset, double* stats, mava_stream_t s);

/* Row index list of one minibatch from a permutation (ff_mappo.py:272-280): for replica u and
 * position j, rows[u*mb + j] = t*NE + u*E + e with (t,e) = divmod(perm[mb_index*mb + j], E). */
int mava_ppo_minibatch_rows(const int32_t* perm, int mb_index, int mb_size, int num_replicas,
                            int envs_per_replica, int32_t* rows, mava_stream_t s);

/* Gradients of the actor and critic losses for one minibatch, averaged over the
 * update_batch_size replicas on this GPU (value_and_grad + pmean("batch"), :205-226,232-234).
 * rows[num_replicas*mb_size] index env-steps of the rollout buffers (leading [T*NE]).
 * grad_out: [actor grads | critic grads | total_actor, actor_loss, entropy, total_critic,
 * value_loss | pad] f32, the buffer that is all-reduced across GPUs.
 * workspace: at least mava_ppo_workspace_bytes(...) bytes. */
int64_t mava_ppo_workspace_bytes(const mava_mlp_desc* actor_host, const mava_mlp_desc* critic_host,
                                 int rows_total);
int mava_ppo_loss_grad(const mava_mlp_desc* actor_host, const float* actor_params,
                       const mava_mlp_desc* critic_host, const float* critic_params,
                       const mava_ppo_hyper* hyper_host, const int8_t* view, const uint8_t* mask,
                       const int8_t* action, const float* old_logp, const float* old_value,
                       const float* adv, const float* targets, const int32_t* rows,
                       int num_replicas, int mb_size, float* grad_out, void* workspace,
                       mava_stream_t s);

/* optax.chain(clip_by_global_norm(max_norm), adam(lr, eps=1e-5)) + apply_updates for ONE
 * network (ff_mappo.py:240-250,359-366).  grad is the summed buffer after the cross-GPU
 * all-reduce; grad_scale = 1/world_size turns the sum into pmean("device").  count[1] (int32,
 * device) is optax's step count, incremented here.  lr_decay_num_updates > 0 enables the linear
 * schedule of mava/utils/training.py:39-47 with steps_per_update = ppo_epochs*num_minibatches. */
int mava_clip_adam(float* params, float* mu, float* nu, int32_t* count, const float* grad,
                   int64_t n, float grad_scale, float lr, float max_norm, int lr_decay_num_updates,
                   int steps_per_update, mava_stream_t s);

/* Both networks in two launches (squared norms, update): params/mu/nu/grad hold [actor | critic],
 * counts[2].  Kept for hosts that run their own all-reduce; NOT re-entrant across streams of one
 * device (the norm accumulator is a device global) -- mava_reduce_clip_adam_pair below is, and is what
 * the learners of this repository call. */
int mava_clip_adam_pair(float* params, float* mu, float* nu, int32_t* counts, const float* grad,
                        int64_t n_actor, int64_t n_critic, float grad_scale, float lr_actor,
                        float lr_critic, float max_norm, int lr_decay_num_updates,
                        int steps_per_update, mava_stream_t s);

/* Same step for networks that also live as packed bf16 operand images (mava_mlp_pack_bf16): the
 * thread that updates a parameter refreshes its bf16 copy, so no packing launch follows the
 * optimiser (optax.apply_updates, ff_mappo.py:243-250, then the next minibatch's forward pass).
 * Both networks must be [128, 128] torsos; the images must have been packed once before. */
int mava_clip_adam_pair_pack(float* params, float* mu, float* nu, int32_t* counts, const float* grad,
                             const mava_mlp_desc* actor, void* actor_image,
                             const mava_mlp_desc* critic, void* critic_image, float grad_scale,
                             float lr_actor, float lr_critic, float max_norm,
                             int lr_decay_num_updates, int steps_per_update, mava_stream_t s);

/* ------------------------------------------------------------------------------------------
 * pmean("device") fused with the optimiser step (ff_mappo.py:224-250; rec_mappo.py:283-305):
 * one launch = all-reduce(sum) of the ranks' gradient buffers over peer-mapped memory (NVLink /
 * NVSwitch) -> clip_by_global_norm -> adam -> apply_updates -> bf16 image refresh -> loss metrics.
 *
 * Every rank of the node owns ONE exchange buffer: [n_grad floats: actor grads | critic grads |
 * 8 loss scalars][flag block][receive area: 2 call parities x 8 source ranks x 128-byte lines].
 * mava_peer_alloc creates it (cudaMalloc, zeroed) and returns the 64-byte CUDA IPC handle the host
 * layer sends to the other ranks (torch.distributed / MPI / a pipe - not this library's business);
 * mava_peer_open maps a peer's buffer.  The loss kernels write their gradients straight into the
 * rank's own buffer (grad_out = buf), so nothing is copied.  Inside the kernel every rank pushes its
 * vector into its slot of every peer's receive area as 128-byte lines that carry the call number
 * (one one-way NVLink traversal, no flag round trip; slots double buffered by call parity, so the
 * rank may overwrite its gradients right after the call), then adds the W vectors in rank order:
 * all ranks obtain bit-identical sums, like the reference's psum.  Environment MAVA_PEER_PULL=1
 * selects the checker protocol instead (flag handshake, peers' vectors read over NVLink, second
 * handshake).  Lines / flags that do not arrive within 2 s raise the buffer's error word
 * (mava_peer_status) instead of hanging the device.  All ranks must make the same sequence of calls.
 * world == 1 degenerates to clip + Adam on buf[0], which must still be a mava_peer_alloc buffer: the
 * kernel keeps its call counter, its grid-barrier word and the two squared-norm accumulators in the
 * flag block (no device-global scratch: re-entrant across learners and streams). */
#define MAVA_PEER_MAX_RANKS 8
typedef struct mava_peer_group {
  int32_t rank, world;
  void* buf[MAVA_PEER_MAX_RANKS]; /* this process's mapping of every rank's exchange buffer */
} mava_peer_group;
int64_t mava_peer_buffer_bytes(int64_t n_grad);
int mava_peer_alloc(int64_t bytes, void** buf_out, void* ipc_handle64_host);
int mava_peer_open(const void* ipc_handle64_host, void** buf_out);
int mava_peer_close(void* buf);
int mava_peer_free(void* buf);
/* Calls completed on this rank's buffer and its error word (synchronises the stream). */
int mava_peer_status(const void* buf, int64_t n_grad, uint32_t* seq_out_host, uint32_t* err_out_host,
                     mava_stream_t s);
/* gsum: local scratch of n_actor + n_critic floats.  actor/critic + images: optional (both NULL for
 * networks without packed bf16 images, e.g. the recurrent ones).  grad_scale = 1/world_size.
 * loss_out5 (optional) receives the 5 loss scalars behind the gradients, averaged over the ranks. */
int mava_reduce_clip_adam_pair(float* params, float* mu, float* nu, int32_t* counts,
                               const mava_peer_group* group_host, float* gsum, int64_t n_actor,
                               int64_t n_critic, const mava_mlp_desc* actor, void* actor_image,
                               const mava_mlp_desc* critic, void* critic_image, float grad_scale,
                               float lr_actor, float lr_critic, float max_norm,
                               int lr_decay_num_updates, int steps_per_update, float* loss_out5,
                               mava_stream_t s);

/* The pair above without three graph nodes per minibatch: mava_ppo_loss_grad_bf16_acc ADDS into
 * grad_out and into the loss accumulators of its workspace (it clears nothing and does not finalise
 * the loss metrics); mava_reduce_clip_adam_pair_acc computes the five loss scalars from those
 * accumulators (loss_workspace = the loss call's workspace, loss_rows = num_replicas * mb_size *
 * num_agents, hyper = the loss call's), and leaves the rank's gradient vector and the accumulators
 * ZERO for the next minibatch.  Both start from zeroed buffers (mava_peer_alloc zeroes; zero the
 * first 256 bytes of the workspace once).  Only with one rank or the push protocol (nobody else reads
 * the rank's vector): MAVA_E_UNSUPPORTED under MAVA_PEER_PULL=1 with world > 1. */
int mava_ppo_loss_grad_bf16_acc(const mava_mlp_desc* actor, const float* actor_params,
                                const void* actor_image, const mava_mlp_desc* critic,
                                const float* critic_params, const void* critic_image,
                                const mava_ppo_hyper* hyper, const int8_t* view, const uint8_t* mask,
                                const int8_t* action, const float* old_logp, const float* old_value,
                                const float* adv, const float* targets, const int32_t* rows,
                                int num_replicas, int mb_size, const double* adv_stats,
                                float* grad_out, void* workspace, mava_stream_t stream);
int mava_reduce_clip_adam_pair_acc(float* params, float* mu, float* nu, int32_t* counts,
                                   const mava_peer_group* group_host, float* gsum, int64_t n_actor,
                                   int64_t n_critic, const mava_mlp_desc* actor, void* actor_image,
                                   const mava_mlp_desc* critic, void* critic_image,
                                   float grad_scale, float lr_actor, float lr_critic, float max_norm,
                                   int lr_decay_num_updates, int steps_per_update, float* loss_out5,
                                   void* loss_workspace, const mava_ppo_hyper* hyper,
                                   int64_t loss_rows, mava_stream_t s);

/* ------------------------------------------------------------------------------------------
 * bf16 tensor-core path (tcgen05 + TMEM).  Same regions as mava_ff_act / mava_ppo_loss_grad with
 * bf16 operands and fp32 accumulation (tolerance 2e-2, BASELINE.json).  Requires h1 == h2 == 128.
 * The fp32 parameters are first packed into a bf16 image in the shared-memory operand format;
 * re-pack whenever the parameters change.
 * ---------------------------------------------------------------------------------------- */
int64_t mava_mlp_pack_bytes(const mava_mlp_desc* d_host);
int mava_mlp_pack_bf16(const mava_mlp_desc* d_host, const float* params, void* image,
                       mava_stream_t s);
/* actor_host == NULL evaluates the critic only (bootstrap value). */
int mava_ff_act_bf16(const mava_mlp_desc* actor_host, const float* actor_params,
                     const void* actor_image, const mava_mlp_desc* critic_host,
                     const float* critic_params, const void* critic_image, const int8_t* view,
                     const uint8_t* mask, const uint32_t* policy_key, int envs_per_replica,
                     int num_envs, int greedy, const int8_t* actions_in, int8_t* action,
                     float* logp, float* value, mava_stream_t s);

/* Fused forward + loss + backward of one minibatch on the tensor cores; same contract and
 * grad_out layout as mava_ppo_loss_grad.  workspace >= mava_ppo_workspace_bytes_bf16(...). */
int64_t mava_ppo_workspace_bytes_bf16(const mava_mlp_desc* actor_host,
                                      const mava_mlp_desc* critic_host, int rows_total);
int mava_ppo_loss_grad_bf16(const mava_mlp_desc* actor_host, const float* actor_params,
                            const void* actor_image, const mava_mlp_desc* critic_host,
                            const float* critic_params, const void* critic_image,
                            const mava_ppo_hyper* hyper_host, const int8_t* view,
                            const uint8_t* mask, const int8_t* action, const float* old_logp,
                            const float* old_value, const float* adv, const float* targets,
                            const int32_t* rows, int num_replicas, int mb_size, float* grad_out,
                            void* workspace, mava_stream_t s);

/* The per-replica advantage statistics of a minibatch (sum and sum of squares over its rows and
 * agents: the mean / std of `gae` in _actor_loss_fn, ff_mappo.py:164) depend only on the rollout and
 * the row list, not on the parameters: mava_ppo_adv_stats computes them ahead of time (stats: 16
 * doubles, [2 * replica] = sum, [2 * replica + 1] = sum of squares) and
 * mava_ppo_loss_grad_bf16_stats takes them instead of recomputing them, so that the row lists and
 * statistics of all minibatches of an update can be prepared off the critical path. */
int mava_ppo_adv_stats(const float* adv, const int32_t* rows, int num_replicas, int mb_size,
                       int num_agents, double* stats, mava_stream_t stream);
int mava_ppo_loss_grad_bf16_stats(const mava_mlp_desc* actor, const float* actor_params,
                                  const void* actor_image, const mava_mlp_desc* critic,
                                  const float* critic_params, const void* critic_image,
                                  const mava_ppo_hyper* hyper, const int8_t* view,
                                  const uint8_t* mask, const int8_t* action, const float* old_logp,
                                  const float* old_value, const float* adv, const float* targets,
                                  const int32_t* rows, int num_replicas, int mb_size,
                                  const double* adv_stats, float* grad_out, void* workspace,
                                  mava_stream_t stream);

/* The whole rollout scan of ff_ippo / ff_mappo (ff_mappo.py:76-106) in one persistent kernel:
 * rollout_length x (actor forward + masked categorical sample + env step through the wrapper
 * stack), env records, observation rows and the actor weights resident in shared memory for the
 * whole rollout.  view / mask have rollout_length + 1 time slots: slot 0 holds the observation the
 * first step acts on, slot t + 1 receives the observation after step t.  policy_keys[T][2] are the
 * per-step sampling keys (mava_prng_split_chain).  action, logp, reward: [T][NE][A]; done,
 * ep_return, ep_length: [T][NE].  The critic values are computed afterwards with one batched
 * mava_ff_act_bf16(actor = NULL) call over the T + 1 observation slots.
 * Returns MAVA_E_UNSUPPORTED unless the env is RobotWarehouse with 2, 4 or 8 agents and
 * sensor_range 1 (callers then fall back to mava_ff_act_bf16 + mava_env_step per step). */
int mava_ff_rollout_bf16(mava_env_t env, const mava_mlp_desc* actor_host, const float* actor_params,
                         const void* actor_image, uint8_t* state, int8_t* view, uint8_t* mask,
                         const uint32_t* policy_keys, int envs_per_replica, int num_envs,
                         int rollout_length, int8_t* action, float* logp, float* reward,
                         uint8_t* done, float* ep_return, int32_t* ep_length, mava_stream_t s);
/* The same kernel as the evaluator runs it (mava/evaluator.py:80-172): auto_reset = 0 is the
 * evaluation env (no AutoResetWrapper, make_env.py:79-81), greedy != 0 takes pi.mode()
 * (evaluator.py:183), record = 0 keeps only done / ep_return / ep_length per step ([T][NE]) and
 * writes observation, mask, action, log-prob and reward of every step into slot 0 of their
 * buffers, which then need one time slot only. */
int mava_ff_rollout_bf16_ex(mava_env_t env, const mava_mlp_desc* actor_host, const float* actor_params,
                            const void* actor_image, uint8_t* state, int8_t* view, uint8_t* mask,
                            const uint32_t* policy_keys, int envs_per_replica, int num_envs,
                            int rollout_length, int auto_reset, int greedy, int record,
                            int8_t* action, float* logp, float* reward, uint8_t* done,
                            float* ep_return, int32_t* ep_length, mava_stream_t s);
/* The evaluator's metric pick (evaluator.py:143-150): per env, episode return and length at the
 * FIRST terminal step of done[T][NE] (argmax of the flag; step 0 if the episode never ended). */
int mava_episode_first_terminal(const uint8_t* done, const float* ep_return,
                                const int32_t* ep_length, int T, int num_envs, float* out_return,
                                int32_t* out_length, mava_stream_t s);

/* ------------------------------------------------------------------------------------------
 * Recurrent systems - RecurrentActor / RecurrentValueNet / ScannedRNN (mava/networks.py:238-331)
 * as rec_ippo / rec_mappo use them (mava/systems/ppo/rec_mappo.py:91-149,208-293,334-360).
 * Network: Dense(in,H)+relu -> reset-masked GRU(H) -> Dense(H,post)+relu -> Dense(post,out).
 * Parameters are one flat f32 vector per network:
 *   [pre.kernel (in,H) | pre.bias | Wi (H,3H) = [ir|iz|in] kernels | bi (3H) = [ir|iz|in] biases |
 *    Wh (H,3H) = [hr|hz|hn] kernels | hn.bias (H) | post.kernel (H,post) | post.bias |
 *    head.kernel (post,out) | head.bias]            (flax GRUCell: hr and hz carry no bias)
 * Hidden states are f32 [num_envs * rows_per_env][H].
 * ---------------------------------------------------------------------------------------- */
#define MAVA_IN_DENSE 2 /* x = obs[step][row][:] given as f32 (e.g. SMAX-shaped observations)   */

typedef struct mava_rnn_desc {
  int32_t input_mode;   /* MAVA_IN_AGENT_VIEW | MAVA_IN_GLOBAL | MAVA_IN_DENSE                   */
  int32_t add_agent_id; /* AGENT_VIEW only                                                       */
  int32_t num_agents, view_dim;
  int32_t in_dim;       /* derived for the view modes; the feature count for DENSE               */
  int32_t rows_per_env; /* A, or 1 for a centralised critic (identical rows are evaluated once)  */
  int32_t hidden;       /* pre_torso width == GRU width (flax: features = ins.shape[-1])         */
  int32_t post;         /* post_torso width                                                      */
  int32_t out_dim;      /* action_dim for the actor, 1 for the critic                            */
  int32_t precision;    /* 0: fp32 contractions (rtol 1e-5 path); 1: bf16 operands on the tcgen05
                           tensor cores with fp32 accumulation (tolerance 2e-2, BASELINE.json)   */
} mava_rnn_desc;

int64_t mava_rnn_param_count(const mava_rnn_desc* d_host);

/* One acting step (rec_mappo.py:91-134): both networks advance their hidden state from
 * h_*_in to h_*_out (out may alias in, or be NULL to discard the new state - the bootstrap value
 * of rec_mappo.py:165 does that), the carry being zeroed first where done_in[env] is set
 * (ScannedRNN, networks.py:249-253).  actor_host == NULL evaluates the critic only.
 * view is the int8 observation of the env kernels; obs_actor / obs_critic are the f32 inputs of
 * MAVA_IN_DENSE networks ([num_envs][rows_per_env][in_dim]).  mask holds uint8 entries (bit k =
 * action k legal) for out_dim <= 8 and uint16 entries above.  Sampling as mava_ff_act. */
int64_t mava_rec_act_workspace_bytes(const mava_rnn_desc* actor_host,
                                     const mava_rnn_desc* critic_host, int num_envs);
int mava_rec_act(const mava_rnn_desc* actor_host, const float* actor_params,
                 const mava_rnn_desc* critic_host, const float* critic_params, const int8_t* view,
                 const float* obs_actor, const float* obs_critic, const void* mask,
                 const uint8_t* done_in, const float* h_actor_in, float* h_actor_out,
                 const float* h_critic_in, float* h_critic_out, const uint32_t* policy_key,
                 int envs_per_replica, int num_envs, int greedy, const int8_t* actions_in,
                 int8_t* action, float* logp, float* value, void* workspace, mava_stream_t s);

/* Gradients of the recurrent actor and critic losses for one minibatch (rec_mappo.py:208-312),
 * averaged over the update_batch_size replicas on this GPU.  The batch is the rollout reshaped
 * exactly like the reference does it (rec_mappo.py:339-349): (T, E) -> (chunk, E*num_chunks), so
 * column col = c*E + e holds env e at times t = l*num_chunks + c, l = 0..chunk-1, and starts from
 * the hidden state stored at time c.  cols[mb_cols] are the (shuffled) columns of this minibatch,
 * shared by all replicas.  Rollout buffers have a leading [T] axis over NE = num_replicas *
 * envs_per_replica envs; done_in[T][NE] is the flag ENTERING each step; hs_* are the hidden
 * states entering steps 0..num_chunks-1: [num_chunks][NE*rows_per_env][H].
 * grad_out: [actor grads | critic grads | total_actor, actor_loss, entropy, total_critic,
 * value_loss | pad]. */
int64_t mava_rec_ppo_workspace_bytes(const mava_rnn_desc* actor_host,
                                     const mava_rnn_desc* critic_host, int seq_envs_total,
                                     int chunk);
int mava_rec_ppo_loss_grad(const mava_rnn_desc* actor_host, const float* actor_params,
                           const mava_rnn_desc* critic_host, const float* critic_params,
                           const mava_ppo_hyper* hyper_host, const int8_t* view,
                           const float* obs_actor, const float* obs_critic, const void* mask,
                           const int8_t* action, const float* old_logp, const float* old_value,
                           const float* adv, const float* targets, const uint8_t* done_in,
                           const float* hs_actor, const float* hs_critic, const int32_t* cols,
                           int num_replicas, int envs_per_replica, int mb_cols, int chunk,
                           int num_chunks, float* grad_out, void* workspace, mava_stream_t s);

/* ------------------------------------------------------------------------------------------
 * Synthetic SMAX-shaped step source (benchmark only; BASELINE.json configs[3]).  SMAX's dynamics
 * live in jaxmarl, which is neither under the reference tree nor installable here, so the
 * rec_mappo workload is driven by device-generated tensors of the SMAX 3s5z shape: f32 per-agent
 * observations [NE][A][obs_dim] ~ U(0,1), one world-state row per env [NE][state_dim], uint16
 * action masks (bits ~ Bernoulli(0.7), actions 0..4 always legal), team reward ~ N(0, reward_std),
 * done ~ Bernoulli(done_prob), with RecordEpisodeMetrics applied for real.  `key` (2 words, device)
 * seeds the step; the actions are accepted and ignored.  state: uint8 [NE][16].
 * ---------------------------------------------------------------------------------------- */
typedef struct mava_synth_config {
  int32_t num_agents, obs_dim, state_dim, num_actions;
  float done_prob, reward_std;
} mava_synth_config;

int mava_synth_reset(const mava_synth_config* cfg_host, const uint32_t* key, uint8_t* state,
                     float* obs_actor, float* obs_critic, uint16_t* mask, int num_envs,
                     mava_stream_t s);
int mava_synth_step(const mava_synth_config* cfg_host, const uint32_t* key, uint8_t* state,
                    const int8_t* action, float* obs_actor, float* obs_critic, uint16_t* mask,
                    float* reward, uint8_t* done, float* ep_return, int32_t* ep_length,
                    int num_envs, mava_stream_t s);

/* ------------------------------------------------------------------------------------------
 * Diagnostics.  One 128 x N x K bf16 GEMM on the tcgen05 tensor cores in each operand arrangement
 * the fused MLP kernels use (0: X W, 1: dZ W^T, 2: H^T dZ); A, B, D are row-major f32.
 * ---------------------------------------------------------------------------------------- */
int mava_tc_selftest(int mode, const float* A, const float* B, float* D, int N, int K,
                     mava_stream_t s);
/* The dense contraction the recurrent path is built from, exposed for tests:
 *   C[M][N] (mode 0: =, 1: +=, 2: atomic +=) opA(A)[M][K] opB(B)[K][N] (+ bias[N]) (relu)
 *   (zeroed where relu_ref[M][ldr] <= 0);  ta: A is stored [K][lda];  tb: B is stored [N][ldb].
 * use_tc = 0: fp32 SIMT kernel; 1: bf16 tcgen05 kernel.  k_splits > 1 splits K over CTAs (mode 2). */
int mava_gemm(int use_tc, const float* A, int ta, int64_t lda, const float* B, int tb, int64_t ldb,
              float* C, int64_t ldc, int M, int N, int K, const float* bias, int relu,
              const float* relu_ref, int64_t ldr, int mode, int k_splits, mava_stream_t s);

#ifdef __cplusplus
}
#endif
#endif /* MAVA_B200_H_ */
