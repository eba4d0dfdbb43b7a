// XLA FFI custom-call handlers over the C ABI of libmava_b200.so (include/mava_b200.h): what
// BASELINE.json's "host code stays Python and calls hand-written CUDA through jax.ffi" binds.
//
// One handler per XLA-compiled region of mava/systems/ppo/ff_mappo.py that the library replaces
// (the table in INTEGRATION.md).  Each handler is a few lines: the ABI already takes raw device
// pointers and a stream.  Conventions:
//   * opaque handles (the env, the peer group) arrive as int64 ATTRIBUTES holding the pointer the
//     plain C calls mava_env_create / mava_peer_alloc returned to Python -- no process-global state;
//   * network descriptors arrive as int32[8] array attributes (the fields of mava_mlp_desc);
//   * buffers the kernels update in place (env state, parameters, optimiser moments, ...) are
//     donated: the Python side passes input_output_aliases, so operand and result share memory and
//     the handler only uses the result pointer;
//   * workspaces are extra results (XLA allocates them, the caller drops them).
//
// Build (needs jaxlib's headers, which are NOT in this image -- `python -c "import jax.ffi;
// print(jax.ffi.include_dir())"` on a machine that has jax):
//   g++ -std=c++17 -shared -fPIC -I$JAX_FFI_INCLUDE -Iinclude -I/usr/local/cuda/include
//       mava_b200/csrc/xla_ffi_shim.cc -Lmava_b200 -lmava_b200 -o mava_b200/libmava_b200_xla.so
// Without the header this file compiles to an empty translation unit (mava_b200/build.py skips it);
// tests/test_host_cpu.py type-checks it against the C header with a small stand-in of the FFI API
// (tests/ffi_stub/), which proves the calls into the C ABI are well formed, not that XLA accepts it.
#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#define MAVA_HAVE_XLA_FFI 1
#endif
#endif

#ifdef MAVA_HAVE_XLA_FFI
#include <cstdint>

#include "mava_b200.h"
#include "xla/ffi/api/ffi.h"

#ifndef MAVA_FFI_STREAM_T
#include <cuda_runtime_api.h>
#define MAVA_FFI_STREAM_T cudaStream_t
#endif

namespace ffi = xla::ffi;
using Stream = MAVA_FFI_STREAM_T;
template <ffi::DataType T>
using In = ffi::Buffer<T>;
template <ffi::DataType T>
using Out = ffi::ResultBuffer<T>;
using I32s = ffi::Span<const int32_t>;

namespace {

ffi::Error status(int rc) {
  return rc == 0 ? ffi::Error::Success() : ffi::Error::Internal(mava_error_string(rc));
}

bool desc_from(I32s a, mava_mlp_desc* d) {
  if (a.size() != 8) return false;
  d->input_mode = a[0]; d->add_agent_id = a[1]; d->num_agents = a[2]; d->view_dim = a[3];
  d->in_dim = a[4]; d->h1 = a[5]; d->h2 = a[6]; d->out_dim = a[7];
  return true;
}

mava_env_t env_of(int64_t handle) { return reinterpret_cast<mava_env_t>(static_cast<intptr_t>(handle)); }

// jax.vmap(env.reset) through the wrapper stack (ff_mappo.py:395)
ffi::Error EnvReset(Stream stream, In<ffi::U32> keys, Out<ffi::U8> state, Out<ffi::S8> view,
                    Out<ffi::U8> mask, int64_t env) {
  const int num_envs = static_cast<int>(keys.element_count() / 2);
  return status(mava_env_reset(env_of(env), keys.typed_data(), state->typed_data(),
                               view->typed_data(), mask->typed_data(), num_envs, stream));
}

// jax.vmap(env.step) (ff_mappo.py:88); the state is donated (input_output_aliases={0: 0})
ffi::Error EnvStep(Stream stream, In<ffi::U8> /*state*/, In<ffi::S8> action, Out<ffi::U8> state,
                   Out<ffi::S8> view, Out<ffi::U8> mask, Out<ffi::F32> reward, Out<ffi::U8> done,
                   Out<ffi::F32> ep_return, Out<ffi::S32> ep_length, int64_t env,
                   int32_t auto_reset) {
  const int num_envs = static_cast<int>(done->element_count());
  return status(mava_env_step(env_of(env), state->typed_data(), action.typed_data(),
                              view->typed_data(), mask->typed_data(), reward->typed_data(),
                              done->typed_data(), ep_return->typed_data(), ep_length->typed_data(),
                              num_envs, auto_reset, stream));
}

// actor_apply_fn + critic_apply_fn + sample + log_prob (ff_mappo.py:82-85), bf16 tensor-core path
ffi::Error FfAct(Stream stream, In<ffi::F32> actor_params, In<ffi::U8> actor_image,
                 In<ffi::F32> critic_params, In<ffi::U8> critic_image, In<ffi::S8> view,
                 In<ffi::U8> mask, In<ffi::U32> policy_key, Out<ffi::S8> action,
                 Out<ffi::F32> logp, Out<ffi::F32> value, I32s actor_desc, I32s critic_desc,
                 int32_t envs_per_replica, int32_t greedy) {
  mava_mlp_desc a, c;
  if (!desc_from(actor_desc, &a) || !desc_from(critic_desc, &c))
    return ffi::Error::InvalidArgument("actor_desc / critic_desc must hold 8 int32");
  const int num_envs = static_cast<int>(mask.element_count() / a.num_agents);
  return status(mava_ff_act_bf16(&a, actor_params.typed_data(), actor_image.typed_data(), &c,
                                 critic_params.typed_data(), critic_image.typed_data(),
                                 view.typed_data(), mask.typed_data(), policy_key.typed_data(),
                                 envs_per_replica, num_envs, greedy, nullptr, action->typed_data(),
                                 logp->typed_data(), value->typed_data(), stream));
}

// the whole jax.lax.scan(_env_step, ..., rollout_length) (ff_mappo.py:104-106).  view / mask are the
// [T+1] stacks whose slot 0 holds the observation the rollout starts from (donated, aliases 1, 2).
ffi::Error FfRollout(Stream stream, In<ffi::U8> /*state*/, In<ffi::S8> /*view*/, In<ffi::U8> /*mask*/,
                     In<ffi::F32> actor_params, In<ffi::U8> actor_image, In<ffi::U32> policy_keys,
                     Out<ffi::U8> state, Out<ffi::S8> view, Out<ffi::U8> mask, Out<ffi::S8> action,
                     Out<ffi::F32> logp, Out<ffi::F32> reward, Out<ffi::U8> done,
                     Out<ffi::F32> ep_return, Out<ffi::S32> ep_length, int64_t env, I32s actor_desc,
                     int32_t envs_per_replica) {
  mava_mlp_desc a;
  if (!desc_from(actor_desc, &a)) return ffi::Error::InvalidArgument("actor_desc must hold 8 int32");
  const int T = static_cast<int>(policy_keys.element_count() / 2);
  const int num_envs = static_cast<int>(done->element_count() / T);
  return status(mava_ff_rollout_bf16(env_of(env), &a, actor_params.typed_data(),
                                     actor_image.typed_data(), state->typed_data(),
                                     view->typed_data(), mask->typed_data(), policy_keys.typed_data(),
                                     envs_per_replica, num_envs, T, action->typed_data(),
                                     logp->typed_data(), reward->typed_data(), done->typed_data(),
                                     ep_return->typed_data(), ep_length->typed_data(), stream));
}

// _calculate_gae (ff_mappo.py:112-139; rec flavour rec_mappo.py:177-199 when last_done is given)
ffi::Error Gae(Stream stream, In<ffi::F32> reward, In<ffi::F32> value, In<ffi::U8> done,
               In<ffi::F32> last_val, In<ffi::U8> last_done, Out<ffi::F32> adv,
               Out<ffi::F32> targets, float gamma, float gae_lambda, int32_t rec) {
  const auto dims = reward.dimensions();  // [T, NE, A]
  if (dims.size() != 3) return ffi::Error::InvalidArgument("reward must be [T, NE, A]");
  return status(mava_gae(reward.typed_data(), value.typed_data(), done.typed_data(),
                         last_val.typed_data(), rec ? last_done.typed_data() : nullptr, gamma,
                         gae_lambda, static_cast<int>(dims[0]), static_cast<int>(dims[1]),
                         static_cast<int>(dims[2]), rec, adv->typed_data(), targets->typed_data(),
                         stream));
}

// both value_and_grad + pmean("batch") on one minibatch (ff_mappo.py:150-234).  rows = the
// minibatch's env-step indices (mava_ppo_minibatch_rows of the permutation); grad = [actor | critic |
// 8 loss scalars], the buffer pmean("device") reduces; workspace = mava_ppo_workspace_bytes_bf16.
ffi::Error PpoLossGrad(Stream stream, In<ffi::F32> actor_params, In<ffi::U8> actor_image,
                       In<ffi::F32> critic_params, In<ffi::U8> critic_image, In<ffi::S8> view,
                       In<ffi::U8> mask, In<ffi::S8> action, In<ffi::F32> old_logp,
                       In<ffi::F32> old_value, In<ffi::F32> adv, In<ffi::F32> targets,
                       In<ffi::S32> rows, Out<ffi::F32> grad, Out<ffi::U8> workspace,
                       I32s actor_desc, I32s critic_desc, int32_t num_replicas, float clip_eps,
                       float ent_coef, float vf_coef) {
  mava_mlp_desc a, c;
  if (!desc_from(actor_desc, &a) || !desc_from(critic_desc, &c))
    return ffi::Error::InvalidArgument("actor_desc / critic_desc must hold 8 int32");
  const mava_ppo_hyper hyper{clip_eps, ent_coef, vf_coef};
  const int mb_size = static_cast<int>(rows.element_count() / num_replicas);
  return status(mava_ppo_loss_grad_bf16(&a, actor_params.typed_data(), actor_image.typed_data(), &c,
                                        critic_params.typed_data(), critic_image.typed_data(), &hyper,
                                        view.typed_data(), mask.typed_data(), action.typed_data(),
                                        old_logp.typed_data(), old_value.typed_data(),
                                        adv.typed_data(), targets.typed_data(), rows.typed_data(),
                                        num_replicas, mb_size, grad->typed_data(),
                                        workspace->typed_data(), stream));
}

// pmean("device") + optax update + apply_updates + bf16 image refresh + loss metrics
// (ff_mappo.py:228-250,260-265).  params / mu / nu / counts / images are donated (aliases 0..5);
// `group` = address of a host mava_peer_group whose buffers the loss kernel wrote into.
ffi::Error ReduceClipAdam(Stream stream, In<ffi::F32> /*params*/, In<ffi::F32> /*mu*/,
                          In<ffi::F32> /*nu*/, In<ffi::S32> /*counts*/, In<ffi::U8> /*actor_image*/,
                          In<ffi::U8> /*critic_image*/, Out<ffi::F32> params, Out<ffi::F32> mu,
                          Out<ffi::F32> nu, Out<ffi::S32> counts, Out<ffi::U8> actor_image,
                          Out<ffi::U8> critic_image, Out<ffi::F32> gsum, Out<ffi::F32> loss5,
                          int64_t group, I32s actor_desc, I32s critic_desc, float grad_scale,
                          float lr_actor, float lr_critic, float max_norm,
                          int32_t lr_decay_num_updates, int32_t steps_per_update) {
  mava_mlp_desc a, c;
  if (!desc_from(actor_desc, &a) || !desc_from(critic_desc, &c))
    return ffi::Error::InvalidArgument("actor_desc / critic_desc must hold 8 int32");
  const auto* g = reinterpret_cast<const mava_peer_group*>(static_cast<intptr_t>(group));
  return status(mava_reduce_clip_adam_pair(
      params->typed_data(), mu->typed_data(), nu->typed_data(), counts->typed_data(), g,
      gsum->typed_data(), mava_mlp_param_count(&a), mava_mlp_param_count(&c), &a,
      actor_image->typed_data(), &c, critic_image->typed_data(), grad_scale, lr_actor, lr_critic,
      max_norm, lr_decay_num_updates, steps_per_update, loss5->typed_data(), stream));
}

}  // namespace

#define MAVA_STREAM Ctx<ffi::PlatformStream<Stream>>()

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    MavaEnvReset, EnvReset,
    ffi::Ffi::Bind().MAVA_STREAM.Arg<In<ffi::U32>>().Ret<In<ffi::U8>>().Ret<In<ffi::S8>>()
        .Ret<In<ffi::U8>>().Attr<int64_t>("env"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    MavaEnvStep, EnvStep,
    ffi::Ffi::Bind().MAVA_STREAM.Arg<In<ffi::U8>>().Arg<In<ffi::S8>>().Ret<In<ffi::U8>>()
        .Ret<In<ffi::S8>>().Ret<In<ffi::U8>>().Ret<In<ffi::F32>>().Ret<In<ffi::U8>>()
        .Ret<In<ffi::F32>>().Ret<In<ffi::S32>>().Attr<int64_t>("env").Attr<int32_t>("auto_reset"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    MavaFfAct, FfAct,
    ffi::Ffi::Bind().MAVA_STREAM.Arg<In<ffi::F32>>().Arg<In<ffi::U8>>().Arg<In<ffi::F32>>()
        .Arg<In<ffi::U8>>().Arg<In<ffi::S8>>().Arg<In<ffi::U8>>().Arg<In<ffi::U32>>()
        .Ret<In<ffi::S8>>().Ret<In<ffi::F32>>().Ret<In<ffi::F32>>().Attr<I32s>("actor_desc")
        .Attr<I32s>("critic_desc").Attr<int32_t>("envs_per_replica").Attr<int32_t>("greedy"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    MavaFfRollout, FfRollout,
    ffi::Ffi::Bind().MAVA_STREAM.Arg<In<ffi::U8>>().Arg<In<ffi::S8>>().Arg<In<ffi::U8>>()
        .Arg<In<ffi::F32>>().Arg<In<ffi::U8>>().Arg<In<ffi::U32>>().Ret<In<ffi::U8>>()
        .Ret<In<ffi::S8>>().Ret<In<ffi::U8>>().Ret<In<ffi::S8>>().Ret<In<ffi::F32>>()
        .Ret<In<ffi::F32>>().Ret<In<ffi::U8>>().Ret<In<ffi::F32>>().Ret<In<ffi::S32>>()
        .Attr<int64_t>("env").Attr<I32s>("actor_desc").Attr<int32_t>("envs_per_replica"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    MavaGae, Gae,
    ffi::Ffi::Bind().MAVA_STREAM.Arg<In<ffi::F32>>().Arg<In<ffi::F32>>().Arg<In<ffi::U8>>()
        .Arg<In<ffi::F32>>().Arg<In<ffi::U8>>().Ret<In<ffi::F32>>().Ret<In<ffi::F32>>()
        .Attr<float>("gamma").Attr<float>("gae_lambda").Attr<int32_t>("rec"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    MavaPpoLossGrad, PpoLossGrad,
    ffi::Ffi::Bind().MAVA_STREAM.Arg<In<ffi::F32>>().Arg<In<ffi::U8>>().Arg<In<ffi::F32>>()
        .Arg<In<ffi::U8>>().Arg<In<ffi::S8>>().Arg<In<ffi::U8>>().Arg<In<ffi::S8>>()
        .Arg<In<ffi::F32>>().Arg<In<ffi::F32>>().Arg<In<ffi::F32>>().Arg<In<ffi::F32>>()
        .Arg<In<ffi::S32>>().Ret<In<ffi::F32>>().Ret<In<ffi::U8>>().Attr<I32s>("actor_desc")
        .Attr<I32s>("critic_desc").Attr<int32_t>("num_replicas").Attr<float>("clip_eps")
        .Attr<float>("ent_coef").Attr<float>("vf_coef"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    MavaReduceClipAdam, ReduceClipAdam,
    ffi::Ffi::Bind().MAVA_STREAM.Arg<In<ffi::F32>>().Arg<In<ffi::F32>>().Arg<In<ffi::F32>>()
        .Arg<In<ffi::S32>>().Arg<In<ffi::U8>>().Arg<In<ffi::U8>>().Ret<In<ffi::F32>>()
        .Ret<In<ffi::F32>>().Ret<In<ffi::F32>>().Ret<In<ffi::S32>>().Ret<In<ffi::U8>>()
        .Ret<In<ffi::U8>>().Ret<In<ffi::F32>>().Ret<In<ffi::F32>>().Attr<int64_t>("group")
        .Attr<I32s>("actor_desc").Attr<I32s>("critic_desc").Attr<float>("grad_scale")
        .Attr<float>("lr_actor").Attr<float>("lr_critic").Attr<float>("max_norm")
        .Attr<int32_t>("lr_decay_num_updates").Attr<int32_t>("steps_per_update"));

#endif  // MAVA_HAVE_XLA_FFI
