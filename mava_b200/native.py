"""Torch-tensor level calls into libmava_b200.so.

torch is used only as the owner of device memory and of the CUDA stream: every function here
validates its tensors (device, dtype, contiguity, size) and forwards raw device pointers and the
current stream to the C ABI in include/mava_b200.h.  Nothing here computes.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import (EnvDims, LbfConfig, MlpDesc, PpoHyper, RnnDesc, RwareConfig, SynthConfig,
                   check)

ENV_RWARE, ENV_LBF = 1, 2
# Number of mava_b200 kernels launched through this module (the benchmark's `gpu_launches`).
LAUNCHES = 0


def _count(n: int) -> None:
    global LAUNCHES
    LAUNCHES += n

IN_AGENT_VIEW, IN_GLOBAL, IN_DENSE = 0, 1, 2
OMAX = 16


def _p(t: Optional[torch.Tensor], dtype: torch.dtype, numel: Optional[int] = None, name: str = ""):
    if t is None:
        return None
    if not t.is_cuda:
        raise ValueError(f"{name}: expected a CUDA tensor (mava_b200 has no CPU path)")
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name}: tensor must be contiguous")
    if numel is not None and t.numel() < numel:
        raise ValueError(f"{name}: needs {numel} elements, has {t.numel()}")
    return C.c_void_p(t.data_ptr())


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def device_info():
    out = (C.c_int * 3)()
    check(_lib.load().mava_device_info(out), "mava_device_info")
    return tuple(out)


def mlp_desc(input_mode: int, add_agent_id: bool, num_agents: int, view_dim: int, h1: int, h2: int,
             out_dim: int) -> MlpDesc:
    in_dim = num_agents * view_dim if input_mode == IN_GLOBAL else view_dim + (
        num_agents if add_agent_id else 0)
    return MlpDesc(input_mode, int(add_agent_id), num_agents, view_dim, in_dim, h1, h2, out_dim)


def mlp_param_count(d: MlpDesc) -> int:
    return int(_lib.load().mava_mlp_param_count(C.byref(d)))


class Env:
    """Opaque native env handle + its dimensions."""

    def __init__(self, kind: int, cfg):
        lib = _lib.load()
        self._h = C.c_void_p()
        check(lib.mava_env_create(kind, C.byref(cfg), C.sizeof(cfg), C.byref(self._h)),
              "mava_env_create")
        self.dims = EnvDims()
        check(lib.mava_env_dims_of(self._h, C.byref(self.dims)), "mava_env_dims_of")
        self.kind = kind

    @classmethod
    def rware(cls, column_height=8, shelf_rows=1, shelf_columns=3, num_agents=4, sensor_range=1,
              request_queue_size=4, time_limit=500) -> "Env":
        return cls(ENV_RWARE, RwareConfig(column_height, shelf_rows, shelf_columns, num_agents,
                                          sensor_range, request_queue_size, time_limit))

    @classmethod
    def lbf(cls, grid_size=8, fov=8, num_agents=2, num_food=2, max_agent_level=2, force_coop=True,
            time_limit=100, use_individual_rewards=False) -> "Env":
        return cls(ENV_LBF, LbfConfig(grid_size, fov, num_agents, num_food, max_agent_level,
                                      int(force_coop), time_limit, int(use_individual_rewards)))

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                _lib.load().mava_env_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # -- shapes -------------------------------------------------------------------------------
    @property
    def num_agents(self) -> int:
        return self.dims.num_agents

    @property
    def view_dim(self) -> int:
        return self.dims.view_dim

    @property
    def num_actions(self) -> int:
        return self.dims.num_actions

    @property
    def state_stride(self) -> int:
        return self.dims.state_stride

    def alloc_state(self, num_envs: int, device) -> torch.Tensor:
        return torch.zeros(num_envs, self.state_stride, dtype=torch.uint8, device=device)

    # -- calls --------------------------------------------------------------------------------
    def reset(self, keys, state, view, mask, num_envs: int) -> None:
        A, F = self.num_agents, self.view_dim
        _count(1)
        check(_lib.load().mava_env_reset(
            self._h, _p(keys, torch.uint32, 2 * num_envs, "keys"),
            _p(state, torch.uint8, num_envs * self.state_stride, "state"),
            _p(view, torch.int8, num_envs * A * F, "view"),
            _p(mask, torch.uint8, num_envs * A, "mask"), num_envs, _stream()), "mava_env_reset")

    def step(self, state, action, view, mask, reward, done, ep_return, ep_length, num_envs: int,
             auto_reset: bool = True) -> None:
        A, F = self.num_agents, self.view_dim
        _count(1)
        check(_lib.load().mava_env_step(
            self._h, _p(state, torch.uint8, num_envs * self.state_stride, "state"),
            _p(action, torch.int8, num_envs * A, "action"),
            _p(view, torch.int8, num_envs * A * F, "view"),
            _p(mask, torch.uint8, num_envs * A, "mask"),
            _p(reward, torch.float32, num_envs * A, "reward"),
            _p(done, torch.uint8, num_envs, "done"),
            _p(ep_return, torch.float32, num_envs, "ep_return"),
            _p(ep_length, torch.int32, num_envs, "ep_length"), num_envs, int(auto_reset),
            _stream()), "mava_env_step")

    def peek(self, state, field: int, num_envs: int) -> torch.Tensor:
        A = self.num_agents
        last = self.dims.aux0 if self.kind == ENV_LBF else self.dims.aux1  # eaten flags | queue
        width = {0: 1, 1: 2, 2: 4 * A, 3: 3 * self.dims.aux0, 4: last}[field]
        out = torch.zeros(num_envs, width, dtype=torch.int32, device=state.device)
        check(_lib.load().mava_env_peek(self._h, _p(state, torch.uint8, None, "state"), field,
                                        _p(out, torch.int32), num_envs, _stream()), "mava_env_peek")
        return out


def prng_split_chain(key_io: torch.Tensor, subkeys: torch.Tensor, n: int) -> None:
    _count(1)
    check(_lib.load().mava_prng_split_chain(_p(key_io, torch.uint32, 2, "key"),
                                            _p(subkeys, torch.uint32, 2 * n, "subkeys"), n,
                                            _stream()), "mava_prng_split_chain")


def prng_split(key: torch.Tensor, out: torch.Tensor, num: int) -> None:
    _count(1)
    check(_lib.load().mava_prng_split(_p(key, torch.uint32, 2, "key"),
                                      _p(out, torch.uint32, 2 * num, "out"), num, _stream()),
          "mava_prng_split")


def prng_random_bits(key: torch.Tensor, out: torch.Tensor, n: int) -> None:
    _count(1)
    check(_lib.load().mava_prng_random_bits(_p(key, torch.uint32, 2, "key"),
                                            _p(out, torch.uint32, n, "out"), n, _stream()),
          "mava_prng_random_bits")


def sort_workspace_bytes(n: int) -> int:
    return int(_lib.load().mava_sort_workspace_bytes(n))


def sort_by_key(keys, val_in, val_out, n: int, workspace, overflow) -> None:
    """val_out = val_in stably sorted by uint32 keys (one round of jax.random.permutation)."""
    _count(4)
    check(_lib.load().mava_sort_by_key(
        _p(keys, torch.uint32, n, "keys"), _p(val_in, torch.int32, n, "val_in"),
        _p(val_out, torch.int32, n, "val_out"), n,
        _p(workspace, torch.uint8, sort_workspace_bytes(n), "workspace"),
        _p(overflow, torch.int32, 1, "overflow"), _stream()), "mava_sort_by_key")


def ff_act(actor: MlpDesc, actor_params, critic: Optional[MlpDesc], critic_params, view, mask,
           policy_key, envs_per_replica: int, num_envs: int, action, logp, value=None,
           greedy: bool = False, actions_in=None) -> None:
    A = actor.num_agents
    _count(2 if value is not None else 1)
    check(_lib.load().mava_ff_act(
        C.byref(actor), _p(actor_params, torch.float32, mlp_param_count(actor), "actor_params"),
        C.byref(critic) if critic is not None else None,
        _p(critic_params, torch.float32, mlp_param_count(critic) if critic is not None else None,
           "critic_params"),
        _p(view, torch.int8, num_envs * A * actor.view_dim, "view"),
        _p(mask, torch.uint8, num_envs * A, "mask"), _p(policy_key, torch.uint32, 2, "policy_key"),
        envs_per_replica, num_envs, int(greedy), _p(actions_in, torch.int8, num_envs * A, "actions_in"),
        _p(action, torch.int8, num_envs * A, "action"), _p(logp, torch.float32, num_envs * A, "logp"),
        _p(value, torch.float32, num_envs * A, "value"), _stream()), "mava_ff_act")


def ff_value(critic: MlpDesc, critic_params, view, num_envs: int, value) -> None:
    A = critic.num_agents
    _count(1)
    check(_lib.load().mava_ff_value(
        C.byref(critic), _p(critic_params, torch.float32, mlp_param_count(critic), "critic_params"),
        _p(view, torch.int8, num_envs * A * critic.view_dim, "view"), num_envs,
        _p(value, torch.float32, num_envs * A, "value"), _stream()), "mava_ff_value")


def gae(reward, value, done, last_val, gamma: float, gae_lambda: float, T: int, num_envs: int,
        num_agents: int, adv, targets, last_done=None) -> None:
    n = T * num_envs * num_agents
    _count(1)
    check(_lib.load().mava_gae(
        _p(reward, torch.float32, n, "reward"), _p(value, torch.float32, n, "value"),
        _p(done, torch.uint8, T * num_envs, "done"),
        _p(last_val, torch.float32, num_envs * num_agents, "last_val"),
        _p(last_done, torch.uint8, num_envs, "last_done"), gamma, gae_lambda, T, num_envs,
        num_agents, int(last_done is not None), _p(adv, torch.float32, n, "adv"),
        _p(targets, torch.float32, n, "targets"), _stream()), "mava_gae")


def episode_stats(done, ep_return, ep_length, n: int, reset: bool, stats) -> None:
    """Accumulate the finished-episode statistics of n env-steps into stats[10] (float64)."""
    _count(1 + int(reset))
    check(_lib.load().mava_episode_stats(
        _p(done, torch.uint8, n, "done"), _p(ep_return, torch.float32, n, "ep_return"),
        _p(ep_length, torch.int32, n, "ep_length"), n, int(reset),
        _p(stats, torch.float64, 10, "stats"), _stream()), "mava_episode_stats")


def ppo_minibatch_rows(perm, mb_index: int, mb_size: int, num_replicas: int, envs_per_replica: int,
                       rows) -> None:
    _count(1)
    check(_lib.load().mava_ppo_minibatch_rows(
        _p(perm, torch.int32, (mb_index + 1) * mb_size, "perm"), mb_index, mb_size, num_replicas,
        envs_per_replica, _p(rows, torch.int32, num_replicas * mb_size, "rows"), _stream()),
        "mava_ppo_minibatch_rows")


def ppo_workspace_bytes(actor: MlpDesc, critic: MlpDesc, rows_total: int) -> int:
    return int(_lib.load().mava_ppo_workspace_bytes(C.byref(actor), C.byref(critic), rows_total))


def ppo_loss_grad(actor: MlpDesc, actor_params, critic: MlpDesc, critic_params, hyper: PpoHyper,
                  view, mask, action, old_logp, old_value, adv, targets, rows, num_replicas: int,
                  mb_size: int, grad_out, workspace, precision: str = "auto") -> None:
    if precision not in ("auto", "fp32"):
        raise ValueError(f"precision '{precision}' is not available in this build (fp32 only)")
    na, nc = mlp_param_count(actor), mlp_param_count(critic)
    need = ppo_workspace_bytes(actor, critic, num_replicas * mb_size)
    _count(14)
    check(_lib.load().mava_ppo_loss_grad(
        C.byref(actor), _p(actor_params, torch.float32, na, "actor_params"), C.byref(critic),
        _p(critic_params, torch.float32, nc, "critic_params"), C.byref(hyper),
        _p(view, torch.int8, None, "view"), _p(mask, torch.uint8, None, "mask"),
        _p(action, torch.int8, None, "action"), _p(old_logp, torch.float32, None, "old_logp"),
        _p(old_value, torch.float32, None, "old_value"), _p(adv, torch.float32, None, "adv"),
        _p(targets, torch.float32, None, "targets"),
        _p(rows, torch.int32, num_replicas * mb_size, "rows"), num_replicas, mb_size,
        _p(grad_out, torch.float32, na + nc + 8, "grad_out"),
        _p(workspace, torch.uint8, need, "workspace"), _stream()), "mava_ppo_loss_grad")


def clip_adam(params, mu, nu, count, grad, n: int, grad_scale: float, lr: float, max_norm: float,
              lr_decay_num_updates: int = 0, steps_per_update: int = 1) -> None:
    _count(1)
    check(_lib.load().mava_clip_adam(
        _p(params, torch.float32, n, "params"), _p(mu, torch.float32, n, "mu"),
        _p(nu, torch.float32, n, "nu"), _p(count, torch.int32, 1, "count"),
        _p(grad, torch.float32, n, "grad"), n, grad_scale, lr, max_norm, lr_decay_num_updates,
        steps_per_update, _stream()), "mava_clip_adam")


def tc_selftest(mode: int, A, B, D, N: int, K: int) -> None:
    _count(1)
    check(_lib.load().mava_tc_selftest(mode, _p(A, torch.float32, None, "A"),
                                       _p(B, torch.float32, None, "B"),
                                       _p(D, torch.float32, 128 * N, "D"), N, K, _stream()),
          "mava_tc_selftest")


def mlp_pack_bytes(d: MlpDesc) -> int:
    n = int(_lib.load().mava_mlp_pack_bytes(C.byref(d)))
    if n < 0:
        raise _lib.MavaNativeError("mava_mlp_pack_bytes: network not supported by the bf16 path "
                                   "(needs two hidden layers of width 128)")
    return n


def mlp_pack_bf16(d: MlpDesc, params, image) -> None:
    _count(1)
    check(_lib.load().mava_mlp_pack_bf16(
        C.byref(d), _p(params, torch.float32, mlp_param_count(d), "params"),
        _p(image, torch.uint8, mlp_pack_bytes(d), "image"), _stream()), "mava_mlp_pack_bf16")


def ff_act_bf16(actor: Optional[MlpDesc], actor_params, actor_image, critic: Optional[MlpDesc],
                critic_params, critic_image, view, mask, policy_key, envs_per_replica: int,
                num_envs: int, action, logp, value=None, greedy: bool = False,
                actions_in=None) -> None:
    """actor=None evaluates the critic only (the bootstrap value)."""
    ref = actor if actor is not None else critic
    A = ref.num_agents
    _count(1)
    check(_lib.load().mava_ff_act_bf16(
        C.byref(actor) if actor is not None else None,
        _p(actor_params, torch.float32, mlp_param_count(actor) if actor is not None else None,
           "actor_params"),
        _p(actor_image, torch.uint8, mlp_pack_bytes(actor) if actor is not None else None,
           "actor_image"),
        C.byref(critic) if critic is not None else None,
        _p(critic_params, torch.float32, mlp_param_count(critic) if critic is not None else None,
           "critic_params"),
        _p(critic_image, torch.uint8, mlp_pack_bytes(critic) if critic is not None else None,
           "critic_image"),
        _p(view, torch.int8, num_envs * A * ref.view_dim, "view"),
        _p(mask, torch.uint8, num_envs * A, "mask"), _p(policy_key, torch.uint32, 2, "policy_key"),
        envs_per_replica, num_envs, int(greedy), _p(actions_in, torch.int8, num_envs * A, "actions_in"),
        _p(action, torch.int8, num_envs * A, "action"), _p(logp, torch.float32, num_envs * A, "logp"),
        _p(value, torch.float32, num_envs * A, "value"), _stream()), "mava_ff_act_bf16")


def ff_rollout_supported(env: "Env") -> bool:
    """Whether mava_ff_rollout_bf16 handles this env (RobotWarehouse, 2/4/8 agents, range 1)."""
    return (env.kind == ENV_RWARE and env.num_agents in (2, 4, 8) and env.view_dim == 66)


def ff_rollout_bf16(env: "Env", actor: MlpDesc, actor_params, actor_image, state, view, mask,
                    policy_keys, envs_per_replica: int, num_envs: int, T: int, action, logp, reward,
                    done, ep_return, ep_length) -> None:
    """The whole rollout scan in one persistent kernel (see include/mava_b200.h)."""
    A, F = actor.num_agents, actor.view_dim
    _count(1)
    check(_lib.load().mava_ff_rollout_bf16(
        env._h, C.byref(actor), _p(actor_params, torch.float32, mlp_param_count(actor), "actor_params"),
        _p(actor_image, torch.uint8, mlp_pack_bytes(actor), "actor_image"),
        _p(state, torch.uint8, num_envs * env.state_stride, "state"),
        _p(view, torch.int8, (T + 1) * num_envs * A * F, "view"),
        _p(mask, torch.uint8, (T + 1) * num_envs * A, "mask"),
        _p(policy_keys, torch.uint32, 2 * T, "policy_keys"), envs_per_replica, num_envs, T,
        _p(action, torch.int8, T * num_envs * A, "action"),
        _p(logp, torch.float32, T * num_envs * A, "logp"),
        _p(reward, torch.float32, T * num_envs * A, "reward"),
        _p(done, torch.uint8, T * num_envs, "done"),
        _p(ep_return, torch.float32, T * num_envs, "ep_return"),
        _p(ep_length, torch.int32, T * num_envs, "ep_length"), _stream()), "mava_ff_rollout_bf16")


def ff_rollout_bf16_ex(env: "Env", actor: MlpDesc, actor_params, actor_image, state, view, mask,
                       policy_keys, envs_per_replica: int, num_envs: int, T: int, auto_reset: bool,
                       greedy: bool, record: bool, action, logp, reward, done, ep_return,
                       ep_length) -> None:
    """The fused rollout kernel as the evaluator runs it (no auto-reset, optional pi.mode(), and --
    record=False -- per-step outputs other than done / episode metrics folded into slot 0)."""
    A, F = actor.num_agents, actor.view_dim
    slots, slots1 = (T, T + 1) if record else (1, 1)
    _count(1)
    check(_lib.load().mava_ff_rollout_bf16_ex(
        env._h, C.byref(actor), _p(actor_params, torch.float32, mlp_param_count(actor), "actor_params"),
        _p(actor_image, torch.uint8, mlp_pack_bytes(actor), "actor_image"),
        _p(state, torch.uint8, num_envs * env.state_stride, "state"),
        _p(view, torch.int8, slots1 * num_envs * A * F, "view"),
        _p(mask, torch.uint8, slots1 * num_envs * A, "mask"),
        _p(policy_keys, torch.uint32, 2 * T, "policy_keys"), envs_per_replica, num_envs, T,
        int(auto_reset), int(greedy), int(record),
        _p(action, torch.int8, slots * num_envs * A, "action"),
        _p(logp, torch.float32, slots * num_envs * A, "logp"),
        _p(reward, torch.float32, slots * num_envs * A, "reward"),
        _p(done, torch.uint8, T * num_envs, "done"),
        _p(ep_return, torch.float32, T * num_envs, "ep_return"),
        _p(ep_length, torch.int32, T * num_envs, "ep_length"), _stream()), "mava_ff_rollout_bf16_ex")


def episode_first_terminal(done, ep_return, ep_length, T: int, num_envs: int, out_return,
                           out_length) -> None:
    """Per env: episode return / length at the first terminal step (evaluator.py:143-150)."""
    _count(1)
    check(_lib.load().mava_episode_first_terminal(
        _p(done, torch.uint8, T * num_envs, "done"), _p(ep_return, torch.float32, T * num_envs, "ep_return"),
        _p(ep_length, torch.int32, T * num_envs, "ep_length"), T, num_envs,
        _p(out_return, torch.float32, num_envs, "out_return"),
        _p(out_length, torch.int32, num_envs, "out_length"), _stream()), "mava_episode_first_terminal")


def ppo_workspace_bytes_bf16(actor: MlpDesc, critic: MlpDesc, rows_total: int) -> int:
    return int(_lib.load().mava_ppo_workspace_bytes_bf16(C.byref(actor), C.byref(critic),
                                                         rows_total))


def ppo_loss_grad_bf16(actor: MlpDesc, actor_params, actor_image, critic: MlpDesc, critic_params,
                       critic_image, hyper: PpoHyper, view, mask, action, old_logp, old_value, adv,
                       targets, rows, num_replicas: int, mb_size: int, grad_out, workspace) -> None:
    na, nc = mlp_param_count(actor), mlp_param_count(critic)
    need = ppo_workspace_bytes_bf16(actor, critic, num_replicas * mb_size)
    _count(4)  # adv stats, fused fwd+bwd, first-layer wgrad, loss finalize
    check(_lib.load().mava_ppo_loss_grad_bf16(
        C.byref(actor), _p(actor_params, torch.float32, na, "actor_params"),
        _p(actor_image, torch.uint8, mlp_pack_bytes(actor), "actor_image"), C.byref(critic),
        _p(critic_params, torch.float32, nc, "critic_params"),
        _p(critic_image, torch.uint8, mlp_pack_bytes(critic), "critic_image"), C.byref(hyper),
        _p(view, torch.int8, None, "view"), _p(mask, torch.uint8, None, "mask"),
        _p(action, torch.int8, None, "action"), _p(old_logp, torch.float32, None, "old_logp"),
        _p(old_value, torch.float32, None, "old_value"), _p(adv, torch.float32, None, "adv"),
        _p(targets, torch.float32, None, "targets"),
        _p(rows, torch.int32, num_replicas * mb_size, "rows"), num_replicas, mb_size,
        _p(grad_out, torch.float32, na + nc + 8, "grad_out"),
        _p(workspace, torch.uint8, need, "workspace"), _stream()), "mava_ppo_loss_grad_bf16")


def ppo_adv_stats(adv, rows, num_replicas: int, mb_size: int, num_agents: int, stats) -> None:
    """Per-replica advantage sums of one minibatch (16 doubles), ahead of the loss kernels."""
    _count(1)
    check(_lib.load().mava_ppo_adv_stats(
        _p(adv, torch.float32, None, "adv"), _p(rows, torch.int32, num_replicas * mb_size, "rows"),
        num_replicas, mb_size, num_agents, _p(stats, torch.float64, 16, "stats"), _stream()),
        "mava_ppo_adv_stats")


def ppo_loss_grad_bf16_stats(actor: MlpDesc, actor_params, actor_image, critic: MlpDesc,
                             critic_params, critic_image, hyper: PpoHyper, view, mask, action,
                             old_logp, old_value, adv, targets, rows, num_replicas: int,
                             mb_size: int, adv_stats, grad_out, workspace) -> None:
    """ppo_loss_grad_bf16 with the advantage statistics computed beforehand (ppo_adv_stats)."""
    na, nc = mlp_param_count(actor), mlp_param_count(critic)
    need = ppo_workspace_bytes_bf16(actor, critic, num_replicas * mb_size)
    _count(3)  # fused fwd+bwd, first-layer wgrad, loss finalize
    check(_lib.load().mava_ppo_loss_grad_bf16_stats(
        C.byref(actor), _p(actor_params, torch.float32, na, "actor_params"),
        _p(actor_image, torch.uint8, mlp_pack_bytes(actor), "actor_image"), C.byref(critic),
        _p(critic_params, torch.float32, nc, "critic_params"),
        _p(critic_image, torch.uint8, mlp_pack_bytes(critic), "critic_image"), C.byref(hyper),
        _p(view, torch.int8, None, "view"), _p(mask, torch.uint8, None, "mask"),
        _p(action, torch.int8, None, "action"), _p(old_logp, torch.float32, None, "old_logp"),
        _p(old_value, torch.float32, None, "old_value"), _p(adv, torch.float32, None, "adv"),
        _p(targets, torch.float32, None, "targets"),
        _p(rows, torch.int32, num_replicas * mb_size, "rows"), num_replicas, mb_size,
        _p(adv_stats, torch.float64, 16, "adv_stats"),
        _p(grad_out, torch.float32, na + nc + 8, "grad_out"),
        _p(workspace, torch.uint8, need, "workspace"), _stream()), "mava_ppo_loss_grad_bf16_stats")


def ppo_loss_grad_bf16_acc(actor: MlpDesc, actor_params, actor_image, critic: MlpDesc,
                           critic_params, critic_image, hyper: PpoHyper, view, mask, action,
                           old_logp, old_value, adv, targets, rows, num_replicas: int,
                           mb_size: int, adv_stats, grad_out, workspace) -> None:
    """ppo_loss_grad_bf16_stats that ADDS into ``grad_out`` and the loss accumulators of its
    workspace (both left zero by ``reduce_clip_adam_pair_acc``): no memsets, no finalize launch."""
    na, nc = mlp_param_count(actor), mlp_param_count(critic)
    need = ppo_workspace_bytes_bf16(actor, critic, num_replicas * mb_size)
    _count(2)  # fused fwd+bwd, first-layer wgrad
    check(_lib.load().mava_ppo_loss_grad_bf16_acc(
        C.byref(actor), _p(actor_params, torch.float32, na, "actor_params"),
        _p(actor_image, torch.uint8, mlp_pack_bytes(actor), "actor_image"), C.byref(critic),
        _p(critic_params, torch.float32, nc, "critic_params"),
        _p(critic_image, torch.uint8, mlp_pack_bytes(critic), "critic_image"), C.byref(hyper),
        _p(view, torch.int8, None, "view"), _p(mask, torch.uint8, None, "mask"),
        _p(action, torch.int8, None, "action"), _p(old_logp, torch.float32, None, "old_logp"),
        _p(old_value, torch.float32, None, "old_value"), _p(adv, torch.float32, None, "adv"),
        _p(targets, torch.float32, None, "targets"),
        _p(rows, torch.int32, num_replicas * mb_size, "rows"), num_replicas, mb_size,
        _p(adv_stats, torch.float64, 16, "adv_stats"),
        _p(grad_out, torch.float32, na + nc + 8, "grad_out"),
        _p(workspace, torch.uint8, need, "workspace"), _stream()), "mava_ppo_loss_grad_bf16_acc")


def clip_adam_pair(params, mu, nu, counts, grad, n_actor: int, n_critic: int, grad_scale: float,
                   lr_actor: float, lr_critic: float, max_norm: float,
                   lr_decay_num_updates: int = 0, steps_per_update: int = 1) -> None:
    n = n_actor + n_critic
    _count(2)  # squared-norm pass + update pass
    check(_lib.load().mava_clip_adam_pair(
        _p(params, torch.float32, n, "params"), _p(mu, torch.float32, n, "mu"),
        _p(nu, torch.float32, n, "nu"), _p(counts, torch.int32, 2, "counts"),
        _p(grad, torch.float32, n, "grad"), n_actor, n_critic, grad_scale, lr_actor, lr_critic,
        max_norm, lr_decay_num_updates, steps_per_update, _stream()), "mava_clip_adam_pair")


def clip_adam_pair_pack(params, mu, nu, counts, grad, actor: MlpDesc, actor_image,
                        critic: MlpDesc, critic_image, grad_scale: float, lr_actor: float,
                        lr_critic: float, max_norm: float, lr_decay_num_updates: int = 0,
                        steps_per_update: int = 1) -> None:
    """clip_by_global_norm -> adam -> apply_updates that also refreshes the packed bf16 images."""
    n = mlp_param_count(actor) + mlp_param_count(critic)
    _count(2)
    check(_lib.load().mava_clip_adam_pair_pack(
        _p(params, torch.float32, n, "params"), _p(mu, torch.float32, n, "mu"),
        _p(nu, torch.float32, n, "nu"), _p(counts, torch.int32, 2, "counts"),
        _p(grad, torch.float32, n, "grad"), C.byref(actor),
        _p(actor_image, torch.uint8, mlp_pack_bytes(actor), "actor_image"), C.byref(critic),
        _p(critic_image, torch.uint8, mlp_pack_bytes(critic), "critic_image"), grad_scale, lr_actor,
        lr_critic, max_norm, lr_decay_num_updates, steps_per_update, _stream()),
        "mava_clip_adam_pair_pack")


# ---------------------------------------------------------------------------------------------
# recurrent systems (rec_ippo / rec_mappo)
# ---------------------------------------------------------------------------------------------
def rnn_desc(input_mode: int, add_agent_id: bool, num_agents: int, view_dim: int, hidden: int,
             post: int, out_dim: int, dense_in_dim: int = 0, rows_per_env: Optional[int] = None,
             precision: int = 0) -> RnnDesc:
    if input_mode == IN_DENSE:
        in_dim = int(dense_in_dim)
        rpe = num_agents if rows_per_env is None else int(rows_per_env)
    elif input_mode == IN_GLOBAL:
        in_dim, rpe = num_agents * view_dim, 1
    else:
        in_dim, rpe = view_dim + (num_agents if add_agent_id else 0), num_agents
    return RnnDesc(input_mode, int(add_agent_id), num_agents, view_dim, in_dim, rpe, hidden, post,
                   out_dim, int(precision))


def _p2d(t: torch.Tensor, name: str):
    """Row-major 2-D float32 matrix whose rows may be strided (leading dimension = stride(0))."""
    if not t.is_cuda or t.dtype != torch.float32 or t.dim() != 2 or t.stride(1) != 1:
        raise ValueError(f"{name}: expected a CUDA float32 matrix with unit column stride")
    return C.c_void_p(t.data_ptr())


def gemm(use_tc: bool, A, ta: bool, B, tb: bool, C, M: int, N: int, K: int, bias=None,
         relu: bool = False, relu_ref=None, mode: int = 0, k_splits: int = 1) -> None:
    """The recurrent path's dense contraction on row-major 2-D tensors (tests / diagnostics)."""
    _count(1)
    check(_lib.load().mava_gemm(
        int(use_tc), _p2d(A, "A"), int(ta), A.stride(0), _p2d(B, "B"), int(tb), B.stride(0),
        _p2d(C, "C"), C.stride(0), M, N, K, _p(bias, torch.float32, None, "bias"), int(relu),
        _p2d(relu_ref, "relu_ref") if relu_ref is not None else None,
        relu_ref.stride(0) if relu_ref is not None else 0, mode, k_splits, _stream()), "mava_gemm")


def _pmask(mask, actor: Optional[RnnDesc]):
    """Action masks are uint8 per (env, agent) up to 8 actions and uint16 above."""
    if mask is None:
        return None
    wide = actor is not None and actor.out_dim > 8
    return _p(mask, torch.uint16 if wide else torch.uint8, None, "mask")


class SynthEnv:
    """Synthetic SMAX-shaped step source (include/mava_b200.h, benchmark only)."""

    def __init__(self, num_agents=8, obs_dim=205, state_dim=168, num_actions=13,
                 done_prob=0.01, reward_std=0.1):
        self.cfg = SynthConfig(num_agents, obs_dim, state_dim, num_actions, done_prob, reward_std)
        self.num_agents, self.obs_dim, self.state_dim = num_agents, obs_dim, state_dim
        self.num_actions = num_actions
        self.state_stride = 16

    def alloc_state(self, num_envs: int, device) -> torch.Tensor:
        return torch.zeros(num_envs, 16, dtype=torch.uint8, device=device)

    def reset(self, key, state, obs_actor, obs_critic, mask, num_envs: int) -> None:
        _count(1)
        check(_lib.load().mava_synth_reset(
            C.byref(self.cfg), _p(key, torch.uint32, 2, "key"),
            _p(state, torch.uint8, num_envs * 16, "state"),
            _p(obs_actor, torch.float32, num_envs * self.num_agents * self.obs_dim, "obs_actor"),
            _p(obs_critic, torch.float32, num_envs * self.state_dim, "obs_critic"),
            _p(mask, torch.uint16, num_envs * self.num_agents, "mask"), num_envs, _stream()),
            "mava_synth_reset")

    def step(self, key, state, action, obs_actor, obs_critic, mask, reward, done, ep_return,
             ep_length, num_envs: int) -> None:
        A = self.num_agents
        _count(1)
        check(_lib.load().mava_synth_step(
            C.byref(self.cfg), _p(key, torch.uint32, 2, "key"),
            _p(state, torch.uint8, num_envs * 16, "state"),
            _p(action, torch.int8, num_envs * A, "action"),
            _p(obs_actor, torch.float32, num_envs * A * self.obs_dim, "obs_actor"),
            _p(obs_critic, torch.float32, num_envs * self.state_dim, "obs_critic"),
            _p(mask, torch.uint16, num_envs * A, "mask"),
            _p(reward, torch.float32, num_envs * A, "reward"), _p(done, torch.uint8, num_envs, "done"),
            _p(ep_return, torch.float32, num_envs, "ep_return"),
            _p(ep_length, torch.int32, num_envs, "ep_length"), num_envs, _stream()),
            "mava_synth_step")


def rnn_param_count(d: RnnDesc) -> int:
    return int(_lib.load().mava_rnn_param_count(C.byref(d)))


def rec_act_workspace_bytes(actor: Optional[RnnDesc], critic: RnnDesc, num_envs: int) -> int:
    return int(_lib.load().mava_rec_act_workspace_bytes(
        C.byref(actor) if actor is not None else None, C.byref(critic), num_envs))


def rec_act(actor: Optional[RnnDesc], actor_params, critic: Optional[RnnDesc], critic_params, view,
            obs_actor, obs_critic, mask, done_in, h_actor_in, h_actor_out, h_critic_in,
            h_critic_out, policy_key, envs_per_replica: int, num_envs: int, action, logp, value,
            workspace, greedy: bool = False, actions_in=None) -> None:
    """One acting step of the recurrent systems (actor=None: critic only, the bootstrap value)."""
    _count((9 if actor is not None else 0) + (9 if critic is not None and value is not None else 0))
    check(_lib.load().mava_rec_act(
        C.byref(actor) if actor is not None else None,
        _p(actor_params, torch.float32, rnn_param_count(actor) if actor is not None else None,
           "actor_params"),
        C.byref(critic) if critic is not None else None,
        _p(critic_params, torch.float32, rnn_param_count(critic) if critic is not None else None,
           "critic_params"),
        _p(view, torch.int8, None, "view"), _p(obs_actor, torch.float32, None, "obs_actor"),
        _p(obs_critic, torch.float32, None, "obs_critic"), _pmask(mask, actor),
        _p(done_in, torch.uint8, num_envs, "done_in"),
        _p(h_actor_in, torch.float32, None, "h_actor_in"),
        _p(h_actor_out, torch.float32, None, "h_actor_out"),
        _p(h_critic_in, torch.float32, None, "h_critic_in"),
        _p(h_critic_out, torch.float32, None, "h_critic_out"),
        _p(policy_key, torch.uint32, 2, "policy_key"), envs_per_replica, num_envs, int(greedy),
        _p(actions_in, torch.int8, None, "actions_in"), _p(action, torch.int8, None, "action"),
        _p(logp, torch.float32, None, "logp"), _p(value, torch.float32, None, "value"),
        _p(workspace, torch.uint8, None, "workspace"), _stream()), "mava_rec_act")


def rec_ppo_workspace_bytes(actor: RnnDesc, critic: RnnDesc, seq_envs_total: int, chunk: int) -> int:
    return int(_lib.load().mava_rec_ppo_workspace_bytes(C.byref(actor), C.byref(critic),
                                                        seq_envs_total, chunk))


def rec_ppo_loss_grad(actor: RnnDesc, actor_params, critic: RnnDesc, critic_params,
                      hyper: PpoHyper, view, obs_actor, obs_critic, mask, action, old_logp,
                      old_value, adv, targets, done_in, hs_actor, hs_critic, cols,
                      num_replicas: int, envs_per_replica: int, mb_cols: int, chunk: int,
                      num_chunks: int, grad_out, workspace) -> None:
    na, nc = rnn_param_count(actor), rnn_param_count(critic)
    need = rec_ppo_workspace_bytes(actor, critic, num_replicas * mb_cols, chunk)
    _count(2 * (4 * chunk + 24) + 3)
    check(_lib.load().mava_rec_ppo_loss_grad(
        C.byref(actor), _p(actor_params, torch.float32, na, "actor_params"), C.byref(critic),
        _p(critic_params, torch.float32, nc, "critic_params"), C.byref(hyper),
        _p(view, torch.int8, None, "view"), _p(obs_actor, torch.float32, None, "obs_actor"),
        _p(obs_critic, torch.float32, None, "obs_critic"), _pmask(mask, actor),
        _p(action, torch.int8, None, "action"), _p(old_logp, torch.float32, None, "old_logp"),
        _p(old_value, torch.float32, None, "old_value"), _p(adv, torch.float32, None, "adv"),
        _p(targets, torch.float32, None, "targets"), _p(done_in, torch.uint8, None, "done_in"),
        _p(hs_actor, torch.float32, None, "hs_actor"), _p(hs_critic, torch.float32, None, "hs_critic"),
        _p(cols, torch.int32, mb_cols, "cols"), num_replicas, envs_per_replica, mb_cols, chunk,
        num_chunks, _p(grad_out, torch.float32, na + nc + 8, "grad_out"),
        _p(workspace, torch.uint8, need, "workspace"), _stream()), "mava_rec_ppo_loss_grad")


# ---------------------------------------------------------------------------------------------
# pmean("device") fused with the optimiser step (csrc/peer.cu)
# ---------------------------------------------------------------------------------------------
def reduce_clip_adam_pair_acc(params, mu, nu, counts, group, gsum, n_actor: int, n_critic: int,
                              actor: Optional[MlpDesc], actor_image, critic: Optional[MlpDesc],
                              critic_image, grad_scale: float, lr_actor: float, lr_critic: float,
                              max_norm: float, lr_decay_num_updates: int, steps_per_update: int,
                              loss_out, loss_workspace, hyper: PpoHyper, loss_rows: int) -> None:
    """``reduce_clip_adam_pair`` paired with ``ppo_loss_grad_bf16_acc``: the loss metrics come from
    the loss kernel's accumulators, and the rank's gradient vector and the accumulators are left
    zero for the next minibatch."""
    n = n_actor + n_critic
    _count(1)
    check(_lib.load().mava_reduce_clip_adam_pair_acc(
        _p(params, torch.float32, n, "params"), _p(mu, torch.float32, n, "mu"),
        _p(nu, torch.float32, n, "nu"), _p(counts, torch.int32, 2, "counts"),
        C.byref(group.struct), _p(gsum, torch.float32, n, "gsum"), n_actor, n_critic,
        C.byref(actor) if actor is not None else None,
        _p(actor_image, torch.uint8, None, "actor_image"),
        C.byref(critic) if critic is not None else None,
        _p(critic_image, torch.uint8, None, "critic_image"), grad_scale, lr_actor, lr_critic,
        max_norm, lr_decay_num_updates, steps_per_update,
        _p(loss_out, torch.float32, 5, "loss_out"), _p(loss_workspace, torch.uint8, 256, "workspace"),
        C.byref(hyper), int(loss_rows), _stream()), "mava_reduce_clip_adam_pair_acc")


def reduce_clip_adam_pair(params, mu, nu, counts, group, gsum, n_actor: int, n_critic: int,
                          actor: Optional[MlpDesc], actor_image, critic: Optional[MlpDesc],
                          critic_image, grad_scale: float, lr_actor: float, lr_critic: float,
                          max_norm: float, lr_decay_num_updates: int = 0, steps_per_update: int = 1,
                          loss_out=None) -> None:
    """all-reduce(sum) over the ranks' exchange buffers of ``group`` (a ``peer.PeerGroup``) ->
    clip_by_global_norm -> adam -> apply_updates (+ bf16 image refresh, + loss metrics): one launch."""
    n = n_actor + n_critic
    _count(1)
    check(_lib.load().mava_reduce_clip_adam_pair(
        _p(params, torch.float32, n, "params"), _p(mu, torch.float32, n, "mu"),
        _p(nu, torch.float32, n, "nu"), _p(counts, torch.int32, 2, "counts"),
        C.byref(group.struct), _p(gsum, torch.float32, n, "gsum"), n_actor, n_critic,
        C.byref(actor) if actor is not None else None,
        _p(actor_image, torch.uint8, None, "actor_image"),
        C.byref(critic) if critic is not None else None,
        _p(critic_image, torch.uint8, None, "critic_image"), grad_scale, lr_actor, lr_critic,
        max_norm, lr_decay_num_updates, steps_per_update,
        _p(loss_out, torch.float32, 5, "loss_out"), _stream()), "mava_reduce_clip_adam_pair")
