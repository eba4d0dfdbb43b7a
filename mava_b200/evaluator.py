"""Evaluator mirroring mava/evaluator.py:64-209 on the native env/actor kernels."""
from __future__ import annotations

import math
import time
import warnings
from typing import Callable, Dict

import numpy as np
import torch

from . import native, prng
from .systems.ppo.anakin import world


def get_num_eval_envs(config, absolute_metric: bool) -> int:
    """evaluator.py:64-77."""
    _, n_devices = world()
    n_parallel_envs = config.arch.num_envs * n_devices
    eval_episodes = (config.arch.num_absolute_metric_eval_episodes if absolute_metric
                     else config.arch.num_eval_episodes)
    if eval_episodes <= n_parallel_envs:
        return math.ceil(eval_episodes / n_devices)
    return int(config.arch.num_envs)


def make_ff_eval_act_fn(actor_desc, config) -> Callable:
    """evaluator.py:175-186: sample (or mode() when evaluation_greedy) from the actor."""
    greedy = bool(config.arch.evaluation_greedy)

    def eval_act_fn(params, view, mask, key, num_envs, action, logp):
        native.ff_act(actor_desc, params, None, None, view, mask, key, num_envs, num_envs, action,
                      logp, None, greedy=greedy)

    # what the evaluator needs to run whole episodes in ONE launch of the fused rollout kernel
    eval_act_fn.actor_desc = actor_desc  # type: ignore[attr-defined]
    eval_act_fn.greedy = greedy  # type: ignore[attr-defined]
    eval_act_fn.precision = str(config.arch.get("precision", "auto"))  # type: ignore[attr-defined]
    return eval_act_fn


def make_rec_eval_act_fn(actor_desc, config) -> Callable:
    """evaluator.py:189-209: the recurrent actor carries a hidden state through the episode
    (zeros at the start, evaluator's ``init_act_state``) and resets it on ``timestep.last()``."""
    greedy = bool(config.arch.evaluation_greedy)
    st: Dict[str, torch.Tensor] = {}

    def reset(num_envs: int, device) -> None:
        if "h" not in st or st["h"].shape[0] != num_envs * actor_desc.num_agents:
            st["h"] = torch.zeros(num_envs * actor_desc.num_agents, actor_desc.hidden, device=device)
            st["ws"] = torch.zeros(native.rec_act_workspace_bytes(actor_desc, actor_desc, num_envs),
                                   dtype=torch.uint8, device=device)
            st["zero"] = torch.zeros(num_envs, dtype=torch.uint8, device=device)
        st["h"].zero_()

    def eval_act_fn(params, view, mask, key, num_envs, action, logp, last_done=None):
        native.rec_act(actor_desc, params, None, None, view, None, None, mask,
                       st["zero"] if last_done is None else last_done, st["h"], st["h"], None, None,
                       key, num_envs, num_envs, action, logp, None, st["ws"], greedy=greedy)

    eval_act_fn.reset = reset  # type: ignore[attr-defined]
    eval_act_fn.recurrent = True  # type: ignore[attr-defined]
    return eval_act_fn


def get_eval_fn(env, act_fn: Callable, config, absolute_metric: bool):
    """evaluator.py:80-172.  Returns ``evaluator(params, key, init_act_state) -> metrics`` where
    ``key`` is this rank's uint32[2] evaluation key (host array)."""
    _, n_devices = world()
    eval_episodes = (config.arch.num_absolute_metric_eval_episodes if absolute_metric
                     else config.arch.num_eval_episodes)
    n_envs = get_num_eval_envs(config, absolute_metric)
    n_parallel = n_envs * n_devices
    episode_loops = math.ceil(eval_episodes / n_parallel)
    if eval_episodes % n_parallel != 0:
        warnings.warn(f"Number of evaluation episodes ({eval_episodes}) is not divisible by "
                      f"`num_envs` * `num_devices` ({n_parallel}); "
                      f"{episode_loops * n_parallel} episodes will be run.", stacklevel=2)
    dev = env.device
    A, FR, T = env.num_agents, env.native.view_dim, int(env.time_limit)
    # Fast path: the episode loop `while not timestep.last()` (evaluator.py:95-130) as ONE launch of
    # the fused rollout kernel (actor MLP on the tensor cores + sample / mode + env step, csrc/
    # rollout_tc.cu) with auto_reset = 0, then the first-terminal metric pick on the device.  Same
    # key schedule and noise layout as the step-wise path below, which stays for everything the
    # fused kernel does not cover (LBF, recurrent actors, arch.precision=fp32).
    desc = getattr(act_fn, "actor_desc", None)
    fused = (desc is not None and getattr(act_fn, "precision", "auto") != "fp32"
             and not getattr(act_fn, "recurrent", False) and native.ff_rollout_supported(env.native)
             and desc.h1 == 128 and desc.h2 == 128 and desc.out_dim <= 16
             and bool(config.arch.get("fused_rollout", True)))
    state = env.native.alloc_state(n_envs, dev)
    slots = T if not fused else 1
    view = torch.zeros(n_envs, A, FR, dtype=torch.int8, device=dev)
    mask = torch.zeros(n_envs, A, dtype=torch.uint8, device=dev)
    action = torch.zeros(n_envs, A, dtype=torch.int8, device=dev)
    logp = torch.zeros(n_envs, A, device=dev)
    reward = torch.zeros(n_envs, A, device=dev)
    done = torch.zeros(T, n_envs, dtype=torch.uint8, device=dev)
    ep_ret = torch.zeros(T, n_envs, device=dev)
    ep_len = torch.zeros(T, n_envs, dtype=torch.int32, device=dev)
    act_keys = torch.zeros(T, 2, dtype=torch.uint32, device=dev)
    key_dev = torch.zeros(2, dtype=torch.uint32, device=dev)
    image = torch.zeros(native.mlp_pack_bytes(desc), dtype=torch.uint8, device=dev) if fused else None

    def timed_eval_fn(params: torch.Tensor, key: np.ndarray, init_act_state=None,
                      record: Dict | None = None) -> Dict:
        """``record`` (tests): a dict that receives the sampled actions [T][n_envs][A] and the reset
        keys of every episode loop, so the episodes can be replayed through the oracle."""
        start = time.time()
        rets, lens = [], []
        if fused:
            native.mlp_pack_bf16(desc, params, image)
        for _ in range(episode_loops):  # _episode, evaluator.py:132-150
            key, reset_key = prng.split(key)
            reset_keys = prng.split(reset_key, n_envs)
            env.native.reset(torch.from_numpy(reset_keys.copy()).to(dev), state, view, mask, n_envs)
            key_dev.copy_(torch.from_numpy(np.ascontiguousarray(key)).to(dev))
            native.prng_split_chain(key_dev, act_keys, T)  # key, act_key = split(key) per step
            if fused:
                rec = record is not None
                if rec:  # full [T] stacks so the test can read the actions back
                    v = torch.zeros(T + 1, n_envs, A, FR, dtype=torch.int8, device=dev)
                    m = torch.zeros(T + 1, n_envs, A, dtype=torch.uint8, device=dev)
                    v[0].copy_(view)
                    m[0].copy_(mask)
                    acts = torch.zeros(T, n_envs, A, dtype=torch.int8, device=dev)
                    lp, rw = torch.zeros(T, n_envs, A, device=dev), torch.zeros(T, n_envs, A, device=dev)
                else:
                    v, m, acts, lp, rw = view, mask, action, logp, reward
                native.ff_rollout_bf16_ex(env.native, desc, params, image, state, v, m, act_keys,
                                          n_envs, n_envs, T, False, bool(act_fn.greedy), rec, acts,
                                          lp, rw, done, ep_ret, ep_len)
                if rec:
                    record.setdefault("actions", []).append(acts.cpu().numpy())
                    record.setdefault("reset_keys", []).append(reset_keys.copy())
                r = torch.empty(n_envs, device=dev)
                ln = torch.empty(n_envs, dtype=torch.int32, device=dev)
                native.episode_first_terminal(done, ep_ret, ep_len, T, n_envs, r, ln)
                rets.append(r)
                lens.append(ln)
                key = key_dev.cpu().numpy()
                continue
            recurrent = bool(getattr(act_fn, "recurrent", False))
            if recurrent:
                act_fn.reset(n_envs, dev)
            for t in range(T):
                if recurrent:
                    act_fn(params, view, mask, act_keys[t], n_envs, action, logp,
                           done[t - 1] if t > 0 else None)
                else:
                    act_fn(params, view, mask, act_keys[t], n_envs, action, logp)
                if record is not None:
                    record.setdefault("step_actions", []).append(action.cpu().numpy().copy())
                env.native.step(state, action, view, mask, reward, done[t], ep_ret[t], ep_len[t],
                                n_envs, False)
            if record is not None:
                record.setdefault("actions", []).append(np.stack(record.pop("step_actions")))
                record.setdefault("reset_keys", []).append(reset_keys.copy())
            key = key_dev.cpu().numpy()
            r = torch.empty(n_envs, device=dev)
            ln = torch.empty(n_envs, dtype=torch.int32, device=dev)
            native.episode_first_terminal(done, ep_ret, ep_len, T, n_envs, r, ln)
            rets.append(r)
            lens.append(ln)
        metrics = {"episode_return": torch.cat(rets), "episode_length": torch.cat(lens)}
        torch.cuda.synchronize(dev)
        total = float(metrics["episode_length"].sum().item())
        metrics["steps_per_second"] = torch.tensor(total / max(time.time() - start, 1e-9))
        return metrics

    timed_eval_fn.fused = fused  # type: ignore[attr-defined]
    return timed_eval_fn
