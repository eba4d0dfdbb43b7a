"""Environment factory mirroring mava/utils/make_env.py:215-240 for the native env kernels."""
from __future__ import annotations

from typing import Tuple

import torch

from .. import native
from ..wrappers import NativeMarlEnv

_jumanji_registry = {"RobotWarehouse-v0": "rware", "LevelBasedForaging-v0": "lbf"}


def add_extra_wrappers(config) -> bool:
    """make_env.py:69-83: the agent-id decision (the wrappers themselves are fused in-kernel)."""
    config.system.add_agent_id = bool(config.system.add_agent_id) and not bool(
        config.env.implicit_agent_id)
    return config.system.add_agent_id


def make_jumanji_env(env_name: str, config, add_global_state: bool = False,
                     device: torch.device = None) -> Tuple[NativeMarlEnv, NativeMarlEnv]:
    """make_env.py:86-116."""
    device = device or torch.device("cuda", torch.cuda.current_device())
    task = dict(config.env.scenario.task_config)
    env_kwargs = {**dict(config.env.kwargs), **dict(config.env.scenario.env_kwargs or {})}
    if _jumanji_registry[env_name] == "rware":
        handle = native.Env.rware(time_limit=int(env_kwargs.get("time_limit", 500)), **task)
    else:
        handle = native.Env.lbf(
            time_limit=int(env_kwargs.get("time_limit", 100)),
            use_individual_rewards=bool(config.env.get("use_individual_rewards", False)), **task)
    add_id = add_extra_wrappers(config)
    train_env = NativeMarlEnv(handle, add_global_state, add_id, auto_reset=True, device=device)
    eval_env = NativeMarlEnv(handle, add_global_state, add_id, auto_reset=False, device=device)
    return train_env, eval_env


def make(config, add_global_state: bool = False, device: torch.device = None
         ) -> Tuple[NativeMarlEnv, NativeMarlEnv]:
    """Create the training and evaluation environments (make_env.py:215-240)."""
    env_name = config.env.scenario.name
    if env_name in _jumanji_registry:
        return make_jumanji_env(env_name, config, add_global_state, device)
    if env_name == "SmaxSynthetic":
        from ..wrappers import SyntheticSmaxEnv

        device = device or torch.device("cuda", torch.cuda.current_device())
        tc = dict(config.env.scenario.task_config)
        add_extra_wrappers(config)
        mk = lambda: SyntheticSmaxEnv(int(tc["num_agents"]), int(tc["obs_dim"]),
                                      int(tc["state_dim"]), int(tc["num_actions"]),
                                      int(config.env.kwargs.get("time_limit", 100)), device)
        return mk(), mk()
    raise ValueError(
        f"{env_name} is not supported by the mava_b200 env kernels "
        f"(supported: {sorted(_jumanji_registry)})")
