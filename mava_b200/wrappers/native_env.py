"""The MarlEnv contract (mava/types.py:34-108) over the fused CUDA env kernels.

One object stands for the whole reference wrapper stack
``RecordEpisodeMetrics(AutoResetWrapper(AgentIDWrapper(RwareWrapper | LbfWrapper(env))))``
(mava/utils/make_env.py:69-83,104-115): ``auto_reset=True`` is the training env, ``False`` the
evaluation env.  Unlike the reference, ``reset``/``step`` are natively batched over environments
(no vmap) and the state is a packed device buffer that ``step`` advances in place.

``reset``/``step`` here are the convenience API (they allocate their outputs and decode the int8
observation into the reference's float32 ``Observation``); the learner drives the same kernels
through ``self.native`` with preallocated rollout buffers.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Tuple, Union

import torch

from .. import native
from ..types import Observation, ObservationGlobalState, StepType, TimeStep


@dataclass
class EnvState:
    buf: torch.Tensor  # uint8 [num_envs, state_stride]
    view: torch.Tensor  # int8 [num_envs, A, FR]  current raw observation
    mask: torch.Tensor  # uint8 [num_envs, A]

    @property
    def num_envs(self) -> int:
        return self.buf.shape[0]


@dataclass
class ObservationSpec:
    num_agents: int
    num_obs_features: int
    num_actions: int
    global_state_dim: int  # 0 when there is no global state
    time_limit: int

    def generate_value(self) -> Union[Observation, ObservationGlobalState]:
        A = self.num_agents
        view = torch.zeros(A, self.num_obs_features)
        mask = torch.zeros(A, self.num_actions, dtype=torch.bool)
        step = torch.zeros(A, dtype=torch.int32)
        if self.global_state_dim:
            return ObservationGlobalState(view, mask, torch.zeros(A, self.global_state_dim), step)
        return Observation(view, mask, step)


class NativeMarlEnv:
    def __init__(self, env: native.Env, add_global_state: bool, add_agent_id: bool,
                 auto_reset: bool, device: torch.device):
        self.native = env
        self.add_global_state = add_global_state
        self.add_agent_id = add_agent_id
        self.auto_reset = auto_reset
        self.device = device
        self.num_agents = env.num_agents
        self.time_limit = env.dims.time_limit
        self.action_dim = env.num_actions

    # -- specs --------------------------------------------------------------------------------
    def observation_spec(self) -> ObservationSpec:
        A, FR = self.num_agents, self.native.view_dim
        return ObservationSpec(A, FR + (A if self.add_agent_id else 0), self.action_dim,
                               A * FR if self.add_global_state else 0, self.time_limit)

    def action_spec(self) -> Dict[str, int]:
        return {"num_agents": self.num_agents, "num_values": self.action_dim}

    # -- decoding -----------------------------------------------------------------------------
    def decode_observation(self, view: torch.Tensor, mask: torch.Tensor, step_count: torch.Tensor
                           ) -> Union[Observation, ObservationGlobalState]:
        """int8 view + mask bits -> the reference's float32 Observation (jumanji.py:135-144,
        observation.py:41-53, jumanji.py:53-59)."""
        A = self.num_agents
        raw = view.to(torch.float32)
        agents_view = raw
        if self.add_agent_id:
            ids = torch.eye(A, device=view.device).expand(*view.shape[:-2], A, A)
            agents_view = torch.cat([ids, raw], dim=-1)
        bits = torch.arange(self.action_dim, device=view.device)
        action_mask = ((mask.to(torch.int32).unsqueeze(-1) >> bits) & 1).bool()
        step = step_count.to(torch.int32).unsqueeze(-1).expand(*view.shape[:-2], A)
        if self.add_global_state:
            gs = raw.reshape(*view.shape[:-2], 1, -1).expand(*view.shape[:-2], A, -1)
            return ObservationGlobalState(agents_view, action_mask, gs, step)
        return Observation(agents_view, action_mask, step)

    def step_count(self, state: EnvState) -> torch.Tensor:
        return self.native.peek(state.buf, 0, state.num_envs)[:, 0]

    # -- MarlEnv API --------------------------------------------------------------------------
    def reset(self, keys: torch.Tensor) -> Tuple[EnvState, TimeStep]:
        """vmap(env.reset)(keys); keys uint32 [num_envs, 2]."""
        n = keys.shape[0]
        A, FR = self.num_agents, self.native.view_dim
        st = EnvState(self.native.alloc_state(n, self.device),
                      torch.zeros(n, A, FR, dtype=torch.int8, device=self.device),
                      torch.zeros(n, A, dtype=torch.uint8, device=self.device))
        self.native.reset(keys.contiguous(), st.buf, st.view, st.mask, n)
        obs = self.decode_observation(st.view, st.mask, torch.zeros(n, device=self.device))
        extras = {"episode_metrics": {
            "episode_return": torch.zeros(n, device=self.device),
            "episode_length": torch.zeros(n, dtype=torch.int32, device=self.device),
            "is_terminal_step": torch.zeros(n, dtype=torch.bool, device=self.device)}}
        ts = TimeStep(torch.full((n,), StepType.FIRST, dtype=torch.int8, device=self.device),
                      torch.zeros(n, A, device=self.device), torch.ones(n, A, device=self.device),
                      obs, extras)
        return st, ts

    def step(self, state: EnvState, action: torch.Tensor) -> Tuple[EnvState, TimeStep]:
        """vmap(env.step)(state, action); the state buffers are advanced in place."""
        n, A = state.num_envs, self.num_agents
        dev = self.device
        reward = torch.empty(n, A, device=dev)
        done = torch.empty(n, dtype=torch.uint8, device=dev)
        ep_ret = torch.empty(n, device=dev)
        ep_len = torch.empty(n, dtype=torch.int32, device=dev)
        self.native.step(state.buf, action.to(torch.int8).contiguous(), state.view, state.mask,
                         reward, done, ep_ret, ep_len, n, self.auto_reset)
        last = done.bool()
        obs = self.decode_observation(state.view, state.mask, self.step_count(state))
        extras = {"episode_metrics": {"episode_return": ep_ret, "episode_length": ep_len,
                                      "is_terminal_step": last}}
        step_type = torch.where(last, StepType.LAST, StepType.MID).to(torch.int8)
        discount = (~last).float().unsqueeze(-1).expand(n, A)
        return state, TimeStep(step_type, reward, discount, obs, extras)
