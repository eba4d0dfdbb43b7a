"""Host-side records mirroring mava/types.py:111-160 and mava/systems/ppo/types.py:25-91.

Arrays are torch CUDA tensors.  Unlike the JAX reference the learner state is a set of device
buffers that ``learn`` advances IN PLACE; the records below are views onto those buffers, so a
state returned by ``learn`` aliases the one passed in.
"""
from __future__ import annotations

from typing import Any, Callable, Dict, Generic, NamedTuple, Optional, TypeVar

import torch

Metrics = Dict[str, torch.Tensor]
State = Any


class Observation(NamedTuple):
    """mava/types.py:111-121."""

    agents_view: torch.Tensor  # (..., num_agents, num_obs_features) float32
    action_mask: torch.Tensor  # (..., num_agents, num_actions) bool
    step_count: torch.Tensor  # (..., num_agents) int32


class ObservationGlobalState(NamedTuple):
    """mava/types.py:124-134."""

    agents_view: torch.Tensor
    action_mask: torch.Tensor
    global_state: torch.Tensor  # (..., num_agents, num_agents * num_obs_features)
    step_count: torch.Tensor


class StepType:
    FIRST, MID, LAST = 0, 1, 2


class TimeStep(NamedTuple):
    """jumanji.types.TimeStep as the wrappers emit it (batched over envs)."""

    step_type: torch.Tensor  # (E,) int8
    reward: torch.Tensor  # (E, A) float32
    discount: torch.Tensor  # (E, A) float32
    observation: Any
    extras: Dict[str, Any]

    def last(self) -> torch.Tensor:
        return self.step_type == StepType.LAST


MavaState = TypeVar("MavaState")


class ExperimentOutput(NamedTuple, Generic[MavaState]):
    """mava/types.py:146-151."""

    learner_state: MavaState
    episode_metrics: Metrics
    train_metrics: Metrics


LearnerFn = Callable[[MavaState], ExperimentOutput]


class Params(NamedTuple):
    actor_params: torch.Tensor  # flat f32, flax order (see include/mava_b200.h)
    critic_params: torch.Tensor


class OptStates(NamedTuple):
    actor_opt_state: Dict[str, torch.Tensor]  # {"mu", "nu", "count"}
    critic_opt_state: Dict[str, torch.Tensor]


class HiddenStates(NamedTuple):
    policy_hidden_state: torch.Tensor
    critic_hidden_state: torch.Tensor


class LearnerState(NamedTuple):
    """mava/systems/ppo/types.py:47-54."""

    params: Params
    opt_states: OptStates
    key: torch.Tensor  # uint32[2]
    env_state: State
    timestep: TimeStep


class RNNLearnerState(NamedTuple):
    """mava/systems/ppo/types.py:57-66."""

    params: Params
    opt_states: OptStates
    key: torch.Tensor
    env_state: State
    timestep: TimeStep
    dones: torch.Tensor
    hstates: Optional[HiddenStates]
