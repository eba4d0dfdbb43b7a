"""SMAX-shaped synthetic step source behind the env seam the recurrent learner uses.

``configs/env/smax.yaml`` selects JaxMARL's ``HeuristicEnemySMAX`` through ``SmaxWrapper``
(mava/wrappers/jaxmarl.py:326-373, mava/utils/make_env.py:137-147).  Its float physics live in
jaxmarl (third party, unpinned, not under /root/reference), so there is no SMAX env kernel here;
``env=smax_synthetic`` stands in for it in benchmarks with tensors of the same SHAPE (per-agent
observation, world state, 5 + n_enemies actions) generated on the device.  Training on it learns
nothing and ``run_experiment`` refuses it: it exists to time rec_mappo at SMAX's shapes
(BASELINE.json configs[3])."""
from __future__ import annotations

import torch

from .. import native


class SyntheticSmaxEnv:
    dense = True  # observations are f32 rows (MAVA_IN_DENSE), not the env kernels' int8 view

    def __init__(self, num_agents: int, obs_dim: int, state_dim: int, num_actions: int,
                 time_limit: int, device: torch.device):
        self.native = native.SynthEnv(num_agents, obs_dim, state_dim, num_actions,
                                      done_prob=1.0 / max(1, time_limit))
        self.device = device
        self.num_agents, self.action_dim, self.time_limit = num_agents, num_actions, time_limit
        self.obs_dim, self.state_dim = obs_dim, state_dim
