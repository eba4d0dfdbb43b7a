"""mava/utils/total_timestep_checker.py:21-49 with n_devices = number of ranks."""
from __future__ import annotations


def check_total_timesteps(config, n_devices: int):
    s = config.system
    if s.total_timesteps is None:
        s.num_updates = int(s.num_updates)
        s.total_timesteps = int(n_devices * s.num_updates * s.rollout_length * s.update_batch_size
                                * config.arch.num_envs)
    else:
        s.total_timesteps = int(s.total_timesteps)
        s.num_updates = int(s.total_timesteps // s.rollout_length // s.update_batch_size
                            // config.arch.num_envs // n_devices)
        print(f"Changing the number of updates to {s.num_updates}: if you want to train for a "
              "specific number of updates, please set total_timesteps to None!")
    return config
