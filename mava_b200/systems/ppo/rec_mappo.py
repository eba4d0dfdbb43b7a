"""rec_mappo on the B200 kernels: drop-in for mava/systems/ppo/rec_mappo.py (same entry points:
``get_learner_fn``, ``learner_setup``, ``run_experiment``, ``hydra_entry_point``).

    python -m mava_b200.systems.ppo.rec_mappo env=rware env/scenario=tiny-4ag arch.num_envs=256

Actor and critic are pre-MLP -> reset-masked GRU -> post-MLP -> head (mava/networks.py:269-331);
the critic is centralised (global state, rec_mappo.py:457,575)."""
from __future__ import annotations

from typing import Tuple

from ...config import compose, parse_cli
from ...evaluator import make_rec_eval_act_fn
from . import _runner, anakin_rec

CENTRALISED_CRITIC = True
CONFIG_NAME = "default_rec_mappo.yaml"

get_learner_fn = anakin_rec.get_learner_fn


def learner_setup(env, keys: Tuple, config):
    """Initialise learner_fn, network, optimiser, environment and states."""
    return anakin_rec.learner_setup(env, keys, config, centralised_critic=CENTRALISED_CRITIC)


def run_experiment(_config) -> float:
    """Runs experiment; returns the final evaluation metric."""
    return _runner.run_experiment(
        _config, learner_setup, lambda learner, cfg: make_rec_eval_act_fn(learner.actor_desc, cfg),
        add_global_state=CENTRALISED_CRITIC, recurrent=True)


def hydra_entry_point(argv=None) -> float:
    """Experiment entry point: Hydra-style overrides on the command line."""
    overrides, config_dir = parse_cli(argv)
    cfg = compose(CONFIG_NAME, overrides, config_dir)
    eval_performance = run_experiment(cfg)
    print("Recurrent MAPPO experiment completed")
    return eval_performance


if __name__ == "__main__":
    hydra_entry_point()
