"""``run_experiment`` shared by the PPO systems: mava/systems/ppo/ff_mappo.py:435-553 (ff_ippo,
rec_ippo and rec_mappo run the same loop)."""
from __future__ import annotations

import copy
import os
import time
from typing import Callable

import numpy as np
import torch
import torch.distributed as dist

from ... import prng
from ...evaluator import get_eval_fn
from ...utils import make_env as environments
from ...utils.logger import LogEvent, MavaLogger, get_final_step_metrics
from ...utils.total_timestep_checker import check_total_timesteps
from .anakin import episode_summary, world


def init_distributed() -> torch.device:
    """One process per GPU.  Under torchrun (RANK/WORLD_SIZE set) join the NCCL group."""
    if not torch.cuda.is_available():
        raise RuntimeError("mava_b200 needs a CUDA device: there is no CPU fallback")
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local_rank)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
    return torch.device("cuda", local_rank)


def run_experiment(_config, learner_setup: Callable, make_eval_act_fn: Callable,
                   add_global_state: bool, recurrent: bool = False) -> float:
    config = copy.deepcopy(_config)
    device = init_distributed()
    rank, n_devices = world()

    if recurrent:  # rec_mappo.py:586-591
        if config.system.recurrent_chunk_size is None:
            config.system.recurrent_chunk_size = config.system.rollout_length
        else:
            assert config.system.rollout_length % config.system.recurrent_chunk_size == 0, \
                "Rollout length must be divisible by recurrent chunk size."

    env, eval_env = environments.make(config=config, add_global_state=add_global_state,
                                      device=device)
    if getattr(env, "dense", False):
        raise ValueError("env=smax_synthetic is a benchmark-only step source (no dynamics, no "
                         "evaluator): use bench.py --workload rec_mappo_smax")
    key, key_e, actor_net_key, critic_net_key = prng.split(prng.PRNGKey(config.system.seed), 4)
    learn, actor_network, learner_state = learner_setup(
        env, (key, actor_net_key, critic_net_key), config)
    learner = learn.learner

    eval_keys = prng.split(key_e, n_devices)
    eval_act_fn = make_eval_act_fn(learner, config)
    evaluator = get_eval_fn(eval_env, eval_act_fn, config, absolute_metric=False)

    config = check_total_timesteps(config, n_devices)
    assert config.system.num_updates > config.arch.num_evaluation, \
        "Number of updates per evaluation must be less than total number of updates."
    config.system.num_updates_per_eval = config.system.num_updates // config.arch.num_evaluation
    learner.config = config
    steps_per_rollout = (n_devices * config.system.num_updates_per_eval
                         * config.system.rollout_length * config.system.update_batch_size
                         * config.arch.num_envs)

    logger = MavaLogger(config, rank)
    # Set up checkpointer (ff_mappo.py:482-489); rank 0 writes, the state is replicated
    save_checkpoint = bool(config.logger.checkpointing.save_model)
    if save_checkpoint:
        from ...utils.checkpointing import Checkpointer

        checkpointer = Checkpointer(metadata=config, model_name=config.logger.system_name,
                                    **dict(config.logger.checkpointing.save_args))
    max_episode_return = -np.inf
    best_params = None
    eval_metrics = {}
    eval_step = 0
    for eval_step in range(config.arch.num_evaluation):
        start_time = time.time()
        # the parameters evaluated are those BEFORE this learn call (ff_mappo.py:513 vs :535)
        trained_params = learner_state.params.actor_params.clone()
        learner_output = learn(learner_state)
        torch.cuda.synchronize(device)
        elapsed_time = time.time() - start_time

        t = int(steps_per_rollout * (eval_step + 1))
        # the finished-episode statistics were reduced on the device (mava_episode_stats): 80 bytes
        # cross the bus instead of ExperimentOutput.episode_metrics (three [updates][T][envs] arrays,
        # still returned for callers that want them -- get_final_step_metrics works on them)
        summary, ep_completed = episode_summary(learner)
        logger.log({"timestep": t}, t, eval_step, LogEvent.MISC)
        if ep_completed:
            logger.log_summary(summary, {"steps_per_second": steps_per_rollout / elapsed_time}, t,
                               eval_step, LogEvent.ACT)
        logger.log(learner_output.train_metrics, t, eval_step, LogEvent.TRAIN)

        ks = prng.split(key_e, n_devices + 1)
        key_e, eval_keys = ks[0], ks[1:]
        eval_metrics = evaluator(trained_params, eval_keys[rank], {})
        eval_metrics = _gather_metrics(eval_metrics, n_devices)
        logger.log(eval_metrics, t, eval_step, LogEvent.EVAL)
        episode_return = float(eval_metrics["episode_return"].float().mean().item())

        if save_checkpoint and rank == 0:  # ff_mappo.py:522-528
            checkpointer.save(timestep=t, unreplicated_learner_state=_checkpoint_tree(learner),
                              episode_return=episode_return)

        if config.arch.absolute_metric and max_episode_return <= episode_return:
            best_params = trained_params.clone()
            max_episode_return = episode_return
        learner_state = learner_output.learner_state

    eval_performance = float(eval_metrics[config.env.eval_metric].float().mean().item())

    if config.arch.absolute_metric:
        abs_evaluator = get_eval_fn(eval_env, eval_act_fn, config, absolute_metric=True)
        eval_keys = prng.split(key, n_devices)
        abs_metrics = _gather_metrics(abs_evaluator(best_params, eval_keys[rank], {}), n_devices)
        t = int(steps_per_rollout * (eval_step + 1))
        logger.log(abs_metrics, t, eval_step, LogEvent.ABSOLUTE)
    logger.stop()
    learner.release()
    return eval_performance


def _checkpoint_tree(learner):
    """The unreplicated learner state as the reference's checkpoint tree (params in flax naming,
    optimiser moments in optax naming)."""
    from ...utils.checkpointing import learner_tree

    actor_net, critic_net = learner.networks
    na = learner.na
    p = learner.params.cpu().numpy()
    mu, nu = learner.mu.cpu().numpy(), learner.nu.cpu().numpy()
    cnt = learner.counts.cpu().numpy()
    ain, cin = learner.actor_desc.in_dim, learner.critic_desc.in_dim
    opt = {"actor_opt_state": {"count": cnt[0], "mu": actor_net.to_flax_tree(mu[:na], ain),
                               "nu": actor_net.to_flax_tree(nu[:na], ain)},
           "critic_opt_state": {"count": cnt[1], "mu": critic_net.to_flax_tree(mu[na:], cin),
                                "nu": critic_net.to_flax_tree(nu[na:], cin)}}
    return learner_tree(actor_net, critic_net, p[:na], p[na:], ain, cin, opt)


def _gather_metrics(metrics, n_devices: int):
    """Concatenate per-rank evaluation metrics (the reference's pmap output axis)."""
    if n_devices == 1:
        return metrics
    out = {}
    for k, v in metrics.items():
        v = v.cuda() if not v.is_cuda else v
        v = v.reshape(-1) if v.dim() else v.reshape(1)
        bufs = [torch.empty_like(v) for _ in range(n_devices)]
        dist.all_gather(bufs, v)
        out[k] = torch.cat(bufs)
    return out
