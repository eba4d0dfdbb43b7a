"""Evaluator parity (mava/evaluator.py:64-172): the episodes the evaluator plays -- whole episodes in
one launch of the fused rollout kernel with auto_reset = 0, or step by step where that kernel does
not apply -- are replayed action by action through the C port of the env oracle; the metrics the
evaluator reports must be the oracle's episode return / length at the FIRST terminal step of every
env (evaluator.py:143-150), for every episode loop, and the number of episodes must follow the
`n_vmapped_envs` rule (:64-77)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("system,precision,greedy,episodes,expect_fused", [
    ("ff_mappo", "auto", False, 24, True), ("ff_ippo", "auto", True, 10, True),
    ("ff_mappo", "fp32", False, 8, False)])
def test_evaluator_metrics_match_oracle_replay(lib_built, system, precision, greedy, episodes,
                                               expect_fused):
    import importlib

    from mava_b200 import prng
    from mava_b200.config import compose
    from mava_b200.evaluator import get_eval_fn, make_ff_eval_act_fn
    from mava_b200.utils import make_env
    from oracle.rware_c import RwareC

    torch.cuda.set_device(0)
    mod = importlib.import_module(f"mava_b200.systems.ppo.{system}")
    time_limit, num_envs = 40, 8
    cfg = compose(mod.CONFIG_NAME, [
        "env/scenario=tiny-4ag", f"arch.num_envs={num_envs}", "system.rollout_length=8",
        f"env.kwargs.time_limit={time_limit}", f"arch.num_eval_episodes={episodes}",
        f"arch.evaluation_greedy={greedy}", f"+arch.precision={precision}",
        "logger.use_console=False"])
    env, eval_env = make_env.make(cfg, add_global_state=mod.CENTRALISED_CRITIC)
    key, key_e, ak, ck = prng.split(prng.PRNGKey(17), 4)
    learn, _, _ = mod.learner_setup(env, (key, ak, ck), cfg)
    L = learn.learner
    evaluator = get_eval_fn(eval_env, make_ff_eval_act_fn(L.actor_desc, cfg), cfg,
                            absolute_metric=False)
    assert evaluator.fused == expect_fused
    rec = {}
    metrics = evaluator(L.params[:L.na].clone(), key_e, {}, record=rec)
    # evaluator.py:64-77: 24 episodes > 8 envs -> 8 envs x 3 loops; 10 -> 8 x 2 (16 episodes run);
    # 8 -> 8 x 1
    loops = -(-episodes // num_envs)
    assert len(rec["actions"]) == loops
    assert metrics["episode_return"].numel() == loops * num_envs
    got_ret = metrics["episode_return"].cpu().numpy().reshape(loops, num_envs)
    got_len = metrics["episode_length"].cpu().numpy().reshape(loops, num_envs)

    oc = RwareC(time_limit=time_limit, **dict(cfg.env.scenario.task_config))
    for k in range(loops):
        acts = rec["actions"][k]
        assert acts.shape == (time_limit, num_envs, 4) and acts.min() >= 0 and acts.max() <= 4
        state, _, _ = oc.reset(rec["reset_keys"][k])
        first = np.full(num_envs, -1)
        ret, ln = np.zeros(num_envs, np.float32), np.zeros(num_envs, np.int32)
        for t in range(time_limit):
            _, _, _, done, er, el = oc.step(state, acts[t], auto_reset=False)
            new = (done != 0) & (first < 0)
            ret[new], ln[new], first[new] = er[new], el[new], t
        assert (first >= 0).all()  # every episode ends by the time limit at the latest
        np.testing.assert_array_equal(got_ret[k], ret)
        np.testing.assert_array_equal(got_len[k], ln)
        assert (ln == first + 1).all()
    assert float(metrics["steps_per_second"]) > 0


def test_evaluator_lbf_stepwise_matches_oracle_replay(lib_built):
    """Level-Based Foraging goes through the step-wise evaluator path (the fused rollout kernel is
    RobotWarehouse only): same check against the numpy oracle without AutoResetWrapper."""
    from mava_b200 import prng
    from mava_b200.config import compose
    from mava_b200.evaluator import get_eval_fn, make_ff_eval_act_fn
    from mava_b200.systems.ppo import ff_ippo
    from mava_b200.utils import make_env
    from oracle import lbf as olbf

    torch.cuda.set_device(0)
    time_limit, num_envs, episodes = 25, 6, 12
    cfg = compose(ff_ippo.CONFIG_NAME, [
        "env=lbf", "env/scenario=2s-8x8-2p-2f-coop", f"arch.num_envs={num_envs}",
        "system.rollout_length=8", f"env.kwargs.time_limit={time_limit}",
        f"arch.num_eval_episodes={episodes}", "logger.use_console=False"])
    env, eval_env = make_env.make(cfg, add_global_state=False)
    key, key_e, ak, ck = prng.split(prng.PRNGKey(23), 4)
    learn, _, _ = ff_ippo.learner_setup(env, (key, ak, ck), cfg)
    L = learn.learner
    evaluator = get_eval_fn(eval_env, make_ff_eval_act_fn(L.actor_desc, cfg), cfg,
                            absolute_metric=False)
    assert not evaluator.fused
    rec = {}
    metrics = evaluator(L.params[:L.na].clone(), key_e, {}, record=rec)
    loops = episodes // num_envs
    got_ret = metrics["episode_return"].cpu().numpy().reshape(loops, num_envs)
    got_len = metrics["episode_length"].cpu().numpy().reshape(loops, num_envs)
    spec = olbf.make_spec(time_limit=time_limit, **dict(cfg.env.scenario.task_config))
    oenv = olbf.MavaLbf(spec, add_global_state=False, add_agent_id=False, auto_reset=False)
    for k in range(loops):
        acts = rec["actions"][k]
        for e in range(num_envs):
            state, _ = oenv.reset(rec["reset_keys"][k][e])
            first = None
            for t in range(time_limit):
                state, ts = oenv.step(state, acts[t, e])
                if ts["done"] and first is None:
                    first = (ts["metrics"]["episode_return"], ts["metrics"]["episode_length"], t)
            assert first is not None
            assert got_ret[k, e] == np.float32(first[0]) and got_len[k, e] == first[1] == first[2] + 1
