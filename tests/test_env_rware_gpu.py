"""Parity (bit-exact) of the fused RWARE env kernels against the numpy oracle on replayed actions."""
import numpy as np
import pytest
import torch

from oracle import rware as orw
from oracle import threefry as tf

pytestmark = pytest.mark.gpu

SCENARIOS = {
    "tiny-2ag": dict(column_height=8, shelf_rows=1, shelf_columns=3, num_agents=2, sensor_range=1,
                     request_queue_size=2),
    "tiny-4ag": dict(column_height=8, shelf_rows=1, shelf_columns=3, num_agents=4, sensor_range=1,
                     request_queue_size=4),
    "small-4ag": dict(column_height=8, shelf_rows=2, shelf_columns=3, num_agents=4, sensor_range=1,
                      request_queue_size=4),
    "tiny-6ag-q8": dict(column_height=8, shelf_rows=1, shelf_columns=3, num_agents=6,
                        sensor_range=1, request_queue_size=8),
}


def _mask_bits(mask_bool):
    return (mask_bool.astype(np.int64) << np.arange(mask_bool.shape[-1])).sum(-1).astype(np.uint8)


def _compare_state(env, state, ostates, NE):
    A = env.num_agents
    ag = env.peek(state, 2, NE).cpu().numpy().reshape(NE, A, 4)
    sh = env.peek(state, 3, NE).cpu().numpy().reshape(NE, env.dims.aux0, 3)
    qu = env.peek(state, 4, NE).cpu().numpy()
    ky = env.peek(state, 1, NE).cpu().numpy().astype(np.uint32)
    st = env.peek(state, 0, NE).cpu().numpy()[:, 0]
    for e, os_ in enumerate(ostates):
        inner = os_["inner"]
        np.testing.assert_array_equal(ag[e, :, 0], inner["ax"], err_msg=f"env {e} ax")
        np.testing.assert_array_equal(ag[e, :, 1], inner["ay"], err_msg=f"env {e} ay")
        np.testing.assert_array_equal(ag[e, :, 2], inner["adir"], err_msg=f"env {e} dir")
        np.testing.assert_array_equal(ag[e, :, 3], inner["carry"], err_msg=f"env {e} carry")
        np.testing.assert_array_equal(sh[e, :, 0], inner["sx"], err_msg=f"env {e} sx")
        np.testing.assert_array_equal(sh[e, :, 1], inner["sy"], err_msg=f"env {e} sy")
        np.testing.assert_array_equal(sh[e, :, 2], inner["req"], err_msg=f"env {e} req")
        np.testing.assert_array_equal(qu[e], inner["queue"], err_msg=f"env {e} queue")
        np.testing.assert_array_equal(ky[e], inner["key"], err_msg=f"env {e} key")
        assert st[e] == inner["step"]


@pytest.mark.parametrize("name,NE,T,time_limit", [
    ("tiny-4ag", 131, 120, 40),
    ("tiny-2ag", 70, 150, 60),
    ("small-4ag", 65, 100, 500),
    ("tiny-6ag-q8", 37, 80, 30),
])
def test_rware_step_matches_oracle(lib_built, name, NE, T, time_limit):
    from mava_b200 import native

    dev = torch.device("cuda:0")
    cfg = dict(SCENARIOS[name], time_limit=time_limit)
    spec = orw.make_spec(**cfg)
    oenv = orw.MavaRware(spec, add_global_state=False, add_agent_id=False)
    env = native.Env.rware(**cfg)
    A, FR = env.num_agents, env.view_dim
    assert FR == spec.num_obs_features and A == spec.A

    keys = tf.split(tf.prng_key(7), NE)
    state = env.alloc_state(NE, dev)
    view = torch.zeros(NE, A, FR, dtype=torch.int8, device=dev)
    mask = torch.zeros(NE, A, dtype=torch.uint8, device=dev)
    env.reset(torch.from_numpy(keys.copy()).to(dev), state, view, mask, NE)
    ostates, ots = zip(*[oenv.reset(keys[e]) for e in range(NE)])
    ostates = list(ostates)
    oview = np.stack([t["obs"]["agents_view"] for t in ots]).astype(np.int8)
    omask = np.stack([_mask_bits(t["obs"]["action_mask"]) for t in ots])
    np.testing.assert_array_equal(view.cpu().numpy(), oview)
    np.testing.assert_array_equal(mask.cpu().numpy(), omask)
    _compare_state(env, state, ostates, NE)

    rng = np.random.default_rng(3)
    reward = torch.zeros(NE, A, dtype=torch.float32, device=dev)
    done = torch.zeros(NE, dtype=torch.uint8, device=dev)
    ep_ret = torch.zeros(NE, dtype=torch.float32, device=dev)
    ep_len = torch.zeros(NE, dtype=torch.int32, device=dev)
    n_done = n_rew = 0
    for t in range(T):
        # biased towards moving so that shelves get carried to the goals
        act = rng.choice(5, size=(NE, A), p=[0.05, 0.5, 0.15, 0.15, 0.15]).astype(np.int8)
        env.step(state, torch.from_numpy(act).to(dev), view, mask, reward, done, ep_ret, ep_len, NE,
                 auto_reset=True)
        res = [oenv.step(ostates[e], act[e]) for e in range(NE)]
        ostates = [r[0] for r in res]
        ots = [r[1] for r in res]
        np.testing.assert_array_equal(
            view.cpu().numpy(), np.stack([x["obs"]["agents_view"] for x in ots]).astype(np.int8),
            err_msg=f"view t={t}")
        np.testing.assert_array_equal(
            mask.cpu().numpy(), np.stack([_mask_bits(x["obs"]["action_mask"]) for x in ots]),
            err_msg=f"mask t={t}")
        np.testing.assert_array_equal(reward.cpu().numpy(), np.stack([x["reward"] for x in ots]))
        od = np.array([x["done"] for x in ots])
        np.testing.assert_array_equal(done.cpu().numpy().astype(bool), od)
        np.testing.assert_array_equal(
            ep_ret.cpu().numpy(), np.array([x["metrics"]["episode_return"] for x in ots], np.float32))
        np.testing.assert_array_equal(
            ep_len.cpu().numpy(), np.array([x["metrics"]["episode_length"] for x in ots], np.int32))
        n_done += int(od.sum())
        n_rew += int(sum(x["reward"][0] for x in ots))
    _compare_state(env, state, ostates, NE)
    assert n_done > 0, "the replay never exercised the in-kernel auto-reset"
    print(f"{name}: {n_done} episode ends, {n_rew} deliveries")


def test_rware_eval_env_does_not_reset(lib_built):
    """auto_reset=0 is the evaluation env (make_env.py:79-81): after done the state keeps going."""
    from mava_b200 import native

    dev = torch.device("cuda:0")
    cfg = dict(SCENARIOS["tiny-2ag"], time_limit=5)
    spec = orw.make_spec(**cfg)
    oenv = orw.MavaRware(spec, add_global_state=False, add_agent_id=False, auto_reset=False)
    env = native.Env.rware(**cfg)
    NE, A, FR = 9, env.num_agents, env.view_dim
    keys = tf.split(tf.prng_key(1), NE)
    state = env.alloc_state(NE, dev)
    view = torch.zeros(NE, A, FR, dtype=torch.int8, device=dev)
    mask = torch.zeros(NE, A, dtype=torch.uint8, device=dev)
    env.reset(torch.from_numpy(keys.copy()).to(dev), state, view, mask, NE)
    ostates = [oenv.reset(keys[e])[0] for e in range(NE)]
    reward = torch.zeros(NE, A, dtype=torch.float32, device=dev)
    done = torch.zeros(NE, dtype=torch.uint8, device=dev)
    ep_ret = torch.zeros(NE, dtype=torch.float32, device=dev)
    ep_len = torch.zeros(NE, dtype=torch.int32, device=dev)
    rng = np.random.default_rng(0)
    for t in range(8):
        act = rng.integers(0, 5, size=(NE, A)).astype(np.int8)
        env.step(state, torch.from_numpy(act).to(dev), view, mask, reward, done, ep_ret, ep_len, NE,
                 auto_reset=False)
        res = [oenv.step(ostates[e], act[e]) for e in range(NE)]
        ostates = [r[0] for r in res]
        np.testing.assert_array_equal(
            view.cpu().numpy(),
            np.stack([r[1]["obs"]["agents_view"] for r in res]).astype(np.int8))
        np.testing.assert_array_equal(done.cpu().numpy().astype(bool),
                                      np.array([r[1]["done"] for r in res]))
