"""Parity (bit-exact) of the fused LBF env kernels against the numpy oracle on replayed actions."""
import numpy as np
import pytest
import torch

from oracle import lbf as olbf
from oracle import threefry as tf

pytestmark = pytest.mark.gpu

SCENARIOS = {
    "8x8-2p-2f-coop": dict(grid_size=8, fov=8, num_agents=2, num_food=2, max_agent_level=2,
                           force_coop=True),
    "2s-8x8-2p-2f-coop": dict(grid_size=8, fov=2, num_agents=2, num_food=2, max_agent_level=2,
                              force_coop=True),
    "2s-10x10-3p-3f": dict(grid_size=10, fov=2, num_agents=3, num_food=3, max_agent_level=2,
                           force_coop=False),
    "15x15-4p-5f": dict(grid_size=15, fov=15, num_agents=4, num_food=5, max_agent_level=2,
                        force_coop=False),
}


def _bits(mask_bool):
    return (mask_bool.astype(np.int64) << np.arange(mask_bool.shape[-1])).sum(-1).astype(np.uint8)


def _compare_state(env, state, ostates, NE):
    A, NF = env.num_agents, env.dims.aux0
    ag = env.peek(state, 2, NE).cpu().numpy().reshape(NE, A, 4)
    fo = env.peek(state, 3, NE).cpu().numpy().reshape(NE, NF, 3)
    ea = env.peek(state, 4, NE).cpu().numpy()
    ky = env.peek(state, 1, NE).cpu().numpy().astype(np.uint32)
    st = env.peek(state, 0, NE).cpu().numpy()[:, 0]
    for e, os_ in enumerate(ostates):
        inner = os_["inner"]
        np.testing.assert_array_equal(ag[e, :, 0], inner["ax"], err_msg=f"env {e} ax")
        np.testing.assert_array_equal(ag[e, :, 1], inner["ay"], err_msg=f"env {e} ay")
        np.testing.assert_array_equal(ag[e, :, 2], inner["alvl"], err_msg=f"env {e} level")
        np.testing.assert_array_equal(fo[e, :, 0], inner["fx"])
        np.testing.assert_array_equal(fo[e, :, 1], inner["fy"])
        np.testing.assert_array_equal(fo[e, :, 2], inner["flvl"])
        np.testing.assert_array_equal(ea[e].astype(bool), inner["eaten"])
        np.testing.assert_array_equal(ky[e], inner["key"])
        assert st[e] == inner["step"]


@pytest.mark.parametrize("name,NE,T,time_limit,individual", [
    ("8x8-2p-2f-coop", 131, 150, 30, False),
    ("2s-8x8-2p-2f-coop", 70, 120, 25, False),
    ("2s-10x10-3p-3f", 65, 120, 40, True),
    ("15x15-4p-5f", 37, 100, 50, False),
])
def test_lbf_step_matches_oracle(lib_built, name, NE, T, time_limit, individual):
    from mava_b200 import native

    dev = torch.device("cuda:0")
    cfg = dict(SCENARIOS[name], time_limit=time_limit)
    spec = olbf.make_spec(**cfg)
    oenv = olbf.MavaLbf(spec, add_global_state=False, add_agent_id=False,
                        use_individual_rewards=individual)
    env = native.Env.lbf(use_individual_rewards=individual, **cfg)
    A, FR = env.num_agents, env.view_dim
    assert FR == spec.num_obs_features and env.num_actions == 6

    keys = tf.split(tf.prng_key(11), NE)
    state = env.alloc_state(NE, dev)
    view = torch.zeros(NE, A, FR, dtype=torch.int8, device=dev)
    mask = torch.zeros(NE, A, dtype=torch.uint8, device=dev)
    env.reset(torch.from_numpy(keys.copy()).to(dev), state, view, mask, NE)
    ostates, ots = zip(*[oenv.reset(keys[e]) for e in range(NE)])
    ostates = list(ostates)
    np.testing.assert_array_equal(
        view.cpu().numpy(), np.stack([t["obs"]["agents_view"] for t in ots]).astype(np.int8))
    np.testing.assert_array_equal(
        mask.cpu().numpy(), np.stack([_bits(t["obs"]["action_mask"]) for t in ots]))
    _compare_state(env, state, ostates, NE)

    rng = np.random.default_rng(5)
    reward = torch.zeros(NE, A, dtype=torch.float32, device=dev)
    done = torch.zeros(NE, dtype=torch.uint8, device=dev)
    ep_ret = torch.zeros(NE, dtype=torch.float32, device=dev)
    ep_len = torch.zeros(NE, dtype=torch.int32, device=dev)
    omask = np.stack([t["obs"]["action_mask"] for t in ots])
    n_done = 0
    total_reward = 0.0
    for t in range(T):
        # mostly legal actions (loads when possible), sometimes arbitrary ones
        act = np.zeros((NE, A), np.int8)
        for e in range(NE):
            for a in range(A):
                legal = np.flatnonzero(omask[e, a])
                if omask[e, a, 5] and rng.random() < 0.7:
                    act[e, a] = 5
                elif rng.random() < 0.9:
                    act[e, a] = rng.choice(legal)
                else:
                    act[e, a] = rng.integers(0, 6)
        env.step(state, torch.from_numpy(act).to(dev), view, mask, reward, done, ep_ret, ep_len, NE,
                 auto_reset=True)
        res = [oenv.step(ostates[e], act[e]) for e in range(NE)]
        ostates = [r[0] for r in res]
        ots = [r[1] for r in res]
        omask = np.stack([x["obs"]["action_mask"] for x in ots])
        np.testing.assert_array_equal(
            view.cpu().numpy(), np.stack([x["obs"]["agents_view"] for x in ots]).astype(np.int8),
            err_msg=f"view t={t}")
        np.testing.assert_array_equal(mask.cpu().numpy(), np.stack([_bits(m) for m in omask]),
                                      err_msg=f"mask t={t}")
        np.testing.assert_array_equal(reward.cpu().numpy(), np.stack([x["reward"] for x in ots]),
                                      err_msg=f"reward t={t}")
        od = np.array([x["done"] for x in ots])
        np.testing.assert_array_equal(done.cpu().numpy().astype(bool), od)
        np.testing.assert_array_equal(
            ep_ret.cpu().numpy(), np.array([x["metrics"]["episode_return"] for x in ots], np.float32))
        np.testing.assert_array_equal(
            ep_len.cpu().numpy(), np.array([x["metrics"]["episode_length"] for x in ots], np.int32))
        n_done += int(od.sum())
        total_reward += float(sum(x["reward"].sum() for x in ots))
    _compare_state(env, state, ostates, NE)
    assert n_done > 0 and total_reward > 0, "the replay exercised neither eating nor auto-reset"
    print(f"{name}: {n_done} episode ends, reward {total_reward:.2f}")
