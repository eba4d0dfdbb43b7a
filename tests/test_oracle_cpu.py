"""CPU tests of the oracle: published known-answer vectors, golden fixtures, invariants, C port."""
import json
import os

import numpy as np
import pytest

from oracle import ppo as oppo
from oracle import rware as orw
from oracle import threefry as tf

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_threefry_known_answers():
    kat = json.load(open(os.path.join(GOLD, "threefry_kat.json")))
    for v in kat["threefry2x32"]:
        k = [int(x, 16) for x in v["key"]]
        c = [int(x, 16) for x in v["count"]]
        o0, o1 = tf.threefry2x32(k[0], k[1], [c[0]], [c[1]])
        assert [hex(int(o0[0])), hex(int(o1[0]))] == v["out"]
    for v in kat["split"]:
        np.testing.assert_array_equal(tf.split(tf.prng_key(v["seed"])), np.array(v["out"], np.uint32))
    for v in kat["uniform_scalar"]:
        assert abs(float(tf.uniform(tf.prng_key(v["seed"]), ())) - v["out"]) < 1e-8


def test_product_prng_matches_oracle():
    from mava_b200 import prng

    for seed in (0, 42, 2**33 + 5):
        np.testing.assert_array_equal(prng.PRNGKey(seed), tf.prng_key(seed))
        for n in (2, 3, 4, 17):
            np.testing.assert_array_equal(prng.split(prng.PRNGKey(seed), n),
                                          tf.split(tf.prng_key(seed), n))


def test_permutation_is_a_permutation_and_stable():
    for n in (1, 2, 10, 110, 4096):
        p = tf.permutation(tf.prng_key(n), n)
        assert sorted(p.tolist()) == list(range(n))
    c = tf.choice_no_replace(tf.prng_key(3), np.arange(32), 4)
    assert len(set(c.tolist())) == 4
    r = tf.randint(tf.prng_key(1), (1000,), 0, 4)
    assert r.min() == 0 and r.max() == 3


def _replay(name, env_factory):
    g = np.load(os.path.join(GOLD, "rware_golden.npz"))
    keys, actions = g[f"{name}/keys"], g[f"{name}/actions"]
    return g, keys, actions


@pytest.mark.parametrize("name", ["tiny-2ag", "tiny-4ag", "small-4ag"])
def test_rware_oracle_reproduces_golden(name):
    from tests.golden.make_rware_golden import SCENARIOS

    g, keys, actions = _replay(name, None)
    spec = orw.make_spec(time_limit=25, **SCENARIOS[name])
    env = orw.MavaRware(spec, add_global_state=False, add_agent_id=False)
    states, ts = zip(*[env.reset(k) for k in keys])
    states = list(states)
    np.testing.assert_array_equal(np.stack([t["obs"]["agents_view"] for t in ts]).astype(np.int8),
                                  g[f"{name}/views"][0])
    for t in range(actions.shape[0]):
        res = [env.step(states[e], actions[t, e]) for e in range(len(keys))]
        states = [r[0] for r in res]
        np.testing.assert_array_equal(
            np.stack([r[1]["obs"]["agents_view"] for r in res]).astype(np.int8),
            g[f"{name}/views"][t + 1])
        np.testing.assert_array_equal(np.array([r[1]["done"] for r in res]), g[f"{name}/dones"][t])
        np.testing.assert_array_equal(np.stack([r[1]["reward"] for r in res]),
                                      g[f"{name}/rewards"][t])


@pytest.mark.parametrize("name", ["tiny-2ag", "tiny-4ag", "small-4ag"])
def test_c_port_reproduces_golden(name):
    from oracle.rware_c import RwareC
    from tests.golden.make_rware_golden import SCENARIOS

    g, keys, actions = _replay(name, None)
    env = RwareC(time_limit=25, **SCENARIOS[name])
    state, view, mask = env.reset(keys)
    np.testing.assert_array_equal(view, g[f"{name}/views"][0])
    bits = (g[f"{name}/masks"].astype(np.int64) << np.arange(5)).sum(-1).astype(np.uint8)
    np.testing.assert_array_equal(mask, bits[0])
    for t in range(actions.shape[0]):
        view, mask, rew, done, er, el = env.step(state, actions[t])
        np.testing.assert_array_equal(view, g[f"{name}/views"][t + 1], err_msg=f"t={t}")
        np.testing.assert_array_equal(mask, bits[t + 1])
        np.testing.assert_array_equal(rew, g[f"{name}/rewards"][t])
        np.testing.assert_array_equal(done.astype(bool), g[f"{name}/dones"][t])
        np.testing.assert_array_equal(er, g[f"{name}/ep_returns"][t])
        np.testing.assert_array_equal(el, g[f"{name}/ep_lengths"][t])


def test_rware_layout_and_invariants():
    tiny = orw.make_spec(num_agents=4, request_queue_size=4)
    small = orw.make_spec(shelf_rows=2)
    assert (tiny.H, tiny.W, tiny.n_shelves, tiny.num_obs_features) == (11, 10, 32, 66)
    assert (small.H, small.W, small.n_shelves) == (20, 10, 80)
    env = orw.MavaRware(tiny, add_global_state=True)
    st, ts = env.reset(tf.prng_key(9))
    assert ts["obs"]["agents_view"].shape == (4, 70)
    assert ts["obs"]["global_state"].shape == (4, 264)
    # ids are prepended; the global state carries no ids and is the same row for every agent
    np.testing.assert_array_equal(ts["obs"]["agents_view"][:, :4], np.eye(4))
    np.testing.assert_array_equal(ts["obs"]["global_state"][0],
                                  ts["obs"]["agents_view"][:, 4:].reshape(-1))
    rng = np.random.default_rng(0)
    for _ in range(300):
        st, ts = env.step(st, rng.integers(0, 5, 4))
        inner = st["inner"]
        assert inner["req"].sum() == tiny.Q and len(set(inner["queue"].tolist())) == tiny.Q
        assert (inner["req"][inner["queue"]] == 1).all()
        if not ts["done"]:  # no two agents share a cell inside an episode
            cells = set(zip(inner["ax"].tolist(), inner["ay"].tolist()))
            assert len(cells) == 4
        assert ts["obs"]["action_mask"][:, [0, 2, 3, 4]].all()


def test_time_limit_and_metrics():
    spec = orw.make_spec(num_agents=2, request_queue_size=2, time_limit=7)
    env = orw.MavaRware(spec, add_global_state=False)
    st, _ = env.reset(tf.prng_key(4))
    lens = []
    for t in range(30):
        st, ts = env.step(st, np.zeros(2, np.int64))  # noop never collides
        if ts["done"]:
            lens.append(ts["metrics"]["episode_length"])
            assert ts["obs"]["step_count"][0] == 0  # auto-reset: observation of the new episode
    assert lens == [7, 7, 7, 7]


def test_gae_against_closed_form():
    T, n = 5, 3
    r = np.ones((T, n), np.float32)
    v = np.zeros((T, n), np.float32)
    d = np.zeros((T, n), bool)
    adv, tgt = oppo.gae_ff(r, v, d, np.zeros(n, np.float32), 0.5, 1.0)
    np.testing.assert_allclose(adv[:, 0], [1.9375, 1.875, 1.75, 1.5, 1.0], rtol=1e-6)
    d[2] = True
    adv, _ = oppo.gae_ff(r, v, d, np.zeros(n, np.float32), 0.5, 1.0)
    np.testing.assert_allclose(adv[:, 0], [1.75, 1.5, 1.0, 1.5, 1.0], rtol=1e-6)


def test_clip_adam_first_step():
    p = np.zeros(4, np.float32)
    g = np.array([3.0, -4.0, 0.0, 0.0], np.float32)  # norm 5 -> clipped to 0.5
    p1, mu, nu = oppo.clip_adam(p, g, np.zeros(4, np.float32), np.zeros(4, np.float32), 0, 1e-3, 0.5)
    gc = g / 5 * 0.5
    np.testing.assert_allclose(mu, 0.1 * gc, rtol=1e-6)
    expect = -1e-3 * gc / (np.abs(gc) + 1e-5)
    np.testing.assert_allclose(p1[:2], expect[:2], rtol=1e-4)


@pytest.mark.parametrize("name", ["8x8-2p-2f-coop", "2s-10x10-3p-3f"])
def test_lbf_oracle_reproduces_golden(name):
    """tests/golden/lbf_golden.npz (make_oracle_golden.py, or make_reference_golden.py where the real
    jumanji exists) against the numpy LBF restatement."""
    from oracle import lbf as olbf
    from tests.golden import golden_inputs as gi

    g = np.load(os.path.join(GOLD, "lbf_golden.npz"))
    spec = olbf.make_spec(time_limit=gi.LBF_TIME_LIMIT, **gi.LBF_SCENARIOS[name])
    env = olbf.MavaLbf(spec, add_global_state=False, add_agent_id=False)
    keys, actions = g[f"{name}/keys"], g[f"{name}/actions"]
    states, ts = zip(*[env.reset(k) for k in keys])
    states = list(states)
    np.testing.assert_array_equal(np.stack([t["obs"]["agents_view"] for t in ts]).astype(np.int8),
                                  g[f"{name}/views"][0])
    np.testing.assert_array_equal(np.stack([t["obs"]["action_mask"] for t in ts]),
                                  g[f"{name}/masks"][0])
    for t in range(actions.shape[0]):
        res = [env.step(states[e], actions[t, e]) for e in range(len(keys))]
        states = [r[0] for r in res]
        np.testing.assert_array_equal(
            np.stack([r[1]["obs"]["agents_view"] for r in res]).astype(np.int8),
            g[f"{name}/views"][t + 1], err_msg=f"t={t}")
        np.testing.assert_array_equal(np.stack([r[1]["obs"]["action_mask"] for r in res]),
                                      g[f"{name}/masks"][t + 1])
        np.testing.assert_array_equal(np.stack([r[1]["reward"] for r in res]), g[f"{name}/rewards"][t])
        np.testing.assert_array_equal(np.array([r[1]["done"] for r in res]), g[f"{name}/dones"][t])
        np.testing.assert_array_equal(
            np.array([r[1]["metrics"]["episode_return"] for r in res], np.float32),
            g[f"{name}/ep_returns"][t])


def test_ppo_oracle_reproduces_golden():
    """tests/golden/ppo_golden.npz: jax.random.permutation, both GAE flavours, the losses and
    gradients of one ff_mappo minibatch and three optax steps, against oracle/ppo.py and
    oracle/threefry.py (float tolerances: the file may come from XLA, see make_reference_golden.py)."""
    import torch

    from oracle import ppo as oppo
    from tests.golden import golden_inputs as gi

    g = np.load(os.path.join(GOLD, "ppo_golden.npz"))
    for n in gi.PERM_SIZES:
        np.testing.assert_array_equal(tf.permutation(tf.prng_key(gi.PERM_SEED + n), n), g[f"perm/{n}"])
    x = gi.gae_inputs()
    A = x["reward"].shape[2]
    done_a = np.repeat(x["done"][:, :, None], A, 2)
    adv, tgt = oppo.gae_ff(x["reward"], x["value"], done_a, x["last_val"], x["gamma"], x["gae_lambda"])
    np.testing.assert_allclose(adv, g["gae/ff_adv"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(tgt, g["gae/ff_targets"], rtol=1e-5, atol=1e-6)
    adv, tgt = oppo.gae_rec(x["reward"], x["value"], done_a, x["last_val"],
                            np.repeat(x["last_done"][:, None], A, 1), x["gamma"], x["gae_lambda"])
    np.testing.assert_allclose(adv, g["gae/rec_adv"], rtol=1e-5, atol=1e-6)

    li = gi.loss_inputs()
    S, A, FR = li["view"].shape
    mk = lambda ps: [torch.tensor(p, dtype=torch.float64, requires_grad=True) for p in ps]  # noqa: E731
    at, ct = mk(li["actor"]), mk(li["critic"])
    lay = lambda t: [(t[0], t[1]), (t[2], t[3]), (t[4], t[5])]  # noqa: E731
    v = li["view"].astype(np.float64)
    xa = torch.tensor(np.concatenate([np.broadcast_to(np.eye(A), (S, A, A)), v], -1))
    xg = torch.tensor(np.repeat(v.reshape(S, 1, A * FR), A, 1))
    logits = oppo.actor_logits(lay(at), xa, torch.tensor(li["mask"]))
    ta, la, ent = oppo.actor_loss(logits, torch.tensor(li["action"]),
                                  torch.tensor(li["old_logp"], dtype=torch.float64),
                                  torch.tensor(li["adv"], dtype=torch.float64), 0.2, 0.01)
    val = oppo.critic_value(lay(ct), xg)
    tc, vl = oppo.critic_loss(val, torch.tensor(li["old_value"], dtype=torch.float64),
                              torch.tensor(li["targets"], dtype=torch.float64), 0.2, 0.5)
    ta.backward()
    tc.backward()
    legal = li["mask"]
    np.testing.assert_allclose(logits.detach().numpy()[legal], g["loss/logits"][legal], rtol=1e-4,
                               atol=1e-5)
    np.testing.assert_allclose(val.detach().numpy(), g["loss/value"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose([ta.item(), la.item(), ent.item(), tc.item(), vl.item()],
                               g["loss/scalars"], rtol=1e-4, atol=1e-6)
    for ts, key in ((at, "loss/actor_grad"), (ct, "loss/critic_grad")):
        got = np.concatenate([p.grad.numpy().ravel() for p in ts])
        np.testing.assert_allclose(got, g[key], rtol=1e-3, atol=1e-4 * np.abs(g[key]).max())
    ai = gi.adam_inputs()
    p, mu, nu = ai["params"].copy(), np.zeros_like(ai["params"]), np.zeros_like(ai["params"])
    for c, gr in enumerate(ai["grads"]):
        p, mu, nu = oppo.clip_adam(p, gr, mu, nu, c, float(ai["lr"]), float(ai["max_norm"]))
    np.testing.assert_allclose(p, g["adam/params"], rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(nu, g["adam/nu"], rtol=1e-5, atol=1e-12)
