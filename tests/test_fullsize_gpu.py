"""Size-independent properties at BASELINE.json's full sizes (the oracle cannot run there):

* RWARE env-step kernel at 2^18 small-4ag envs (config 5's regime): conservation and consistency
  invariants of the packed state and the emitted observations after hundreds of random steps with
  in-kernel auto-reset;
* the fused rollout kernel against the per-step kernels on 2048 envs x 128 steps (config 2's shape):
  bit-identical env trajectories when the same actions are replayed;
* GAE at 16.8 M elements: linearity in (reward, value) and the one-step recursion.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_rware_invariants_at_scale(lib_built):
    from mava_b200 import native, prng

    dev = torch.device("cuda:0")
    env = native.Env.rware(shelf_rows=2, num_agents=4, request_queue_size=4, time_limit=500)
    E, A, FR, n, Q = 1 << 18, 4, env.view_dim, env.dims.aux0, env.dims.aux1
    H, W = env.dims.grid_h, env.dims.grid_w
    keys = torch.from_numpy(prng.split(prng.PRNGKey(11), E).copy()).to(dev)
    state = env.alloc_state(E, dev)
    view = torch.zeros(E, A, FR, dtype=torch.int8, device=dev)
    mask = torch.zeros(E, A, dtype=torch.uint8, device=dev)
    reward = torch.zeros(E, A, device=dev)
    done = torch.zeros(E, dtype=torch.uint8, device=dev)
    ep_ret = torch.zeros(E, device=dev)
    ep_len = torch.zeros(E, dtype=torch.int32, device=dev)
    env.reset(keys, state, view, mask, E)
    g = torch.Generator(device=dev).manual_seed(5)
    total_done = total_reward = 0
    prev_len = torch.zeros(E, dtype=torch.int32, device=dev)
    for t in range(300):
        act = torch.randint(0, 5, (E, A), generator=g, device=dev, dtype=torch.int8)
        env.step(state, act, view, mask, reward, done, ep_ret, ep_len, E, True)
        total_done += int(done.sum())
        total_reward += float(reward[:, 0].sum())
        # shared reward: identical for the agents of an env, a small non-negative integer
        assert bool((reward == reward[:, :1]).all()) and bool((reward >= 0).all())
        assert bool((reward <= 2).all())
        # RecordEpisodeMetrics: the reported length only changes on terminal steps
        changed = ep_len != prev_len
        assert bool((changed <= done.bool()).all())
        prev_len = ep_len.clone()
        if t % 50 == 49 or t == 0:
            ag = env.peek(state, 2, E).reshape(E, A, 4)
            sh = env.peek(state, 3, E).reshape(E, n, 3)
            qu = env.peek(state, 4, E)
            st = env.peek(state, 0, E)[:, 0]
            x, y, d, carry = ag[..., 0], ag[..., 1], ag[..., 2], ag[..., 3]
            assert bool(((x >= 0) & (x < H) & (y >= 0) & (y < W) & (d >= 0) & (d < 4)).all())
            assert bool(((carry == 0) | (carry == 1)).all())
            cell = x * W + y
            srt = cell.sort(dim=1).values
            assert bool((srt[:, 1:] != srt[:, :-1]).all()), "two agents share a cell in a live state"
            # every shelf is on the grid exactly once (conservation), Q distinct requested shelves
            scell = sh[..., 0] * W + sh[..., 1]
            ssrt = scell.sort(dim=1).values
            assert bool((ssrt[:, 1:] != ssrt[:, :-1]).all()), "two shelves share a cell"
            assert bool((sh[..., 2].sum(1) == Q).all())
            qs = qu.sort(dim=1).values
            assert bool((qs[:, 1:] != qs[:, :-1]).all())
            assert bool(sh[..., 2].gather(1, qu.long()).bool().all()), "queue <-> requested flags"
            # a carrying agent stands on a shelf
            on_shelf = (cell.unsqueeze(-1) == scell.unsqueeze(1)).any(-1)
            assert bool((on_shelf | (carry == 0)).all())
            assert bool(((st >= 0) & (st < 500)).all())
            # observation header of every agent mirrors the state
            v = view.to(torch.int32)
            assert bool((v[..., 0] == x).all() and (v[..., 1] == y).all() and (v[..., 2] == carry).all())
            assert bool((v[..., 3:7].argmax(-1) == d).all() and (v[..., 3:7].sum(-1) == 1).all())
            # shelf block: 9 cells x (present, requested); the own cell entry matches `on_shelf`
            own = v[..., 48 + 2 * 4]
            assert bool((own.bool() == on_shelf).all())
            # mask: everything but FORWARD is always legal
            assert bool(((mask & 0x1D) == 0x1D).all())
    assert total_done > E // 10 and total_reward > 0, (total_done, total_reward)


def test_fused_rollout_equals_stepwise_at_config2_shape(lib_built):
    """2048 envs x 128 steps (BASELINE.json configs[1]): replaying the fused kernel's sampled
    actions through the stand-alone env kernel gives bit-identical trajectories."""
    from mava_b200 import prng
    from mava_b200.config import compose
    from mava_b200.systems.ppo import ff_mappo
    from mava_b200.utils import make_env

    torch.cuda.set_device(0)
    cfg = compose(ff_mappo.CONFIG_NAME, [
        "env/scenario=tiny-4ag", "arch.num_envs=1024", "system.update_batch_size=2",
        "system.rollout_length=128", "+arch.use_cuda_graph=False", "logger.use_console=False"])
    env, _ = make_env.make(cfg, add_global_state=True)
    key, _, ak, ck = prng.split(prng.PRNGKey(9), 4)
    learn, _, state = ff_mappo.learner_setup(env, (key, ak, ck), cfg)
    L = learn.learner
    assert L.fused_rollout
    buf0 = L.env_buf.clone()
    view0, mask0 = L.view[0].clone(), L.mask[0].clone()
    native_env = env.native
    L._pack()
    from mava_b200 import native

    native.prng_split_chain(L.key, L.policy_keys, L.T)
    L._rollout()
    torch.cuda.synchronize()
    # stepwise replay
    st = buf0.clone()
    v = torch.zeros_like(view0)
    m = torch.zeros_like(mask0)
    r = torch.zeros(L.NE, L.A, device=st.device)
    d = torch.zeros(L.NE, dtype=torch.uint8, device=st.device)
    er = torch.zeros(L.NE, device=st.device)
    el = torch.zeros(L.NE, dtype=torch.int32, device=st.device)
    for t in range(L.T):
        native_env.step(st, L.action[t], v, m, r, d, er, el, L.NE, True)
        assert torch.equal(v, L.view[t + 1]), t
        assert torch.equal(m, L.mask[t + 1]), t
        assert torch.equal(r, L.reward[t]) and torch.equal(d, L.done[t]), t
        assert torch.equal(er, L.ep_ret[t]) and torch.equal(el, L.ep_len[t]), t
    assert torch.equal(st, L.env_buf)
    # sampled actions are always legal, log-probs are finite and <= 0
    mk = L.mask[:L.T].to(torch.int32)
    assert bool((((mk >> L.action.to(torch.int32)) & 1) == 1).all())
    assert bool(torch.isfinite(L.logp).all()) and bool((L.logp <= 1e-6).all())
    assert int(L.done.sum()) > 0


def test_gae_linearity_and_recursion_at_scale(lib_built):
    from mava_b200 import native

    dev = torch.device("cuda:0")
    T, NE, A = 128, 32768, 4  # 16.8 M elements
    g = torch.Generator(device=dev).manual_seed(3)
    r1, r2 = (torch.randn(T, NE, A, generator=g, device=dev) for _ in range(2))
    v1, v2 = (torch.randn(T, NE, A, generator=g, device=dev) for _ in range(2))
    l1, l2 = (torch.randn(NE, A, generator=g, device=dev) for _ in range(2))
    done = (torch.rand(T, NE, generator=g, device=dev) < 0.02).to(torch.uint8)
    last_done = (torch.rand(NE, generator=g, device=dev) < 0.02).to(torch.uint8)
    gamma, lam = 0.99, 0.95

    def run(r, v, lv, rec):
        adv, tgt = torch.empty_like(r), torch.empty_like(r)
        native.gae(r, v, done, lv, gamma, lam, T, NE, A, adv, tgt,
                   **(dict(last_done=last_done) if rec else {}))
        return adv, tgt

    for rec in (False, True):
        a1, t1 = run(r1, v1, l1, rec)
        a2, _ = run(r2, v2, l2, rec)
        a12, t12 = run(r1 + r2, v1 + v2, l1 + l2, rec)
        scale = float(a12.abs().max())
        assert float((a12 - (a1 + a2)).abs().max()) < 2e-5 * scale       # linear in (r, V)
        assert torch.allclose(t1, a1 + v1, rtol=0, atol=1e-5 * scale)   # targets = adv + V
        # one-step recursion: gae_t = delta_t + gamma * lambda * m_t * gae_{t+1}
        nd = 1.0 - done.float()
        if rec:  # the flag entering step t+1 gates the bootstrap of step t
            m = torch.cat([nd[1:], (1.0 - last_done.float()).unsqueeze(0)], 0).unsqueeze(-1)
        else:
            m = nd.unsqueeze(-1)
        nv = torch.cat([v1[1:], l1.unsqueeze(0)], 0)
        na = torch.cat([a1[1:], torch.zeros_like(a1[:1])], 0)
        want = r1 + gamma * nv * m - v1 + gamma * lam * m * na
        assert float((a1 - want).abs().max()) < 2e-5 * float(a1.abs().max())
