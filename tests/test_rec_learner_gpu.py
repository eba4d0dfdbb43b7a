"""End-to-end parity of one recurrent Anakin update (rec_ippo / rec_mappo) against the oracle:
the sampled actions are replayed through the numpy env oracle (bit-exact), the GRU networks,
GAE (next_done flavour), the chunked minibatches and the optimiser through the torch float64
restatement of mava/systems/ppo/rec_mappo.py."""
import numpy as np
import pytest
import torch

from oracle import ppo as oppo
from oracle import rware as orw
from oracle import threefry as tf

pytestmark = pytest.mark.gpu
F32_MIN = float(np.finfo(np.float32).min)


@pytest.mark.parametrize("system,chunk,use_graph", [
    ("rec_mappo", None, False), ("rec_ippo", 4, False), ("rec_mappo", 8, True)])
def test_one_recurrent_update_matches_oracle(lib_built, system, chunk, use_graph):
    import importlib

    from mava_b200 import prng
    from mava_b200.config import compose
    from mava_b200.utils import make_env

    torch.cuda.set_device(0)
    mod = importlib.import_module(f"mava_b200.systems.ppo.{system}")
    over = ["env/scenario=tiny-2ag", "arch.num_envs=8", "system.rollout_length=16",
            "system.ppo_epochs=2", "system.num_minibatches=2", "system.update_batch_size=2",
            "env.kwargs.time_limit=10", f"+arch.use_cuda_graph={use_graph}",
            "+arch.precision=fp32", "network.hidden_state_dim=32", "network.actor_network.pre_torso.layer_sizes=[32]",
            "network.actor_network.post_torso.layer_sizes=[24]",
            "network.critic_network.pre_torso.layer_sizes=[32]",
            "network.critic_network.post_torso.layer_sizes=[24]"]
    if chunk is not None:
        over.append(f"system.recurrent_chunk_size={chunk}")
    cfg = compose(f"default_{system}.yaml", over)
    central = system == "rec_mappo"
    env, _ = make_env.make(cfg, add_global_state=central)
    key, _, ak, ck = prng.split(prng.PRNGKey(5), 4)
    learn, actor_net, state = mod.learner_setup(env, (key, ak, ck), cfg)
    L = learn.learner
    T, U, E, NE, A, FR, N, H = L.T, L.U, L.E, L.NE, L.A, L.FR, L.N, L.H
    chunk = L.chunk
    nc = T // chunk
    cfg.system.num_updates_per_eval = 1
    p0 = L.params.cpu().numpy().copy()
    key0 = L.key.cpu().numpy().copy()
    view0 = L.view[0].cpu().numpy().copy()
    all_keys = tf.split(key, U * E + 1)
    spec = orw.make_spec(**dict(cfg.env.scenario.task_config), time_limit=10)
    oenv = orw.MavaRware(spec, add_global_state=False, add_agent_id=False)
    ostates, ots0 = zip(*[oenv.reset(all_keys[1 + e]) for e in range(NE)])
    ostates = list(ostates)

    out = learn(state)
    torch.cuda.synchronize()

    # ---- rollout: replay the sampled actions through the oracle env (bit-exact)
    act = L.action.cpu().numpy()
    views = L.view.cpu().numpy()
    o_views = [view0]
    o_masks = [np.stack([t_["obs"]["action_mask"] for t_ in ots0])]
    rew = np.zeros((T, NE, A), np.float32)
    don = np.zeros((T + 1, NE), bool)  # flag ENTERING each step; slot T = last_done
    for t in range(T):
        res = [oenv.step(ostates[e], act[t, e]) for e in range(NE)]
        ostates = [r[0] for r in res]
        o_views.append(np.stack([r[1]["obs"]["agents_view"] for r in res]).astype(np.int8))
        o_masks.append(np.stack([r[1]["obs"]["action_mask"] for r in res]))
        rew[t] = np.stack([r[1]["reward"] for r in res])
        don[t + 1] = [r[1]["done"] for r in res]
        np.testing.assert_array_equal(views[t + 1], o_views[-1], err_msg=f"view t={t}")
    np.testing.assert_array_equal(L.reward.cpu().numpy(), rew)
    np.testing.assert_array_equal(L.done_in.cpu().numpy()[1:].astype(bool), don[1:])
    o_views, o_masks = np.stack(o_views), np.stack(o_masks)

    # ---- networks along the rollout
    def actor_in(v):
        ids = np.broadcast_to(np.eye(A), v.shape[:-2] + (A, A))
        return torch.tensor(np.concatenate([ids, v.astype(np.float64)], -1))

    def critic_in(v):
        if central:
            return torch.tensor(v.astype(np.float64).reshape(v.shape[:-2] + (1, A * FR)))
        return actor_in(v)

    rpc = 1 if central else A
    Q = L.actor_desc.post

    def nets(flat):
        pa_flat = torch.tensor(flat[:L.na], dtype=torch.float64, requires_grad=True)
        pc_flat = torch.tensor(flat[L.na:], dtype=torch.float64, requires_grad=True)
        return (pa_flat, pc_flat, oppo.rnn_unflatten(pa_flat, L.actor_desc.in_dim, H, Q, N),
                oppo.rnn_unflatten(pc_flat, L.critic_desc.in_dim, H, Q, 1))

    _, _, pa, pc = nets(p0)
    with torch.no_grad():
        xa = actor_in(o_views).reshape(T + 1, NE * A, -1)
        xc = critic_in(o_views).reshape(T + 1, NE * rpc, -1)
        ra = torch.tensor(np.repeat(don, A, 1))
        rc = torch.tensor(np.repeat(don, rpc, 1))
        ha, hc = torch.zeros(NE * A, H, dtype=torch.float64), torch.zeros(NE * rpc, H, dtype=torch.float64)
        hs_a, hs_c, lps, vals = [], [], [], []
        for t in range(T):
            hs_a.append(ha)
            hs_c.append(hc)
            ha, lg = oppo.rec_net(pa, ha, xa[t:t + 1], ra[t:t + 1])
            hc, v = oppo.rec_net(pc, hc, xc[t:t + 1], rc[t:t + 1])
            lg = torch.where(torch.tensor(o_masks[t]).reshape(1, NE * A, N), lg,
                             torch.full_like(lg, F32_MIN))
            lps.append(oppo.categorical_log_prob(lg, torch.tensor(act[t].reshape(1, -1).astype(np.int64))))
            vals.append(v)
        _, lv = oppo.rec_net(pc, hc, xc[T:T + 1], rc[T:T + 1])
    lp = torch.cat(lps).reshape(T, NE, A).numpy()
    val = torch.cat(vals).reshape(T, NE, rpc).numpy()
    val = np.repeat(val, A, 2) if rpc == 1 else val
    last_val = lv.reshape(NE, rpc).numpy()
    last_val = np.repeat(last_val, A, 1) if rpc == 1 else last_val
    np.testing.assert_allclose(L.logp.cpu().numpy(), lp, rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(L.value.cpu().numpy(), val, rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(L.last_val.cpu().numpy(), last_val, rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(L.h_actor.cpu().numpy(), ha.numpy(), rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(L.hs_actor.cpu().numpy(), torch.stack(hs_a[:nc]).numpy(), rtol=1e-4,
                               atol=2e-5)

    # ---- GAE, next_done flavour
    gval = L.value.cpu().numpy()
    adv, tgt = oppo.gae_rec(rew, gval, np.repeat(don[:T, :, None], A, 2), L.last_val.cpu().numpy(),
                            np.repeat(don[T][:, None], A, 1), cfg.system.gamma, cfg.system.gae_lambda)
    np.testing.assert_allclose(L.adv.cpu().numpy(), adv, rtol=1e-5, atol=2e-5)
    np.testing.assert_allclose(L.targets.cpu().numpy(), tgt, rtol=1e-5, atol=2e-5)

    # ---- epochs: reference key schedule, chunk reshape, permutation over columns
    k = key0
    for _ in range(T):
        k = tf.split(k)[0]
    params = p0.copy()
    mu, nu = np.zeros_like(params), np.zeros_like(params)
    cnt, na = 0, L.na
    g_logp, g_adv, g_tgt = (torch.tensor(x.cpu().numpy(), dtype=torch.float64)
                            for x in (L.logp, L.adv, L.targets))
    g_val = torch.tensor(gval, dtype=torch.float64)
    t_act = torch.tensor(act.astype(np.int64))
    t_mask = torch.tensor(o_masks[:T])
    t_done = torch.tensor(don[:T])
    HA = torch.stack(hs_a).reshape(T, NE, A, H)
    HC = torch.stack(hs_c).reshape(T, NE, rpc, H)
    XA, XC = actor_in(o_views[:T]), critic_in(o_views[:T])
    mbc = E * nc // 2
    losses = []
    for ep in range(2):
        ks = tf.split(k, 3)
        k, shuffle_key = ks[0], ks[1]
        perm = tf.permutation(shuffle_key, E * nc)
        for m in range(2):
            pa_flat, pc_flat, pa, pc = nets(params)
            cols = torch.tensor(perm[m * mbc:(m + 1) * mbc].astype(np.int64))
            tot_a = tot_c = 0.0
            info = np.zeros(5)
            for u in range(U):
                sl = slice(u * E, (u + 1) * E)
                cb = lambda x: oppo.rec_chunk_batch(x[:, sl], chunk, cols)
                xa, xc, dn = cb(XA), cb(XC), cb(t_done)
                mb = xa.shape[1]
                # the stored hidden states of THIS rollout came from p0; the reference re-uses
                # them as constants for every epoch (rec_mappo.py:221,254)
                h0a, h0c = cb(HA)[0].reshape(mb * A, H), cb(HC)[0].reshape(mb * rpc, H)
                _, lg = oppo.rec_net(pa, h0a, xa.reshape(chunk, mb * A, -1),
                                     dn[:, :, None].expand(chunk, mb, A).reshape(chunk, -1))
                _, v = oppo.rec_net(pc, h0c, xc.reshape(chunk, mb * rpc, -1),
                                    dn[:, :, None].expand(chunk, mb, rpc).reshape(chunk, -1))
                lg = lg.reshape(chunk, mb, A, N)
                lg = torch.where(cb(t_mask), lg, torch.full_like(lg, F32_MIN))
                v = v.reshape(chunk, mb, rpc)
                v = v.expand(chunk, mb, A) if rpc == 1 else v
                ta, la, en = oppo.actor_loss(lg, cb(t_act), cb(g_logp), cb(g_adv),
                                             cfg.system.clip_eps, cfg.system.ent_coef)
                tc, vl = oppo.critic_loss(v, cb(g_val), cb(g_tgt), cfg.system.clip_eps,
                                          cfg.system.vf_coef)
                tot_a, tot_c = tot_a + ta / U, tot_c + tc / U
                info += np.array([ta.item(), la.item(), en.item(), tc.item(), vl.item()]) / U
            ga, = torch.autograd.grad(tot_a, pa_flat)
            gc, = torch.autograd.grad(tot_c, pc_flat)
            params[:na], mu[:na], nu[:na] = oppo.clip_adam(
                params[:na], ga.numpy().astype(np.float32), mu[:na], nu[:na], cnt,
                cfg.system.actor_lr, 0.5)
            params[na:], mu[na:], nu[na:] = oppo.clip_adam(
                params[na:], gc.numpy().astype(np.float32), mu[na:], nu[na:], cnt,
                cfg.system.critic_lr, 0.5)
            cnt += 1
            losses.append(info)
    np.testing.assert_array_equal(L.key.cpu().numpy(), k)
    got = L.params.cpu().numpy()
    moved = np.abs(params - p0).max()
    assert moved > 1e-4
    np.testing.assert_allclose(got, params, rtol=0, atol=0.02 * moved)
    assert np.mean(np.abs(got - params) < 1e-3 * moved) > 0.99
    losses = np.array(losses).reshape(2, 2, 5)
    tm = out.train_metrics
    np.testing.assert_allclose(tm["actor_loss"][0].cpu().numpy(), losses[..., 1], rtol=2e-4, atol=1e-6)
    np.testing.assert_allclose(tm["entropy"][0].cpu().numpy(), losses[..., 2], rtol=2e-4, atol=1e-6)
    np.testing.assert_allclose(tm["value_loss"][0].cpu().numpy(), losses[..., 4], rtol=2e-4, atol=1e-6)
    em = out.episode_metrics
    assert em["episode_return"].shape == (1, U, T, E)
    np.testing.assert_array_equal(
        em["is_terminal_step"][0].permute(1, 0, 2).reshape(T, NE).cpu().numpy(), don[1:])


def test_run_recurrent_experiment_smoke(lib_built):
    """test/integration_test.py:35-46 for the recurrent systems."""
    from mava_b200.config import compose
    from mava_b200.systems.ppo import rec_ippo, rec_mappo

    for mod in (rec_ippo, rec_mappo):
        cfg = compose(mod.CONFIG_NAME, [
            "env/scenario=tiny-2ag", "arch.num_envs=4", "system.rollout_length=8",
            "system.num_updates=4", "arch.num_evaluation=2", "arch.num_eval_episodes=4",
            "arch.num_absolute_metric_eval_episodes=4", "env.kwargs.time_limit=12",
            "logger.use_console=False", "network.hidden_state_dim=32",
            "network.actor_network.pre_torso.layer_sizes=[32]",
            "network.critic_network.pre_torso.layer_sizes=[32]"])
        perf = mod.run_experiment(cfg)
        assert isinstance(perf, float) and np.isfinite(perf)


def test_recurrent_learner_on_synthetic_smax_shapes(lib_built):
    """env=smax_synthetic (benchmark-only step source): dense f32 observations, uint16 masks for 13
    actions, centralised critic on the world-state row.  Checks the plumbing: finite losses, masked
    actions, parameters move, episode metrics follow RecordEpisodeMetrics."""
    from mava_b200 import prng
    from mava_b200.config import compose
    from mava_b200.systems.ppo import rec_mappo
    from mava_b200.utils import make_env

    torch.cuda.set_device(0)
    cfg = compose(rec_mappo.CONFIG_NAME, [
        "env=smax_synthetic", "arch.num_envs=16", "system.rollout_length=8",
        "system.update_batch_size=2", "system.ppo_epochs=2", "env.kwargs.time_limit=6",
        "env.scenario.task_config.obs_dim=21", "env.scenario.task_config.state_dim=17",
        "network.hidden_state_dim=32", "network.actor_network.pre_torso.layer_sizes=[32]",
        "network.critic_network.pre_torso.layer_sizes=[32]", "+arch.use_cuda_graph=True"])
    env, _ = make_env.make(cfg, add_global_state=True)
    key, _, ak, ck = prng.split(prng.PRNGKey(0), 4)
    learn, _, state = rec_mappo.learner_setup(env, (key, ak, ck), cfg)
    L = learn.learner
    cfg.system.num_updates_per_eval = 2
    p0 = L.params.clone()
    out = learn(state)
    torch.cuda.synchronize()
    for name in ("total_loss", "value_loss", "actor_loss", "entropy"):
        assert torch.isfinite(out.train_metrics[name]).all()
    assert float((L.params - p0).abs().max()) > 1e-5
    act = L.action.cpu().numpy().astype(np.int64)
    mk = L.mask[:L.T].cpu().numpy().astype(np.int64)
    # slot 0 was overwritten with slot T at the end of the update: compare steps 1 .. T-1
    assert (((mk[1:] >> act[1:]) & 1) == 1).all(), "a masked action was sampled"
    assert ((mk & 0x1F) == 0x1F).all() and (mk < (1 << 13)).all()
    oa = L.obs_a.cpu().numpy()
    assert 0.45 < oa.mean() < 0.55 and oa.min() >= 0.0 and oa.max() < 1.0
    em = out.episode_metrics
    assert bool(em["is_terminal_step"].any())
    done_len = em["episode_length"][em["is_terminal_step"]]
    assert (done_len >= 1).all()
