"""GPU parity of the fp32 actor/critic kernels (acting, values, PPO loss and gradients)."""
import numpy as np
import pytest
import torch

from oracle import ppo as oppo
from oracle import threefry as tf

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def make_params(rng, in_dim, h1, h2, out, scale_out=0.3):
    shapes = [(in_dim, h1), (h1,), (h1, h2), (h2,), (h2, out), (out,)]
    ps = []
    for i, s in enumerate(shapes):
        sc = scale_out if i >= 4 else (1.0 / np.sqrt(s[0]) if len(s) == 2 else 0.1)
        ps.append((rng.normal(size=s) * sc).astype(np.float32))
    return ps


def flat(ps):
    return np.concatenate([p.ravel() for p in ps]).astype(np.float32)


def as_layers(ps, dtype):
    t = [torch.tensor(p, dtype=dtype, requires_grad=True) for p in ps]
    return t, [(t[0], t[1]), (t[2], t[3]), (t[4], t[5])]


def build_inputs(view, A, add_id, mode):
    """agents_view / global_state exactly as the wrapper stack presents them."""
    S = view.shape[0]
    v = view.astype(np.float64)
    if mode == "global":
        return np.repeat(v.reshape(S, 1, -1), A, axis=1)  # (S, A, A*FR)
    if add_id:
        ids = np.broadcast_to(np.eye(A), (S, A, A))
        return np.concatenate([ids, v], -1)
    return v


def random_batch(rng, S, A, FR, N):
    view = rng.integers(0, 12, size=(S, A, FR)).astype(np.int8)
    mask_bool = rng.random((S, A, N)) < 0.7
    mask_bool[..., 0] = True
    mask = (mask_bool.astype(np.int64) << np.arange(N)).sum(-1).astype(np.uint8)
    return view, mask_bool, mask


@pytest.mark.parametrize("A,FR,N,h,critic_mode,add_id", [
    (4, 66, 5, 128, "global", True),
    (2, 66, 5, 128, "agent", True),
    (3, 12, 6, 64, "agent", False),
])
def test_act_and_value(lib_built, A, FR, N, h, critic_mode, add_id):
    from mava_b200 import native

    rng = np.random.default_rng(0)
    NE, E = 150, 75
    view, mask_bool, mask = random_batch(rng, NE, A, FR, N)
    actor = native.mlp_desc(native.IN_AGENT_VIEW, add_id, A, FR, h, h, N)
    cmode = native.IN_GLOBAL if critic_mode == "global" else native.IN_AGENT_VIEW
    critic = native.mlp_desc(cmode, add_id, A, FR, h, h, 1)
    ap = make_params(rng, actor.in_dim, h, h, N)
    cp = make_params(rng, critic.in_dim, h, h, 1)
    key = tf.prng_key(99)

    action = torch.zeros(NE, A, dtype=torch.int8, device=DEV)
    logp = torch.zeros(NE, A, device=DEV)
    value = torch.zeros(NE, A, device=DEV)
    native.ff_act(actor, torch.from_numpy(flat(ap)).to(DEV), critic,
                  torch.from_numpy(flat(cp)).to(DEV), torch.from_numpy(view).to(DEV),
                  torch.from_numpy(mask).to(DEV),
                  torch.from_numpy(key.copy()).to(DEV), E, NE, action, logp, value)

    _, al = as_layers(ap, torch.float64)
    _, cl = as_layers(cp, torch.float64)
    x = torch.tensor(build_inputs(view, A, add_id, "agent"))
    logits = oppo.actor_logits(al, x, torch.tensor(mask_bool)).detach()
    xc = torch.tensor(build_inputs(view, A, add_id, critic_mode))
    v = oppo.critic_value(cl, xc).detach().numpy()
    np.testing.assert_allclose(value.cpu().numpy(), v, rtol=1e-5, atol=1e-5)

    # sampling: same threefry noise layout (E, A, N) reused by each replica
    g = tf.gumbel(key, (E, A, N))
    g = np.concatenate([g] * (NE // E), 0)
    lg32 = logits.numpy().astype(np.float32)
    exp_action = np.argmax(g + lg32, -1)
    got = action.cpu().numpy()
    # fp32 rounding may flip near-ties: require agreement except where the top-2 gap is tiny
    z = np.sort(g + lg32, -1)
    safe = (z[..., -1] - z[..., -2]) > 1e-4
    assert safe.mean() > 0.99
    np.testing.assert_array_equal(got[safe], exp_action[safe])
    lp = torch.log_softmax(logits, -1).numpy()
    exp_lp = np.take_along_axis(lp, got[..., None].astype(np.int64), -1)[..., 0]
    np.testing.assert_allclose(logp.cpu().numpy(), exp_lp, rtol=1e-5, atol=1e-5)
    assert (np.take_along_axis(mask_bool, got[..., None].astype(np.int64), -1)).all()

    # greedy and replay
    native.ff_act(actor, torch.from_numpy(flat(ap)).to(DEV), None, None,
                  torch.from_numpy(view).to(DEV), torch.from_numpy(mask).to(DEV), None, E, NE,
                  action, logp, None, greedy=True)
    gl = np.sort(lg32, -1)
    safe = (gl[..., -1] - gl[..., -2]) > 1e-4
    np.testing.assert_array_equal(action.cpu().numpy()[safe], np.argmax(lg32, -1)[safe])
    replay = torch.from_numpy(exp_action.astype(np.int8)).to(DEV)
    native.ff_act(actor, torch.from_numpy(flat(ap)).to(DEV), None, None,
                  torch.from_numpy(view).to(DEV), torch.from_numpy(mask).to(DEV), None, E, NE,
                  action, logp, None, actions_in=replay)
    np.testing.assert_array_equal(action.cpu().numpy(), exp_action)


@pytest.mark.parametrize("A,FR,N,h,critic_mode,U,mb", [
    (4, 66, 5, 128, "global", 2, 96),
    (2, 66, 5, 128, "agent", 1, 200),
    (3, 12, 6, 64, "agent", 2, 67),
])
def test_ppo_loss_grad(lib_built, A, FR, N, h, critic_mode, U, mb):
    from mava_b200 import native
    from mava_b200._lib import PpoHyper

    rng = np.random.default_rng(1)
    T, E = 12, 40
    NE = U * E
    S = T * NE
    add_id = True
    view, mask_bool, mask = random_batch(rng, S, A, FR, N)
    actor = native.mlp_desc(native.IN_AGENT_VIEW, add_id, A, FR, h, h, N)
    cmode = native.IN_GLOBAL if critic_mode == "global" else native.IN_AGENT_VIEW
    critic = native.mlp_desc(cmode, add_id, A, FR, h, h, 1)
    ap = make_params(rng, actor.in_dim, h, h, N)
    cp = make_params(rng, critic.in_dim, h, h, 1)
    # legal old actions, plausible old log-probs/values
    action = np.zeros((S, A), np.int8)
    for idx in np.ndindex(S, A):
        action[idx] = rng.choice(np.flatnonzero(mask_bool[idx]))
    old_logp = (-rng.random((S, A)) * 2.0).astype(np.float32)
    old_value = rng.normal(size=(S, A)).astype(np.float32)
    adv = rng.normal(size=(S, A)).astype(np.float32)
    targets = (old_value + rng.normal(size=(S, A)) * 0.5).astype(np.float32)
    perm = rng.permutation(T * E).astype(np.int32)
    hyper = PpoHyper(0.2, 0.01, 0.5)

    rows = torch.zeros(U * mb, dtype=torch.int32, device=DEV)
    native.ppo_minibatch_rows(torch.from_numpy(perm).to(DEV), 1, mb, U, E, rows)
    na, nc = native.mlp_param_count(actor), native.mlp_param_count(critic)
    grad = torch.zeros(na + nc + 8, device=DEV)
    ws = torch.zeros(native.ppo_workspace_bytes(actor, critic, U * mb), dtype=torch.uint8, device=DEV)
    native.ppo_loss_grad(actor, torch.from_numpy(flat(ap)).to(DEV), critic,
                         torch.from_numpy(flat(cp)).to(DEV), hyper, torch.from_numpy(view).to(DEV),
                         torch.from_numpy(mask).to(DEV), torch.from_numpy(action).to(DEV),
                         torch.from_numpy(old_logp).to(DEV), torch.from_numpy(old_value).to(DEV),
                         torch.from_numpy(adv).to(DEV), torch.from_numpy(targets).to(DEV), rows, U,
                         mb, grad, ws)
    grad = grad.cpu().numpy()

    # oracle: per-replica value_and_grad, then the mean over replicas (pmean "batch")
    r = rows.cpu().numpy().reshape(U, mb)
    at, al = as_layers(ap, torch.float64)
    ct, cl = as_layers(cp, torch.float64)
    tot_a = tot_c = 0.0
    infos = np.zeros(5)
    for u in range(U):
        idx = r[u]
        x = torch.tensor(build_inputs(view[idx], A, add_id, "agent"))
        logits = oppo.actor_logits(al, x, torch.tensor(mask_bool[idx]))
        ta, la, ent = oppo.actor_loss(logits, torch.tensor(action[idx]),
                                      torch.tensor(old_logp[idx], dtype=torch.float64),
                                      torch.tensor(adv[idx], dtype=torch.float64), 0.2, 0.01)
        xc = torch.tensor(build_inputs(view[idx], A, add_id, critic_mode))
        val = oppo.critic_value(cl, xc)
        tc, vl = oppo.critic_loss(val, torch.tensor(old_value[idx], dtype=torch.float64),
                                  torch.tensor(targets[idx], dtype=torch.float64), 0.2, 0.5)
        tot_a = tot_a + ta / U
        tot_c = tot_c + tc / U
        infos += np.array([ta.item(), la.item(), ent.item(), tc.item(), vl.item()]) / U
    tot_a.backward()
    tot_c.backward()
    ga = np.concatenate([p.grad.numpy().ravel() for p in at])
    gc = np.concatenate([p.grad.numpy().ravel() for p in ct])
    # tolerance: rtol 1e-5 fp32 on the scale of each gradient block
    np.testing.assert_allclose(grad[na + nc:na + nc + 5], infos, rtol=2e-5, atol=1e-6)
    for got, exp, name in ((grad[:na], ga, "actor"), (grad[na:na + nc], gc, "critic")):
        scale = np.abs(exp).max()
        np.testing.assert_allclose(got, exp, rtol=1e-4, atol=2e-5 * scale, err_msg=name)
