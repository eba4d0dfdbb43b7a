"""The CUDA kernels against the committed golden files DIRECTLY (no oracle in between):
tests/golden/lbf_golden.npz and ppo_golden.npz -- written by the oracle today, by the real
jax + jumanji + mava stack once tests/golden/make_reference_golden.py has been run somewhere that has
it (the files record which in `generator`).  RobotWarehouse: tests/test_bf16_path_gpu.py::
test_golden_file_on_gpu."""
import os

import numpy as np
import pytest
import torch

from tests.golden import golden_inputs as gi

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", sorted(gi.LBF_SCENARIOS))
def test_lbf_golden_on_gpu(lib_built, name):
    from mava_b200 import native

    g = np.load(os.path.join(GOLD, "lbf_golden.npz"))
    env = native.Env.lbf(time_limit=gi.LBF_TIME_LIMIT, **gi.LBF_SCENARIOS[name])
    keys, actions = g[f"{name}/keys"], g[f"{name}/actions"]
    E, A, FR, N = keys.shape[0], env.num_agents, env.view_dim, env.num_actions
    bits = (g[f"{name}/masks"].astype(np.int64) << np.arange(N)).sum(-1).astype(np.uint8)
    state = env.alloc_state(E, DEV)
    view = torch.zeros(E, A, FR, dtype=torch.int8, device=DEV)
    mask = torch.zeros(E, A, dtype=torch.uint8, device=DEV)
    reward = torch.zeros(E, A, device=DEV)
    done = torch.zeros(E, dtype=torch.uint8, device=DEV)
    er, el = torch.zeros(E, device=DEV), torch.zeros(E, dtype=torch.int32, device=DEV)
    env.reset(torch.from_numpy(keys.astype(np.uint32)).to(DEV), state, view, mask, E)
    np.testing.assert_array_equal(view.cpu().numpy(), g[f"{name}/views"][0])
    np.testing.assert_array_equal(mask.cpu().numpy(), bits[0])
    for t in range(actions.shape[0]):
        env.step(state, torch.from_numpy(actions[t]).to(DEV), view, mask, reward, done, er, el, E, True)
        np.testing.assert_array_equal(view.cpu().numpy(), g[f"{name}/views"][t + 1], err_msg=f"t={t}")
        np.testing.assert_array_equal(mask.cpu().numpy(), bits[t + 1], err_msg=f"t={t}")
        np.testing.assert_array_equal(reward.cpu().numpy(), g[f"{name}/rewards"][t])
        np.testing.assert_array_equal(done.cpu().numpy().astype(bool), g[f"{name}/dones"][t])
        np.testing.assert_array_equal(er.cpu().numpy(), g[f"{name}/ep_returns"][t])
        np.testing.assert_array_equal(el.cpu().numpy(), g[f"{name}/ep_lengths"][t])


def test_ppo_golden_on_gpu(lib_built):
    """Permutation (threefry bits + stable sort), GAE (both flavours), the fp32 loss / gradient
    kernels and the optimiser kernel against ppo_golden.npz."""
    import math

    from mava_b200 import native
    from mava_b200._lib import PpoHyper
    from mava_b200.peer import PeerGroup

    g = np.load(os.path.join(GOLD, "ppo_golden.npz"))
    dev = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(DEV)  # noqa: E731
    # ---- jax.random.permutation: rounds of sort_key_val(random_bits(subkey), x)
    from oracle import threefry as tf  # key splitting on the host only (mava_b200.prng is the same)

    for n in gi.PERM_SIZES:
        rounds = int(math.ceil(3 * math.log(max(1, n)) / math.log(2 ** 32 - 1)))
        key = dev(tf.prng_key(gi.PERM_SEED + n))
        src = torch.arange(n, dtype=torch.int32, device=DEV)
        ws = torch.zeros(native.sort_workspace_bytes(n), dtype=torch.uint8, device=DEV)
        ovf = torch.zeros(1, dtype=torch.int32, device=DEV)
        bits = torch.zeros(n, dtype=torch.uint32, device=DEV)
        k2 = torch.zeros(2, 2, dtype=torch.uint32, device=DEV)
        for _ in range(rounds):
            native.prng_split(key, k2, 2)
            key = k2[0].clone()
            native.prng_random_bits(k2[1], bits, n)
            dst = torch.empty_like(src)
            native.sort_by_key(bits, src, dst, n, ws, ovf)
            src = dst
        assert int(ovf.item()) == 0
        np.testing.assert_array_equal(src.cpu().numpy(), g[f"perm/{n}"])
    # ---- GAE
    x = gi.gae_inputs()
    T, NE, A = x["reward"].shape
    adv, tgt = torch.zeros(T, NE, A, device=DEV), torch.zeros(T, NE, A, device=DEV)
    native.gae(dev(x["reward"]), dev(x["value"]), dev(x["done"].astype(np.uint8)), dev(x["last_val"]),
               float(x["gamma"]), float(x["gae_lambda"]), T, NE, A, adv, tgt)
    np.testing.assert_allclose(adv.cpu().numpy(), g["gae/ff_adv"], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(tgt.cpu().numpy(), g["gae/ff_targets"], rtol=1e-5, atol=2e-6)
    native.gae(dev(x["reward"]), dev(x["value"]), dev(x["done"].astype(np.uint8)), dev(x["last_val"]),
               float(x["gamma"]), float(x["gae_lambda"]), T, NE, A, adv, tgt,
               last_done=dev(x["last_done"].astype(np.uint8)))
    np.testing.assert_allclose(adv.cpu().numpy(), g["gae/rec_adv"], rtol=1e-5, atol=2e-6)
    # ---- losses and gradients (fp32 kernels), one replica, the whole batch as one minibatch
    li = gi.loss_inputs()
    S, A, FR = li["view"].shape
    N = li["mask"].shape[-1]
    actor = native.mlp_desc(native.IN_AGENT_VIEW, True, A, FR, 128, 128, N)
    critic = native.mlp_desc(native.IN_GLOBAL, True, A, FR, 128, 128, 1)
    flat = lambda ps: np.concatenate([p.ravel() for p in ps]).astype(np.float32)  # noqa: E731
    na, nc = native.mlp_param_count(actor), native.mlp_param_count(critic)
    mask = (li["mask"].astype(np.int64) << np.arange(N)).sum(-1).astype(np.uint8)
    rows = torch.arange(S, dtype=torch.int32, device=DEV)
    grad = torch.zeros(na + nc + 8, device=DEV)
    ws = torch.zeros(native.ppo_workspace_bytes(actor, critic, S), dtype=torch.uint8, device=DEV)
    native.ppo_loss_grad(actor, dev(flat(li["actor"])), critic, dev(flat(li["critic"])),
                         PpoHyper(float(li["clip_eps"]), float(li["ent_coef"]), float(li["vf_coef"])),
                         dev(li["view"]), dev(mask), dev(li["action"].astype(np.int8)),
                         dev(li["old_logp"]), dev(li["old_value"]), dev(li["adv"]),
                         dev(li["targets"]), rows, 1, S, grad, ws)
    got = grad.cpu().numpy()
    np.testing.assert_allclose(got[na + nc:na + nc + 5], g["loss/scalars"], rtol=1e-4, atol=1e-6)
    for sl, key in ((slice(0, na), "loss/actor_grad"), (slice(na, na + nc), "loss/critic_grad")):
        np.testing.assert_allclose(got[sl], g[key], rtol=1e-3, atol=1e-4 * np.abs(g[key]).max())
    # ---- optax.chain(clip_by_global_norm, adam): three steps of the optimiser kernel
    ai = gi.adam_inputs()
    n = ai["params"].size
    grp = PeerGroup(2 * n + 8, torch.device(DEV))
    p = dev(np.concatenate([ai["params"], ai["params"]]))
    mu, nu, gsum = torch.zeros(2 * n, device=DEV), torch.zeros(2 * n, device=DEV), torch.zeros(2 * n, device=DEV)
    counts = torch.zeros(2, dtype=torch.int32, device=DEV)
    for gr in ai["grads"]:
        grp.grad.copy_(dev(np.concatenate([gr, gr, np.zeros(8, np.float32)])))
        native.reduce_clip_adam_pair(p, mu, nu, counts, grp, gsum, n, n, None, None, None, None, 1.0,
                                     float(ai["lr"]), float(ai["lr"]), float(ai["max_norm"]))
    for half in (slice(0, n), slice(n, 2 * n)):  # both "networks" got the same data
        np.testing.assert_allclose(p[half].cpu().numpy(), g["adam/params"], rtol=1e-6, atol=1e-8)
        # (the clipped gradient g / ||g|| * max_norm carries a rounding of ~1e-6 into mu, twice that
        # into nu = g^2 terms)
        np.testing.assert_allclose(mu[half].cpu().numpy(), g["adam/mu"], rtol=5e-5, atol=1e-10)
        np.testing.assert_allclose(nu[half].cpu().numpy(), g["adam/nu"], rtol=1e-4, atol=1e-12)
    grp.release()
