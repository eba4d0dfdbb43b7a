"""Writes tests/golden/{rware,lbf,ppo}_golden.npz from the REAL reference stack.

    python tests/golden/make_reference_golden.py [--mava /path/to/Mava] [--out tests/golden]

Needs jax (0.4.30 per requirements/requirements.txt:9-10), jumanji (the sash-a fork,
requirements.txt:12), flax, optax, tensorflow_probability and the reference repository itself
(`--mava`, default /root/reference, then baseline/_ref).  None of these is installable in the build
image (no network), so THIS SCRIPT HAS NOT BEEN RUN THERE: it is the pin-ready half of
the oracle.  It writes exactly the keys tests/golden/make_oracle_golden.py and
make_rware_golden.py write, from the same seeded inputs (tests/golden/golden_inputs.py); after
running it,

    python -m pytest tests/test_oracle_cpu.py -q          # the restatement against the real stack
    python -m pytest tests -m gpu -q -k golden           # the CUDA kernels against the real stack

turn "parity unpinned" into reference-pinned parity, or show exactly where the restatement of
Jumanji (oracle/rware.py, oracle/lbf.py) deviates.  The npz files carry `generator = "reference"`
then.

What is taken from where:
* env trajectories: `jumanji.make(...)` wrapped exactly like mava/utils/make_env.py:69-116
  (RwareWrapper | LbfWrapper -> AutoResetWrapper -> RecordEpisodeMetrics; AgentIDWrapper off because
  the files store the raw integer view), `jax.vmap(env.reset)` / `jax.vmap(env.step)`;
* `jax.random.permutation` (ff_mappo.py:273);
* GAE: the scan of ff_mappo.py:112-139 / rec_mappo.py:177-199 (nested functions there, restated
  here line by line in jax so that XLA's float semantics are the reference's);
* losses and gradients: mava.networks.FeedForwardActor / FeedForwardValueNet (flax Dense +
  tfd.Categorical masking) applied to parameters given in flax naming, `_actor_loss_fn` /
  `_critic_loss_fn` of ff_mappo.py:150-201 under jax.value_and_grad;
* optimiser: optax.chain(optax.clip_by_global_norm, optax.adam(eps=1e-5)) (ff_mappo.py:359-366).
"""
from __future__ import annotations

import argparse
import importlib
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from tests.golden import golden_inputs as gi  # noqa: E402


def _probe(mava_path: str):
    missing = []
    for name in ("jax", "jumanji", "flax", "optax", "tensorflow_probability", "chex"):
        try:
            importlib.import_module(name)
        except Exception as e:  # noqa: BLE001
            missing.append(f"{name} ({type(e).__name__})")
    for cand in (mava_path, "/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "mava")):
            sys.path.insert(0, cand)
            break
    else:
        missing.append("the reference repository (a directory holding mava/)")
    return missing


def _load_wrappers():
    """mava.wrappers.{jumanji,auto_reset_wrapper,episode_metrics} without mava/wrappers/__init__.py,
    which imports gigastep / jaxmarl / matrax (not needed for RWARE / LBF)."""
    import mava  # noqa: F401

    pkg_dir = os.path.join(os.path.dirname(importlib.import_module("mava").__file__), "wrappers")
    if "mava.wrappers" not in sys.modules:
        pkg = types.ModuleType("mava.wrappers")
        pkg.__path__ = [pkg_dir]
        sys.modules["mava.wrappers"] = pkg
    jm = importlib.import_module("mava.wrappers.jumanji")
    ar = importlib.import_module("mava.wrappers.auto_reset_wrapper")
    em = importlib.import_module("mava.wrappers.episode_metrics")
    return jm.RwareWrapper, jm.LbfWrapper, ar.AutoResetWrapper, em.RecordEpisodeMetrics


def _rollout(env, keys, actions):
    import jax
    import jax.numpy as jnp

    reset, step = jax.jit(jax.vmap(env.reset)), jax.jit(jax.vmap(env.step))
    state, ts = reset(jnp.asarray(keys))
    views, masks = [np.asarray(ts.observation.agents_view)], [np.asarray(ts.observation.action_mask)]
    rewards, dones, rets, lens = [], [], [], []
    for t in range(actions.shape[0]):
        state, ts = step(state, jnp.asarray(actions[t], jnp.int32))
        views.append(np.asarray(ts.observation.agents_view))
        masks.append(np.asarray(ts.observation.action_mask))
        rewards.append(np.asarray(ts.reward, np.float32))
        dones.append(np.asarray(ts.last()))
        m = ts.extras["episode_metrics"]
        rets.append(np.asarray(m["episode_return"], np.float32))
        lens.append(np.asarray(m["episode_length"], np.int32))
    return dict(views=np.stack(views).astype(np.int8), masks=np.stack(masks).astype(bool),
                rewards=np.stack(rewards), dones=np.stack(dones).astype(bool),
                ep_returns=np.stack(rets), ep_lengths=np.stack(lens))


def env_goldens():
    import jax
    import jumanji
    from jumanji.environments.routing.lbf.generator import RandomGenerator as LbfGen
    from jumanji.environments.routing.robot_warehouse.generator import RandomGenerator as RwareGen

    RwareWrapper, LbfWrapper, AutoReset, Metrics = _load_wrappers()
    rware, lbf = {}, {}
    for name, task in gi.RWARE_SCENARIOS.items():
        env = jumanji.make("RobotWarehouse-v0", generator=RwareGen(**task),
                           time_limit=gi.RWARE_TIME_LIMIT)
        env = Metrics(AutoReset(RwareWrapper(env, add_global_state=False)))
        keys = np.asarray(jax.random.split(jax.random.PRNGKey(gi.RWARE_SEED), gi.RWARE_NE))
        actions = gi.rware_actions(task["num_agents"])
        rware.update({f"{name}/keys": keys.astype(np.uint32), f"{name}/actions": actions})
        rware.update({f"{name}/{k}": v for k, v in _rollout(env, keys, actions).items()})
    for name, task in gi.LBF_SCENARIOS.items():
        env = jumanji.make("LevelBasedForaging-v0", generator=LbfGen(**task),
                           time_limit=gi.LBF_TIME_LIMIT)
        env = Metrics(AutoReset(LbfWrapper(env, add_global_state=False)))
        keys = np.asarray(jax.random.split(jax.random.PRNGKey(gi.LBF_SEED), gi.LBF_NE))
        actions = gi.lbf_actions(task["num_agents"])
        lbf.update({f"{name}/keys": keys.astype(np.uint32), f"{name}/actions": actions})
        lbf.update({f"{name}/{k}": v for k, v in _rollout(env, keys, actions).items()})
    return rware, lbf


def ppo_golden():
    import jax
    import jax.numpy as jnp
    import optax

    from mava.networks import FeedForwardActor, FeedForwardValueNet  # flax modules
    from mava.types import Observation, ObservationGlobalState

    out = {}
    for n in gi.PERM_SIZES:
        out[f"perm/{n}"] = np.asarray(
            jax.random.permutation(jax.random.PRNGKey(gi.PERM_SEED + n), n), np.int32)

    # ---- GAE: ff_mappo.py:112-139 and rec_mappo.py:177-199
    g = gi.gae_inputs()
    A = g["reward"].shape[2]
    done_a = jnp.repeat(jnp.asarray(g["done"])[:, :, None], A, 2)
    gamma, lam = float(g["gamma"]), float(g["gae_lambda"])

    def gae_ff(reward, value, done, last_val):
        def step(carry, x):
            gae, next_value = carry
            d, v, r = x
            delta = r + gamma * next_value * (1 - d) - v
            gae = delta + gamma * lam * (1 - d) * gae
            return (gae, v), gae
        _, adv = jax.lax.scan(step, (jnp.zeros_like(last_val), last_val), (done, value, reward),
                              reverse=True, unroll=16)
        return adv, adv + value

    def gae_rec(reward, value, done, last_val, last_done):
        def step(carry, x):
            gae, next_value, next_done = carry
            d, v, r = x
            delta = r + gamma * next_value * (1 - next_done) - v
            gae = delta + gamma * lam * (1 - next_done) * gae
            return (gae, v, d), gae
        _, adv = jax.lax.scan(step, (jnp.zeros_like(last_val), last_val, last_done),
                              (done, value, reward), reverse=True, unroll=16)
        return adv, adv + value

    adv, tgt = jax.jit(gae_ff)(jnp.asarray(g["reward"]), jnp.asarray(g["value"]),
                               done_a.astype(jnp.float32), jnp.asarray(g["last_val"]))
    out["gae/ff_adv"], out["gae/ff_targets"] = np.asarray(adv), np.asarray(tgt)
    last_done = jnp.repeat(jnp.asarray(g["last_done"])[:, None], A, 1).astype(jnp.float32)
    adv, tgt = jax.jit(gae_rec)(jnp.asarray(g["reward"]), jnp.asarray(g["value"]),
                                done_a.astype(jnp.float32), jnp.asarray(g["last_val"]), last_done)
    out["gae/rec_adv"], out["gae/rec_targets"] = np.asarray(adv), np.asarray(tgt)

    # ---- losses + gradients through the reference's own network modules
    li = gi.loss_inputs()
    S, A, FR = li["view"].shape

    from mava.networks import DiscreteActionHead, MLPTorso

    N = li["mask"].shape[-1]
    actor = FeedForwardActor(torso=MLPTorso((128, 128)), action_head=DiscreteActionHead(action_dim=N))
    critic = FeedForwardValueNet(torso=MLPTorso((128, 128)), centralised_critic=True)
    v = jnp.asarray(li["view"], jnp.float32)
    agents_view = jnp.concatenate([jnp.broadcast_to(jnp.eye(A), (S, A, A)), v], -1)
    global_state = jnp.repeat(v.reshape(S, 1, A * FR), A, 1)
    step_count = jnp.zeros((S, A), jnp.int32)
    obs = Observation(agents_view, jnp.asarray(li["mask"]), step_count)
    gobs = ObservationGlobalState(agents_view, jnp.asarray(li["mask"]), global_state, step_count)
    # the parameter tree names are whatever flax assigns: take the structure from init and fill it
    # leaf by leaf in flax's deterministic order (kernel / bias of Dense_0, Dense_1, head)
    a_init = actor.init(jax.random.PRNGKey(0), obs)
    c_init = critic.init(jax.random.PRNGKey(0), gobs)

    def fill(init_tree, ps):
        leaves, treedef = jax.tree_util.tree_flatten_with_path(init_tree)
        by_shape = {}
        for p in ps:
            by_shape.setdefault(p.shape, []).append(jnp.asarray(p))
        new = []
        for path, leaf in leaves:  # kernels and biases have distinct shapes per layer here
            cands = by_shape[tuple(leaf.shape)]
            new.append(cands.pop(0))
        return jax.tree_util.tree_unflatten(treedef, new), [jax.tree_util.keystr(p) for p, _ in leaves]

    a_params, a_names = fill(a_init, li["actor"])
    c_params, c_names = fill(c_init, li["critic"])
    out["flax/actor_leaf_paths"] = np.array(a_names)
    out["flax/critic_leaf_paths"] = np.array(c_names)
    clip_eps, ent_coef, vf_coef = float(li["clip_eps"]), float(li["ent_coef"]), float(li["vf_coef"])

    def actor_loss(params):  # ff_mappo.py:159-180
        pi = actor.apply(params, obs)
        log_prob = pi.log_prob(jnp.asarray(li["action"]))
        ratio = jnp.exp(log_prob - jnp.asarray(li["old_logp"]))
        gae = jnp.asarray(li["adv"])
        gae = (gae - gae.mean()) / (gae.std() + 1e-8)
        l1 = ratio * gae
        l2 = jnp.clip(ratio, 1.0 - clip_eps, 1.0 + clip_eps) * gae
        loss = -jnp.minimum(l1, l2).mean()
        entropy = pi.entropy(seed=jax.random.PRNGKey(0)).mean()
        return loss - ent_coef * entropy, (loss, entropy, pi.logits if hasattr(pi, "logits") else 0)

    def critic_loss(params):  # ff_mappo.py:190-201
        value = critic.apply(params, gobs)
        vo, tg = jnp.asarray(li["old_value"]), jnp.asarray(li["targets"])
        v_clip = vo + (value - vo).clip(-clip_eps, clip_eps)
        vl = 0.5 * jnp.maximum((value - tg) ** 2, (v_clip - tg) ** 2).mean()
        return vf_coef * vl, (vl, value)

    (ta, (la, ent, logits)), ga = jax.value_and_grad(actor_loss, has_aux=True)(a_params)
    (tc, (vl, value)), gc = jax.value_and_grad(critic_loss, has_aux=True)(c_params)
    out["loss/logits"] = np.asarray(logits, np.float32)
    out["loss/value"] = np.asarray(value, np.float32)
    out["loss/scalars"] = np.array([float(ta), float(la), float(ent), float(tc), float(vl)])
    # gradients in the flat order of include/mava_b200.h (Dense_0 kernel, bias, Dense_1 ..., head)
    order = lambda tree: np.concatenate(  # noqa: E731
        [np.asarray(x).ravel() for _, x in sorted(
            jax.tree_util.tree_flatten_with_path(tree)[0],
            key=lambda kv: ("head" in jax.tree_util.keystr(kv[0]).lower()
                            or "critic_head" in jax.tree_util.keystr(kv[0]).lower(),
                            jax.tree_util.keystr(kv[0]).replace("kernel", "a").replace("bias", "b")))])
    out["loss/actor_grad"] = order(ga).astype(np.float32)
    out["loss/critic_grad"] = order(gc).astype(np.float32)

    # ---- optimiser: ff_mappo.py:359-366
    ai = gi.adam_inputs()
    opt = optax.chain(optax.clip_by_global_norm(float(ai["max_norm"])),
                      optax.adam(float(ai["lr"]), eps=1e-5))
    p = jnp.asarray(ai["params"])
    st = opt.init(p)
    for gr in ai["grads"]:
        upd, st = opt.update(jnp.asarray(gr), st)
        p = optax.apply_updates(p, upd)
    adam_state = st[1][0]
    out["adam/params"] = np.asarray(p)
    out["adam/mu"], out["adam/nu"] = np.asarray(adam_state.mu), np.asarray(adam_state.nu)
    return out


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--mava", default="/root/reference")
    ap.add_argument("--out", default=HERE)
    args = ap.parse_args()
    missing = _probe(args.mava)
    if missing:
        print("make_reference_golden: the reference stack is not importable here:\n  - "
              + "\n  - ".join(missing)
              + "\nThe committed golden files were written by the oracle "
                "(tests/golden/make_oracle_golden.py, make_rware_golden.py): parity stays unpinned.")
        return 2
    rware, lbf = env_goldens()
    files = {"rware_golden.npz": rware, "lbf_golden.npz": lbf, "ppo_golden.npz": ppo_golden()}
    for fname, data in files.items():
        path = os.path.join(args.out, fname)
        np.savez_compressed(path, generator=np.array("reference"), **data)
        print("wrote", path, os.path.getsize(path), "bytes")
    return 0


if __name__ == "__main__":
    sys.exit(main())
