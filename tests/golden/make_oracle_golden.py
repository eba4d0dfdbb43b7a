"""Writes tests/golden/{lbf_golden,ppo_golden}.npz from the CPU restatement (oracle/).

These are SELF-CONSISTENCY pins (regression pins for the oracle, the C port and the CUDA kernels):
the reference's own tests hold no golden vectors (test/integration_test.py:35-46) and the real
jax + jumanji stack is not installable in this image.  tests/golden/make_reference_golden.py writes
the same files from the real stack where it exists.

    python tests/golden/make_oracle_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import lbf as olbf  # noqa: E402
from oracle import ppo as oppo  # noqa: E402
from oracle import threefry as tf  # noqa: E402
from tests.golden import golden_inputs as gi  # noqa: E402


def lbf_golden():
    out = {}
    for name, task in gi.LBF_SCENARIOS.items():
        spec = olbf.make_spec(time_limit=gi.LBF_TIME_LIMIT, **task)
        env = olbf.MavaLbf(spec, add_global_state=False, add_agent_id=False)
        keys = tf.split(tf.prng_key(gi.LBF_SEED), gi.LBF_NE)
        actions = gi.lbf_actions(spec.A)
        states, ts = zip(*[env.reset(k) for k in keys])
        states = list(states)
        views = [np.stack([t["obs"]["agents_view"] for t in ts])]
        masks = [np.stack([t["obs"]["action_mask"] for t in ts])]
        rewards, dones, rets, lens = [], [], [], []
        for t in range(actions.shape[0]):
            res = [env.step(states[e], actions[t, e]) for e in range(gi.LBF_NE)]
            states = [r[0] for r in res]
            views.append(np.stack([r[1]["obs"]["agents_view"] for r in res]))
            masks.append(np.stack([r[1]["obs"]["action_mask"] for r in res]))
            rewards.append(np.stack([r[1]["reward"] for r in res]))
            dones.append(np.array([r[1]["done"] for r in res]))
            rets.append(np.array([r[1]["metrics"]["episode_return"] for r in res], np.float32))
            lens.append(np.array([r[1]["metrics"]["episode_length"] for r in res], np.int32))
        out.update({f"{name}/keys": keys, f"{name}/actions": actions,
                    f"{name}/views": np.stack(views).astype(np.int8),
                    f"{name}/masks": np.stack(masks), f"{name}/rewards": np.stack(rewards),
                    f"{name}/dones": np.stack(dones), f"{name}/ep_returns": np.stack(rets),
                    f"{name}/ep_lengths": np.stack(lens)})
        assert np.stack(dones).sum() > gi.LBF_NE  # several episodes per env
    return out


def _layers(ps):
    t = [torch.tensor(p, dtype=torch.float64, requires_grad=True) for p in ps]
    return t, [(t[0], t[1]), (t[2], t[3]), (t[4], t[5])]


def ppo_golden():
    out = {}
    for n in gi.PERM_SIZES:  # jax.random.permutation(PRNGKey(PERM_SEED + n), n)
        out[f"perm/{n}"] = tf.permutation(tf.prng_key(gi.PERM_SEED + n), n).astype(np.int32)
    g = gi.gae_inputs()
    done_a = np.repeat(g["done"][:, :, None], g["reward"].shape[2], 2)
    adv, tgt = oppo.gae_ff(g["reward"], g["value"], done_a, g["last_val"], g["gamma"],
                           g["gae_lambda"])
    out["gae/ff_adv"], out["gae/ff_targets"] = adv, tgt
    adv, tgt = oppo.gae_rec(g["reward"], g["value"], done_a, g["last_val"],
                            np.repeat(g["last_done"][:, None], g["reward"].shape[2], 1),
                            g["gamma"], g["gae_lambda"])
    out["gae/rec_adv"], out["gae/rec_targets"] = adv, tgt

    li = gi.loss_inputs()
    S, A, FR = li["view"].shape
    at, al = _layers(li["actor"])
    ct, cl = _layers(li["critic"])
    v = li["view"].astype(np.float64)
    x = torch.tensor(np.concatenate([np.broadcast_to(np.eye(A), (S, A, A)), v], -1))
    xg = torch.tensor(np.repeat(v.reshape(S, 1, A * FR), A, 1))
    logits = oppo.actor_logits(al, x, torch.tensor(li["mask"]))
    ta, la, ent = oppo.actor_loss(logits, torch.tensor(li["action"]),
                                  torch.tensor(li["old_logp"], dtype=torch.float64),
                                  torch.tensor(li["adv"], dtype=torch.float64),
                                  float(li["clip_eps"]), float(li["ent_coef"]))
    val = oppo.critic_value(cl, xg)
    tc, vl = oppo.critic_loss(val, torch.tensor(li["old_value"], dtype=torch.float64),
                              torch.tensor(li["targets"], dtype=torch.float64),
                              float(li["clip_eps"]), float(li["vf_coef"]))
    ta.backward()
    tc.backward()
    out["loss/logits"] = logits.detach().numpy().astype(np.float32)
    out["loss/value"] = val.detach().numpy().astype(np.float32)
    out["loss/scalars"] = np.array([ta.item(), la.item(), ent.item(), tc.item(), vl.item()])
    out["loss/actor_grad"] = np.concatenate([p.grad.numpy().ravel() for p in at]).astype(np.float32)
    out["loss/critic_grad"] = np.concatenate([p.grad.numpy().ravel() for p in ct]).astype(np.float32)

    ai = gi.adam_inputs()
    p, mu, nu = ai["params"].copy(), np.zeros_like(ai["params"]), np.zeros_like(ai["params"])
    for c, gr in enumerate(ai["grads"]):
        p, mu, nu = oppo.clip_adam(p, gr, mu, nu, c, float(ai["lr"]), float(ai["max_norm"]))
    out["adam/params"], out["adam/mu"], out["adam/nu"] = p, mu, nu
    return out


def main():
    for fname, data in (("lbf_golden.npz", lbf_golden()), ("ppo_golden.npz", ppo_golden())):
        path = os.path.join(HERE, fname)
        np.savez_compressed(path, generator=np.array("oracle"), **data)
        print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
