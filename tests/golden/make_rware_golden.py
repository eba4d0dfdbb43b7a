"""Generates tests/golden/rware_golden.npz from the numpy oracle (oracle/rware.py).

The reference's RWARE dynamics live in the third-party jumanji package, which is not installable
here (no network), and the reference's own tests hold no golden vectors
(test/integration_test.py:35-46).  These vectors therefore pin the restatement against itself
(regression pins for the oracle, the C port and the CUDA kernel), not against Jumanji.

    python tests/golden/make_rware_golden.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import rware as orw  # noqa: E402
from oracle import threefry as tf  # noqa: E402

from tests.golden.golden_inputs import RWARE_SCENARIOS as SCENARIOS  # noqa: E402

def main():
    out = {}
    for name, task in SCENARIOS.items():
        spec = orw.make_spec(time_limit=25, **task)
        env = orw.MavaRware(spec, add_global_state=False, add_agent_id=False)
        NE, T = 6, 60
        keys = tf.split(tf.prng_key(2024), NE)
        rng = np.random.default_rng(11)
        actions = rng.choice(5, size=(T, NE, spec.A), p=[0.05, 0.5, 0.15, 0.15, 0.15]).astype(np.int8)
        states, ts = zip(*[env.reset(k) for k in keys])
        states = list(states)
        views = [np.stack([t["obs"]["agents_view"] for t in ts]).astype(np.int8)]
        masks, rewards, dones, rets, lens = [], [], [], [], []
        masks.append(np.stack([t["obs"]["action_mask"] for t in ts]))
        for t in range(T):
            res = [env.step(states[e], actions[t, e]) for e in range(NE)]
            states = [r[0] for r in res]
            views.append(np.stack([r[1]["obs"]["agents_view"] for r in res]).astype(np.int8))
            masks.append(np.stack([r[1]["obs"]["action_mask"] for r in res]))
            rewards.append(np.stack([r[1]["reward"] for r in res]))
            dones.append(np.array([r[1]["done"] for r in res]))
            rets.append(np.array([r[1]["metrics"]["episode_return"] for r in res], np.float32))
            lens.append(np.array([r[1]["metrics"]["episode_length"] for r in res], np.int32))
        out[f"{name}/keys"] = keys
        out[f"{name}/actions"] = actions
        out[f"{name}/views"] = np.stack(views)
        out[f"{name}/masks"] = np.stack(masks)
        out[f"{name}/rewards"] = np.stack(rewards)
        out[f"{name}/dones"] = np.stack(dones)
        out[f"{name}/ep_returns"] = np.stack(rets)
        out[f"{name}/ep_lengths"] = np.stack(lens)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "rware_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
