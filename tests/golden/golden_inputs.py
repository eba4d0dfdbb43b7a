"""Seeded inputs and the npz schema shared by tests/golden/make_oracle_golden.py (outputs from the
CPU restatement under oracle/) and tests/golden/make_reference_golden.py (outputs from the real
jax + jumanji + mava stack).  Both write the SAME keys, so the tests that read the files pin whatever
generated them; regenerate with the reference script on a machine that has the stack and the oracle
becomes reference-pinned.

TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import numpy as np

RWARE_SCENARIOS = {
    "tiny-2ag": dict(column_height=8, shelf_rows=1, shelf_columns=3, num_agents=2, sensor_range=1,
                     request_queue_size=2),
    "tiny-4ag": dict(column_height=8, shelf_rows=1, shelf_columns=3, num_agents=4, sensor_range=1,
                     request_queue_size=4),
    "small-4ag": dict(column_height=8, shelf_rows=2, shelf_columns=3, num_agents=4, sensor_range=1,
                      request_queue_size=4),
}
RWARE_TIME_LIMIT, RWARE_NE, RWARE_T, RWARE_SEED = 25, 6, 60, 2024

LBF_SCENARIOS = {
    "8x8-2p-2f-coop": dict(grid_size=8, fov=8, num_agents=2, num_food=2, max_agent_level=2,
                           force_coop=True),
    "2s-10x10-3p-3f": dict(grid_size=10, fov=2, num_agents=3, num_food=3, max_agent_level=2,
                           force_coop=False),
}
LBF_TIME_LIMIT, LBF_NE, LBF_T, LBF_SEED = 20, 6, 200, 77


def rware_actions(A: int) -> np.ndarray:
    rng = np.random.default_rng(11)
    return rng.choice(5, size=(RWARE_T, RWARE_NE, A), p=[0.05, 0.5, 0.15, 0.15, 0.15]).astype(np.int8)


def lbf_actions(A: int) -> np.ndarray:
    rng = np.random.default_rng(13)
    return rng.integers(0, 6, size=(LBF_T, LBF_NE, A)).astype(np.int8)


# ---- PPO arithmetic -----------------------------------------------------------------------------
PERM_SIZES = (7, 128, 1000, 2048)
PERM_SEED = 5


def gae_inputs():
    rng = np.random.default_rng(21)
    T, NE, A = 16, 6, 2
    return dict(reward=rng.normal(size=(T, NE, A)).astype(np.float32),
                value=rng.normal(size=(T, NE, A)).astype(np.float32),
                done=(rng.random((T, NE)) < 0.15),
                last_val=rng.normal(size=(NE, A)).astype(np.float32),
                last_done=(rng.random(NE) < 0.3),
                gamma=np.float32(0.99), gae_lambda=np.float32(0.95))


def loss_inputs():
    """One minibatch of ff_mappo shapes: A = 2 agents, FR = 12 raw features, N = 6 actions,
    [128, 128] torsos, agent ids on, centralised critic."""
    rng = np.random.default_rng(31)
    S, A, FR, N, H = 48, 2, 12, 6, 128
    view = rng.integers(-1, 9, size=(S, A, FR)).astype(np.int8)
    mask = rng.random((S, A, N)) < 0.7
    mask[..., 0] = True
    action = np.array([[rng.choice(np.flatnonzero(m)) for m in row] for row in mask], np.int32)

    def net(in_dim, out, scale_out):
        shapes = [(in_dim, H), (H,), (H, H), (H,), (H, out), (out,)]
        return [(rng.normal(size=s) * (scale_out if i >= 4 else (1 / np.sqrt(s[0]) if len(s) == 2
                                                                 else 0.1))).astype(np.float32)
                for i, s in enumerate(shapes)]

    return dict(view=view, mask=mask, action=action,
                old_logp=(-rng.random((S, A)) * 2).astype(np.float32),
                adv=rng.normal(size=(S, A)).astype(np.float32),
                old_value=rng.normal(size=(S, A)).astype(np.float32),
                targets=rng.normal(size=(S, A)).astype(np.float32),
                actor=net(FR + A, N, 0.05), critic=net(A * FR, 1, 0.3),
                clip_eps=np.float32(0.2), ent_coef=np.float32(0.01), vf_coef=np.float32(0.5))


def adam_inputs():
    rng = np.random.default_rng(41)
    n = 257
    return dict(params=(rng.normal(size=n) * 0.1).astype(np.float32),
                grads=[(rng.normal(size=n) * s).astype(np.float32) for s in (0.3, 0.001, 0.05)],
                lr=np.float32(2.5e-4), max_norm=np.float32(0.5))
