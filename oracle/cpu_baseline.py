"""CPU baseline: the ff_ippo / ff_mappo update (rollout + GAE + PPO epochs) run on host cores.

TEST INFRASTRUCTURE / BASELINE ONLY.  The prescribed baseline - the reference itself on
``JAX_PLATFORMS=cpu`` - cannot run in this image (no jax / jumanji, no network), so
``bench.py``'s ``cpu_baseline`` and ``--impl reference`` legs time this port instead
(``"kind": "port"``): the C port of the env oracle (OpenMP over envs) plus the torch-CPU float32
restatement of the networks, losses and optimiser from ``oracle/ppo.py``, using every host core.
It follows the same schedule as mava/systems/ppo/ff_mappo.py:56-300 (sampling uses torch's RNG:
irrelevant for timing).
"""
from __future__ import annotations

import os
import time
from typing import Dict

import numpy as np
import torch

from . import ppo as oppo
from . import threefry as tf
from .rware_c import RwareC


def _init(in_dim, h, out, gen):
    shapes = [(in_dim, h), (h,), (h, h), (h,), (h, out), (out,)]
    ps = []
    for s in shapes:
        if len(s) == 1:
            ps.append(torch.zeros(s, requires_grad=True))
        else:
            ps.append((torch.randn(s, generator=gen) / np.sqrt(s[0])).requires_grad_())
    return ps


def _layers(ps):
    return [(ps[0], ps[1]), (ps[2], ps[3]), (ps[4], ps[5])]


def run(task: Dict, num_envs: int, rollout_length: int = 128, ppo_epochs: int = 4,
        num_minibatches: int = 2, centralised: bool = True, updates: int = 1, warmup: int = 0,
        time_limit: int = 500, threads: int = 0) -> Dict:
    """Time ``updates`` full updates on ``num_envs`` envs.  Returns env-steps/s and the cores used."""
    cores = threads or len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    # torchrun exports OMP_NUM_THREADS=1 to its workers: the baseline must still use every core
    os.environ["OMP_NUM_THREADS"] = str(cores)
    try:
        import ctypes

        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(int(cores))
    except OSError:
        pass
    env = RwareC(time_limit=time_limit, **task)
    A, FR, N, E, T = env.A, env.FR, 5, num_envs, rollout_length
    gen = torch.Generator().manual_seed(0)
    ap = _init(FR + A, 128, N, gen)
    cp = _init(A * FR if centralised else FR + A, 128, 1, gen)
    opt_a = torch.optim.Adam(ap, lr=2.5e-4, eps=1e-5)
    opt_c = torch.optim.Adam(cp, lr=2.5e-4, eps=1e-5)
    state, view, mask = env.reset(tf.split(tf.prng_key(0), E))
    eye = torch.eye(A).expand(E, A, A)
    bits = torch.arange(N)

    def obs_tensors(view, mask):
        v = torch.from_numpy(view).float()
        x = torch.cat([eye[: v.shape[0]], v], -1)
        g = v.reshape(v.shape[0], 1, A * FR).expand(-1, A, -1) if centralised else x
        m = ((torch.from_numpy(mask).int().unsqueeze(-1) >> bits) & 1).bool()
        return x, g, m

    def one_update():
        nonlocal view, mask
        xs, gs, ms, acts, lps, vals, rews, dones = [], [], [], [], [], [], [], []
        with torch.no_grad():
            for _ in range(T):
                x, g, m = obs_tensors(view, mask)
                logits = oppo.actor_logits(_layers(ap), x, m)
                val = oppo.critic_value(_layers(cp), g)
                u = torch.rand(logits.shape, generator=gen).clamp_min(1e-20)
                a = torch.argmax(logits - torch.log(-torch.log(u)), -1)
                lp = oppo.categorical_log_prob(logits, a)
                view, mask, r, d, _, _ = env.step(state, a.numpy().astype(np.int8))
                xs.append(x); gs.append(g); ms.append(m); acts.append(a); lps.append(lp)
                vals.append(val); rews.append(torch.from_numpy(r)); dones.append(d.copy())
            x, g, m = obs_tensors(view, mask)
            last_val = oppo.critic_value(_layers(cp), g).numpy()
        value = torch.stack(vals).numpy()
        done = np.repeat(np.stack(dones)[:, :, None], A, 2)
        adv, tgt = oppo.gae_ff(torch.stack(rews).numpy(), value, done, last_val, 0.99, 0.95)
        flat = lambda lst: torch.stack(lst).reshape(T * E, *lst[0].shape[1:])
        X, G, M_, AC, LP = flat(xs), flat(gs), flat(ms), flat(acts), flat(lps)
        V = torch.from_numpy(value).reshape(T * E, A)
        ADV = torch.from_numpy(adv).reshape(T * E, A)
        TGT = torch.from_numpy(tgt).reshape(T * E, A)
        mb = T * E // num_minibatches
        for _ in range(ppo_epochs):
            perm = torch.randperm(T * E, generator=gen)
            for k in range(num_minibatches):
                idx = perm[k * mb:(k + 1) * mb]
                logits = oppo.actor_logits(_layers(ap), X[idx], M_[idx])
                la, _, _ = oppo.actor_loss(logits, AC[idx], LP[idx], ADV[idx], 0.2, 0.01)
                lc, _ = oppo.critic_loss(oppo.critic_value(_layers(cp), G[idx]), V[idx], TGT[idx],
                                         0.2, 0.5)
                opt_a.zero_grad(); opt_c.zero_grad()
                la.backward(); lc.backward()
                torch.nn.utils.clip_grad_norm_(ap, 0.5)
                torch.nn.utils.clip_grad_norm_(cp, 0.5)
                opt_a.step(); opt_c.step()

    for _ in range(warmup):
        one_update()
    t0 = time.perf_counter()
    for _ in range(updates):
        one_update()
    dt = time.perf_counter() - t0
    return {"env_steps_per_s": updates * T * E / dt, "seconds": dt, "cores": cores,
            "env_steps": updates * T * E}
