"""Oracle: the PPO arithmetic of mava/systems/ppo/{ff_ippo,ff_mappo,rec_ippo,rec_mappo}.py.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Restated from the reference lines cited on
each function; gradients come from torch autograd on the CPU (the torch fp32/fp64 reference the
floating-point kernels are compared with).  The reference's tests pin no values
(test/integration_test.py:35-46) -> parity unpinned beyond this restatement.

Third-party pieces restated from their published algorithms (not under /root/reference):
flax ``nn.Dense`` (y = x @ kernel + bias, kernel (in, out)), flax ``nn.GRUCell``,
``tfd.Categorical`` log_prob / entropy, ``optax.clip_by_global_norm`` and ``optax.adam``
(all unpinned in requirements/requirements.txt:5,19,23).
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch

F32_MIN = float(np.finfo(np.float32).min)


# ------------------------------------------------------------------------------------------
# networks (mava/networks.py)
# ------------------------------------------------------------------------------------------
def mlp_torso(x: torch.Tensor, layers: Sequence[Tuple[torch.Tensor, torch.Tensor]]) -> torch.Tensor:
    """MLPTorso.__call__, networks.py:49-58 (relu, no layer norm)."""
    for w, b in layers:
        x = torch.relu(x @ w + b)
    return x


def actor_logits(params: List[Tuple[torch.Tensor, torch.Tensor]], agents_view, action_mask):
    """FeedForwardActor + DiscreteActionHead, networks.py:114-124,181-183."""
    emb = mlp_torso(agents_view, params[:-1])
    w, b = params[-1]
    logits = emb @ w + b
    return torch.where(action_mask, logits, torch.full_like(logits, F32_MIN))


def critic_value(params: List[Tuple[torch.Tensor, torch.Tensor]], critic_in):
    """FeedForwardValueNet, networks.py:204-207."""
    emb = mlp_torso(critic_in, params[:-1])
    w, b = params[-1]
    return (emb @ w + b).squeeze(-1)


def categorical_log_prob(logits, action):
    """tfd.Categorical.log_prob: log_softmax gathered at the action."""
    return torch.log_softmax(logits, -1).gather(-1, action.long().unsqueeze(-1)).squeeze(-1)


def categorical_entropy(logits):
    """tfd.Categorical.entropy: -sum p log p with 0 * (-big) := 0."""
    logp = torch.log_softmax(logits, -1)
    p = logp.exp()
    return -(torch.where(p == 0, torch.zeros_like(p), p * logp)).sum(-1)


def gru_cell(h, x, p: Dict[str, torch.Tensor]):
    """flax.linen.GRUCell: r,z = sigmoid(Wi x + bi + Wh h); n = tanh(Win x + bin + r*(Whn h + bhn))."""
    r = torch.sigmoid(x @ p["ir_w"] + p["ir_b"] + h @ p["hr_w"])
    z = torch.sigmoid(x @ p["iz_w"] + p["iz_b"] + h @ p["hz_w"])
    n = torch.tanh(x @ p["in_w"] + p["in_b"] + r * (h @ p["hn_w"] + p["hn_b"]))
    return (1.0 - z) * n + z * h


def scanned_rnn(h, xs, resets, p):
    """ScannedRNN.__call__, networks.py:249-259: zero the carry where reset, then GRUCell."""
    ys = []
    for t in range(xs.shape[0]):
        h = torch.where(resets[t].unsqueeze(-1), torch.zeros_like(h), h)
        h = gru_cell(h, xs[t], p)
        ys.append(h)
    return h, torch.stack(ys)


# ------------------------------------------------------------------------------------------
# GAE (ff_mappo.py:112-139, rec_mappo.py:177-199)
# ------------------------------------------------------------------------------------------
def gae_ff(reward, value, done, last_val, gamma, lam):
    """Reverse scan; ``done`` is the done flag of the transition itself."""
    T = reward.shape[0]
    reward = np.asarray(reward, np.float32)
    value = np.asarray(value, np.float32)
    nd = (1 - np.asarray(done).astype(np.int32)).astype(np.float32)
    adv = np.zeros_like(value)
    gae = np.zeros_like(last_val, dtype=np.float32)
    nxt = np.asarray(last_val, np.float32)
    g, gl = np.float32(gamma), np.float32(gamma * lam)
    for t in range(T - 1, -1, -1):
        delta = reward[t] + g * nxt * nd[t] - value[t]
        gae = delta + gl * nd[t] * gae
        adv[t] = gae
        nxt = value[t]
    return adv, adv + value


def gae_rec(reward, value, done, last_val, last_done, gamma, lam):
    """Recurrent flavour: stored ``done`` is the flag ENTERING the step; carry next_done."""
    T = reward.shape[0]
    reward = np.asarray(reward, np.float32)
    value = np.asarray(value, np.float32)
    adv = np.zeros_like(value)
    gae = np.zeros_like(last_val, dtype=np.float32)
    nxt = np.asarray(last_val, np.float32)
    nxt_nd = (1 - np.asarray(last_done).astype(np.int32)).astype(np.float32)
    g, gl = np.float32(gamma), np.float32(gamma * lam)
    for t in range(T - 1, -1, -1):
        delta = reward[t] + g * nxt * nxt_nd - value[t]
        gae = delta + gl * nxt_nd * gae
        adv[t] = gae
        nxt = value[t]
        nxt_nd = (1 - np.asarray(done[t]).astype(np.int32)).astype(np.float32)
    return adv, adv + value


# ------------------------------------------------------------------------------------------
# losses (ff_mappo.py:150-201)
# ------------------------------------------------------------------------------------------
def actor_loss(logits, action, old_log_prob, gae, clip_eps, ent_coef):
    """_actor_loss_fn, ff_mappo.py:159-180.  Returns (total, loss_actor, entropy)."""
    log_prob = categorical_log_prob(logits, action)
    ratio = torch.exp(log_prob - old_log_prob)
    gae = (gae - gae.mean()) / (gae.std(unbiased=False) + 1e-8)
    l1 = ratio * gae
    l2 = torch.clamp(ratio, 1.0 - clip_eps, 1.0 + clip_eps) * gae
    loss = -torch.minimum(l1, l2).mean()
    ent = categorical_entropy(logits).mean()
    return loss - ent_coef * ent, loss, ent


def critic_loss(value, old_value, targets, clip_eps, vf_coef):
    """_critic_loss_fn, ff_mappo.py:190-201.  Returns (total, value_loss)."""
    v_clip = old_value + (value - old_value).clamp(-clip_eps, clip_eps)
    vl = 0.5 * torch.maximum((value - targets) ** 2, (v_clip - targets) ** 2).mean()
    return vf_coef * vl, vl


# ------------------------------------------------------------------------------------------
# optimiser (ff_mappo.py:240-250,359-366): optax.chain(clip_by_global_norm, adam(eps=1e-5))
# ------------------------------------------------------------------------------------------
def clip_adam(params: np.ndarray, grads: np.ndarray, mu: np.ndarray, nu: np.ndarray, count: int,
              lr: float, max_norm: float, b1=0.9, b2=0.999, eps=1e-5):
    """One optax step on flat float32 vectors.  ``count`` is the number of steps already taken."""
    g = grads.astype(np.float32)
    g_norm = np.float32(np.sqrt(np.sum(g.astype(np.float64) ** 2)))
    if not (g_norm < np.float32(max_norm)):
        g = (g / g_norm) * np.float32(max_norm)
    mu = (np.float32(1 - b1) * g + np.float32(b1) * mu).astype(np.float32)
    nu = (np.float32(1 - b2) * g * g + np.float32(b2) * nu).astype(np.float32)
    c = count + 1
    mu_hat = mu / np.float32(1 - b1 ** c)
    nu_hat = nu / np.float32(1 - b2 ** c)
    upd = mu_hat / (np.sqrt(nu_hat) + np.float32(eps))
    return (params + np.float32(-lr) * upd).astype(np.float32), mu, nu


def linear_lr(init_lr: float, count: int, ppo_epochs: int, num_minibatches: int, num_updates: int):
    """make_learning_rate_schedule, mava/utils/training.py:39-47."""
    return init_lr * (1.0 - (count // (ppo_epochs * num_minibatches)) / num_updates)


# ------------------------------------------------------------------------------------------
# recurrent systems (mava/networks.py:238-331, mava/systems/ppo/rec_mappo.py)
# ------------------------------------------------------------------------------------------
def rnn_unflatten(flat: torch.Tensor, in_dim: int, H: int, Q: int, out: int) -> Dict[str, torch.Tensor]:
    """Split the flat parameter vector of include/mava_b200.h (mava_rnn_desc) into named blocks."""
    shapes = [("pre_w", (in_dim, H)), ("pre_b", (H,)), ("wi", (H, 3 * H)), ("bi", (3 * H,)),
              ("wh", (H, 3 * H)), ("hn_b", (H,)), ("post_w", (H, Q)), ("post_b", (Q,)),
              ("head_w", (Q, out)), ("head_b", (out,))]
    p, off = {}, 0
    for name, s in shapes:
        n = int(np.prod(s))
        p[name] = flat[off:off + n].reshape(s)
        off += n
    assert off == flat.numel()
    return p


def _gru_params(p: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    H = p["hn_b"].shape[0]
    wi, bi, wh = p["wi"], p["bi"], p["wh"]
    return {"ir_w": wi[:, :H], "iz_w": wi[:, H:2 * H], "in_w": wi[:, 2 * H:],
            "ir_b": bi[:H], "iz_b": bi[H:2 * H], "in_b": bi[2 * H:],
            "hr_w": wh[:, :H], "hz_w": wh[:, H:2 * H], "hn_w": wh[:, 2 * H:], "hn_b": p["hn_b"]}


def rec_net(p: Dict[str, torch.Tensor], h0: torch.Tensor, xs: torch.Tensor, resets: torch.Tensor):
    """RecurrentActor / RecurrentValueNet body, networks.py:281-294,314-331: pre_torso ->
    ScannedRNN -> post_torso -> head.  xs (L, S, in), resets (L, S) bool.  Returns (h_L, out)."""
    emb = torch.relu(xs @ p["pre_w"] + p["pre_b"])
    h, ys = scanned_rnn(h0, emb, resets, _gru_params(p))
    post = torch.relu(ys @ p["post_w"] + p["post_b"])
    return h, post @ p["head_w"] + p["head_b"]


def rec_chunk_batch(x: torch.Tensor, chunk: int, cols: torch.Tensor) -> torch.Tensor:
    """The reference's minibatch view of a (T, E, ...) rollout array, rec_mappo.py:339-357:
    reshape to (chunk, E * num_chunks, ...) and take the minibatch's columns."""
    T, E = x.shape[:2]
    nc = T // chunk
    return x.reshape(chunk, E * nc, *x.shape[2:])[:, cols.long()]
