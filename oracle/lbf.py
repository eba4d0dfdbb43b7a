"""Oracle: Jumanji LevelBasedForaging + the Mava wrapper stack, one env at a time, in numpy.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  **Parity unpinned** for the inner environment:
``jumanji.environments.routing.lbf`` lives in the unpinned jumanji fork
(requirements/requirements.txt:12), absent from /root/reference and from this image.  The inner env
restates the published Jumanji algorithm (lbf/env.py, generator.py, observer.py, utils.py); the
reference call sites that anchor it are ``mava/wrappers/jumanji.py:33,158-215`` and
``mava/utils/make_env.py:28-30,55``.

The wrapper stack is restated from the reference:
``LbfWrapper.modify_timestep`` / ``aggregate_rewards`` (mava/wrappers/jumanji.py:180-204),
AgentID / AutoReset / RecordEpisodeMetrics as in oracle/rware.py.

Conventions: positions are (x, y) = (row, col) indices of a grid_size x grid_size grid; actions
0 noop, 1 up (x-1), 2 down (x+1), 3 left (y-1), 4 right (y+1), 5 load.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Tuple

import numpy as np

from . import threefry as tf

MOVES = np.array([[0, 0], [-1, 0], [1, 0], [0, -1], [0, 1], [0, 0]], np.int32)
LOAD = 5
F32 = np.float32


@dataclass(frozen=True)
class LbfSpec:
    S: int  # grid size
    fov: int
    A: int
    NF: int
    max_level: int
    force_coop: bool
    time_limit: int

    @property
    def num_obs_features(self) -> int:
        return 3 * (self.NF + self.A)


def make_spec(grid_size=8, fov=8, num_agents=2, num_food=2, max_agent_level=2, force_coop=True,
              time_limit=100) -> LbfSpec:
    return LbfSpec(grid_size, fov, num_agents, num_food, max_agent_level, bool(force_coop), time_limit)


def _choice_p(key, p: np.ndarray) -> int:
    """jax.random.choice(key, n, shape=(), p=p): inverse-CDF on a float32 cumulative sum."""
    p_cuml = np.cumsum(p.astype(F32), dtype=F32)
    u = tf.uniform(key, ())
    r = F32(p_cuml[-1] * (F32(1.0) - u))
    return int(np.searchsorted(p_cuml, r, side="left"))


def _choice_p_noreplace(key, p: np.ndarray, n: int) -> np.ndarray:
    """jax.random.choice(..., replace=False, p=p): Gumbel top-k."""
    with np.errstate(divide="ignore"):
        g = -tf.gumbel(key, p.shape) - np.log(p.astype(F32))
    return np.argsort(g, kind="stable")[:n].astype(np.int32)


def generator(spec: LbfSpec, key) -> Dict:
    """jumanji lbf/generator.py RandomGenerator.__call__."""
    S = spec.S
    ks = tf.split(key, 5)
    key_food, key_agents, key_food_level, key_agent_level, key = ks
    # food: never on the border, never adjacent to another food
    flat = S * S
    mask = np.ones(flat, bool)
    mask[np.arange(S)] = False
    mask[np.arange(flat - S, flat)] = False
    mask[np.arange(0, flat, S)] = False
    mask[np.arange(S - 1, flat, S)] = False
    pos_keys = tf.split(key_food, spec.NF)
    food = []
    for i in range(spec.NF):
        f = _choice_p(pos_keys[i], mask)
        for adj in (f, f + 1, f - 1, f + S, f - S):
            if 0 <= adj < flat:  # jax drops out-of-bounds scatter indices
                mask[adj] = False
        food.append(f)
    food = np.array(food, np.int32)
    fx, fy = food // S, food % S
    amask = np.ones((S, S), bool)
    amask[fx, fy] = False
    agents = _choice_p_noreplace(key_agents, amask.ravel(), spec.A)
    ax, ay = agents // S, agents % S
    alvl = tf.randint(key_agent_level, (spec.A,), 1, spec.max_level + 1)
    max_food_level = int(np.sort(alvl)[:3].sum())
    if spec.force_coop:
        flvl = np.full(spec.NF, max_food_level, np.int32)
    else:
        flvl = tf.randint(key_food_level, (spec.NF,), 1, max_food_level + 1)
    return dict(ax=ax.astype(np.int32), ay=ay.astype(np.int32), alvl=alvl.astype(np.int32),
                loading=np.zeros(spec.A, bool), fx=fx.astype(np.int32), fy=fy.astype(np.int32),
                flvl=flvl.astype(np.int32), eaten=np.zeros(spec.NF, bool), step=0,
                key=np.asarray(key, np.uint32))


def observe(spec: LbfSpec, st) -> Tuple[np.ndarray, np.ndarray]:
    """VectorObserver.state_to_observation -> (agents_view (A, 3(F+A)) int32, action_mask (A, 6))."""
    A, NF, S = spec.A, spec.NF, spec.S
    view = np.zeros((A, spec.num_obs_features), np.int32)
    mask = np.zeros((A, 6), bool)
    for i in range(A):
        px, py = int(st["ax"][i]), int(st["ay"][i])
        ox, oy = min(spec.fov, px), min(spec.fov, py)
        row = []
        for f in range(NF):
            vis = (abs(px - st["fx"][f]) <= spec.fov and abs(py - st["fy"][f]) <= spec.fov
                   and not st["eaten"][f])
            row += [int(st["fx"][f]) - px + ox, int(st["fy"][f]) - py + oy, int(st["flvl"][f])] if vis \
                else [-1, -1, 0]
        order = [i] + [j for j in range(A) if j != i]
        for j in order:
            vis = abs(px - st["ax"][j]) <= spec.fov and abs(py - st["ay"][j]) <= spec.fov
            row += [int(st["ax"][j]) - px + ox, int(st["ay"][j]) - py + oy, int(st["alvl"][j])] if vis \
                else [-1, -1, 0]
        view[i] = row
        adj_food = False
        for f in range(NF):
            if not st["eaten"][f] and abs(px - st["fx"][f]) + abs(py - st["fy"][f]) == 1:
                adj_food = True
        for a in range(6):
            nx, ny = px + MOVES[a, 0], py + MOVES[a, 1]
            oob = nx < 0 or ny < 0 or nx >= S or ny >= S
            occ = any(j != i and st["ax"][j] == nx and st["ay"][j] == ny for j in range(A))
            foodc = any((not st["eaten"][f]) and st["fx"][f] == nx and st["fy"][f] == ny
                        for f in range(NF))
            mask[i, a] = not (oob or occ or foodc)
        if not adj_food:
            mask[i, LOAD] = False
    return view, mask


def step(spec: LbfSpec, st, action) -> Tuple[Dict, Dict]:
    """jumanji lbf/env.py LevelBasedForaging.step."""
    st = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in st.items()}
    A, NF, S = spec.A, spec.NF, spec.S
    ox, oy = st["ax"].copy(), st["ay"].copy()
    nx, ny = ox.copy(), oy.copy()
    for i in range(A):  # simulate_agent_movement (all agents against the ORIGINAL positions)
        a = int(action[i])
        tx, ty = ox[i] + MOVES[a, 0], oy[i] + MOVES[a, 1]
        oob = tx < 0 or ty < 0 or tx >= S or ty >= S
        occ = any(j != i and ox[j] == tx and oy[j] == ty for j in range(A))
        foodc = any((not st["eaten"][f]) and st["fx"][f] == tx and st["fy"][f] == ty
                    for f in range(NF))
        if not (oob or occ or foodc):
            nx[i], ny[i] = tx, ty
    # fix_collisions: every agent whose target cell is shared goes back
    dup = [any(j != i and nx[j] == nx[i] and ny[j] == ny[i] for j in range(A)) for i in range(A)]
    for i in range(A):
        if dup[i]:
            nx[i], ny[i] = ox[i], oy[i]
    st["ax"], st["ay"] = nx, ny
    st["loading"] = np.array([int(a) == LOAD for a in action])
    # eat_food per food, rewards (get_reward, normalize_reward=True, penalty=0)
    total_food_level = int(st["flvl"].sum())
    reward = np.zeros(A, F32)
    new_eaten = st["eaten"].copy()
    for f in range(NF):
        lv = np.array([int(st["alvl"][i]) if (abs(nx[i] - st["fx"][f]) + abs(ny[i] - st["fy"][f]) == 1
                                              and st["loading"][i] and not st["eaten"][f]) else 0
                       for i in range(A)], np.int32)
        s = int(lv.sum())
        eaten_now = s >= int(st["flvl"][f])
        new_eaten[f] = eaten_now or st["eaten"][f]
        num = (lv * int(eaten_now) * int(st["flvl"][f])).astype(F32)
        den = F32(s * total_food_level)
        with np.errstate(divide="ignore", invalid="ignore"):
            r = np.nan_to_num(num / den).astype(F32)
        reward = (reward + r).astype(F32)
    st["eaten"] = new_eaten
    st["step"] = int(st["step"]) + 1
    done = bool(new_eaten.all() or st["step"] >= spec.time_limit)
    view, mask = observe(spec, st)
    return st, dict(agents_view=view, action_mask=mask, step_count=st["step"], reward=reward,
                    done=done)


class MavaLbf:
    """RecordEpisodeMetrics(AutoResetWrapper(AgentIDWrapper(LbfWrapper(LevelBasedForaging))))."""

    def __init__(self, spec: LbfSpec, add_global_state: bool, add_agent_id: bool = True,
                 auto_reset: bool = True, use_individual_rewards: bool = False):
        self.spec, self.add_global_state = spec, add_global_state
        self.add_agent_id, self.auto_reset = add_agent_id, auto_reset
        self.individual = use_individual_rewards
        self.num_agents, self.action_dim, self.time_limit = spec.A, 6, spec.time_limit

    def _observation(self, raw_view, mask, step_count) -> Dict:
        A = self.spec.A
        view = raw_view.astype(np.float32)
        obs = dict(action_mask=mask.copy(), step_count=np.full(A, step_count, np.int32))
        if self.add_global_state:
            obs["global_state"] = np.tile(view.reshape(-1), (A, 1))
        if self.add_agent_id:
            view = np.concatenate([np.eye(A, dtype=np.float32), view], axis=-1)
        obs["agents_view"] = view
        return obs

    def reset(self, key):
        key, reset_key = tf.split(key)
        inner = generator(self.spec, reset_key)
        view, mask = observe(self.spec, inner)
        state = dict(inner=inner, key=np.asarray(key, np.uint32), run_ret=F32(0.0), run_len=0,
                     ep_ret=F32(0.0), ep_len=0)
        ts = dict(obs=self._observation(view, mask, 0), reward=np.zeros(self.spec.A, F32),
                  done=False, metrics=dict(episode_return=F32(0.0), episode_length=0,
                                           is_terminal_step=False))
        return state, ts

    def step(self, state, action):
        inner, raw = step(self.spec, state["inner"], action)
        done = raw["done"]
        if done and self.auto_reset:
            key, _ = tf.split(inner["key"])
            inner = generator(self.spec, key)
            view, mask = observe(self.spec, inner)
            obs = self._observation(view, mask, 0)
        else:
            obs = self._observation(raw["agents_view"], raw["action_mask"], raw["step_count"])
        reward = raw["reward"]
        if not self.individual:  # aggregate_rewards: team reward = sum, repeated (jumanji.py:180-187)
            team = F32(0.0)
            for r in reward:
                team = F32(team + r)
            reward = np.full(self.spec.A, team, F32)
        mean = F32(0.0)
        for r in reward:
            mean = F32(mean + r)
        mean = F32(mean / F32(self.spec.A))
        nd, dd = F32(0.0 if done else 1.0), F32(1.0 if done else 0.0)
        new_ret = F32(state["run_ret"] + mean)
        new_len = state["run_len"] + 1
        ep_ret = F32(state["ep_ret"] * nd + new_ret * dd)
        ep_len = new_len if done else state["ep_len"]
        nstate = dict(inner=inner, key=state["key"], run_ret=F32(new_ret * nd),
                      run_len=0 if done else new_len, ep_ret=ep_ret, ep_len=ep_len)
        ts = dict(obs=obs, reward=reward, done=done,
                  metrics=dict(episode_return=ep_ret, episode_length=ep_len, is_terminal_step=done))
        return nstate, ts
