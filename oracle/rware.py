"""Oracle: Jumanji RobotWarehouse + the Mava wrapper stack, one env at a time, in numpy.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  **Parity unpinned** for the inner
environment: its arithmetic lives in ``jumanji.environments.routing.robot_warehouse``
(``jumanji @ git+https://github.com/sash-a/jumanji``, requirements/requirements.txt:12, a fork
with no commit pin) which is not under /root/reference and not installed in this image.  The
inner env below restates the published Jumanji algorithm (env.py / generator.py / utils*.py of
that package); the reference call sites that anchor it are
``mava/wrappers/jumanji.py:34,128-155`` and ``mava/utils/make_env.py:31-33,54,104-111``.

The wrapper stack IS under /root/reference and is restated exactly:

* ``RwareWrapper.modify_timestep``            mava/wrappers/jumanji.py:135-144
* ``JumanjiMarlWrapper.get_global_state``     mava/wrappers/jumanji.py:53-59
* ``AgentIDWrapper._add_agent_ids``           mava/wrappers/observation.py:41-53
* ``AutoResetWrapper._auto_reset`` / ``step`` mava/wrappers/auto_reset_wrapper.py:60-101
* ``RecordEpisodeMetrics.reset`` / ``step``   mava/wrappers/episode_metrics.py:59-111
* wrapper order                                mava/utils/make_env.py:69-83,112-115

Conventions (Jumanji's): ``Position(x, y)`` indexes ``grid[layer, x, y]`` so x is the ROW and
y the COLUMN; directions 0 up (x-1), 1 right (y+1), 2 down (x+1), 3 left (y-1); actions
0 noop, 1 forward, 2 turn left, 3 turn right, 4 toggle load.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Tuple

import numpy as np

from . import threefry as tf

SHELVES, AGENTS = 0, 1
NOOP, FORWARD, LEFT, RIGHT, TOGGLE = range(5)


@dataclass(frozen=True)
class RwareSpec:
    """Static scenario constants (jumanji generator.py ``Generator.__init__``)."""

    H: int
    W: int
    A: int
    Q: int
    R: int  # sensor range
    time_limit: int
    highways: np.ndarray  # (H, W) uint8
    shelf_pos: np.ndarray  # (n_shelves, 2) row, col
    goals: Tuple[Tuple[int, int], ...]  # (row, col), scan order

    @property
    def n_shelves(self) -> int:
        return int(self.shelf_pos.shape[0])

    @property
    def num_obs_features(self) -> int:
        loc = (2 * self.R + 1) ** 2
        return 8 + (loc - 1) * 5 + loc * 2


def make_spec(column_height=8, shelf_rows=1, shelf_columns=3, num_agents=4, sensor_range=1,
              request_queue_size=4, time_limit=500) -> RwareSpec:
    """Warehouse layout: the rware/Jumanji rule for highways, shelf cells and the two goals."""
    H = (column_height + 1) * shelf_rows + 2
    W = 3 * shelf_columns + 1
    hw = np.zeros((H, W), np.uint8)
    for r in range(H):
        for c in range(W):
            hw[r, c] = (
                (c % 3 == 0)
                or (r % (column_height + 1) == 0)
                or (r == H - 1)
                or ((r > H - (column_height + 3)) and (c == W // 2 - 1 or c == W // 2))
            )
    shelf_pos = np.argwhere(hw == 0).astype(np.int32)  # row-major order -> shelf ids 0..n-1
    goals = ((H - 1, W // 2 - 1), (H - 1, W // 2))
    return RwareSpec(H, W, num_agents, request_queue_size, sensor_range, time_limit, hw,
                     shelf_pos, goals)


def _action_mask(spec: RwareSpec, grid, ax, ay, adir, carry) -> np.ndarray:
    """jumanji utils.compute_action_mask: only FORWARD can be illegal."""
    mask = np.ones((spec.A, 5), bool)
    for i in range(spec.A):
        nx, ny = _forward(spec, ax[i], ay[i], adir[i])
        stuck = (nx == ax[i]) and (ny == ay[i])
        blocked = bool(carry[i]) and grid[SHELVES, nx, ny] > 0
        mask[i, FORWARD] = not (stuck or blocked)
    return mask


def _forward(spec: RwareSpec, x, y, d):
    """jumanji utils_agent.get_new_position_after_forward (clipped at the border)."""
    if d == 0:
        return max(0, x - 1), y
    if d == 1:
        return x, min(spec.W - 1, y + 1)
    if d == 2:
        return min(spec.H - 1, x + 1), y
    return x, max(0, y - 1)


def generator(spec: RwareSpec, key) -> Dict:
    """jumanji generator.RandomGenerator.__call__ + utils_spawn.spawn_random_entities."""
    key, pos_key = tf.split(key)
    cells = tf.choice_no_replace(pos_key, np.arange(spec.H * spec.W, dtype=np.int32), spec.A)
    ax = (cells // spec.W).astype(np.int32)
    ay = (cells % spec.W).astype(np.int32)
    key, dir_key = tf.split(key)
    adir = tf.randint(dir_key, (spec.A,), 0, 4)
    key, q_key = tf.split(key)
    queue = tf.choice_no_replace(q_key, np.arange(spec.n_shelves, dtype=np.int32), spec.Q)
    req = np.zeros(spec.n_shelves, np.int32)
    req[queue] = 1
    grid = np.zeros((2, spec.H, spec.W), np.int32)
    for s in range(spec.n_shelves):
        grid[SHELVES, spec.shelf_pos[s, 0], spec.shelf_pos[s, 1]] = s + 1
    for i in range(spec.A):  # later agents overwrite (positions are distinct anyway)
        grid[AGENTS, ax[i], ay[i]] = i + 1
    carry = np.zeros(spec.A, np.int32)
    st = dict(grid=grid, ax=ax, ay=ay, adir=adir.astype(np.int32), carry=carry,
              sx=spec.shelf_pos[:, 0].copy(), sy=spec.shelf_pos[:, 1].copy(), req=req,
              queue=queue.astype(np.int32), step=0, key=np.asarray(key, np.uint32))
    st["mask"] = _action_mask(spec, grid, ax, ay, st["adir"], carry)
    return st


def observe(spec: RwareSpec, st) -> np.ndarray:
    """jumanji utils.make_agent_observation for every agent -> (A, num_obs_features) int32."""
    R = spec.R
    out = np.zeros((spec.A, spec.num_obs_features), np.int32)
    grid = st["grid"]
    for i in range(spec.A):
        x, y = int(st["ax"][i]), int(st["ay"][i])
        o = [x, y, int(st["carry"][i])]
        o += [1 if st["adir"][i] == d else 0 for d in range(4)]
        o += [int(spec.highways[x, y])]
        ag, sh = [], []
        for dx in range(-R, R + 1):
            for dy in range(-R, R + 1):
                cx, cy = x + dx, y + dy
                inside = 0 <= cx < spec.H and 0 <= cy < spec.W
                aid = int(grid[AGENTS, cx, cy]) if inside else 0
                sid = int(grid[SHELVES, cx, cy]) if inside else 0
                if not (dx == 0 and dy == 0):  # the agent's own cell carries no "other agent" block
                    if aid == 0:
                        ag += [0, 0, 0, 0, 0]
                    else:
                        ag += [1] + [1 if st["adir"][aid - 1] == d else 0 for d in range(4)]
                if sid == 0:
                    sh += [0, 0]
                else:
                    sh += [1, int(st["req"][sid - 1])]
        out[i] = np.array(o + ag + sh, np.int32)
    return out


def step(spec: RwareSpec, st, action) -> Tuple[Dict, Dict]:
    """jumanji env.RobotWarehouse.step.  Returns (next_state, raw timestep dict)."""
    st = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in st.items()}
    grid, ax, ay, adir, carry = st["grid"], st["ax"], st["ay"], st["adir"], st["carry"]
    n = spec.n_shelves
    # invalid actions -> noop (utils.get_valid_actions)
    act = [int(a) if st["mask"][i, int(a)] else NOOP for i, a in enumerate(action)]
    # agents are updated sequentially in index order (lax.scan over agents in env.step)
    for i in range(spec.A):
        a = act[i]
        x, y = int(ax[i]), int(ay[i])
        if a == LEFT:
            adir[i] = (adir[i] - 1) % 4
        elif a == RIGHT:
            adir[i] = (adir[i] + 1) % 4
        elif a == FORWARD:
            nx, ny = _forward(spec, x, y, int(adir[i]))
            ax[i], ay[i] = nx, ny
            grid[AGENTS, x, y] = 0
            grid[AGENTS, nx, ny] = i + 1
            if carry[i]:
                sid = int(grid[SHELVES, x, y])
                s = (sid - 1) % n  # jax .at[-1] wraps to the last shelf
                st["sx"][s], st["sy"][s] = nx, ny
                grid[SHELVES, x, y] = 0
                grid[SHELVES, nx, ny] = sid
        elif a == TOGGLE:
            sid = int(grid[SHELVES, x, y])
            if not carry[i]:
                if sid != 0:
                    carry[i] = 1
            elif not spec.highways[x, y]:
                carry[i] = 0
    # collisions (utils.is_collision)
    collision = any(int(grid[AGENTS, ax[i], ay[i]]) != i + 1 for i in range(spec.A))
    # deliveries (env._update_reward_and_request_queue), goals scanned in order
    reward = np.float32(0.0)
    key = st["key"]
    for (gx, gy) in spec.goals:
        sid = int(grid[SHELVES, gx, gy])
        if sid != 0 and st["req"][sid - 1] == 1:
            key, rkey = tf.split(key)
            in_q = np.zeros(n, bool)
            in_q[st["queue"]] = True
            not_in_queue = np.arange(n, dtype=np.int32)[~in_q]  # setdiff1d: sorted
            new_req = int(tf.choice_no_replace(rkey, not_in_queue, 1)[0])
            slot = int(np.argmax(st["queue"] == sid - 1))
            st["queue"][slot] = new_req
            reward = np.float32(reward + np.float32(1.0))
            st["req"][sid - 1] = 0
            st["req"][new_req] = 1
    st["key"] = np.asarray(key, np.uint32)
    st["step"] = int(st["step"]) + 1
    done = bool(collision or st["step"] >= spec.time_limit)
    st["mask"] = _action_mask(spec, grid, ax, ay, adir, carry)
    ts = dict(agents_view=observe(spec, st), action_mask=st["mask"].copy(), step_count=st["step"],
              reward=reward, done=done, discount=np.float32(0.0 if done else 1.0))
    return st, ts


# --------------------------------------------------------------------------------------------
# Mava wrapper stack
# --------------------------------------------------------------------------------------------
class MavaRware:
    """RecordEpisodeMetrics(AutoResetWrapper(AgentIDWrapper(RwareWrapper(RobotWarehouse))))."""

    def __init__(self, spec: RwareSpec, add_global_state: bool, add_agent_id: bool = True,
                 auto_reset: bool = True):
        self.spec, self.add_global_state = spec, add_global_state
        self.add_agent_id, self.auto_reset = add_agent_id, auto_reset
        self.num_agents, self.action_dim, self.time_limit = spec.A, 5, spec.time_limit

    def _observation(self, raw_view, mask, step_count) -> Dict:
        A = self.spec.A
        view = raw_view.astype(np.float32)  # RwareWrapper: astype(float)
        obs = dict(action_mask=mask.copy(), step_count=np.full(A, step_count, np.int32))
        if self.add_global_state:  # built BEFORE ids are added (jumanji.py:53-59,77-91)
            obs["global_state"] = np.tile(view.reshape(-1), (A, 1))
        if self.add_agent_id:  # ids are PREPENDED (observation.py:45-50)
            view = np.concatenate([np.eye(A, dtype=np.float32), view], axis=-1)
        obs["agents_view"] = view
        return obs

    def reset(self, key):
        key, reset_key = tf.split(key)  # episode_metrics.py:61
        inner = generator(self.spec, reset_key)
        obs = self._observation(observe(self.spec, inner), inner["mask"], 0)
        state = dict(inner=inner, key=np.asarray(key, np.uint32), run_ret=np.float32(0.0),
                     run_len=0, ep_ret=np.float32(0.0), ep_len=0)
        ts = dict(obs=obs, reward=np.zeros(self.spec.A, np.float32), done=False,
                  discount=np.ones(self.spec.A, np.float32),
                  metrics=dict(episode_return=np.float32(0.0), episode_length=0,
                               is_terminal_step=False))
        return state, ts

    def step(self, state, action):
        inner, raw = step(self.spec, state["inner"], action)
        done = raw["done"]
        if done and self.auto_reset:
            key, _ = tf.split(inner["key"])  # auto_reset_wrapper.py:74
            inner = generator(self.spec, key)
            obs = self._observation(observe(self.spec, inner), inner["mask"], 0)
        else:
            obs = self._observation(raw["agents_view"], raw["action_mask"], raw["step_count"])
        reward = np.full(self.spec.A, raw["reward"], np.float32)  # jumanji.py:142
        # RecordEpisodeMetrics.step, episode_metrics.py:78-111 (float32 / int32 arithmetic)
        not_done = np.float32(0.0 if done else 1.0)
        d = np.float32(1.0 if done else 0.0)
        new_ret = np.float32(state["run_ret"] + np.float32(reward.mean()))
        new_len = state["run_len"] + 1
        ep_ret = np.float32(state["ep_ret"] * not_done + new_ret * d)
        ep_len = state["ep_len"] * (0 if done else 1) + new_len * (1 if done else 0)
        nstate = dict(inner=inner, key=state["key"], run_ret=np.float32(new_ret * not_done),
                      run_len=new_len * (0 if done else 1), ep_ret=ep_ret, ep_len=ep_len)
        ts = dict(obs=obs, reward=reward, done=done,
                  discount=np.full(self.spec.A, raw["discount"], np.float32),
                  metrics=dict(episode_return=ep_ret, episode_length=ep_len,
                               is_terminal_step=done))
        return nstate, ts
