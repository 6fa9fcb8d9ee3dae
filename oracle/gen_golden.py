#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by EXECUTING the unmodified
reference environment (`/root/reference/gym_aqua`) under the `gym` stand-in.

TEST INFRASTRUCTURE ONLY.  Runs in the build container only (the GPU box has no
/root/reference); the resulting .npz files are committed.  Nothing here imports
the C oracle or the product package: the draws come from `oracle/draws.py`
(pure Python Philox), the dynamics from the reference itself.

Injection points:
  * `observation_space.sample` / `wave_space.sample` (instance attributes) feed the
    reset draws of aqua.py:103-115,124;
  * the module global `np` of `gym_aqua.envs.aqua` is wrapped so that
    `np.random.uniform` (aqua.py:188) returns the wave increment of the step.

    python oracle/gen_golden.py            # (re)writes tests/golden/*.npz
    python oracle/gen_golden.py wide       # only the 256-env x 1000-step fixtures (wide_*.npz)
"""
import contextlib
import io
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REFERENCE = os.environ.get("AQUA_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "gym_standin"))
sys.path.insert(0, REFERENCE)
sys.path.insert(0, HERE)

import draws  # noqa: E402
import gym  # noqa: E402  (stand-in)
import gym_aqua  # noqa: E402,F401  (the reference: registers the six ids)
from gym_aqua.envs import aqua as ref_aqua  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
EPISODE_NONE = 0x1FFFFF
EPISODE_MASK = 0x1FFFFF


# ---------------------------------------------------------------- numpy proxy
class _RandomProxy(object):
    def __init__(self, real):
        self._real = real
        self.next_uniform = None

    def uniform(self, low, high, size=None):
        assert self.next_uniform is not None, "wave noise not injected"
        v, self.next_uniform = self.next_uniform, None
        return np.array(v, dtype=np.float64)

    def __getattr__(self, k):
        return getattr(self._real, k)


class _NumpyProxy(object):
    def __init__(self, real):
        self._real = real
        self.random = _RandomProxy(real.random)

    def __getattr__(self, k):
        return getattr(self._real, k)


PROXY = _NumpyProxy(np)
ref_aqua.np = PROXY  # aqua.py resolves `np.random.uniform` through its module global


# ---------------------------------------------------------------- layouts
def dense64_table():
    """The frozen obstacle-dense stress map of SURVEY.md section 8d (C4), float32-rounded."""
    rs = np.random.RandomState(64)
    rows = []
    for k in range(64):
        cx, cy = rs.uniform(8.0, 92.0, 2)
        if k % 2 == 0:
            rows.append([cx, cy, 0.0, rs.uniform(1.5, 4.0), 0.0])
        else:
            w, h = rs.uniform(3.0, 8.0, 2)
            rows.append([cx, cy, 1.0, w, h])
    return np.array(rows, dtype=np.float32).astype(np.float64)


def table_to_reference_list(tab):
    out = []
    for cx, cy, typ, a, b in tab:
        if typ == 0.0:
            out.append((np.array([cx, cy]), "c", a))
        else:
            out.append((np.array([cx, cy]), "r", (a, b)))
    return out


def reference_list_to_table(obstacles):
    rows = []
    for pos, kind, dims in obstacles:
        if kind == "c":
            rows.append([pos[0], pos[1], 0.0, dims, 0.0])
        else:
            rows.append([pos[0], pos[1], 1.0, dims[0], dims[1]])
    return np.array(rows, dtype=np.float64).reshape(-1, 5)


# ---------------------------------------------------------------- one reference env + injected draws
class RefEnv(object):
    def __init__(self, env_id, kwargs, seed, global_id):
        self.env = gym.make(env_id, **kwargs)
        self.seed, self.gid = seed, global_id
        self.episode = EPISODE_NONE
        self.n = 0
        self.env.observation_space.sample = self._sample_obs
        self.env.wave_space.sample = self._sample_wave

    def _sample_obs(self):
        v = draws.reset_sample(self.seed, self.gid, self.episode, self.n)
        self.n += 1
        return v

    def _sample_wave(self):
        return draws.wave_init(self.seed, self.gid, self.episode, bool(self.env.has_waves))

    def reset(self):
        self.episode = (self.episode + 1) & EPISODE_MASK
        self.n = 0
        obs = self.env.reset()
        return obs, self.n

    def step(self, action, noise=None):
        e = self.env
        if noise is None:
            noise = draws.wave_noise(self.seed, self.gid, self.episode, e.time, bool(e.has_waves))
        PROXY.random.next_uniform = noise
        with contextlib.redirect_stdout(io.StringIO()):  # the out-of-bounds print, aqua.py:146
            obs, reward, done, info = e.step(action)
        PROXY.random.next_uniform = None
        flags = (1 if done else 0) | (2 if info['Termination.collided'] else 0) | \
                (4 if info['Termination.time'] else 0) | (8 if info['Termination.success'] else 0)
        return obs, float(reward), flags, noise

    def margin(self):
        e = self.env
        m = abs(e._distance_boat_from_nearest_border())          # aqua.py:404-407
        m = min(m, abs(e._distance_boat_from_nearest_obstacle()))  # aqua.py:409-418 (inf if none)
        return min(m, abs(e._distance_boat_from_goal()))           # aqua.py:401-402


# ---------------------------------------------------------------- policies
def seek_action(obs):
    """Bearing controller (own rule, used only to obtain goal arrivals)."""
    x, y, th, gx, gy = obs
    want = math.atan2(-(gx - x), gy - y)
    d = (want - th + math.pi) % (2 * math.pi) - math.pi
    if d > 0.14:
        return 0
    if d < -0.14:
        return 1
    return 2


def make_action(policy, continuous, seed, gid, step, obs):
    if continuous:
        a = draws.action_continuous(seed, gid, step)
        if policy == "oob":  # out-of-range thrusts exercise the clip branch aqua.py:145-150
            a = (np.float32(0.1) + (a - np.float32(0.2)) * np.float32(5.0 / 3.0)).astype(np.float32)
        return a
    if policy == "circle":
        return 0
    if policy == "seek" and draws.action_discrete(seed ^ 0x5EE, gid, step) != 0:
        return seek_action(obs)
    return int(draws.action_discrete(seed, gid, step))


# ---------------------------------------------------------------- rollout fixtures
def gen_rollout(name, env_id, kwargs, policy, n_envs, steps, seed, env_id_base, table):
    continuous = "Continuous" in env_id
    envs = [RefEnv(env_id, kwargs, seed, env_id_base + i) for i in range(n_envs)]
    T, N = steps, n_envs
    actions = np.zeros((T, N, 2), np.float32) if continuous else np.zeros((T, N), np.int32)
    noise = np.zeros((T, N, 2), np.float64)
    obs_step = np.zeros((T, N, 5), np.float64)    # obs returned by step (terminal obs on done)
    wave_step = np.zeros((T, N, 2), np.float64)   # wave speed after the step
    time_step = np.zeros((T, N), np.int32)
    reward = np.zeros((T, N), np.float64)
    flags = np.zeros((T, N), np.uint8)
    margin = np.zeros((T, N), np.float64)
    init_state = np.zeros((9, N), np.float64)
    resets = []  # rows {t, env, x, y, theta, wx, wy, gx, gy, episode, n_samples}
    cur = []
    for i, e in enumerate(envs):
        obs, n = e.reset()
        cur.append(obs)
        init_state[:, i] = [obs[0], obs[1], obs[2], e.env._wave_speed[0], e.env._wave_speed[1],
                            obs[3], obs[4], e.env.time, e.episode]
        resets.append([-1, i, obs[0], obs[1], obs[2], e.env._wave_speed[0], e.env._wave_speed[1],
                       obs[3], obs[4], e.episode, n])
    for t in range(T):
        for i, e in enumerate(envs):
            a = make_action(policy, continuous, seed, e.gid, t, cur[i])
            actions[t, i] = a
            act = np.array(a, dtype=np.float64) if continuous else int(a)
            obs, r, f, nz = e.step(act)
            noise[t, i] = nz
            obs_step[t, i] = obs
            wave_step[t, i] = e.env._wave_speed
            time_step[t, i] = e.env.time
            reward[t, i] = r
            flags[t, i] = f
            margin[t, i] = e.margin()
            if f & 1:
                obs, n = e.reset()
                resets.append([t, i, obs[0], obs[1], obs[2], e.env._wave_speed[0],
                               e.env._wave_speed[1], obs[3], obs[4], e.episode, n])
            cur[i] = obs
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"),
        kind="rollout", env_id=env_id, continuous=continuous, policy=policy,
        waves=bool(kwargs.get("waves", True)), random_boat=bool(kwargs.get("random_boat", True)),
        random_goal=bool(kwargs.get("random_goal", True)), seed=np.uint64(seed),
        env_id_base=np.uint64(env_id_base), table=table, init_state=init_state, actions=actions,
        noise=noise.astype(np.float32), obs_step=obs_step, wave_step=wave_step,
        time_step=time_step, reward=reward, flags=flags, margin=margin,
        resets=np.array(resets, dtype=np.float64))
    n_done = int((flags & 1).sum())
    print("%-34s T=%d N=%d done=%d collided=%d time=%d success=%d min-margin=%.3g" % (
        name, T, N, n_done, int(((flags & 2) > 0).sum()), int(((flags & 4) > 0).sum()),
        int(((flags & 8) > 0).sum()), margin.min()))


# ---------------------------------------------------------------- wide rollout fixtures
def gen_wide(name, env_id, kwargs, n_envs, steps, seed, env_id_base, table, every=25):
    """256 envs x 1000 steps per layout in compact form: the flags of EVERY
    step, every reset (state, episode, samples drawn), and full snapshots (state after auto-reset,
    reward) every `every` steps.  A free-running rollout (own Philox draws, own resets) can only
    reproduce these if every step and every reset in between was right.  Actions: the synthetic
    random policy (draws.action_discrete / action_continuous keyed by (seed, env id, step)); the
    discrete ones are stored, the continuous ones are regenerated by the tests."""
    continuous = "Continuous" in env_id
    envs = [RefEnv(env_id, kwargs, seed, env_id_base + i) for i in range(n_envs)]
    T, N = steps, n_envs
    actions = np.zeros((T, N), np.int8)
    flags = np.zeros((T, N), np.uint8)
    snaps = np.zeros((T // every, 8, N), np.float64)   # x, y, theta, wx, wy, time, episode, reward of that step
    init_state = np.zeros((9, N), np.float64)
    resets = []
    min_margin = np.inf
    for i, e in enumerate(envs):
        obs, n = e.reset()
        init_state[:, i] = [obs[0], obs[1], obs[2], e.env._wave_speed[0], e.env._wave_speed[1],
                            obs[3], obs[4], e.env.time, e.episode]
    for t in range(T):
        for i, e in enumerate(envs):
            if continuous:
                a = draws.action_continuous(seed, e.gid, t)
                act = np.array(a, dtype=np.float64)
            else:
                a = int(draws.action_discrete(seed, e.gid, t))
                actions[t, i] = a
                act = a
            obs, r, f, _ = e.step(act)
            flags[t, i] = f
            min_margin = min(min_margin, e.margin())
            if f & 1:
                obs, n = e.reset()
                resets.append([t, i, obs[0], obs[1], obs[2], e.env._wave_speed[0], e.env._wave_speed[1],
                               obs[3], obs[4], e.episode, n])
            if (t + 1) % every == 0:
                snaps[t // every, :, i] = [obs[0], obs[1], obs[2], e.env._wave_speed[0], e.env._wave_speed[1],
                                           e.env.time, e.episode, r]
    goals = np.array([[e.env._goal_state[0], e.env._goal_state[1]] for e in envs]).T
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"),
        kind="wide", env_id=env_id, continuous=continuous, policy="random", every=every,
        waves=bool(kwargs.get("waves", True)), random_boat=True, random_goal=True, seed=np.uint64(seed),
        env_id_base=np.uint64(env_id_base), table=table, init_state=init_state, actions=actions, flags=flags,
        snaps=snaps, final_goal=goals, resets=np.array(resets, dtype=np.float64), min_margin=min_margin)
    print("%-34s T=%d N=%d done=%d collided=%d time=%d success=%d min-margin=%.3g" % (
        name, T, N, int((flags & 1).sum()), int(((flags & 2) > 0).sum()), int(((flags & 4) > 0).sum()),
        int(((flags & 8) > 0).sum()), min_margin), flush=True)


# ---------------------------------------------------------------- single-step known-answer fixtures
def gen_single_step(name, env_id, kwargs, n_cases, seed, table):
    """Random (and deliberately edgy) states pushed straight into the reference's
    attributes, one `step` each: in = [x, y, theta, wx, wy, gx, gy, time], action, noise."""
    continuous = "Continuous" in env_id
    rs = np.random.RandomState(seed)
    e = RefEnv(env_id, kwargs, seed, 0)
    e.reset()
    env = e.env
    waves = bool(env.has_waves)
    state_in = np.zeros((n_cases, 8), np.float64)
    actions = np.zeros((n_cases, 2), np.float32) if continuous else np.zeros(n_cases, np.int32)
    noise = np.zeros((n_cases, 2), np.float32)
    state_out = np.zeros((n_cases, 8), np.float64)
    reward = np.zeros(n_cases, np.float64)
    flags = np.zeros(n_cases, np.uint8)
    margin = np.zeros(n_cases, np.float64)
    f32 = lambda v: np.float64(np.float32(v))
    for k in range(n_cases):
        mode = k % 8
        gx, gy = f32(rs.uniform(2.5, 97.5)), f32(rs.uniform(2.5, 97.5))
        x, y = f32(rs.uniform(2.0, 98.0)), f32(rs.uniform(2.0, 98.0))
        th = f32(rs.uniform(-3.1415925, 3.1415925))
        tm = int(rs.randint(0, 1000))
        if mode == 1:    # hugging a border
            side = rs.randint(4)
            off = f32(2.5 + rs.uniform(-0.6, 0.6))
            if side == 0: x = off
            elif side == 1: x = f32(100.0 - off)
            elif side == 2: y = off
            else: y = f32(100.0 - off)
        elif mode == 2:  # around the goal rim (goal radius + boat radius = 5)
            ang, rad = rs.uniform(0, 2 * np.pi), 5.0 + rs.uniform(-0.6, 0.6)
            x, y = f32(gx + rad * np.cos(ang)), f32(gy + rad * np.sin(ang))
        elif mode == 3 and len(table):  # around an obstacle rim
            o = table[rs.randint(len(table))]
            ang = rs.uniform(0, 2 * np.pi)
            ext = (o[3] if o[2] == 0 else 0.5 * max(o[3], o[4])) + 2.5 + rs.uniform(-0.7, 0.7)
            x, y = f32(o[0] + ext * np.cos(ang)), f32(o[1] + ext * np.sin(ang))
        elif mode == 4:  # time limit: 1000 -> 1001 triggers Termination.time (aqua.py:203)
            tm = int(rs.choice([999, 1000, 1001]))
            if rs.rand() < 0.5:  # inside the goal at the deadline: time wins over goal
                x, y = f32(gx + rs.uniform(-1, 1)), f32(gy + rs.uniform(-1, 1))
        elif mode == 5:  # heading next to the +-pi seam
            th = f32(rs.choice([-1, 1]) * (3.1415925 - rs.uniform(0, 0.13)))
        wx, wy = (f32(rs.uniform(-0.05, 0.05)), f32(rs.uniform(-0.05, 0.05))) if waves else (0., 0.)
        if waves and mode == 6:  # wave speed at the clip (aqua.py:189-191)
            wx = f32(rs.choice([-0.05, 0.05]))
        nz = np.array([np.float32(rs.uniform(-0.001, 0.001)),
                       np.float32(rs.uniform(-0.001, 0.001))]) if waves else np.zeros(2, np.float32)
        if continuous:
            lo, hi = (0.2, 0.5) if mode != 7 else (0.0, 0.8)
            a = np.array([rs.uniform(lo, hi), rs.uniform(lo, hi)], dtype=np.float32)
            if mode == 0 and rs.rand() < 0.3:
                a[1] = a[0]  # equal thrusts: the epsilon branch aqua.py:159-160
            act = a.astype(np.float64)
        else:
            a = int(rs.randint(3))
            act = a
        env._boat_state = np.array([x, y, th], dtype=np.float64)
        env._goal_state = np.array([gx, gy], dtype=np.float64)
        env._wave_speed = np.array([wx, wy], dtype=np.float64)
        env.time = tm
        obs, r, f, _ = e.step(act, noise=nz.astype(np.float64))
        state_in[k] = [x, y, th, wx, wy, gx, gy, tm]
        actions[k] = a
        noise[k] = nz
        state_out[k] = [obs[0], obs[1], obs[2], env._wave_speed[0], env._wave_speed[1],
                        obs[3], obs[4], env.time]
        reward[k], flags[k], margin[k] = r, f, e.margin()
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"), kind="single_step", env_id=env_id,
        continuous=continuous, waves=waves, table=table, state_in=state_in, actions=actions,
        noise=noise, state_out=state_out, reward=reward, flags=flags, margin=margin)
    print("%-34s cases=%d done=%d collided=%d time=%d success=%d" % (
        name, n_cases, int((flags & 1).sum()), int(((flags & 2) > 0).sum()),
        int(((flags & 4) > 0).sum()), int(((flags & 8) > 0).sum())))


def main_wide():
    dense = dense64_table()
    dense_list = table_to_reference_list(dense)
    none = np.zeros((0, 5))
    v1 = reference_list_to_table(gym.make("AquaEnv-v1").obstacles)
    v2 = reference_list_to_table(gym.make("AquaEnv-v2").obstacles)
    W = gen_wide
    W("wide_v0_discrete", "AquaEnv-v0", {}, 256, 1000, 31, 0, none)
    W("wide_v1_discrete", "AquaEnv-v1", {}, 256, 1000, 32, 70000, v1)
    W("wide_v2_discrete", "AquaEnv-v2", {}, 256, 1000, 33, 0, v2)
    W("wide_dense64_discrete", "AquaEnv-v0", {"obstacles": dense_list}, 256, 1000, 34, 1 << 20, dense)
    W("wide_v0_continuous", "AquaContinuousEnv-v0", {}, 256, 1000, 35, 0, none)
    W("wide_v1_continuous", "AquaContinuousEnv-v1", {}, 256, 1000, 36, 5, v1)
    W("wide_v2_continuous", "AquaContinuousEnv-v2", {}, 256, 1000, 37, 0, v2)
    W("wide_dense64_continuous", "AquaContinuousEnv-v0", {"obstacles": dense_list}, 256, 1000, 38, 3000000000, dense)
    W("wide_v1_discrete_nowaves", "AquaEnv-v1", {"waves": False}, 256, 1000, 39, 0, v1)


def main():
    os.makedirs(OUT, exist_ok=True)
    dense = dense64_table()
    np.save(os.path.join(OUT, "dense64_table.npy"), dense)
    dense_list = table_to_reference_list(dense)
    none = np.zeros((0, 5))
    v1 = reference_list_to_table(gym.make("AquaEnv-v1").obstacles)          # aqua.py:59-66
    v2 = reference_list_to_table(gym.make("AquaEnv-v2").obstacles)          # gym_aqua/__init__.py:24-31
    np.save(os.path.join(OUT, "v1_table.npy"), v1)
    np.save(os.path.join(OUT, "v2_table.npy"), v2)

    R = gen_rollout
    R("roll_v0_discrete_random", "AquaEnv-v0", {}, "random", 8, 1000, 11, 0, none)
    R("roll_v1_discrete_random", "AquaEnv-v1", {}, "random", 8, 1000, 12, 1000, v1)
    R("roll_v2_discrete_random", "AquaEnv-v2", {}, "random", 8, 1000, 13, 5, v2)
    R("roll_dense64_discrete_random", "AquaEnv-v0", {"obstacles": dense_list}, "random", 8, 1000,
      14, 123456, dense)
    R("roll_v0_continuous_random", "AquaContinuousEnv-v0", {}, "random", 8, 1000, 15, 0, none)
    R("roll_v1_continuous_random", "AquaContinuousEnv-v1", {}, "random", 8, 1000, 16, 77, v1)
    R("roll_v2_continuous_oob", "AquaContinuousEnv-v2", {}, "oob", 8, 1000, 17, 0, v2)
    R("roll_dense64_continuous_random", "AquaContinuousEnv-v0", {"obstacles": dense_list},
      "random", 8, 1000, 18, 4000000000, dense)
    R("roll_v0_discrete_seek", "AquaEnv-v0", {}, "seek", 8, 1000, 19, 0, none)
    R("roll_v1_discrete_seek", "AquaEnv-v1", {}, "seek", 8, 1000, 20, 0, v1)
    R("roll_v1_discrete_nowaves", "AquaEnv-v1", {"waves": False}, "random", 8, 1000, 21, 0, v1)
    R("roll_v0_fixed_circle_timeout", "AquaEnv-v0",
      {"waves": False, "random_boat": False, "random_goal": False}, "circle", 2, 1100, 22, 0, none)
    R("roll_v1_fixed_goal_seek", "AquaEnv-v1", {"random_goal": False}, "seek", 4, 1000, 23, 0, v1)

    S = gen_single_step
    S("step_v0_discrete", "AquaEnv-v0", {}, 4096, 101, none)
    S("step_v1_discrete", "AquaEnv-v1", {}, 4096, 102, v1)
    S("step_v2_discrete", "AquaEnv-v2", {}, 4096, 103, v2)
    S("step_dense64_discrete", "AquaEnv-v0", {"obstacles": dense_list}, 4096, 104, dense)
    S("step_v1_continuous", "AquaContinuousEnv-v1", {}, 4096, 105, v1)
    S("step_dense64_continuous", "AquaContinuousEnv-v0", {"obstacles": dense_list}, 4096, 106, dense)
    S("step_v1_discrete_nowaves", "AquaEnv-v1", {"waves": False}, 2048, 107, v1)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "wide":   # only the 256-env fixtures (about 10 minutes)
        main_wide()
    else:
        main()
        main_wide()
