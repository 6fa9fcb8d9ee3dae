"""A gym-style vector env over the UNMODIFIED reference environment, for the CPU baseline.

TEST / BASELINE INFRASTRUCTURE ONLY (bench.py's `--impl reference` arm and `cpu_baseline` leg,
tests/).  It never imports the product package `aquaticgymenv_b200`.

The reference ships no vector env and `gym` is not installed, so this is
the stand-in for `gym.vector.AsyncVectorEnv` the north star asks for: P worker processes, each
owning M instances of the reference `gym_aqua` env (`gym.make(id, **kwargs)` under the `gym`
stand-in of oracle/gym_standin), stepped in lock-step from the parent: one `step(actions)` call
sends every worker its M actions over a pipe, the worker steps its envs one after the other
(aqua.py:135-213), resets the ones that finished (aqua.py:100-126, gym.vector's auto-reset) and
sends back observations, rewards and dones.

Where the reference comes from: `baseline/_ref/` (a verbatim copy of the reference's `gym_aqua`
and `main` packages made by `__graft_entry__.build()`; git-ignored, it travels to the GPU box
with the repo snapshot), else `/root/reference` (the build container).
"""
import multiprocessing as mp
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def reference_root():
    """Directory that holds the reference's `gym_aqua` package, or None."""
    for cand in (os.environ.get("AQUA_REFERENCE"), os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if cand and os.path.isdir(os.path.join(cand, "gym_aqua")):
            return cand
    return None


def import_reference():
    """-> the stand-in `gym` module with the reference's six ids registered (gym_aqua/__init__.py)."""
    root = reference_root()
    if root is None:
        raise ImportError("reference gym_aqua not found (baseline/_ref is created by __graft_entry__.build())")
    for p in (root, os.path.join(HERE, "gym_standin")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import gym
    import gym_aqua  # noqa: F401  (registers the ids)
    return gym


def _worker(conn, env_id, kwargs, m, seed):
    try:
        gym = import_reference()
        np.random.seed(seed)
        envs = [gym.make(env_id, **kwargs) for _ in range(m)]
        for i, e in enumerate(envs):
            e.observation_space.seed(seed * 100003 + i)
            e.wave_space.seed(seed * 100019 + i)
        obs = np.stack([e.reset() for e in envs])
        conn.send(("ready", obs))
        rew = np.zeros(m)
        done = np.zeros(m, bool)
        while True:
            cmd, arg = conn.recv()
            if cmd == "step":
                for i, e in enumerate(envs):
                    o, r, d, _ = e.step(arg[i])
                    if d:  # gym.vector auto-reset: the returned observation opens the next episode
                        o = e.reset()
                    obs[i] = o
                    rew[i] = r
                    done[i] = d
                conn.send((obs, rew, done))
            elif cmd == "run":  # `arg` steps with the env's own action sampler, no traffic: the upper bound
                t0 = time.perf_counter()
                n_done = 0
                for _ in range(arg):
                    for e in envs:
                        _, _, d, _ = e.step(e.action_space.sample())
                        if d:
                            e.reset()
                            n_done += 1
                conn.send((time.perf_counter() - t0, n_done))
            else:
                break
    except KeyboardInterrupt:
        pass
    finally:
        conn.close()


class RefVectorEnv(object):
    """P processes x M reference envs; `step(actions)` -> (obs [P*M, 5], reward, done)."""

    def __init__(self, env_id, num_envs, processes, seed=0, **kwargs):
        self.p = max(1, min(int(processes), int(num_envs)))
        self.m = (int(num_envs) + self.p - 1) // self.p
        self.num_envs = self.p * self.m
        ctx = mp.get_context("fork")
        self.conns, self.procs = [], []
        for w in range(self.p):
            a, b = ctx.Pipe()
            pr = ctx.Process(target=_worker, args=(b, env_id, kwargs, self.m, seed + w), daemon=True)
            pr.start()
            b.close()
            self.conns.append(a)
            self.procs.append(pr)
        self.obs = np.concatenate([c.recv()[1] for c in self.conns])

    def step(self, actions):
        for w, c in enumerate(self.conns):
            c.send(("step", actions[w * self.m:(w + 1) * self.m]))
        outs = [c.recv() for c in self.conns]
        self.obs = np.concatenate([o[0] for o in outs])
        return self.obs, np.concatenate([o[1] for o in outs]), np.concatenate([o[2] for o in outs])

    def run_free(self, steps):
        """Every worker steps its envs `steps` times on its own; -> (slowest worker's seconds, episodes)."""
        for c in self.conns:
            c.send(("run", int(steps)))
        outs = [c.recv() for c in self.conns]
        return max(o[0] for o in outs), sum(o[1] for o in outs)

    def close(self):
        for c in self.conns:
            try:
                c.send(("close", None))
                c.close()
            except Exception:
                pass
        for p in self.procs:
            p.join(timeout=5)
            if p.is_alive():
                p.terminate()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def time_reference(env_id, num_envs, processes, steps, warmup, seed=0, continuous=False, **kwargs):
    """Random-action rollout through the vector env: -> (env-steps/s, seconds, envs actually stepped)."""
    venv = RefVectorEnv(env_id, num_envs, processes, seed=seed, **kwargs)
    rs = np.random.RandomState(1234)
    n = venv.num_envs

    def actions():
        if continuous:
            return rs.uniform(0.2, 0.5, (n, 2))
        return rs.randint(0, 3, n)

    pool = [actions() for _ in range(4)]
    for k in range(warmup):
        venv.step(pool[k % 4])
    t0 = time.perf_counter()
    for k in range(steps):
        venv.step(pool[k % 4])
    dt = time.perf_counter() - t0
    venv.close()
    return n * steps / dt, dt, n
