"""Pin the C oracle against golden vectors produced by executing the unmodified
reference (oracle/gen_golden.py).  The oracle follows the reference's own
formulation in float64, so the residue is libm-vs-NumPy rounding only: the ICC of a
straight move lies 1.25e8 px away (aqua.py:159-174), i.e. ~1.5e-8 px per step."""
import numpy as np
import pytest

from oracle import aqua_oracle as ao
from tests.helpers import (SCALE_POS, SCALE_REWARD, SCALE_WAVE, err_angle, err_scaled,
                           golden_names, load_golden, wide_actions)
from tests.conftest import GOLDEN

TOL = 1e-7  # scaled-relative; flags must be identical


def _make(g, n, base=None):
    return ao.OracleVecEnv(n, continuous=bool(g["continuous"]), obstacles=g["table"],
                           waves=bool(g["waves"]),
                           random_boat=bool(g.get("random_boat", True)),
                           random_goal=bool(g.get("random_goal", True)),
                           seed=int(g.get("seed", 0)),
                           env_id_base=int(g.get("env_id_base", 0)) if base is None else base)


@pytest.mark.parametrize("name", golden_names(GOLDEN, "step_"))
def test_single_step_golden(name):
    g = load_golden(GOLDEN, name)
    n = g["state_in"].shape[0]
    env = _make(g, n)
    st = np.zeros((9, n))
    st[:8] = g["state_in"].T
    env.set_state(st)
    _, reward, flags, margin = env.step(g["actions"], noise=g["noise"].astype(np.float64).T,
                                        autoreset=False, want_margin=True)
    out = env.get_state()
    assert np.array_equal(flags, g["flags"])
    ref = g["state_out"].T
    assert err_scaled(out[0], ref[0], SCALE_POS).max() < TOL
    assert err_scaled(out[1], ref[1], SCALE_POS).max() < TOL
    assert err_angle(out[2], ref[2]).max() < TOL
    assert err_scaled(out[3:5], ref[3:5], SCALE_WAVE).max() < TOL
    assert np.array_equal(out[5:7], ref[5:7])
    assert np.array_equal(out[7], ref[7])
    assert err_scaled(reward, g["reward"], SCALE_REWARD).max() < TOL
    assert np.abs(margin - g["margin"]).max() < 1e-6


@pytest.mark.parametrize("name", golden_names(GOLDEN, "roll_"))
def test_rollout_golden_free_running(name):
    """Free-running lock-step over the whole rollout, resets included (the oracle draws its
    own noise / reset samples from the Philox streams; nothing is teacher-forced)."""
    g = load_golden(GOLDEN, name)
    T, n = g["flags"].shape
    env = _make(g, n)
    obs = env.reset()
    init = g["init_state"]
    assert np.array_equal(env.get_state()[[0, 1, 2, 3, 4, 5, 6, 7, 8]], init)
    resets = {(int(r[0]), int(r[1])): r for r in g["resets"]}
    worst = 0.0
    for t in range(T):
        # autoreset=False first to look at the terminal state, then reset where done
        _, reward, flags, margin = env.step(g["actions"][t], autoreset=False, want_margin=True)
        st = env.get_state()
        assert np.array_equal(flags, g["flags"][t]), (name, t)
        o = g["obs_step"][t]
        e = max(err_scaled(st[0], o[:, 0], SCALE_POS).max(), err_scaled(st[1], o[:, 1], SCALE_POS).max(),
                err_angle(st[2], o[:, 2]).max(),
                err_scaled(st[3], g["wave_step"][t][:, 0], SCALE_WAVE).max(),
                err_scaled(st[4], g["wave_step"][t][:, 1], SCALE_WAVE).max(),
                err_scaled(reward, g["reward"][t], SCALE_REWARD).max())
        worst = max(worst, e)
        assert e < TOL, (name, t, e)
        assert np.array_equal(st[5], o[:, 3]) and np.array_equal(st[6], o[:, 4])
        assert np.array_equal(st[7], g["time_step"][t])
        assert np.abs(margin - g["margin"][t]).max() < 1e-5
        done = (flags & 1).astype(np.uint8)
        if done.any():
            env.reset(mask=done)
            st = env.get_state()
            for i in np.nonzero(done)[0]:
                r = resets[(t, int(i))]
                assert np.array_equal(st[[0, 1, 2, 3, 4, 5, 6, 8], i], r[2:10]), (name, t, i)
    assert worst < TOL


@pytest.mark.parametrize("name", golden_names(GOLDEN, "wide_"))
def test_oracle_reproduces_wide_golden(name):
    """256 envs x 1000 steps per layout recorded from the executed reference: the C oracle,
    free-running with its own draws and resets, gives the same flags at every step, the same state
    at every 25th step (to the reference's own round-off) and the same final goals."""
    g = load_golden(GOLDEN, name)
    T, n = g["flags"].shape
    every = int(g["every"])
    ref = ao.OracleVecEnv(n, continuous=bool(g["continuous"]), obstacles=np.asarray(g["table"], np.float64),
                          waves=bool(g["waves"]), seed=int(g["seed"]), env_id_base=int(g["env_id_base"]))
    ref.reset()
    assert np.array_equal(ref.get_state(), g["init_state"])
    for c in range(T // every):
        acts = wide_actions(g, c * every, (c + 1) * every)
        for k in range(every):
            _, r, f, _ = ref.step(acts[k], autoreset=True)
            assert np.array_equal(f, g["flags"][c * every + k]), (name, c, k)
        st, snap = ref.get_state(), g["snaps"][c]
        assert np.abs(st[:5] - snap[:5]).max() < 1e-7 and np.abs(r - snap[7]).max() < 1e-7
        assert np.array_equal(st[7], snap[5]) and np.array_equal(st[8], snap[6])
    assert np.array_equal(ref.get_state()[5:7], g["final_goal"])
