"""Pure-Python restatement of the draw streams (Philox4x32-10 + float32 mapping).

TEST INFRASTRUCTURE ONLY.  Independent of the C oracle so that the golden-vector
generator (`gen_golden.py`) does not depend on the code it pins; the two are
cross-checked in tests/test_philox.py.  The layout is specified in DESIGN.md
section 4:

    key     = (seed & 0xffffffff, seed >> 32)
    counter = (global_env_id, c1, c2, stream_tag)
      RESET      c1 = episode, c2 = try index n   -> one `observation_space.sample()`
      WAVE_INIT  c1 = episode, c2 = 0             -> `wave_space.sample()`
      WAVE_NOISE c1 = episode, c2 = t >> 1        -> two steps of `np.random.uniform(-var, var, 2)`
      ACTION     c1 = step lo, c2 = step hi       -> synthetic random policy

All values are computed in float32 (single fused multiply-add where noted) and
then widened to float64, so a float32 kernel, a float64 kernel and the float64
oracle consume bit-identical numbers.
"""
import numpy as np

M0, M1 = 0xD2511F53, 0xCD9E8D57
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK32 = 0xFFFFFFFF

STREAM_RESET, STREAM_WAVE_INIT, STREAM_WAVE_NOISE, STREAM_ACTION = 1, 2, 3, 4

PI_DN_F = np.float32(3.1415925)      # largest float32 below pi
TWO_PI_DN_F = np.float32(6.283185)   # 2 * PI_DN_F exactly


def philox4x32_10(ctr, key):
    c0, c1, c2, c3 = (int(v) & MASK32 for v in ctr)
    k0, k1 = (int(v) & MASK32 for v in key)
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & MASK32, p1 & MASK32, \
                         ((p0 >> 32) ^ c3 ^ k1) & MASK32, p0 & MASK32
        k0 = (k0 + W0) & MASK32
        k1 = (k1 + W1) & MASK32
    return c0, c1, c2, c3


def _philox_env(seed, env_id, c1, c2, tag):
    return philox4x32_10((env_id, c1, c2, tag), (seed & MASK32, (seed >> 32) & MASK32))


def u01(u):
    """24 high bits -> float32 in [0, 1)."""
    return np.float32(u >> 8) * np.float32(2.0 ** -24)


def _fmaf(a, b, c):
    """float32 fused multiply-add.  a has <= 24 significant bits and b, c are
    float32, so a*b + c is exact in float64 (<= 53 bits) and one rounding to
    float32 equals fmaf(a, b, c)."""
    return np.float32(np.float64(a) * np.float64(np.float32(b)) + np.float64(np.float32(c)))


def reset_sample(seed, env_id, episode, n):
    r = _philox_env(seed, env_id, episode, n, STREAM_RESET)
    hundred = np.float32(100.0)
    return np.array([
        np.float32(u01(r[0]) * hundred),
        np.float32(u01(r[1]) * hundred),
        _fmaf(u01(r[2]), TWO_PI_DN_F, -PI_DN_F),
        np.float32(u01(r[2]) * hundred),
        np.float32(u01(r[3]) * hundred)], dtype=np.float64)


def wave_init(seed, env_id, episode, waves=True):
    if not waves:
        return np.zeros(2)
    r = _philox_env(seed, env_id, episode, 0, STREAM_WAVE_INIT)
    return np.array([_fmaf(u01(r[0]), 0.1, -0.05), _fmaf(u01(r[1]), 0.1, -0.05)],
                    dtype=np.float64)


def wave_noise(seed, env_id, episode, t, waves=True):
    """t = env time before the step's increment."""
    if not waves:
        return np.zeros(2)
    r = _philox_env(seed, env_id, episode, t >> 1, STREAM_WAVE_NOISE)
    a, b = (r[2], r[3]) if (t & 1) else (r[0], r[1])
    return np.array([_fmaf(u01(a), 0.002, -0.001), _fmaf(u01(b), 0.002, -0.001)],
                    dtype=np.float64)


def action_discrete(seed, env_id, step):
    r = _philox_env(seed, env_id, step & MASK32, (step >> 32) & MASK32, STREAM_ACTION)
    return (r[0] * 3) >> 32


def action_continuous(seed, env_id, step):
    r = _philox_env(seed, env_id, step & MASK32, (step >> 32) & MASK32, STREAM_ACTION)
    return np.array([_fmaf(u01(r[0]), 0.3, 0.2), _fmaf(u01(r[1]), 0.3, 0.2)], dtype=np.float32)
