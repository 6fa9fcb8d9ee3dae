"""Pure-Python restatement of the draw streams (Philox4x32-10 + float32 mapping).

TEST INFRASTRUCTURE ONLY.  Independent of the C oracle so that the golden-vector
generator (`gen_golden.py`) does not depend on the code it pins; the two are
cross-checked in tests/test_philox.py.  The layout is specified in DESIGN.md
section 4:

    Rare draws (4 words each): Philox4x32-10
      key     = (seed & 0xffffffff, seed >> 32)
      counter = (global_env_id, c1, c2, stream_tag)
        RESET      c1 = episode, c2 = try index n   -> one `observation_space.sample()`
    Per-step draws (2 words each): Philox2x32-10
      key32(tag) = Philox4x32-10(counter = (tag, 0, 0, 0), key as above)[0]
        WAVE_INIT  counter = (global_env_id, episode), key = key32(2)  -> `wave_space.sample()`
        WAVE_NOISE counter = (global_env_id, (episode << 11) | t), key = key32(3)
                   -> `np.random.uniform(-var, var, 2)` of the step taken at env time t
        ACTION     counter = (global_env_id, step lo), key = key32(4) + step hi
                   -> synthetic random policy

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


M2 = 0xD256D193


def philox2x32_10(ctr, key):
    c0, c1 = (int(v) & MASK32 for v in ctr)
    k = int(key) & MASK32
    for _ in range(10):
        p = M2 * c0
        c0, c1 = ((p >> 32) ^ k ^ c1) & MASK32, p & MASK32
        k = (k + W0) & MASK32
    return c0, c1


def key32(seed, tag):
    """32-bit key of a per-step stream, derived from the 64-bit seed."""
    return philox4x32_10((tag, 0, 0, 0), (seed & MASK32, (seed >> 32) & MASK32))[0]


TIME_BITS = 11


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
    r = philox2x32_10((env_id, episode), key32(seed, STREAM_WAVE_INIT))
    return np.array([_fmaf(u01(r[0]), 0.1, -0.05), _fmaf(u01(r[1]), 0.1, -0.05)],
                    dtype=np.float64)


def wave_noise(seed, env_id, episode, t, waves=True):
    """t = env time before the step's increment."""
    if not waves:
        return np.zeros(2)
    tw = ((episode << TIME_BITS) | min(t, (1 << TIME_BITS) - 1)) & MASK32
    a, b = philox2x32_10((env_id, tw), key32(seed, STREAM_WAVE_NOISE))
    return np.array([_fmaf(u01(a), 0.002, -0.001), _fmaf(u01(b), 0.002, -0.001)],
                    dtype=np.float64)


def _action_words(seed, env_id, step):
    key = (key32(seed, STREAM_ACTION) + ((step >> 32) & MASK32)) & MASK32
    return philox2x32_10((env_id, step & MASK32), key)


def action_discrete(seed, env_id, step):
    r = _action_words(seed, env_id, step)
    return (r[0] * 3) >> 32


def action_continuous(seed, env_id, step):
    r = _action_words(seed, env_id, step)
    return np.array([_fmaf(u01(r[0]), 0.3, 0.2), _fmaf(u01(r[1]), 0.3, 0.2)], dtype=np.float32)
