"""Philox4x32-10 known-answer tests (Random123 kat_vectors) for the
C oracle and the pure-Python restatement, and cross-checks of every draw stream."""
import ctypes as C

import numpy as np

from oracle import aqua_oracle as ao
from oracle import draws

KATS = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


KATS_2X32 = [  # Random123 kat_vectors, philox2x32 10
    ((0, 0), 0, (0xff1dae59, 0x6cd10df2)),
    ((0xffffffff, 0xffffffff), 0xffffffff, (0x2c3f628b, 0xab4fd7ad)),
    ((0x243f6a88, 0x85a308d3), 0x13198a2e, (0xdd7ce038, 0xf62a4c12)),
]


def test_philox2x32_kat():
    L = ao.lib()
    for ctr, key, want in KATS_2X32:
        assert draws.philox2x32_10(ctr, key) == want
        c = (C.c_uint32 * 2)(*ctr)
        o = (C.c_uint32 * 2)()
        L.aqo_philox2x32_10(c, key, o)
        assert tuple(o) == want
    for seed, tag in [(0, 3), (2 ** 63 + 12345, 4), (77, 3)]:
        assert L.aqo_key32(seed, tag) == draws.key32(seed, tag)


def _c_philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    ao.lib().aqo_philox4x32_10(c, k, o)
    return tuple(o)


def test_philox_kat_python():
    for ctr, key, want in KATS:
        assert draws.philox4x32_10(ctr, key) == want


def test_philox_kat_c():
    for ctr, key, want in KATS:
        assert _c_philox(ctr, key) == want


def test_draw_streams_c_equals_python():
    L = ao.lib()
    rs = np.random.RandomState(0)
    for _ in range(300):
        seed = int(rs.randint(0, 2 ** 62))
        env = int(rs.randint(0, 2 ** 32))
        ep = int(rs.randint(0, 2 ** 21))
        n = int(rs.randint(0, 50))
        t = int(rs.randint(0, 1002))
        out5 = (C.c_double * 5)()
        L.aqo_draw_reset_sample(seed, env, ep, n, out5)
        assert np.array_equal(np.array(out5), draws.reset_sample(seed, env, ep, n))
        out2 = (C.c_double * 2)()
        L.aqo_draw_wave_init(seed, env, ep, 1, out2)
        assert np.array_equal(np.array(out2), draws.wave_init(seed, env, ep))
        L.aqo_draw_wave_noise(seed, env, ep, t, 1, out2)
        assert np.array_equal(np.array(out2), draws.wave_noise(seed, env, ep, t))
        step = int(rs.randint(0, 2 ** 40))
        assert L.aqo_draw_action_discrete(seed, env, step) == draws.action_discrete(seed, env, step)
        f2 = (C.c_float * 2)()
        L.aqo_draw_action_continuous(seed, env, step, f2)
        assert np.array_equal(np.array(f2, dtype=np.float32), draws.action_continuous(seed, env, step))


def test_draw_ranges():
    rs = np.random.RandomState(1)
    v = np.array([draws.reset_sample(5, int(e), 3, 0) for e in rs.randint(0, 2 ** 32, 2000)])
    assert v[:, [0, 1, 3, 4]].min() >= 0.0 and v[:, [0, 1, 3, 4]].max() < 100.0
    assert v[:, 2].min() >= -np.pi and v[:, 2].max() < np.pi
    # extreme words
    assert float(draws.u01(0xffffffff)) < 1.0 and float(draws.u01(0)) == 0.0
    w = np.array([draws.wave_noise(5, int(e), 3, 7) for e in rs.randint(0, 2 ** 32, 2000)])
    assert np.abs(w).max() <= 0.001
    a = np.array([draws.action_continuous(5, int(e), 9) for e in rs.randint(0, 2 ** 32, 2000)])
    assert a.min() >= 0.2 and a.max() <= 0.5
    d = np.array([draws.action_discrete(5, int(e), 9) for e in rs.randint(0, 2 ** 32, 3000)])
    assert set(d.tolist()) == {0, 1, 2}
