"""Host-side check of the obstacle culling grid (aqua_obstacle_grid, no GPU): the kernels test a
dense table (more than 12 obstacles) only against the obstacles its grid lists for the cell a
position falls in, and skip the test altogether in cells marked blocked.  Both shortcuts must be
conservative with respect to the exact test of aqua.py:429-451, for every point of a cell."""
import os

import numpy as np
import pytest

from aquaticgymenv_b200 import _native as N

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
G = N.GRID_DIM


def _random_table(rs, k):
    rows = []
    for _ in range(k):
        cx, cy = rs.uniform(-3.0, 103.0, 2)
        if rs.rand() < 0.5:
            rows.append([cx, cy, 0.0, rs.choice([0.0, 0.3, 2.0, 6.0, 11.0]) + rs.uniform(0.0, 2.0), 0.0])
        else:
            w, h = rs.choice([0.2, 1.0, 5.0, 14.0], 2) + rs.uniform(0.0, 3.0, 2)
            rows.append([cx, cy, 1.0, w, h])
    return np.asarray(rows, np.float64).reshape(k, 5)


def _collides(rows, px, py):
    """[P, K] exact test in float64 (rows as returned by obstacle_grid)."""
    dx = np.maximum(np.abs(px[:, None] - rows[None, :, 0]) - rows[None, :, 2], 0.0)
    dy = np.maximum(np.abs(py[:, None] - rows[None, :, 1]) - rows[None, :, 3], 0.0)
    return dx * dx + dy * dy <= rows[None, :, 4]


def _cell_index(px, py):
    """the device's cell index: float32 position times float32(G / 100), truncated, clamped"""
    s = np.float32(G / 100.0)
    cx = np.clip((px.astype(np.float32) * s).astype(np.int32), 0, G - 1)
    cy = np.clip((py.astype(np.float32) * s).astype(np.int32), 0, G - 1)
    return cx, cy


def _sample_points(rs, rows, n):
    """uniform points, points on cell borders, and points hugging every obstacle's rim"""
    pts = [rs.uniform(0.0, 100.0, (n, 2))]
    edge = rs.randint(0, G + 1, (n // 4, 2)) * (100.0 / G) + rs.choice([-1e-4, 0.0, 1e-4], (n // 4, 2))
    edge[:, 1] = np.where(rs.rand(n // 4) < 0.5, edge[:, 1], rs.uniform(0, 100, n // 4))
    pts.append(edge)
    for r in rows:
        ang = rs.uniform(0, 2 * np.pi, 400)
        reach = np.sqrt(r[4]) + rs.choice([-1e-3, -1e-6, 0.0, 1e-6, 1e-3], 400)
        # a point at distance `reach` from the box: walk out from a point of the box
        bx, by = r[0] + rs.uniform(-1, 1, 400) * r[2], r[1] + rs.uniform(-1, 1, 400) * r[3]
        ox, oy = np.cos(ang), np.sin(ang)
        # push the start point to the box border along the outward direction component-wise
        bx = np.where(np.abs(ox) > 0, r[0] + np.sign(ox) * r[2], bx)
        by = np.where(np.abs(oy) > 0, r[1] + np.sign(oy) * r[3], by)
        pts.append(np.stack([bx + reach * ox, by + reach * oy], 1))
    p = np.concatenate(pts)
    p = p[(p[:, 0] >= 0) & (p[:, 0] < 100) & (p[:, 1] >= 0) & (p[:, 1] < 100)]
    # positions as the float32 kernels hold them
    return p[:, 0].astype(np.float32).astype(np.float64), p[:, 1].astype(np.float32).astype(np.float64)


def _check(table, expect_grid):
    rows, cells, use_grid = N.obstacle_grid(table)
    assert use_grid == expect_grid
    k = len(rows)
    rs = np.random.RandomState(k + 7)
    px, py = _sample_points(rs, rows, 60000)
    hit = _collides(rows, px, py)                      # [P, K]
    cx, cy = _cell_index(px, py)
    mask = cells[cy, cx]                               # [P] uint64
    blocked = mask == np.uint64(N.CELL_BLOCKED)
    # every point of a blocked cell collides
    assert hit[blocked].any(axis=1).all()
    # elsewhere the mask lists every obstacle the point collides with
    bits = ((mask[:, None] >> np.arange(k, dtype=np.uint64)[None, :]) & np.uint64(1)).astype(bool)
    missed = hit & ~bits & ~blocked[:, None]
    assert not missed.any(), np.argwhere(missed)[:5]
    return blocked.mean(), bits[~blocked].sum(1).mean()


@pytest.mark.parametrize("k", [13, 24, 40, 64])
def test_random_tables(k):
    rs = np.random.RandomState(3000 + k)
    for _ in range(4):
        _check(_random_table(rs, k), expect_grid=True)


def test_dense64_fixture():
    frac_blocked, mean_mask = _check(np.load(os.path.join(GOLDEN, "dense64_table.npy")), expect_grid=True)
    assert frac_blocked > 0.1 and mean_mask < 2.0  # the grid does cull on the stress map


@pytest.mark.parametrize("k", [0, 1, 5, 12])
def test_small_tables_are_tested_by_brute_force(k):
    rs = np.random.RandomState(k)
    rows, cells, use_grid = N.obstacle_grid(_random_table(rs, k))
    assert not use_grid and len(rows) == k


def test_full_mask_is_not_read_as_blocked():
    """64 identical circles: rim cells are reached by all 64 obstacles and lie inside none; their
    mask equals the blocked mark, so such a table must not use the grid."""
    table = np.tile(np.array([[50.0, 50.0, 0.0, 6.0, 0.0]]), (64, 1))
    rows, cells, use_grid = N.obstacle_grid(table)
    assert not use_grid
    # 63 of them plus one elsewhere: no cell is reached by all 64, the grid is fine again
    table[63] = [10.0, 10.0, 0.0, 1.0, 0.0]
    _check(table, expect_grid=True)


def test_rows_are_circles_first_with_squared_reach():
    table = np.array([[65.0, 85.0, 1.0, 5.0, 5.0], [15.0, 75.0, 0.0, 5.0, 0.0]])  # aqua.py:59-66 style
    rows, _, _ = N.obstacle_grid(table)
    assert np.allclose(rows[0], [15.0, 75.0, 0.0, 0.0, 7.5 ** 2])        # circle: radius + 2.5
    assert np.allclose(rows[1], [65.0, 85.0, 2.5, 2.5, 2.5 ** 2])        # rectangle: half extents, 2.5


@pytest.mark.parametrize("k", [13, 40, 64])
def test_cell16_words_follow_the_masks(k):
    """The 16-bit word the kernels read per position (aqua_obstacle_grid_ex) is the cell's mask in
    compact form: its first two obstacles (both always tested; 64 = the entry that never collides),
    the blocked mark, and the overflow bit exactly where the mask lists more than two."""
    rs = np.random.RandomState(900 + k)
    table = _random_table(rs, k) if k < 64 else np.load(os.path.join(GOLDEN, "dense64_table.npy"))
    rows, cells, cells16, use_grid = N.obstacle_grid16(table)
    assert use_grid
    overflow_seen = 0
    for cy in range(G):
        for cx in range(G):
            m, c = int(cells[cy, cx]), int(cells16[cy, cx])
            i0, i1 = c & 127, (c >> 7) & 127
            if m == N.CELL_BLOCKED:
                assert c & N.CELL16_BLOCKED and not c & N.CELL16_OVERFLOW and i0 == i1 == N.CELL16_NONE
                continue
            assert not c & N.CELL16_BLOCKED
            bits = [i for i in range(len(rows)) if (m >> i) & 1]
            want = (bits + [N.CELL16_NONE, N.CELL16_NONE])[:2]
            assert [i0, i1] == want
            assert bool(c & N.CELL16_OVERFLOW) == (len(bits) > 2)
            overflow_seen += len(bits) > 2
    if k == 64:  # the stress map: overflow cells exist (the fallback is exercised) but are rare
        assert 0 < overflow_seen < 0.03 * G * G
