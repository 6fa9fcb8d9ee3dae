// aqua_tables.cu -- host side of the kernels' constant data: the key schedule of the Philox draw
// streams and the obstacle tables (aqua.py:56-68, 373-390) with the culling grid of dense maps.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "aqua_device.cuh"

namespace aqua {

// ------------------------------------------------------------------------------------------
// host-side key schedule of the draw streams
// ------------------------------------------------------------------------------------------
static void philox4x32_10_host(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void make_philox_keys(uint64_t seed, PhiloxKeys* out) {
    const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    for (int r = 0; r < 10; ++r) {
        out->rk4[2 * r] = key[0] + (uint32_t)r * 0x9E3779B9u;
        out->rk4[2 * r + 1] = key[1] + (uint32_t)r * 0xBB67AE85u;
    }
    uint32_t* streams[3] = {out->winit, out->noise, out->action};
    const uint32_t tags[3] = {kStreamWaveInit, kStreamWaveNoise, kStreamAction};
    for (int s = 0; s < 3; ++s) {
        const uint32_t ctr[4] = {tags[s], 0u, 0u, 0u};
        uint32_t o[4];
        philox4x32_10_host(ctr, key, o);
        for (int r = 0; r < 10; ++r) streams[s][r] = o[0] + (uint32_t)r * 0x9E3779B9u;
    }
}

// ------------------------------------------------------------------------------------------
// host-side obstacle tables (aqua.py:56-68, 373-390)
// ------------------------------------------------------------------------------------------
int envs_per_thread(int precision) { return precision == AQUA_F64 ? Prec<double>::V : Prec<float>::V; }

size_t obstacle_table_bytes(int precision) {
    return precision == AQUA_F64 ? sizeof(ObstacleTable<double>) : sizeof(ObstacleTable<float>);
}

// circles first (stable), then rectangles; returns the number of circles
static int sorted_rows(int k, const double* rows5, double* out5) {
    int n = 0;
    for (int pass = 0; pass < 2; ++pass)
        for (int i = 0; i < k; ++i)
            if ((rows5[5 * i + 2] == 0.0) == (pass == 0)) {
                memcpy(out5 + 5 * n, rows5 + 5 * i, 5 * sizeof(double));
                ++n;
            }
    int k_circ = 0;
    for (int i = 0; i < k; ++i) k_circ += rows5[5 * i + 2] == 0.0;
    return k_circ;
}

// Classification of one grid cell (float64, 0.05 px safety margin that also covers the rounding of
// the cell index on the device).  *blocked: the whole cell lies inside one obstacle's clearance
// region (obstacle inflated by the radius 2.5) shrunk by the margin -- the distance to a convex
// obstacle is a convex function of the position, so its maximum over the cell is attained at a
// corner.  Otherwise the return value is the mask of the obstacles whose clearance region, grown
// by the margin, touches the cell.  (With 64 obstacles a genuine mask can equal kCellBlocked, hence
// the separate flag.)
static unsigned long long cell_mask(int k, const double* rows5, int cx, int cy, bool* blocked) {
    constexpr double kMargin = 0.05;
    const double cw = 100.0 / kGridDim;
    const double x0 = cx * cw, x1 = x0 + cw, y0 = cy * cw, y1 = y0 + cw;
    unsigned long long m = 0ull;
    for (int i = 0; i < k; ++i) {
        const double* o = rows5 + 5 * i;
        const bool circle = o[2] == 0.0;
        const double hx = circle ? 0.0 : o[3] / 2, hy = circle ? 0.0 : o[4] / 2;
        const double reach = circle ? o[3] + 2.5 : 2.5;
        // largest distance of a cell corner to the obstacle's box
        double far2 = 0.0;
        for (int c = 0; c < 4; ++c) {
            const double px = (c & 1) ? x1 : x0, py = (c & 2) ? y1 : y0;
            const double ex = std::max(std::fabs(px - o[0]) - hx, 0.0), ey = std::max(std::fabs(py - o[1]) - hy, 0.0);
            far2 = std::max(far2, ex * ex + ey * ey);
        }
        if (reach > kMargin && far2 <= (reach - kMargin) * (reach - kMargin)) {
            *blocked = true;
            return kCellBlocked;
        }
        // distance between the obstacle's box and the cell rectangle
        const double dx = std::max(std::max(x0 - (o[0] + hx), (o[0] - hx) - x1), 0.0);
        const double dy = std::max(std::max(y0 - (o[1] + hy), (o[1] - hy) - y1), 0.0);
        if (dx * dx + dy * dy <= (reach + kMargin) * (reach + kMargin)) m |= 1ull << i;
    }
    *blocked = false;
    return m;
}

// Returns whether the culling grid can be used: the table is dense enough to pay for it, and no
// cell that is not blocked carries a full mask (which would read as kCellBlocked; only possible
// with AQUA_MAX_OBSTACLES obstacles all reaching one 2 px cell -- such a table is tested by brute
// force).
template <typename R> static bool fill_table(int k, const double* rows5, ObstacleTable<R>* t) {
    memset(t, 0, sizeof(*t));
    bool grid_ok = k > kGridMinObstacles;
    for (int cy = 0; cy < kGridDim; ++cy)
        for (int cx = 0; cx < kGridDim; ++cx) {
            bool blocked = false;
            const unsigned long long m = cell_mask(k, rows5, cx, cy, &blocked);
            if (!blocked && m == kCellBlocked) grid_ok = false;  // a genuine full mask: ambiguous
            t->cell[cy * kGridDim + cx] = m;
            // the 16-bit word the kernels read: the first two listed obstacles, blocked, overflow
            uint32_t c = kCellNone | (kCellNone << 7);
            if (blocked) {
                c |= kCell16Blocked;
            } else {
                int found = 0;
                for (int i = 0; i < k; ++i)
                    if ((m >> i) & 1ull) {
                        if (found == 0) c = (c & ~127u) | (uint32_t)i;
                        else if (found == 1) c = (c & ~(127u << 7)) | ((uint32_t)i << 7);
                        ++found;
                    }
                if (found > 2) c |= kCell16Overflow;
            }
            t->cell16[cy * kGridDim + cx] = (unsigned short)c;
        }
    // unified entries: a circle is a box of zero extents and reach radius + 2.5
    for (int i = 0; i <= AQUA_MAX_OBSTACLES; ++i) {
        t->ubox[i].cx = t->ubox[i].cy = t->ubox[i].a = t->ubox[i].b = R(0);
        t->nr2[i] = R(1e30);  // never collides (kCellNone, unused rows)
    }
    for (int i = 0; i < k; ++i) {
        const double* o = rows5 + 5 * i;
        const bool circle = o[2] == 0.0;
        t->ubox[i].cx = (R)o[0];
        t->ubox[i].cy = (R)o[1];
        t->ubox[i].a = circle ? R(0) : (R)(o[3] / 2);
        t->ubox[i].b = circle ? R(0) : (R)(o[4] / 2);
        const double reach = circle ? o[3] + 2.5 : 2.5;
        t->nr2[i] = circle ? -(R)(reach * reach) : R(-kRectR2);
    }
    for (int i = 0; i < k; ++i) {
        const double* o = rows5 + 5 * i;
        const bool circle = o[2] == 0.0;
        t->box[i].cx = (R)o[0];
        t->box[i].cy = (R)o[1];
        if (circle) {
            const double reach = o[3] + 2.5;  // boat / goal radius 2.5, aqua.py:72,75
            t->box[i].a = (R)(reach * reach);
            t->box[i].b = R(0);
        } else {
            t->box[i].a = (R)(o[3] / 2);
            t->box[i].b = (R)(o[4] / 2);
        }
    }
    return grid_ok;
}

int build_obstacle_tables(int precision, int k, const double* rows5, void* tab_host, double* tab64_host,
                          int* use_grid) {
    double sorted[AQUA_MAX_OBSTACLES * 5];
    const int k_circ = sorted_rows(k, rows5, sorted);
    bool grid_ok;
    if (precision == AQUA_F64) grid_ok = fill_table<double>(k, sorted, reinterpret_cast<ObstacleTable<double>*>(tab_host));
    else grid_ok = fill_table<float>(k, sorted, reinterpret_cast<ObstacleTable<float>*>(tab_host));
    *use_grid = grid_ok ? 1 : 0;
    for (int i = 0; i < k; ++i) {
        const double* o = sorted + 5 * i;
        const bool circle = o[2] == 0.0;
        double* d = tab64_host + 5 * i;
        d[0] = o[0];
        d[1] = o[1];
        d[2] = circle ? 0.0 : o[3] / 2;
        d[3] = circle ? 0.0 : o[4] / 2;
        const double reach = circle ? o[3] + 2.5 : 2.5;
        d[4] = reach * reach;
    }
    return k_circ;
}

void fill_small_table(int precision, int k, const double* rows5, int use_grid, World* w) {
    // entries that never collide: a circle of squared reach -1e30, a rectangle far outside the world
    for (int i = 0; i < 4; ++i) {
        const float cdummy[4] = {0.f, 0.f, 1e30f, 0.f}, rdummy[4] = {1e4f, 1e4f, 0.f, 0.f};
        memcpy(w->small.c[i], cdummy, sizeof cdummy);
        memcpy(w->small.r[i], rdummy, sizeof rdummy);
    }
    w->small_mode = 0;
    if (precision != AQUA_F32 || k <= 0 || use_grid) return;
    ObstacleTable<float>* t = new ObstacleTable<float>;
    double sorted[AQUA_MAX_OBSTACLES * 5];
    const int k_circ = sorted_rows(k, rows5, sorted);
    fill_table<float>(k, sorted, t);  // the very values the generic kernels read
    if (k_circ <= 4 && k - k_circ <= 4) {
        for (int i = 0; i < k_circ; ++i) {
            w->small.c[i][0] = t->box[i].cx;
            w->small.c[i][1] = t->box[i].cy;
            w->small.c[i][2] = -t->box[i].a;
        }
        for (int i = k_circ; i < k; ++i) {
            float* r = w->small.r[i - k_circ];
            r[0] = t->box[i].cx;
            r[1] = t->box[i].cy;
            r[2] = t->box[i].a;
            r[3] = t->box[i].b;
        }
        w->small_mode = (k - k_circ) > 2 ? 2 : 1;
    }
    delete t;
}

void copy_table_cells16(int precision, const void* tab_host, unsigned short* out) {
    const unsigned short* cell = precision == AQUA_F64 ? reinterpret_cast<const ObstacleTable<double>*>(tab_host)->cell16
                                                        : reinterpret_cast<const ObstacleTable<float>*>(tab_host)->cell16;
    memcpy(out, cell, sizeof(unsigned short) * kGridCells);
}

void copy_table_cells(int precision, const void* tab_host, unsigned long long* out) {
    static_assert(kGridDim == AQUA_GRID_DIM && kCellBlocked == AQUA_CELL_BLOCKED, "header and kernels agree on the grid");
    const unsigned long long* cell = precision == AQUA_F64 ? reinterpret_cast<const ObstacleTable<double>*>(tab_host)->cell
                                                            : reinterpret_cast<const ObstacleTable<float>*>(tab_host)->cell;
    memcpy(out, cell, sizeof(unsigned long long) * kGridCells);
}

}  // namespace aqua
