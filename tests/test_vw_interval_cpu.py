"""Own-cell row intervals of the DecentralizedCVT window kernels (csrc/vwindow.cuh vw_row_interval, used by
k_voronoi_window and k_cnn_fused8<VOR>; replaces reference voronoi.cpp:55-150 as called from
decentralized_cvt.h:85-143) on the HOST, through ccb_selftest_vw_row_interval -- the kernels' own code compiled
for the host.  Stated rule: a window cell belongs to the robot unless some neighbour is STRICTLY nearer
(d2 = dx*dx + dy*dy in fp64 without FMA contraction; ties go to the robot).  For a row the kernel does not mark
"exact" the interval must be that set bit for bit; on an "exact" row the kernel re-tests every cell of the
interval with the rule, so the interval only has to contain the set.  No GPU, no oracle."""
import ctypes as C

import numpy as np
import pytest


def _window(p, L, N, res):
    """map_utils.h:51-87 reduced to the clipped window, as common.cuh make_window"""
    def rnd(v):  # std::round, half away from zero
        return int(np.floor(abs(v) + 0.5) * np.sign(v))
    cx, cy = rnd(p[0] / res), rnd(p[1] / res)
    ox, oy = cx - L // 2, cy - L // 2
    x0, y0 = max(ox, 0), max(oy, 0)
    x1, y1 = min(cx + L // 2, N), min(cy + L // 2, N)
    return ox, oy, x0, max(x1, x0), y0, max(y1, y0)


def _intervals(res, x0, x1, y0, ncols, sites, wbase=0, want_words=False):
    from coveragecontrol_b200 import _lib

    lib = _lib.load()
    px = np.ascontiguousarray(sites[:, 0], np.float64)
    py = np.ascontiguousarray(sites[:, 1], np.float64)
    n = x1 - x0
    lo, hi, ex = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.int32)
    words = np.zeros(9, np.uint32)
    ip = C.POINTER(C.c_int)
    rc = lib.ccb_selftest_vw_row_interval(float(res), x0, x1, y0, ncols, len(sites) - 1,
                                          px.ctypes.data_as(C.POINTER(C.c_double)), py.ctypes.data_as(C.POINTER(C.c_double)),
                                          lo.ctypes.data_as(ip), hi.ctypes.data_as(ip), ex.ctypes.data_as(ip), wbase,
                                          words.ctypes.data_as(C.POINTER(C.c_uint32)) if want_words else None)
    assert rc == 0
    return lo, hi, ex, words


def _rule_own(res, rows, cols, sites):
    cx = (rows.astype(np.float64) + 0.5) * res
    cy = (cols.astype(np.float64) + 0.5) * res
    dx = cx[:, None, None] - sites[None, None, :, 0]
    dy = cy[None, :, None] - sites[None, None, :, 1]
    d2 = dx * dx + dy * dy
    return ~np.any(d2[:, :, 1:] < d2[:, :, :1], axis=2)  # [rows, cols]


def _check(sites, L, N, res=1.0):
    sites = np.asarray(sites, np.float64)
    ox, oy, x0, x1, y0, y1 = _window(sites[0], L, N, res)
    ncols = y1 - y0
    if x1 <= x0 or ncols <= 0:
        return 0, 0
    lo, hi, ex, _ = _intervals(res, x0, x1, y0, ncols, sites)
    rows, cols = np.arange(x0, x1), np.arange(y0, y1)
    own = _rule_own(res, rows, cols, sites)
    inside = (cols[None, :] >= y0 + lo[:, None]) & (cols[None, :] < y0 + hi[:, None])
    plain = ex == 0
    bad = np.argwhere(inside[plain] != own[plain])
    assert len(bad) == 0, f"{len(bad)} cells of unflagged rows differ from the rule, first: row {rows[plain][bad[0][0]]} col {cols[bad[0][1]]}"
    assert not np.any(own[~plain] & ~inside[~plain]), "an 'exact' row's interval misses a cell of the robot"
    return int((~plain).sum()), len(rows)


@pytest.mark.parametrize("N,L,R,seed", [(1024, 256, 32, 0), (512, 256, 12, 1), (512, 128, 8, 2), (256, 64, 20, 3), (300, 100, 5, 4)])
def test_random_neighbourhoods(N, L, R, seed):
    rng = np.random.default_rng(seed)
    pos = rng.uniform(0, N, (R, 2))
    for r in range(R):
        d = np.hypot(*(pos - pos[r]).T)
        nb = [j for j in range(R) if j != r and d[j] < 0.4 * N]
        _check(np.vstack([pos[r:r + 1], pos[nb]]), L, N)


def test_no_neighbours_owns_the_whole_window():
    lo, hi, ex, _ = _intervals(1.0, 10, 20, 5, 77, np.array([[15.0, 40.0]]))
    assert np.all(lo == 0) and np.all(hi == 77) and np.all(ex == 0)


@pytest.mark.parametrize("res", [0.5, 2.0, 0.25])
def test_resolution(res):
    rng = np.random.default_rng(7)
    N, L = 400, 128
    for _ in range(10):
        _check(rng.uniform(0, N * res, (9, 2)), L, N, res)


def test_window_clipped_by_the_map_border():
    rng = np.random.default_rng(8)
    N, L = 256, 256
    for corner in ([1e-4, 1e-4], [N - 1e-4, 3.0], [5.0, N - 1e-4], [N - 1e-4, N - 1e-4], [128.0, 0.3]):
        sites = np.vstack([[corner], rng.uniform(0, N, (10, 2))])
        _check(sites, L, N)


def test_quantised_positions_ties_and_aligned_bisectors():
    """neighbours at the same row / column as the robot, positions on cell centres / corners / quarter cells:
    bisectors run along rows and through cell centres -- the kernel must flag or decide like the rule"""
    rng = np.random.default_rng(9)
    N, L = 512, 256
    nex = 0
    for q in (1.0, 0.5, 0.25):
        for _ in range(12):
            sites = np.round(rng.uniform(100, 400, (10, 2)) / q) * q + (0.5 if q == 1.0 and rng.random() < 0.5 else 0.0)
            sites[1] = sites[0] + [0.0, 16.0]   # bisector = a line of constant column
            sites[2] = sites[0] + [16.0, 0.0]   # bisector runs along a row (through its cell centres when on .5)
            sites[3] = sites[0] + [7.0, 0.0]
            sites[4] = sites[0] + [12.0, 12.0]
            e, _ = _check(sites, L, N)
            nex += e
    assert nex > 0  # the aligned cases do reach the per-cell path


def test_twins_and_near_twins_of_the_robot():
    rng = np.random.default_rng(10)
    N, L = 512, 256
    for eps in (0.0, 1e-15, 1e-13, 1e-11, 1e-9, 1e-6):
        sites = rng.uniform(150, 350, (8, 2))
        sites[1] = sites[0] + [eps, -eps]
        sites[2] = sites[0] + [0.0, eps]
        sites[3] = sites[0] + [eps, 0.0]
        _check(sites, L, N)


def test_bisectors_through_cell_centres():
    rng = np.random.default_rng(11)
    N, L = 512, 256
    for _ in range(150):
        K = int(rng.integers(1, 12))
        sites = rng.uniform(100, 400, (K + 1, 2))
        c = np.floor(sites[0] + rng.uniform(-100, 100, 2)) + 0.5
        th = rng.uniform(0, np.pi)
        nrm = np.array([np.cos(th), np.sin(th)])
        sites[1] = sites[0] - 2 * np.dot(sites[0] - c, nrm) * nrm + rng.normal(0, 10.0 ** (-int(rng.integers(6, 17))), 2)
        _check(sites, L, N)


def test_far_neighbours_large_coordinates():
    rng = np.random.default_rng(12)
    N, L = 16384, 256
    for _ in range(10):
        p = rng.uniform(1000, 15000, 2)
        sites = np.vstack([[p], p + rng.uniform(-128, 128, (15, 2))])
        _check(sites, L, N)


def test_interval_mask_words():
    """vw_range_bits: bits [y0 + lo, y0 + hi) in the 9 mask words starting at column word wbase"""
    sites = np.array([[200.3, 300.7], [230.0, 280.0], [170.0, 330.5]])
    ox, oy, x0, x1, y0, y1 = _window(sites[0], 256, 1024, 1.0)
    wbase = oy >> 5
    for row in (x0, x0 + 100, x1 - 1):
        lo, hi, ex, words = _intervals(1.0, row, row + 1, y0, y1 - y0, sites, wbase, True)
        bits = np.zeros(9 * 32, bool)
        a, b = y0 + lo[0] - 32 * wbase, y0 + hi[0] - 32 * wbase
        bits[max(a, 0):max(b, 0)] = True
        got = ((words[:, None] >> np.arange(32, dtype=np.uint32)[None, :]) & 1).astype(bool).reshape(-1)
        assert np.array_equal(got, bits)


def test_make_window_matches_the_oracle_offsets():
    """common.cuh make_window (host run) against the oracle's ComputeOffsets (map_utils.h:51-87): half-way
    positions round away from zero, windows clipped on every side, windows that miss the map"""
    from coveragecontrol_b200 import _lib
    from oracle import oracle as orc

    lib = _lib.load()
    rng = np.random.default_rng(3)
    out = np.zeros(6, np.int32)
    cases = [(1.0, 10.5, 11.5, 64, 1024), (1.0, 0.5, 1023.5, 64, 1024), (0.5, 3.25, 3.75, 8, 64), (1.0, 1e-4, 1023.9999, 256, 1024),
             (2.0, 31.0, 33.0, 16, 32), (1.0, 2.5, 3.5, 7, 50)]
    for _ in range(2000):
        res = float(rng.choice([1.0, 0.5, 2.0, 0.3]))
        n = int(rng.integers(8, 300))
        sub = int(rng.integers(1, 2 * n))
        p = rng.uniform(-0.2 * n * res, 1.2 * n * res, 2)
        if rng.random() < 0.3:
            p = (np.round(p / res) + 0.5) * res
        cases.append((res, float(p[0]), float(p[1]), sub, n))
    for res, px, py, sub, n in cases:
        assert lib.ccb_selftest_make_window(res, px, py, sub, n, out.ctypes.data_as(C.POINTER(C.c_int))) == 0
        ix, of = orc.compute_offsets(res, px, py, sub, n)
        assert (out[0], out[1]) == (ix.left, ix.bottom)
        x0, x1 = ix.left + of.left, ix.right + of.right
        y0, y1 = ix.bottom + of.bottom, ix.top + of.top
        # a window that misses the map has a negative extent in the reference and an empty one here
        assert (out[2], out[4]) == (x0, y0)
        assert out[3] == max(x1, x0) and out[5] == max(y1, y0)
