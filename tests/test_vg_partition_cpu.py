"""Row partition of the whole-map Voronoi kernel (csrc/vglobal.cuh k_vg_partition, replacing reference
voronoi.cpp:55-165 for the global cells) on the HOST: ccb_selftest_vg_partition runs the kernel's own row
sweep (vg_partition_row, compiled for the host) and this file holds every row it did NOT flag against the
stated rule -- label = the site with the smallest d2 = (cx - px)^2 + (cy - py)^2 in fp64 without FMA
contraction, strict <, ascending site index (SURVEY 8d "parity gates") -- bit for bit.  Flagged rows are
the ones the kernel hands to the per-cell rule itself (k_vg_fixup), so an unflagged row that differs
would be a wrong label on the GPU.  No GPU, no oracle needed: the rule is three NumPy lines."""
import ctypes as C

import numpy as np
import pytest


BAND_ROWS = 0  # rows per CTA band of the emulated kernel (0: the default, 128; 256 / 512 = CCB_VG_PART_RPT 2 / 4)


def _partition(pos, N, res=1.0, row0=0, row1=None):
    from coveragecontrol_b200 import _lib

    lib = _lib.load()
    pos = np.ascontiguousarray(pos, np.float64)
    R = len(pos)
    row1 = N if row1 is None else row1
    table = np.full((R, N), 0xFFFFFFFF, np.uint32)
    flags = np.full(N, -1, np.int32)
    span = np.zeros((R, 2), np.int32)
    n = lib.ccb_selftest_vg_partition(pos.ctypes.data_as(C.POINTER(C.c_double)), R, N, float(res), row0, row1, BAND_ROWS,
                                      table.ctypes.data_as(C.POINTER(C.c_uint32)),
                                      flags.ctypes.data_as(C.POINTER(C.c_int)), span.ctypes.data_as(C.POINTER(C.c_int)))
    assert n >= 0
    return table, np.sort(flags[:n]), span


def _rule_labels(pos, N, res, rows):
    """the stated rule for the given rows: [len(rows), N] labels"""
    c = (np.arange(N, dtype=np.float64) + 0.5) * res
    cx = c[rows]
    dx = cx[:, None, None] - pos[None, None, :, 0]
    dy = c[None, :, None] - pos[None, None, :, 1]
    d2 = dx * dx + dy * dy  # two roundings per product-sum, no FMA in NumPy
    return np.argmin(d2, axis=2)  # first minimum = lowest index on ties


def _check(pos, N, res=1.0, row0=0, row1=None, max_flag_frac=None):
    pos = np.asarray(pos, np.float64)
    R = len(pos)
    row1 = N if row1 is None else row1
    table, flags, span = _partition(pos, N, res, row0, row1)
    rows = np.setdiff1d(np.arange(row0, row1), flags)
    lo = (table & 0xFFFF).astype(np.int64)
    hi = (table >> 16).astype(np.int64)
    if len(rows):
        lab = np.full((len(rows), N), -1, np.int64)
        cover = np.zeros((len(rows), N), np.int64)
        cols = np.arange(N)
        for k in range(R):
            inside = (cols[None, :] >= lo[k, rows, None]) & (cols[None, :] < hi[k, rows, None])
            lab[inside] = k
            cover += inside
        assert np.all(cover == 1), "intervals of an unflagged row do not tile the row"
        ref = _rule_labels(pos, N, res, rows)
        bad = np.argwhere(lab != ref)
        assert len(bad) == 0, f"{len(bad)} cells differ from the stated rule, first at row {rows[bad[0][0]]} col {bad[0][1]}"
        # row spans: exactly the rows (of the unflagged ones) on which the site owns a cell lie inside
        for k in range(R):
            own = rows[(hi[k, rows] > lo[k, rows])]
            if len(own):
                assert span[k, 0] <= own.min() and span[k, 1] >= own.max() + 1
    # a flagged row leaves an empty table: k_vg_fixup labels it
    if len(flags):
        assert np.all(table[:, flags] == 0)
    if max_flag_frac is not None:
        assert len(flags) <= max_flag_frac * (row1 - row0), f"{len(flags)} of {row1 - row0} rows flagged"
    return len(flags)


@pytest.mark.parametrize("N,R,seed", [(1024, 32, 0), (1024, 64, 1), (512, 3, 2), (256, 1, 3), (300, 17, 4), (2048, 48, 5),
                                      (4, 2, 6), (64, 64, 7)])
def test_random_sites(N, R, seed):
    rng = np.random.default_rng(seed)
    _check(rng.uniform(0, N, (R, 2)), N, max_flag_frac=0.02 if N >= 256 else None)


@pytest.mark.parametrize("res", [0.5, 2.0, 0.1, 3.7])
def test_resolution(res):
    rng = np.random.default_rng(11)
    N = 384
    _check(rng.uniform(0, N * res, (24, 2)), N, res=res, max_flag_frac=0.05)


def test_row_bands():
    rng = np.random.default_rng(12)
    N = 640
    pos = rng.uniform(0, N, (32, 2))
    whole, wf, _ = _partition(pos, N)
    for r0, r1 in ((0, 200), (200, 333), (333, 640), (77, 78)):
        _check(pos, N, row0=r0, row1=r1)
        band, bf, _ = _partition(pos, N, row0=r0, row1=r1)
        # the band cull depends on the band, the intervals must not
        rows = np.setdiff1d(np.setdiff1d(np.arange(r0, r1), wf), bf)
        lo_w, hi_w = whole[:, rows] & 0xFFFF, whole[:, rows] >> 16
        lo_b, hi_b = band[:, rows] & 0xFFFF, band[:, rows] >> 16
        ne = (hi_w > lo_w) | (hi_b > lo_b)  # empty intervals may sit anywhere
        assert np.array_equal(lo_w[ne], lo_b[ne]) and np.array_equal(hi_w[ne], hi_b[ne])


def test_sites_on_cell_centres_and_lattices():
    """ties everywhere: sites on cell centres / corners of a regular lattice put bisectors exactly on cell
    centres -- such rows must be flagged or agree with the lowest-index rule"""
    N = 128
    g = np.arange(8, N, 16, dtype=np.float64)
    for off in (0.0, 0.5):
        pos = np.stack(np.meshgrid(g + off, g + off, indexing="ij"), -1).reshape(-1, 2)
        _check(pos, N)
        _check(pos[::-1].copy(), N)
        rng = np.random.default_rng(5)
        _check(rng.permutation(pos), N)


def test_collinear_sites():
    N = 256
    t = np.linspace(10, 240, 20)
    _check(np.stack([t, np.full_like(t, 100.25)], 1), N)      # same y: one row of sites across the rows
    _check(np.stack([np.full_like(t, 100.25), t], 1), N)      # same x: all bisectors horizontal in y
    _check(np.stack([t, t], 1), N)                            # diagonal
    _check(np.stack([t, 0.5 * t + 3.0], 1), N)


def test_twins_and_near_twins():
    rng = np.random.default_rng(21)
    N = 256
    pos = rng.uniform(0, N, (16, 2))
    pos[5] = pos[2]                     # exact twins: the lower index owns the cell
    pos[9] = pos[2]
    pos[11] = pos[7] + [1e-13, -1e-13]  # inseparable within rounding
    pos[12] = pos[3] + [0.0, 1e-9]
    pos[13] = pos[4] + [1e-9, 0.0]
    _check(pos, N)
    table, flags, _ = _partition(pos, N)
    rows = np.setdiff1d(np.arange(N), flags)
    for twin in (5, 9):
        assert np.all((table[twin, rows] >> 16) <= (table[twin, rows] & 0xFFFF)), "a twin with a higher index owns cells"


def test_sites_outside_and_on_the_border():
    N = 200
    pos = np.array([[-50.0, 100.0], [250.0, 100.0], [100.0, -30.0], [100.0, 260.0], [0.0, 0.0], [200.0, 200.0],
                    [0.0, 200.0], [1e-4, 100.0], [200.0 - 1e-4, 17.0], [100.0, 100.0]])
    _check(pos, N)


def test_clustered_sites():
    rng = np.random.default_rng(31)
    N = 1024
    centres = rng.uniform(100, 900, (4, 2))
    pos = (centres[rng.integers(0, 4, 64)] + rng.normal(0, 3.0, (64, 2))).clip(1e-4, N - 1e-4)
    _check(pos, N)
    # an almost converged CVT: many robots a hair apart
    pos = np.full((32, 2), 512.0) + rng.normal(0, 1e-3, (32, 2))
    _check(pos, N)


def test_positions_after_cvt_steps_are_quantised():
    """robot positions in a roll-out are sums of speed-limited steps: many share coordinates to the last bit"""
    rng = np.random.default_rng(41)
    N = 512
    pos = np.round(rng.uniform(0, N, (40, 2)) * 4) / 4   # quarter-cell lattice: frequent exact ties
    _check(pos, N)
    pos = np.round(rng.uniform(0, N, (40, 2)))           # integer coordinates = cell corners
    _check(pos, N)
    pos = np.floor(rng.uniform(0, N, (40, 2))) + 0.5     # exactly on cell centres
    _check(pos, N)


def test_many_seeds_small_maps():
    total = 0
    for seed in range(60):
        rng = np.random.default_rng(1000 + seed)
        N = int(rng.integers(1, 40)) * 4
        R = int(rng.integers(1, 65))
        total += _check(rng.uniform(-0.1 * N, 1.1 * N, (R, 2)), N)
    assert total >= 0


def test_bisectors_and_vertices_through_cell_centres():
    """a second site mirrored across a line through a cell centre (perturbed by 1e-6 .. 1e-16) and three sites on
    a circle around a cell centre: comparisons inside the rounding error must flag, never guess"""
    rng = np.random.default_rng(77)
    for _ in range(120):
        N = int(rng.integers(8, 80)) * 4
        R = int(rng.integers(2, 20))
        pos = rng.uniform(0, N, (R, 2))
        c = np.floor(rng.uniform(0, N, 2)) + 0.5
        th = rng.uniform(0, np.pi)
        nrm = np.array([np.cos(th), np.sin(th)])
        pos[1] = pos[0] - 2 * np.dot(pos[0] - c, nrm) * nrm + rng.normal(0, 10.0 ** (-int(rng.integers(6, 17))), 2)
        c2, rad = np.floor(rng.uniform(0, N, 2)) + 0.5, rng.uniform(3, 30)
        for k in range(2, min(5, R)):
            a = rng.uniform(0, 2 * np.pi)
            pos[k] = c2 + rad * np.array([np.cos(a), np.sin(a)])
        _check(pos, N)


def test_rejects_bad_arguments():
    from coveragecontrol_b200 import _lib

    lib = _lib.load()
    pos = np.zeros((65, 2))
    t = np.zeros((65, 8), np.uint32)
    f = np.zeros(8, np.int32)
    s = np.zeros((65, 2), np.int32)
    args = (t.ctypes.data_as(C.POINTER(C.c_uint32)), f.ctypes.data_as(C.POINTER(C.c_int)), s.ctypes.data_as(C.POINTER(C.c_int)))
    dp = pos.ctypes.data_as(C.POINTER(C.c_double))
    assert lib.ccb_selftest_vg_partition(dp, 65, 8, 1.0, 0, 8, 0, *args) == -1   # more sites than VG_MAXR
    assert lib.ccb_selftest_vg_partition(dp, 4, 8, 1.0, 0, 9, 0, *args) == -1    # band beyond the map
    assert lib.ccb_selftest_vg_partition(dp, 4, 8, 0.0, 0, 8, 0, *args) == -1    # resolution


@pytest.mark.parametrize("band_rows", [256, 512])
def test_taller_cta_bands(monkeypatch, band_rows):
    """CCB_VG_PART_RPT = 2 / 4 (opt-in): a CTA's band is 256 / 512 rows, so the band cull keeps more sites; the
    intervals of unflagged rows must not change"""
    import sys

    monkeypatch.setattr(sys.modules[__name__], "BAND_ROWS", band_rows)
    rng = np.random.default_rng(50 + band_rows)
    for N, R in ((1024, 32), (640, 64), (300, 9)):
        pos = rng.uniform(0, N, (R, 2))
        _check(pos, N)
        _check(pos, N, row0=100, row1=N - 7)
    g = np.arange(8, 128, 16, dtype=np.float64) + 0.5
    _check(np.stack(np.meshgrid(g, g, indexing="ij"), -1).reshape(-1, 2), 128)


def test_row_prefix_moments_from_the_partition_table():
    """what k_vg_lookup (opt-in CCB_VG_PREFIX=1) computes, restated in NumPy on the host partition table: per
    (site, row) the difference of inclusive fp64 row prefix sums of (v, y v, y^2 v).  Against the direct per-cell
    sums of the stated rule's labels on the unflagged rows: mass, centroid and sum |c - p|^2 v agree to ~1e-12
    (the gate of SURVEY 8d is 1e-4), and sites on an all-zero half of the map get exactly 0."""
    rng = np.random.default_rng(60)
    N, R = 512, 24
    pos = rng.uniform(0, N, (R, 2))
    v = rng.random((N, N)).astype(np.float32).astype(np.float64)
    v[:, : N // 3] = 0.0
    table, flags, _ = _partition(pos, N)
    rows = np.setdiff1d(np.arange(N), flags)
    lo, hi = (table & 0xFFFF).astype(np.int64), (table >> 16).astype(np.int64)
    y = np.arange(N) + 0.5
    x = np.arange(N) + 0.5
    zero = np.zeros((N, 1))
    P = [np.concatenate([zero, np.cumsum(v * w[None, :], axis=1)], axis=1) for w in (np.ones(N), y, y * y)]  # P[c] = sum of cols < c
    lab = _rule_labels(pos, N, 1.0, rows)
    for k in range(R):
        rs, rsy, rsyy = (p[rows, hi[k, rows]] - p[rows, lo[k, rows]] for p in P)
        got = np.array([rs.sum(), (x[rows] * rs).sum(), rsy.sum(), (x[rows] ** 2 * rs + rsyy).sum()])
        m = lab == k
        vv = v[rows]
        ref = np.array([vv[m].sum(), (vv * x[rows, None])[m].sum(), (vv * y[None, :])[m].sum(),
                        (vv * (x[rows, None] ** 2 + y[None, :] ** 2))[m].sum()])
        if ref[0] == 0.0:
            assert np.all(got == 0.0)
            continue
        assert abs(got[0] - ref[0]) <= 1e-12 * ref[0]
        assert abs(got[1] / got[0] - ref[1] / ref[0]) <= 1e-9 and abs(got[2] / got[0] - ref[2] / ref[0]) <= 1e-9
        px, py = pos[k]
        obj = lambda s: s[3] - 2 * (px * s[1] + py * s[2]) + (px * px + py * py) * s[0]  # noqa: E731
        assert abs(obj(got) - obj(ref)) <= 1e-9 * abs(obj(ref))
