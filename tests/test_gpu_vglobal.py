"""Site-ordered global Voronoi (csrc/vglobal.cuh: row partition -> interval streaming -> exact rule on
flagged rows) against the brute-force oracle and against the tile-ordered kernel (csrc/voronoi.cuh),
on random, clustered and lattice site sets (the latter put bisectors through cell centres and along
rows/columns: every such row must be flagged and resolved by the stated rule)."""
import os

import numpy as np
import pytest

from conftest import random_world

pytestmark = pytest.mark.gpu


def _sites(kind, N, R, rng):
    if kind == "random":
        return rng.uniform(1e-4, N - 1e-4, (R, 2))
    if kind == "cluster":  # many near-degenerate vertices, tiny py differences
        return np.concatenate([rng.uniform(1, N - 1, (R // 2, 2)), rng.normal(N / 2, 3.0, (R - R // 2, 2)).clip(1, N - 1)])
    if kind == "lattice":  # integer lattice: bisectors on cell centres, along rows and columns, 4-way vertices
        g = int(np.ceil(np.sqrt(R)))
        xs = (np.arange(g) + 0.5) * (N // g)
        pts = np.array([[x, y] for x in xs for y in xs], np.float64)[:R]
        return pts
    if kind == "halfint":  # sites on cell centres with duplicates
        pts = np.floor(rng.uniform(2, N - 2, (R, 2))) + 0.5
        pts[1] = pts[0]
        pts[3, 1] = pts[2, 1]  # equal py, different px
        pts[5, 0] = pts[4, 0]  # equal px, different py
        return pts
    raise ValueError(kind)


@pytest.mark.parametrize("generic", [0, 1])
@pytest.mark.parametrize("kind,N,R,seed", [("random", 1024, 32, 0), ("random", 512, 64, 1), ("cluster", 256, 48, 2),
                                           ("lattice", 256, 36, 3), ("halfint", 512, 24, 4), ("random", 300, 2, 5),
                                           ("cluster", 1024, 64, 6)])
def test_labels_bit_exact_and_moments(monkeypatch, generic, kind, N, R, seed):
    import coverage_control as cc
    from oracle import oracle as orc

    monkeypatch.setenv("CCB_VG_GENERIC", str(generic))
    rng = np.random.default_rng(seed)
    pos = _sites(kind, N, R, rng)
    p = cc.Parameters()
    p.pWorldMapSize = p.pRobotMapSize = N
    p.pNumRobots = R
    wm = rng.random((N, N)).astype(np.float32)
    env = cc.CoverageSystem(p, cc.WorldIDF(p, wm), cc.PointVector(pos))
    labels = env.voronoi_labels(0)[0]
    ref_labels, ref_sums = orc.voronoi_bruteforce(pos, wm)
    bad = np.argwhere(labels != ref_labels)
    assert len(bad) == 0, f"{len(bad)} labels differ, first at {bad[:5].tolist()}"
    cells = env.GetVoronoiCells()
    assert np.allclose([c.mass for c in cells], ref_sums[:, 0], rtol=1e-9, atol=1e-9)
    assert env.GetObjectiveValue() == pytest.approx(ref_sums[:, 3].sum(), rel=1e-9)


def test_explored_source_matches_tile_kernel(monkeypatch):
    """CentralizedCVT (explored-IDF map): both kernels, same goals, step after step."""
    import coverage_control as cc
    from coveragecontrol_b200 import BatchedCoverageSystem, BatchedCVT

    N, R, E = 512, 20, 3
    p = cc.Parameters()
    p.pWorldMapSize = p.pRobotMapSize = N
    p.pNumRobots = R
    p.pSensorSize = 48
    rng = np.random.default_rng(9)
    worlds = [rng.random((N, N)).astype(np.float32) for _ in range(E)]
    pos = rng.uniform(1, N - 1, (E, R, 2))
    runs = []
    for generic in (0, 1):
        monkeypatch.setenv("CCB_VG_GENERIC", str(generic))
        env = BatchedCoverageSystem(p, E, R)
        for e in range(E):
            env.set_world(e, worlds[e])
        env.reset(pos)
        ctrl = BatchedCVT(env, 1)
        trace = []
        for _ in range(6):
            ctrl.compute_actions()
            trace.append(ctrl.goals().copy())
            env.step_actions(ctrl.actions())
        trace.append(env.voronoi_labels(1))
        runs.append(trace)
    for a, b in zip(runs[0][:-1], runs[1][:-1]):
        assert np.allclose(a, b, rtol=1e-11, atol=1e-9)
    la, lb = runs[0][-1], runs[1][-1]
    assert (la >= 0).all()  # the site-ordered kernel labels every cell
    seen = lb >= 0          # the tile-ordered kernel skips regions without an explored cell
    assert seen.any() and np.array_equal(la[seen], lb[seen])


def test_cell_polygon_with_more_vertices_than_a_warp():
    """40 sites on a circle around a central one: its cell has 40 edges, more than the 32 lanes of the
    warp-parallel clipping (csrc/maps.cuh clip_cell_warp) -> the sequential fallback; same vertex set as the
    oracle's half-plane clipping, and every ring cell (a wedge cut by the map rectangle) as well."""
    import coverage_control as cc
    from oracle import oracle as orc

    N, K = 512, 40
    ang = 2 * np.pi * (np.arange(K) + 0.37) / K
    pos = np.concatenate([[[256.3, 255.6]], np.stack([256.3 + 120 * np.cos(ang), 255.6 + 120 * np.sin(ang)], 1)])
    p = cc.Parameters()
    p.pWorldMapSize = p.pRobotMapSize = N
    p.pNumRobots = len(pos)
    wm = np.random.default_rng(3).random((N, N)).astype(np.float32)
    env = cc.CoverageSystem(p, cc.WorldIDF(p, wm), cc.PointVector(pos))
    cells = env.GetVoronoiCells()
    ocells = orc.voronoi(pos, wm)
    assert len(np.array(cells[0].cell)) == K
    for c, o in zip(cells, ocells):
        a = np.array(sorted(map(tuple, np.round(np.array(c.cell), 6))))
        b = np.array(sorted(map(tuple, np.round(o.polygon(), 6))))
        assert a.shape == b.shape and np.allclose(a, b, atol=1e-5)
        assert c.mass == pytest.approx(o.mass, rel=1e-9)


_MODES = {"fused": ("1", "1", "0"), "lookup": ("1", "0", "0"), "stream": ("0", "0", "0"),
          "bulk": ("0", "0", "1")}  # CCB_VG_PREFIX, CCB_VG_FUSED, CCB_VG_BULK


@pytest.mark.parametrize("mode", list(_MODES))
@pytest.mark.parametrize("kind,N,R,seed", [("random", 1024, 32, 0), ("random", 512, 64, 1), ("cluster", 256, 48, 2),
                                           ("lattice", 256, 36, 3), ("halfint", 512, 24, 4), ("random", 304, 2, 5)])
def test_block_prefix_path_matches_oracle(monkeypatch, mode, kind, N, R, seed):
    """The three ways to the site-ordered world-map moments (vglobal.cuh): "fused" (the default: block prefix sums in
    double-double walked inside the row partition kernel), "lookup" (interval table + site-ordered k_vg_lookup) and
    "stream" (k_vg_reduce reads every cell) and "bulk" (the same with cp.async.bulk + mbarrier staging, the TMA A/B).  Moments within 1e-9 of the oracle (the gate of SURVEY 8d is 1e-4); a
    map with zero stretches (left half, and a band in the middle of every row) keeps the mass of the sites on them at
    exactly 0 (vertex-mean fallback, voronoi.cpp:143-150); setting a new world map refreshes the prefixes."""
    import coverage_control as cc
    from oracle import oracle as orc

    monkeypatch.setenv("CCB_VG_GENERIC", "0")
    monkeypatch.setenv("CCB_VG_PREFIX", _MODES[mode][0])
    monkeypatch.setenv("CCB_VG_FUSED", _MODES[mode][1])
    monkeypatch.setenv("CCB_VG_BULK", _MODES[mode][2])
    rng = np.random.default_rng(seed)
    pos = _sites(kind, N, R, rng)
    p = cc.Parameters()
    p.pWorldMapSize = p.pRobotMapSize = N
    p.pNumRobots = R
    env = None
    for variant in range(3):
        wm = rng.random((N, N)).astype(np.float32)
        if variant == 1:
            wm[:, : N // 2] = 0.0
            wm[:, 5 * N // 8: 7 * N // 8] = 0.0  # zeros AFTER non-zero cells: the prefix entries must not move
        if variant < 2:
            env = cc.CoverageSystem(p, cc.WorldIDF(p, wm), cc.PointVector(pos))
        else:  # a new map in the SAME environment: the stale prefixes must not survive
            env.GetWorldMapMutable()[...] = wm
            env.StepActions(np.zeros((R, 2)))
        _, ref_sums = orc.voronoi_bruteforce(pos, wm)
        cells = env.GetVoronoiCells()
        mass = np.array([c.mass for c in cells])
        assert np.allclose(mass, ref_sums[:, 0], rtol=1e-9, atol=1e-9)
        assert np.all(mass[ref_sums[:, 0] == 0.0] == 0.0)
        assert env.GetObjectiveValue() == pytest.approx(ref_sums[:, 3].sum(), rel=1e-9)
        cen = np.array([list(c.centroid) for c in cells])
        ok = ref_sums[:, 0] >= 1e-10
        assert np.allclose(cen[ok], ref_sums[ok, 1:3] / ref_sums[ok, :1], rtol=1e-9, atol=1e-7)
        if kind in ("random", "cluster"):
            # the reference-faithful scan-line cells as well (vertex-mean fallback included); on the lattice / half-integer
            # sets bisectors run through cell centres, where the scan-line rule (voronoi.cpp:137-141) and the stated
            # lowest-index rule assign the tie cells differently
            ocells = orc.Env(orc.params_from(p), wm, pos).voronoi_cells()
            for c, oc in zip(cells, ocells):
                assert np.allclose(np.array(c.centroid), np.array(oc.centroid[:]), rtol=1e-9, atol=1e-7)


@pytest.mark.parametrize("mode", list(_MODES))
def test_low_mass_cell_behind_a_heavy_row_prefix(monkeypatch, mode):
    """A cell whose whole mass is ~1e-9 / ~1e-11 at the END of rows whose running sums are ~1e3:
    plain fp64 prefix differences would lose it (ulp(1e3) ~ 1e-13) and could flip the `mass < 1e-10` vertex-mean
    fallback of voronoi.cpp:41-53; the double-double entries + direct head / tail sums must not.  The site owns the
    last columns of the map only: heavy cells up to column N - 24, then a few tiny ones."""
    import coverage_control as cc
    from oracle import oracle as orc

    monkeypatch.setenv("CCB_VG_GENERIC", "0")
    monkeypatch.setenv("CCB_VG_PREFIX", _MODES[mode][0])
    monkeypatch.setenv("CCB_VG_FUSED", _MODES[mode][1])
    monkeypatch.setenv("CCB_VG_BULK", _MODES[mode][2])
    N = 1024
    p = cc.Parameters()
    p.pWorldMapSize = p.pRobotMapSize = N
    rng = np.random.default_rng(12)
    # site 0 at the right edge, walled off by a line of sites 40 columns before it: its cell stays in the columns
    # >= N - 28; the other sites share the heavy part of the map
    wall = np.stack([np.arange(384.0, 641.0, 32.0) + 0.3, np.full(9, N - 44.2)], 1)
    pos = np.concatenate([[[512.3, N - 4.2]], wall, [[500.2, 300.6], [130.9, 700.2], [880.5, 520.1], [60.2, 980.4], [990.1, 1000.3]]])
    p.pNumRobots = len(pos)
    for tiny, expect_fallback in ((1e-12, False), (2e-14, True)):
        wm = (rng.random((N, N)).astype(np.float32) * 2.0 + 1.0)  # running row sums ~2e3
        wm[:, N - 40:] = 0.0
        wm[:, N - 13:N - 3] = np.float32(tiny) * (1.0 + rng.random((N, 10)).astype(np.float32))
        env = cc.CoverageSystem(p, cc.WorldIDF(p, wm), cc.PointVector(pos))
        labels, ref_sums = orc.voronoi_bruteforce(pos, wm)
        assert (labels[:, : N - 40] != 0).all()  # the cell of site 0 sees none of the heavy cells
        assert 0 < ref_sums[0, 0] < 1e-8 and (ref_sums[0, 0] < 1e-10) == expect_fallback
        cells = env.GetVoronoiCells()
        ocells = orc.voronoi(pos, wm)
        for k, (c, oc) in enumerate(zip(cells, ocells)):
            assert c.mass == pytest.approx(ref_sums[k, 0], rel=1e-9), k
            # north-star gate (1e-4 on centroids), far inside it here; the fallback cell takes the vertex mean
            assert np.allclose(np.array(c.centroid), np.array(oc.centroid[:]), rtol=1e-9, atol=1e-7), k
        assert env.GetObjectiveValue() == pytest.approx(ref_sums[:, 3].sum(), rel=1e-9)
