"""CPU cross-check of the oracle's Voronoi pieces against NumPy: the brute-force labeller that every GPU label
test is held against (oracle/cc_oracle.c orc_voronoi_bruteforce: the stated rule of SURVEY 8d) and the
polygon-based restatement of reference voronoi.cpp:55-165 (orc_voronoi: convex cell by half-plane clipping,
cell centres inside it, mass / centroid / vertex-mean fallback)."""
import numpy as np
import pytest

from oracle import oracle as orc


def _numpy_labels_and_sums(sites, m, res):
    W, H = m.shape
    cx = (np.arange(W) + 0.5) * res
    cy = (np.arange(H) + 0.5) * res
    dx = cx[:, None, None] - sites[None, None, :, 0]
    dy = cy[None, :, None] - sites[None, None, :, 1]
    d2 = dx * dx + dy * dy
    lab = np.argmin(d2, axis=2)  # first minimum: ties to the lowest index
    best = np.take_along_axis(d2, lab[:, :, None], 2)[:, :, 0]
    v = m.astype(np.float64)
    sums = np.zeros((len(sites), 4))
    for k in range(len(sites)):
        sel = lab == k
        sums[k] = [v[sel].sum(), (v * cx[:, None])[sel].sum(), (v * cy[None, :])[sel].sum(), (v * best)[sel].sum()]
    return lab, sums


@pytest.mark.parametrize("W,R,res,seed", [(128, 9, 1.0, 0), (96, 33, 0.5, 1), (64, 1, 2.0, 2), (100, 70, 1.0, 3)])
def test_bruteforce_labeller_is_the_stated_rule(W, R, res, seed):
    rng = np.random.default_rng(seed)
    m = rng.random((W, W)).astype(np.float32)
    sites = rng.uniform(0, W * res, (R, 2))
    if R > 4:
        sites[3] = sites[1]                    # twins: the lower index wins
        sites[4] = np.floor(sites[4]) + 0.5 * res  # on a cell centre
    lab, sums = orc.voronoi_bruteforce(sites, m, res)
    ref_lab, ref_sums = _numpy_labels_and_sums(sites, m, res)
    assert np.array_equal(lab, ref_lab)
    assert np.allclose(sums, ref_sums, rtol=1e-12, atol=1e-12)
    if R > 4:
        assert not np.any(lab == 3)


def test_bruteforce_ties_on_a_lattice():
    """sites on a regular lattice of cell centres: every bisector passes through cell centres, ties go down"""
    W = 64
    g = np.arange(8, W, 16) + 0.5
    sites = np.stack(np.meshgrid(g, g, indexing="ij"), -1).reshape(-1, 2)
    m = np.ones((W, W), np.float32)
    for order in (np.arange(len(sites)), np.arange(len(sites))[::-1]):
        lab, _ = orc.voronoi_bruteforce(sites[order], m, 1.0)
        ref, _ = _numpy_labels_and_sums(sites[order], m, 1.0)
        assert np.array_equal(lab, ref)


@pytest.mark.parametrize("W,R,seed", [(128, 8, 10), (200, 32, 11), (64, 3, 12)])
def test_polygon_cells_agree_with_the_labels(W, R, seed):
    rng = np.random.default_rng(seed)
    m = rng.random((W, W)).astype(np.float32)
    sites = rng.uniform(1, W - 1, (R, 2))
    cells = orc.voronoi(sites, m)
    _, sums = _numpy_labels_and_sums(sites, m, 1.0)
    area = 0.0
    for k, c in enumerate(cells):
        assert c.mass == pytest.approx(sums[k, 0], rel=1e-9)
        assert c.centroid[0] == pytest.approx(sums[k, 1] / sums[k, 0], rel=1e-9)
        assert c.centroid[1] == pytest.approx(sums[k, 2] / sums[k, 0], rel=1e-9)
        assert c.sum_idf_site_dist_sqr == pytest.approx(sums[k, 3], rel=1e-9)
        poly = c.polygon()
        x, y = poly[:, 0], poly[:, 1]
        a = 0.5 * abs(np.dot(x, np.roll(y, -1)) - np.dot(y, np.roll(x, -1)))
        area += a
        # the site lies inside its own convex cell
        e = np.roll(poly, -1, axis=0) - poly
        cr = e[:, 0] * (sites[k, 1] - poly[:, 1]) - e[:, 1] * (sites[k, 0] - poly[:, 0])
        assert np.all(cr >= -1e-9) or np.all(cr <= 1e-9)
    assert area == pytest.approx(W * W, rel=1e-9)  # the cells tile the world rectangle


def test_zero_mass_cell_falls_back_to_the_vertex_mean():
    """voronoi.cpp:143-150: mass < 1e-10 -> centroid = mean of the cell polygon's vertices"""
    W = 64
    m = np.zeros((W, W), np.float32)
    m[40:, :] = 1.0
    sites = np.array([[10.0, 32.0], [50.0, 32.0], [50.0, 10.0]])
    cells = orc.voronoi(sites, m)
    assert cells[0].mass == 0.0
    assert np.allclose(cells[0].centroid[:], cells[0].polygon().mean(0), atol=1e-12)
    assert cells[1].mass > 0 and cells[1].centroid[0] > 40.0


def test_single_site_owns_the_rectangle():
    """voronoi.cpp:168-181"""
    W = 48
    rng = np.random.default_rng(20)
    m = rng.random((W, W)).astype(np.float32)
    (c,) = orc.voronoi(np.array([[13.3, 31.7]]), m)
    v = m.astype(np.float64)
    i = np.arange(W) + 0.5
    assert c.mass == pytest.approx(v.sum(), rel=1e-12)
    assert c.centroid[0] == pytest.approx((v * i[:, None]).sum() / v.sum(), rel=1e-12)
    assert c.centroid[1] == pytest.approx((v * i[None, :]).sum() / v.sum(), rel=1e-12)
    assert sorted(map(tuple, np.round(c.polygon(), 9).tolist())) == [(0.0, 0.0), (0.0, 48.0), (48.0, 0.0), (48.0, 48.0)]
