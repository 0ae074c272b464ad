"""CPU cross-checks of the oracle's WorldIDF restatement (oracle/cc_oracle.c orc_idf_generate; reference
world_idf.cpp:41-101 = mode 0, generate_world_map.cu:76-280 = mode 1) against an independent NumPy / SciPy fp64
evaluation of the same definition: the integral of each truncated bivariate normal over a
cell from scipy.special.erfc, convex polygons by half-plane tests, normalisation by the map maximum.  Tolerances,
away from the truncation circle (where one fp32 rounding of the mid-point test decides whether a whole term is
present): mode 2 (the fp64 evaluation the GPU's mode 1 is gated against, SURVEY 8d gate ii) |delta| / max <= 1e-6;
modes 0 / 1 follow the reference's fp32 four-corner expression, whose cancellation (products of erfc ~ 1, difference
~ 1e-3) costs ~1e-4 of the maximum: TOL_FP32 = 1e-3, ten times tighter than the reference's own CPU-vs-GPU
tolerance (python/tests/test_parity.py:46, atol 1e-2)."""
import numpy as np
import pytest
from scipy.special import erfc

from oracle import oracle as orc

TOL_FP32 = 1e-3
TOL_FP64 = 1e-6


def _numpy_idf(N, res, bnd, trunc, norm, circular_sigma, polys=()):
    """fp64 truth.  circular_sigma: mode 0 divides both axes by sigma_x (bivariate_normal_distribution.h:152-154)"""
    x1 = np.arange(N) * res
    x2 = x1 + res
    mid = (x1 + x2) / 2
    out = np.zeros((N, N))
    near = np.zeros((N, N), bool)  # cells whose truncation test is within rounding of the circle
    for mx, my, sx, sy, sc, _ in bnd:
        if circular_sigma:
            sy = sx
        tx, ty = (mid - mx) / sx, (mid - my) / sy
        t2 = tx[:, None] ** 2 + ty[None, :] ** 2
        inside = t2 <= trunc * trunc
        near |= np.abs(t2 - trunc * trunc) < 1e-4
        qx = erfc((x1 - mx) / sx / np.sqrt(2)) - erfc((x2 - mx) / sx / np.sqrt(2))
        qy = erfc((x1 - my) / sy / np.sqrt(2)) - erfc((x2 - my) / sy / np.sqrt(2))
        out += np.where(inside, sc * qx[:, None] * qy[None, :] / 4.0, 0.0)
    poly = np.zeros((N, N))
    edge = np.zeros((N, N), bool)
    for pts, imp in polys:  # convex, counter-clockwise
        pts = np.asarray(pts, np.float64)
        ins = np.ones((N, N), bool)
        for a, b in zip(pts, np.roll(pts, -1, axis=0)):
            cr = (b[0] - a[0]) * (mid[None, :] - a[1]) - (b[1] - a[1]) * (mid[:, None] - a[0])
            ins &= cr > 0
            edge |= np.abs(cr) < 1e-3
        poly = np.maximum(poly, np.where(ins, imp, 0.0))
    out += poly
    m = out.max()
    factor = norm if m < 1e-10 else norm / m
    return out * factor, factor, near | edge


def _features(rng, N, F, res=1.0):
    bnd = np.zeros((F, 6))
    bnd[:, :2] = rng.uniform(0, N * res, (F, 2))
    bnd[:, 2] = bnd[:, 3] = rng.uniform(0.04, 0.08, F) * N * res
    bnd[:, 4] = rng.uniform(6, 10, F)
    return bnd


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("N,F,res,seed", [(128, 5, 1.0, 0), (96, 12, 0.5, 1), (64, 1, 2.0, 2)])
def test_gaussian_features_against_scipy(mode, N, F, res, seed):
    rng = np.random.default_rng(seed)
    p = orc.default_params(pWorldMapSize=N, pResolution=res)
    bnd = _features(rng, N, F, res)
    got, factor = orc.idf_generate(p, bnd, mode=mode)
    ref, ref_factor, skip = _numpy_idf(N, res, bnd, p.pTruncationBND, p.pNorm, circular_sigma=(mode == 0))
    tol = TOL_FP64 if mode == 2 else TOL_FP32
    assert factor == pytest.approx(ref_factor, rel=tol)
    assert got.max() == pytest.approx(p.pNorm, rel=1e-6)
    err = np.abs(got - ref)[~skip]
    assert err.max() <= tol * p.pNorm, f"max |delta| = {err.max()}"
    assert skip.mean() < 0.05


def test_anisotropic_sigma_differs_between_the_two_reference_paths():
    """mode 0 (the reference's CPU loop) divides both axes by sigma_x, mode 1 (its CUDA kernel) uses sigma_y for y:
    the oracle keeps that difference instead of hiding it"""
    N = 96
    p = orc.default_params(pWorldMapSize=N)
    bnd = np.array([[48.0, 48.0, 6.0, 12.0, 8.0, 0.0]])
    for mode, circ in ((0, True), (1, False)):
        got, _ = orc.idf_generate(p, bnd, mode=mode)
        ref, _, skip = _numpy_idf(N, 1.0, bnd, p.pTruncationBND, p.pNorm, circular_sigma=circ)
        assert np.abs(got - ref)[~skip].max() <= TOL_FP32
    a, _ = orc.idf_generate(p, bnd, mode=0)
    b, _ = orc.idf_generate(p, bnd, mode=1)
    assert np.abs(a - b).max() > 0.1


@pytest.mark.parametrize("mode", [0, 1])
def test_convex_polygons_against_half_plane_tests(mode):
    rng = np.random.default_rng(5)
    N = 160
    p = orc.default_params(pWorldMapSize=N)
    bnd = _features(rng, N, 3)
    polys = []
    for _ in range(4):
        c, r, n = rng.uniform(30, 130, 2), rng.uniform(8, 25), int(rng.integers(3, 9))
        ang = np.sort(rng.uniform(0, 2 * np.pi, n))
        polys.append((c + r * np.stack([np.cos(ang), np.sin(ang)], 1) + 0.137, float(rng.uniform(0.5, 2.0))))
    got, factor = orc.idf_generate(p, bnd, polys=polys, mode=mode)
    ref, ref_factor, skip = _numpy_idf(N, 1.0, bnd, p.pTruncationBND, p.pNorm, circular_sigma=(mode == 0), polys=polys)
    assert factor == pytest.approx(ref_factor, rel=TOL_FP32)
    assert np.abs(got - ref)[~skip].max() <= TOL_FP32
    # overlapping polygons contribute their maximum, not their sum (world_idf.cpp:66-72, generate_world_map.cu:124-144)
    both = [(np.array([[20.0, 20.0], [60.0, 20.0], [60.0, 60.0], [20.0, 60.0]]) + 0.25, 1.0),
            (np.array([[40.0, 40.0], [90.0, 40.0], [90.0, 90.0], [40.0, 90.0]]) + 0.25, 3.0)]
    got, _ = orc.idf_generate(p, np.zeros((0, 6)), polys=both, mode=mode)
    assert got[30, 30] == pytest.approx(1.0 / 3.0, rel=1e-6) and got[50, 50] == pytest.approx(1.0) and got[80, 80] == pytest.approx(1.0)
    assert got[10, 10] == 0.0


@pytest.mark.parametrize("mode", [0, 1])
def test_normalisation_edge_cases(mode):
    N = 32
    # no feature at all: the maximum is below 1e-10, the factor is pNorm and the map stays zero
    # (world_idf.cpp:83-89, generate_world_map.cu:252-256)
    p = orc.default_params(pWorldMapSize=N, pNorm=3.0)
    got, factor = orc.idf_generate(p, np.zeros((0, 6)), mode=mode)
    assert factor == pytest.approx(3.0) and not got.any()
    # a factor <= 1e-5 is returned but not applied (world_idf.cpp:90, generate_world_map.cu:257)
    p = orc.default_params(pWorldMapSize=N, pNorm=5e-7)
    bnd = np.array([[16.0, 16.0, 4.0, 4.0, 8.0, 0.0]])
    got, factor = orc.idf_generate(p, bnd, mode=mode)
    assert 0 < factor <= 1e-5
    assert got.max() > 0.05  # still the raw integral, not scaled by 6e-6
    # pNorm scales the maximum
    p = orc.default_params(pWorldMapSize=N, pNorm=2.5)
    got, _ = orc.idf_generate(p, bnd, mode=mode)
    assert got.max() == pytest.approx(2.5, rel=1e-6)


def test_truncation_keeps_only_the_disc():
    N = 128
    p = orc.default_params(pWorldMapSize=N)
    bnd = np.array([[64.0, 64.0, 10.0, 10.0, 8.0, 0.0]])
    for mode in (0, 1):
        got, _ = orc.idf_generate(p, bnd, mode=mode)
        i = np.arange(N) + 0.5
        r = np.hypot(i[:, None] - 64.0, i[None, :] - 64.0)
        assert not got[r > 2 * 10.0 + 1e-3].any()
        assert np.all(got[r < 2 * 10.0 - 1e-3] > 0)
