"""Row-band tiling on the GPU (BASELINE config 4 mechanics), single process: the band entry points of
the C ABI give, band by band, exactly what the whole-map calls give."""
import ctypes as C

import numpy as np
import pytest

from conftest import random_world

pytestmark = pytest.mark.gpu


def _setup(N, R, F, seed):
    import coverage_control as cc
    from coveragecontrol_b200 import BatchedCoverageSystem

    p = cc.Parameters()
    p.pWorldMapSize = p.pRobotMapSize = N
    p.pNumRobots = R
    bnd, pos = random_world(seed, N, F, R)
    idf = cc.WorldIDF(p)
    for b in bnd:
        idf.AddNormalDistribution(cc.BivariateNormalDistribution(cc.Point2(b[0], b[1]), float(b[2]), float(b[4])))
    return cc, p, idf, pos, BatchedCoverageSystem


@pytest.mark.parametrize("N,bands", [(512, [(0, 192), (192, 448), (448, 512)]), (300, [(0, 77), (77, 300)])])
def test_idf_bands_equal_whole_map_bit_for_bit(N, bands):
    import torch

    from coveragecontrol_b200 import _lib
    from coveragecontrol_b200.parallel import _wrap_dev

    cc, p, idf, _, _ = _setup(N, 4, 9, 11)
    idf.GenerateMap()
    whole = np.array(idf.GetWorldMap())
    cc2, p2, idf2, _, _ = _setup(N, 4, 9, 11)
    lib = _lib.load()
    # every "rank" generates its band into its own map; the maxima are combined as the all-reduce would
    mx = []
    for r0, r1 in bands:
        _lib.check(lib.ccb_idf_generate_band(idf2._h, 0, r0, r1))
        t = _wrap_dev(lib.ccb_idf_band_max_dev(idf2._h), (1,), torch.int32, torch.device("cuda", 0))
        mx.append(int(t.item()))
    t.fill_(max(mx))
    for r0, r1 in bands:
        _lib.check(lib.ccb_idf_normalize_band(idf2._h, r0, r1))
    idf2._host = None
    banded = np.array(idf2.GetWorldMap())
    assert np.array_equal(whole.view(np.int32), banded.view(np.int32))
    assert idf2.GetNormalizationFactor() == idf.GetNormalizationFactor()


@pytest.mark.parametrize("vg", ["0", "1"])  # site-ordered / tile-ordered global Voronoi kernel
@pytest.mark.parametrize("N,R,nb", [(1024, 32, 4), (512, 64, 3), (300, 7, 2)])
def test_band_partials_sum_to_whole_map_moments(monkeypatch, N, R, nb, vg):
    import torch

    monkeypatch.setenv("CCB_VG_GENERIC", vg)

    from coveragecontrol_b200 import _lib
    from coveragecontrol_b200.parallel import band_rows

    cc, p, idf, pos, Batched = _setup(N, R, 12, 3)
    idf.GenerateMap()
    env = Batched(p, 1, R)
    env.set_world(0, idf)
    env.reset(pos[None])
    mass, cen, _, obj = env.cells(0)
    lib = _lib.load()
    tot = torch.zeros((1, R, 4), dtype=torch.float64, device="cuda")
    part = torch.zeros_like(tot)
    for g in range(nb):
        r0, r1 = band_rows(N, nb, g)
        _lib.check(lib.ccb_env_voronoi_partials_dev(env._h, 0, r0, r1, part.data_ptr()))
        tot += part
    _lib.check(lib.ccb_env_set_band_mode(env._h, 1))
    with pytest.raises(RuntimeError):
        env.cells(0)  # band mode refuses to label the whole map behind the caller's back
    _lib.check(lib.ccb_env_finalize_from_partials_dev(env._h, tot.data_ptr()))
    mass2, cen2, _, obj2 = env.cells(0)
    assert np.allclose(mass2, mass, rtol=1e-12, atol=1e-12)
    assert np.allclose(cen2, cen, rtol=1e-10, atol=1e-9)
    assert np.allclose(obj2, obj, rtol=1e-10, atol=1e-6)


@pytest.mark.parametrize("graph", [False, True])
@pytest.mark.parametrize("kind", [0, 1, 2])
def test_banded_system_step_equals_plain_rollout(kind, graph):
    """BandedCoverageSystem.step (one process: the band is the whole map, the robot shard all robots, the all-reduce a
    no-op) for the three controller families, eager and as ONE captured CUDA graph per step, against the plain roll-out."""
    import torch

    from coveragecontrol_b200 import BatchedCVT
    from coveragecontrol_b200.parallel import BandedCoverageSystem

    N, R = 512, 12
    cc, p, idf, pos, Batched = _setup(N, R, 10, 7)
    idf.GenerateMap()
    stream = torch.cuda.Stream()
    plain = Batched(p, 1, R)
    banded_env = Batched(p, 1, R, stream=stream.cuda_stream)
    for e in (plain, banded_env):
        e.set_world(0, idf)
        e.reset(pos[None])
    c0 = BatchedCVT(plain, kind)
    banded = BandedCoverageSystem(banded_env, 0, stream=stream, graph=graph)
    banded.compute_cells()
    c1 = BatchedCVT(banded_env, kind)
    for it in range(8):
        c0.rollout(1)
        obj = banded.step(c1, with_objective=(it % 2 == 0))
        stream.synchronize()
        assert np.allclose(plain.positions(), banded_env.positions(), rtol=0, atol=1e-9), it
        if obj is not None:
            assert float(obj[0]) == pytest.approx(plain.objective()[0], rel=1e-10)
    assert banded.per_rank_bytes > 4 * N * N


def test_multi_gpu_row_bands_and_robot_shards():
    """scripts/multigpu_check.py under torchrun on the GPUs of the box, two ranks where two are visible, else one
    (profiles/ holds the logs of the 2- and 8-GPU runs): banded IDF bit-identical, the three controllers eager and
    graph-captured with the NCCL all-reduce inside the graph, replicas identical across ranks."""
    import os
    import subprocess
    import sys

    import torch

    from conftest import ROOT

    ranks = 2 if torch.cuda.device_count() >= 2 else 1
    env = dict(os.environ, CHECK_N="1024")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(ranks),
                          "--master-addr", "127.0.0.1", "--master-port", "29533",
                          os.path.join(ROOT, "scripts", "multigpu_check.py")], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0 and "multigpu_check ok" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_idf_band_with_unaligned_start_and_narrow_feature_past_a_bucket_boundary():
    """the kernel picks its feature bucket (256 x 256 cells) from the tile origin, so tiles must
    not straddle a bucket boundary when a band starts at a row that is no multiple of 32.  A narrow Gaussian just
    past row 256 / column 256 (truncation box inside the next bucket only) and a band starting at row 77."""
    import torch

    import coverage_control as cc
    from coveragecontrol_b200 import _lib
    from coveragecontrol_b200.parallel import _wrap_dev

    N = 512
    p = cc.Parameters()
    p.pWorldMapSize = p.pRobotMapSize = N
    feats = [(262.4, 263.1, 1.5, 7.0), (258.2, 40.3, 0.9, 9.0), (100.7, 259.6, 1.2, 8.0), (300.0, 300.0, 30.0, 6.0)]

    def make():
        idf = cc.WorldIDF(p)
        for mx, my, sg, sc in feats:
            idf.AddNormalDistribution(cc.BivariateNormalDistribution(cc.Point2(mx, my), sg, sc))
        return idf

    whole = make()
    whole.GenerateMap()
    want = np.array(whole.GetWorldMap())
    assert want[262, 263] > 0 and want[258, 40] > 0
    lib = _lib.load()
    banded = make()
    bands = [(0, 77), (77, 250), (250, 257), (257, 512)]
    mx = []
    for r0, r1 in bands:
        _lib.check(lib.ccb_idf_generate_band(banded._h, 0, r0, r1))
        t = _wrap_dev(lib.ccb_idf_band_max_dev(banded._h), (1,), torch.int32, torch.device("cuda", 0))
        mx.append(int(t.item()))
    t.fill_(max(mx))
    for r0, r1 in bands:
        _lib.check(lib.ccb_idf_normalize_band(banded._h, r0, r1))
    banded._host = None
    got = np.array(banded.GetWorldMap())
    assert np.array_equal(want.view(np.int32), got.view(np.int32))
