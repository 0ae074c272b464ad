"""The reference's own python/tests/test_env_io.py through the drop-in `coverage_control` package, plus the
text formats of world_idf.h:130-171 / world_idf.cpp:166-195 (CircularBND / BND / Uniform) and the position
file of coverage_system.cpp:368-381."""
import tempfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_env_io():
    """python/tests/test_env_io.py:27-46, statement for statement."""
    import coverage_control as cc

    params = cc.Parameters()
    env = cc.CoverageSystem(params)
    world_idf1 = env.GetWorldIDFObject()
    world_map = world_idf1.GetWorldMap()
    with tempfile.TemporaryDirectory() as tmp_dir:
        env.WriteEnvironment(tmp_dir + "/env.pos", tmp_dir + "/env.idf")
        world_idf2 = cc.WorldIDF(params, tmp_dir + "/env.idf")
        env2 = cc.CoverageSystem(params, world_idf2, tmp_dir + "/env.pos")
        map2 = world_idf2.GetWorldMap()
    assert np.allclose(world_map, map2)
    assert np.array_equal(world_map, map2)  # positions and features are written with full precision
    assert np.array_equal(np.array(env.GetRobotPositions()), np.array(env2.GetRobotPositions()))
    assert env2.GetNumRobots() == env.GetNumRobots()


def test_feature_file_formats_round_trip():
    import coverage_control as cc

    p = cc.Parameters()
    p.pWorldMapSize = p.pRobotMapSize = 256
    p.pNumRobots = 3
    idf = cc.WorldIDF(p)
    idf.AddNormalDistribution(cc.BivariateNormalDistribution(cc.Point2(100.25, 60.5), 14.0, 7.5))                  # CircularBND
    idf.AddNormalDistribution(cc.BivariateNormalDistribution(cc.Point2(180.0, 200.0), cc.Point2(12.0, 20.0), 0.0, 6.0))  # BND
    idf.AddUniformDistributionPolygon(cc.PolygonFeature(cc.PointVector(np.array([[20.5, 30.25], [90.0, 40.0], [50.0, 120.0]])), 0.004))
    idf.GenerateMap()
    assert idf.GetNumFeatures() == 3
    with tempfile.TemporaryDirectory() as d:
        assert idf.WriteDistributions(d + "/f.env") == 0
        text = open(d + "/f.env").read().split()
        assert text.count("CircularBND") == 1 and text.count("BND") == 1 and text.count("Uniform") == 1
        back = cc.WorldIDF(p, d + "/f.env")
        assert back.GetNumFeatures() == 3
        assert np.array_equal(back.GetWorldMap(), idf.GetWorldMap())
        assert back.GetNormalizationFactor() == idf.GetNormalizationFactor()
        # positions: one "x y" pair per line
        pos = np.array([[1.5, 2.25], [200.125, 17.0], [128.0, 128.0]])
        env = cc.CoverageSystem(p, idf, cc.PointVector(pos))
        assert env.WriteRobotPositions(d + "/r.pos") == 0
        assert np.array_equal(np.loadtxt(d + "/r.pos"), pos)
        env2 = cc.CoverageSystem(p, back, d + "/r.pos")
        assert np.array_equal(np.array(env2.GetRobotPositions()), pos)
        assert np.array_equal(env2.GetSystemMap(), env.GetSystemMap())


def test_parity():
    """python/tests/test_parity.py:31-58, statement for statement (GenerateMapCPU is served by the same kernel
    here, so the two maps are identical; the reference only asks for atol 1e-2)."""
    import coverage_control

    params = coverage_control.Parameters()
    assert coverage_control.CudaUtils.IsCudaAvailable()
    env_cuda = coverage_control.CoverageSystem(params)
    world_idf_obj_cuda = env_cuda.GetWorldIDFObject()
    world_idf_obj_cuda.GenerateMapCuda()
    world_map_cuda = world_idf_obj_cuda.GetWorldMap()
    init_robots = env_cuda.GetRobotPositions()
    env_cpu = coverage_control.CoverageSystem(params, world_idf_obj_cuda, init_robots)
    world_idf_obj_cpu = env_cpu.GetWorldIDFObject()
    world_idf_obj_cpu.GenerateMapCPU()
    world_map_cpu = world_idf_obj_cpu.GetWorldMap()
    assert np.allclose(world_map_cuda, world_map_cpu, atol=1e-2)
    assert np.array_equal(world_map_cuda, world_map_cpu)
    # the plot switches of the evaluators must not crash (python/evaluators/eval_single_env.py:85-151)
    assert env_cuda.PlotInitMap("/tmp", "InitMap") is None
    assert env_cuda.RecordPlotData() is None


@pytest.mark.gpu
def test_page_locked_host_buffers_are_copied_directly_and_agree():
    """ccb_ctrl_get_actions / ccb_env_step_actions copy straight from / to a caller's buffer when it is page-locked
    (transfers of 32 KB and more); the results must be those of the staged path with pageable buffers."""
    import ctypes as C

    import torch

    import coveragecontrol_b200 as ccb
    from conftest import random_world
    from coveragecontrol_b200 import BatchedCoverageSystem, BatchedCVT, _lib

    lib = _lib.load()
    E, R, N = 80, 32, 256  # E R 2 doubles = 40 KB
    p = ccb.Parameters()
    p.pWorldMapSize = p.pRobotMapSize = N
    p.pNumRobots = R
    p.pLocalMapSize = 64
    p.pSensorSize = 16
    bnd, _ = random_world(3, N, 6, R)
    idf = ccb.WorldIDF(p, device=0)
    for b in bnd:
        idf.AddNormalDistribution(ccb.BivariateNormalDistribution(ccb.Point2(b[0], b[1]), float(b[2]), float(b[4])))
    idf.GenerateMap()
    pos = np.random.default_rng(5).uniform(1, N - 1, (E, R, 2))
    # Two replicas driven by ONE controller's actions: the whole-map moments of this small batch come from the
    # tile-ordered kernel, whose fp64 atomics make two independent computations differ in the last bits -- only
    # the copies are under test, so both replicas get the same device buffer, once staged and once directly.
    envs = []
    for _ in range(2):
        env = BatchedCoverageSystem(p, E, R, device=0)
        for e in range(E):
            env.set_world(e, idf)
        env.reset(pos)
        envs.append(env)
    e0, e1 = envs
    c0 = BatchedCVT(e0, 0)
    pinned = torch.empty((E, R, 2), dtype=torch.float64).pin_memory()
    pv = pinned.numpy()
    dp = C.POINTER(C.c_double)
    for step in range(4):
        c0.compute_actions()
        a0 = c0.actions()  # pageable destination: staged through the library's pinned slot
        _lib.check(lib.ccb_ctrl_get_actions(c0._h, pv.ctypes.data_as(dp)))  # page-locked destination: direct
        assert np.array_equal(a0, pv), step
        assert np.abs(a0).max() > 0
        e0.step_actions(a0)
        bad = C.c_int(-1)
        _lib.check(lib.ccb_env_step_actions(e1._h, pv.ctypes.data_as(dp), C.byref(bad)))
        pv[:] = -1.0  # the call has consumed the buffer: overwriting it must not change the step
        e1._invalidate()
        assert np.array_equal(e0.positions(), e1.positions()), step
    assert np.allclose(e0.objective(), e1.objective(), rtol=1e-12, atol=0)
