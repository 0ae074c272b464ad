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
