"""The host-side pieces of CoverageEnvUtils (coveragecontrol_b200/coverage_env_utils.py) against the REFERENCE's own
python/coverage_control/coverage_env_utils.py, imported here as a sub-module of this repo's drop-in `coverage_control`
package (its `from .core import ...` resolves to the drop-in value types; torch_geometric, absent from the image, is
stubbed).  Both are driven by the same stand-in environment, so no GPU is needed: to_tensor for every container type,
get_robot_positions, get_voronoi_features, get_weights (host path), resize_maps.  Runs where /root/reference exists
(this container); the device paths of the same functions are covered by tests/test_gpu_observation.py."""
import importlib.util
import os
import sys
import types

import numpy as np
import pytest
import torch

REF = "/root/reference/python/coverage_control/coverage_env_utils.py"
pytestmark = pytest.mark.skipif(not os.path.exists(REF), reason="reference tree absent (GPU box)")


@pytest.fixture(scope="module")
def ref_utils():
    import coverage_control  # noqa: F401  (the drop-in package)

    if "torch_geometric" not in sys.modules:
        tg = types.ModuleType("torch_geometric")
        tg.data = types.ModuleType("torch_geometric.data")
        tg.data.Data = lambda **kw: types.SimpleNamespace(**kw)
        sys.modules["torch_geometric"], sys.modules["torch_geometric.data"] = tg, tg.data
    spec = importlib.util.spec_from_file_location("coverage_control._reference_env_utils", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.CoverageEnvUtils


class StandInEnv:
    """what the host-side functions ask of a CoverageSystem"""

    def __init__(self, pos, feats):
        import coverage_control as cc

        self._pos, self._feats, self._cc = pos, feats, cc

    def GetNumRobots(self):
        return len(self._pos)

    def GetRobotPositions(self):
        return self._cc.PointVector(self._pos)

    def GetRobotVoronoiFeatures(self):
        out = self._cc.DblVectorVector()
        for row in self._feats:
            out.append(self._cc.DblVector([float(v) for v in row]))
        return out


def test_to_tensor_all_container_types(ref_utils):
    import coverage_control as cc
    from coverage_control import CoverageEnvUtils as ours

    rng = np.random.default_rng(0)
    pts = rng.uniform(0, 1024, (7, 2))
    cases = [rng.random((5, 3)), rng.random(4).astype(np.float32), cc.PointVector(pts), cc.DblVector([1.5, 2.25, -3.0])]
    for data in cases:
        a, b = ours.to_tensor(data), ref_utils.to_tensor(data)
        assert a.dtype == b.dtype == torch.float32 and a.shape == b.shape and torch.equal(a, b), type(data)
    with pytest.raises(ValueError):
        ours.to_tensor([1, 2, 3])
    with pytest.raises(ValueError):
        ref_utils.to_tensor([1, 2, 3])


def test_positions_features_and_weights(ref_utils):
    import coverage_control as cc
    from coverage_control import CoverageEnvUtils as ours

    rng = np.random.default_rng(1)
    for R in (3, 32, 40):  # torch.cdist switches to a matrix product above 25 points
        pos = rng.uniform(0, 1024, (R, 2))
        pos[1] = pos[0] + [128.0 / np.sqrt(2), 128.0 / np.sqrt(2)]  # on the 1/e threshold
        env = StandInEnv(pos, rng.normal(size=(R, 7)) * 100)
        p = cc.Parameters()
        assert torch.equal(ours.get_robot_positions(env), ref_utils.get_robot_positions(env))
        assert torch.equal(ours.get_voronoi_features(env), ref_utils.get_voronoi_features(env))
        a, b = ours.get_weights(env, p), ref_utils.get_weights(env, p)
        assert a.dtype == b.dtype and torch.equal(a, b)
        assert not a.diagonal().any() and (a > 0).any()


def test_resize_maps(ref_utils):
    from coverage_control import CoverageEnvUtils as ours

    g = torch.Generator().manual_seed(2)
    maps = torch.rand(5, 256, 256, generator=g)
    assert torch.equal(ours.resize_maps(maps, 32), ref_utils.resize_maps(maps, 32))
