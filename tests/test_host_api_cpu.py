"""Host side of the drop-in boundary that needs no GPU: the Parameters mirror (reference parameters.h:49-175,
parameters.cpp:36-261; the reference's own python/tests/test_parameters.py logic), the value types and opaque
vectors of core_binds.h:63-112, and the import names the reference package exposes.  Expected values come from
tests/golden/params_golden.json, derived from the reference tree by tests/golden/make_params_golden.py."""
import json
import os

import numpy as np
import pytest

from conftest import ROOT

GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "params_golden.json")))


def test_parameter_defaults_match_the_reference_header():
    import coverage_control as cc

    p = cc.Parameters()
    assert set(cc.Parameters._DEFAULTS) == set(GOLD["defaults"])
    for name, want in GOLD["defaults"].items():
        got = getattr(p, name)
        assert type(got) is type(want) and got == want, (name, got, want)


def test_every_parameter_is_a_toml_key_as_in_the_reference_test():
    """python/tests/test_parameters.py:34-67: every p* attribute, minus the p, is a leaf of the TOML file (the only
    exception is NumGaussianFeatures, which parameters.cpp:125 reads but the shipped TOML spells NumFeatures)."""
    import coverage_control as cc

    names = [a[1:] for a in dir(cc.Parameters()) if a.startswith("p") and a[1].isupper()]
    assert len(names) == len(GOLD["defaults"])
    missing = [n for n in names if n not in GOLD["toml_keys"]]
    assert missing == ["NumGaussianFeatures"]
    # and the paths the mirror reads end in exactly those keys
    assert sorted(path[-1] for path in cc.Parameters._TOML) == sorted(names)


def _toml_text(values):
    import coverage_control as cc

    tree = {}
    for path, attr in cc.Parameters._TOML.items():
        node = tree
        for key in path[:-1]:
            node = node.setdefault(key, {})
        node[path[-1]] = values[attr]

    def emit(node, prefix):
        out = []
        for k, v in node.items():
            if not isinstance(v, dict):
                out.append(f"{k} = {str(v).lower() if isinstance(v, bool) else repr(v)}")
        for k, v in node.items():
            if isinstance(v, dict):
                name = ".".join(prefix + [f'"{k}"' if "-" in k else k])
                out.append(f"[{name}]")
                out.extend(emit(v, prefix + [f'"{k}"' if "-" in k else k]))
        return out

    return "\n".join(emit(tree, [])) + "\n"


def test_parameters_load_a_config_file(tmp_path):
    import coverage_control as cc

    # the reference's shipped values, then every value changed
    for bump in (False, True):
        values = dict(GOLD["defaults"])
        values.update(GOLD["toml"])
        if bump:
            values = {k: (not v if isinstance(v, bool) else v + 3 if isinstance(v, int) else v * 1.5 + 0.25) for k, v in values.items()}
        f = tmp_path / f"params_{int(bump)}.toml"
        f.write_text(_toml_text(values))
        p = cc.Parameters(str(f))
        for name, want in values.items():
            assert getattr(p, name) == want, (name, getattr(p, name), want)
        q = cc.Parameters()
        q.SetConfig(str(f))
        assert all(getattr(q, n) == getattr(p, n) for n in values)
        c = p.copy()
        c.pNumRobots += 1
        assert p.pNumRobots == values["pNumRobots"]


def test_parameters_errors(tmp_path):
    import coverage_control as cc

    with pytest.raises(RuntimeError):
        cc.Parameters(str(tmp_path / "absent.toml"))
    bad = tmp_path / "bad.toml"
    bad.write_text("NumRobots = = 3\n")
    with pytest.raises(RuntimeError):
        cc.Parameters(str(bad))
    zero = tmp_path / "zero.toml"
    zero.write_text("NumRobots = 0\n")
    with pytest.raises(RuntimeError):
        cc.Parameters(str(zero))
    partial = tmp_path / "partial.toml"
    partial.write_text("NumRobots = 7\n[RobotModel]\nSensorSize = 32\n")
    p = cc.Parameters(str(partial))
    assert p.pNumRobots == 7 and p.pSensorSize == 32 and p.pWorldMapSize == GOLD["defaults"]["pWorldMapSize"]


def test_value_types_and_vectors():
    """core_binds.h:63-112; the usages of python/tests/test_coverage.py:40-43 and controllers.py:150."""
    import coverage_control as cc

    assert cc.Point2().dtype == np.float64 and np.array_equal(cc.Point2(), [0, 0])
    assert np.array_equal(cc.Point2(3, 4.5), [3.0, 4.5])
    src = np.array([1.0, 2.0])
    cp = cc.Point2(src)
    cp[0] = 9
    assert src[0] == 1.0
    with pytest.raises(TypeError):
        cc.Point2([1, 2, 3])
    from coverage_control import _core  # Point2f / intVector are bound (core_binds.h:71,80) but not re-exported

    assert _core.Point2f(1, 2).dtype == np.float32

    pv = cc.PointVector(np.array([[1.0, 2.0], [3.0, 4.0]]))
    pv.append(np.array([5.0, 6.0]))
    pv.append(cc.Point2(7, 8))
    assert len(pv) == 4 and isinstance(pv, cc.PointVector)
    assert np.array_equal(pv.to_numpy(), [[1, 2], [3, 4], [5, 6], [7, 8]])
    assert np.array_equal(np.array(pv), pv.to_numpy())
    with pytest.raises(TypeError):
        pv.append([1.0, 2.0, 3.0])
    assert len(cc.PointVector(3)) == 3 and np.array_equal(cc.PointVector(3).to_numpy(), np.zeros((3, 2)))

    dv = cc.DblVector([1, 2.5])
    dv.append(3)
    assert list(dv) == [1.0, 2.5, 3.0] and all(isinstance(v, float) for v in dv)
    dvv = cc.DblVectorVector([[1, 2], [3]])
    assert isinstance(dvv[0], cc.DblVector) and list(dvv[1]) == [3.0]
    assert list(_core.intVector([1.0, 2.0])) == [1, 2]

    b = cc.BivariateNormalDistribution()
    assert np.array_equal(b.GetMean(), [0, 0]) and np.array_equal(b.GetSigma(), [1, 1]) and b.GetScale() == 1.0 and b.GetRho() == 0.0
    b = cc.BivariateNormalDistribution(cc.Point2(10, 20), 5.0, 8.0)
    assert np.array_equal(b.GetMean(), [10, 20]) and np.array_equal(b.GetSigma(), [5, 5]) and b.GetScale() == 8.0
    b = cc.BivariateNormalDistribution(cc.Point2(1, 2), cc.Point2(3, 4), 0.5, 6.0)
    assert np.array_equal(b.GetSigma(), [3, 4]) and b.GetScale() == 6.0
    assert b.GetRho() == 0.0  # bivariate_normal_distribution.h:103-116 tests its member rho_, not the argument

    f = cc.PolygonFeature(cc.PointVector([[0, 0], [4, 0], [0, 4]]), 0.25)
    assert f.size == 3 and f.imp == 0.25 and isinstance(f.poly, cc.PointVector)
    assert cc.PolygonFeature().size == 0


def test_import_names_of_the_reference_package():
    """python/coverage_control/__init__.py:11-14, core/__init__.py:7-22, algorithms/__init__.py."""
    import coverage_control as cc
    import coverage_control.algorithms as alg
    import coverage_control.core as core
    from coverage_control import CoverageEnvUtils, IOUtils  # noqa: F401

    for name in ("Point2", "PointVector", "DblVector", "DblVectorVector", "Parameters", "BivariateNormalDistribution",
                 "BNDVector", "WorldIDF", "RobotModel", "CoverageSystem", "PolygonFeature", "VoronoiCell", "VoronoiCells",
                 "CudaUtils"):
        assert hasattr(core, name) or name == "RobotModel", name
        assert hasattr(cc, name) or name == "RobotModel", name
    for name in ("ClairvoyantCVT", "CentralizedCVT", "DecentralizedCVT", "NearOptimalCVT"):
        assert hasattr(alg, name), name
    assert isinstance(cc.__version__, str)
    assert cc.CudaUtils.UseCuda() in (True, False)
    with pytest.raises(RuntimeError):
        cc.CudaUtils.SetUseCuda(False)  # there is no CPU path to switch to


def test_io_utils(tmp_path, monkeypatch):
    """python/coverage_control/io_utils.py:50-132: sanitize_path, load_toml, load_yaml, load_tensor."""
    import torch

    from coverage_control import IOUtils

    monkeypatch.setenv("CCB_TEST_DIR", str(tmp_path))
    assert IOUtils.sanitize_path("$CCB_TEST_DIR/a/../b.toml") == os.path.join(str(tmp_path), "b.toml")
    (tmp_path / "b.toml").write_text("NumRobots = 5\n[RobotModel]\nSensorSize = 32\n")
    assert IOUtils.load_toml("$CCB_TEST_DIR/b.toml") == {"NumRobots": 5, "RobotModel": {"SensorSize": 32}}
    (tmp_path / "c.yaml").write_text("a: 1\nb: [2, 3]\n")
    assert IOUtils.load_yaml(str(tmp_path / "c.yaml")) == {"a": 1, "b": [2, 3]}
    t = torch.arange(6, dtype=torch.float32).reshape(2, 3)
    torch.save(t, tmp_path / "t.pt")
    assert torch.equal(IOUtils.load_tensor(tmp_path / "t.pt"), t)
    for fn in (IOUtils.load_toml, IOUtils.load_yaml, IOUtils.load_tensor):
        with pytest.raises(FileNotFoundError):
            fn(str(tmp_path / "absent"))


def test_point_vector_array_backed_semantics():
    """PointVector (core_binds.h:82 bind_vector<Point2>) keeps one private array when built from an [n, 2] array or
    another PointVector (the per-step GetActions -> StepActions path); it must still behave like a list of Point2."""
    import copy
    import pickle

    from coveragecontrol_b200.core import Point2, PointVector

    a = np.arange(12, dtype=np.float64).reshape(6, 2)
    pv = PointVector(a)
    assert len(pv) == 6 and np.array_equal(pv.to_numpy(), a) and np.array_equal(pv[2], a[2])
    pv[3][0] = 70.0  # through an item: visible in to_numpy, not in the source
    assert pv.to_numpy()[3, 0] == 70.0 and a[3, 0] == 6.0
    out = pv.to_numpy()
    out[0, 0] = -1.0  # to_numpy is a copy
    assert pv[0][0] == 0.0
    pv[4] = Point2(1, 2)  # replacing an item, then growing the list
    pv.append((5, 6))
    pv += [(7, 8)]
    got = pv.to_numpy()
    assert got.shape == (8, 2) and got[4].tolist() == [1, 2] and got[3, 0] == 70.0 and got[-1].tolist() == [7, 8]
    assert pv.pop().tolist() == [7, 8] and pv.to_numpy().shape == (7, 2)
    q = PointVector(pv)
    q[0][1] = 99.0
    assert pv[0][1] == 1.0 and q.to_numpy()[0, 1] == 99.0
    for clone in (copy.copy, copy.deepcopy, lambda x: pickle.loads(pickle.dumps(x))):
        c = clone(q)
        c[1][0] = 5.0
        assert isinstance(c, PointVector) and c.to_numpy()[1, 0] == 5.0 and q.to_numpy()[1, 0] == 2.0
    assert PointVector().to_numpy().shape == (0, 2) and PointVector(3).to_numpy().tolist() == [[0, 0]] * 3
    assert PointVector([(1, 2), [3, 4]]).to_numpy().tolist() == [[1, 2], [3, 4]]
    with pytest.raises(TypeError):
        PointVector([(1, 2, 3)])
