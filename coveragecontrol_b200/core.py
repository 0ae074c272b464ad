"""Host-side mirror of the reference's ``coverage_control._core`` module.

Same class / method names, argument meaning and error behaviour as the pybind11 table in
reference ``cppsrc/python_bindings/core_binds.h:59-394``; every method that touches a map, a
Voronoi cell or a robot state forwards to the C ABI of ``include/ccb200.h`` (hand-written sm_100a
CUDA).  There is no CPU implementation in this package.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import sys
import tomllib

import numpy as np

from . import _lib
from ._lib import CParams, c_double_p, c_float_p, c_int_p

__all__ = [
    "Point2", "Point2f", "PointVector", "Point3Vector", "DblVector", "DblVectorVector", "intVector",
    "MapTypeVector", "PolygonFeature", "PolygonVector", "VoronoiCell", "VoronoiCells",
    "BivariateNormalDistribution", "BNDVector", "WorldIDF", "RobotModel", "CoverageSystem",
    "Parameters", "CudaUtils", "ClairvoyantCVT", "CentralizedCVT", "DecentralizedCVT", "NearOptimalCVT",
    "BatchedCoverageSystem", "BatchedCVT",
]

kEps = 1e-10       # reference constants.h:48
kLargeEps = 1e-4  # reference constants.h:50
_MAX_PRECISION = 19  # reference constants.h:57 kMaxPrecision (long double digits10 + 1)

MAP_WORLD, MAP_SYSTEM, MAP_EXPLORED_IDF, MAP_EXPLORATION, MAP_ROBOT = 0, 1, 2, 3, 4
MAP_SENSOR_VIEW, MAP_LOCAL, MAP_OBSTACLE, MAP_ROBOT_EXPLORATION, MAP_ROBOT_SYSTEM = 5, 6, 7, 8, 9


def _dp(a):
    return a.ctypes.data_as(c_double_p)


def _fp(a):
    return a.ctypes.data_as(c_float_p)


# --------------------------------------------------------------------------------------------
# small value types (core_binds.h:63-86)
# --------------------------------------------------------------------------------------------
def Point2(a=None, b=None) -> np.ndarray:
    """``Point2()``, ``Point2(a, b)`` or ``Point2(p)`` -> float64[2] (Eigen::Vector2d)."""
    if a is None:
        return np.zeros(2, np.float64)
    if b is None:
        p = np.asarray(a, np.float64).reshape(-1)
        if p.size != 2:
            raise TypeError("Point2(p): p must have two entries")
        return p.copy()
    return np.array([a, b], np.float64)


def Point2f(a=None, b=None) -> np.ndarray:
    if a is None:
        return np.zeros(2, np.float32)
    return np.array([a, b], np.float32)


class _TypedList(list):
    """Stand-in for a pybind11 ``bind_vector``: a real class (``isinstance`` works) with list API."""

    @staticmethod
    def _coerce(x):
        return x

    def __init__(self, it=()):
        if isinstance(it, (int, np.integer)):  # bind_vector(n): n default items
            super().__init__(self._coerce(self._default()) for _ in range(int(it)))
        else:
            super().__init__(self._coerce(x) for x in it)

    @staticmethod
    def _default():
        return 0

    def append(self, x):
        super().append(self._coerce(x))

    def extend(self, it):
        super().extend(self._coerce(x) for x in it)

    def __iadd__(self, it):
        self.extend(it)
        return self

    def insert(self, i, x):
        super().insert(i, self._coerce(x))

    def __setitem__(self, i, x):
        if isinstance(i, slice):
            super().__setitem__(i, [self._coerce(v) for v in x])
        else:
            super().__setitem__(i, self._coerce(x))


class PointVector(_TypedList):
    """``std::vector<Point2>``; accepts an ``[R, 2]`` array (core_binds.h:82)."""

    @staticmethod
    def _coerce(x):
        p = np.array(x, np.float64).reshape(-1)
        if p.size != 2:
            raise TypeError("PointVector items must be Point2")
        return p

    @staticmethod
    def _default():
        return (0.0, 0.0)

    def __init__(self, it=()):
        # [n, 2] arrays and other PointVectors (what GetActions returns and StepActions takes every step): one
        # private copy, the items are its rows.  Writes through an item land in the copy; anything that changes
        # the list itself drops it (to_numpy then gathers the items again).
        self._arr = None
        if type(it) is PointVector:
            it = it.to_numpy()
        if isinstance(it, np.ndarray) and it.ndim == 2 and it.shape[1] == 2:
            arr = np.array(it, np.float64, order="C")
            list.__init__(self, arr)
            self._arr = arr
        else:
            super().__init__(it)

    def to_numpy(self) -> np.ndarray:
        if self._arr is not None:
            return self._arr.copy()
        return np.array(self, np.float64).reshape(-1, 2)

    def __reduce__(self):  # copy / deepcopy / pickle rebuild the rows over ONE new array
        return (PointVector, (self.to_numpy(),))

    def _structural(name):  # noqa: N805 - builds the list mutators that invalidate the backing array
        base = getattr(_TypedList, name)

        def method(self, *a, **k):
            self._arr = None
            return base(self, *a, **k)

        method.__name__ = name
        return method

    for _n in ("append", "extend", "insert", "__setitem__", "__delitem__", "pop", "remove", "clear", "sort", "reverse",
               "__iadd__", "__imul__"):
        locals()[_n] = _structural(_n)
    del _n, _structural


class Point3Vector(_TypedList):
    @staticmethod
    def _coerce(x):
        return np.array(x, np.float64).reshape(3)

    @staticmethod
    def _default():
        return (0.0, 0.0, 0.0)


class DblVector(_TypedList):
    @staticmethod
    def _coerce(x):
        return float(x)

    @staticmethod
    def _default():
        return 0.0


class DblVectorVector(_TypedList):
    @staticmethod
    def _coerce(x):
        return x if isinstance(x, DblVector) else DblVector(x)

    @staticmethod
    def _default():
        return ()


class intVector(_TypedList):
    @staticmethod
    def _coerce(x):
        return int(x)


class MapTypeVector(_TypedList):
    @staticmethod
    def _coerce(x):
        return np.asarray(x, np.float32)


class PolygonFeature:
    """reference typedefs.h:57-65."""

    def __init__(self, poly=None, imp=0.0):
        self.poly = PointVector(poly if poly is not None else ())
        self.imp = float(np.float32(imp))
        self.size = len(self.poly)


class PolygonVector(_TypedList):
    pass


class BivariateNormalDistribution:
    """reference bivariate_normal_distribution.h:53-183 (parameters only; the integration is on the GPU)."""

    def __init__(self, mean=None, sigma=None, a=None, b=None):
        if mean is None:  # :68-73
            self._mean, self._sigma, self._rho, self._scale = Point2(0, 0), Point2(1, 1), 0.0, 1.0
            return
        self._mean = Point2(mean)
        if np.ndim(sigma) == 0:  # (mean, sigma, scale=1)  :85-92
            self._sigma = Point2(float(sigma), float(sigma))
            self._rho = 0.0
            self._scale = 1.0 if a is None else float(a)
        else:  # (mean, sigma2, rho, scale=1)  :103-116.  The reference tests its *member* rho_
            # (always 0) instead of the argument, so rho is dropped; kept for parity.
            self._sigma = Point2(sigma)
            self._rho = 0.0
            self._scale = 1.0 if b is None else float(b)

    def GetMean(self):
        return self._mean.copy()

    def GetSigma(self):
        return self._sigma.copy()

    def GetRho(self):
        return self._rho

    def GetScale(self):
        return self._scale


class BNDVector(_TypedList):
    pass


class VoronoiCell:
    """reference voronoi.h:52-103; read-only ``cell``, ``mass``, ``centroid``, ``site`` (core_binds.h:96-100)."""

    __slots__ = ("_cell", "_mass", "_centroid", "_site", "_obj")

    def __init__(self, cell, mass, centroid, site, obj=0.0):
        self._cell, self._mass, self._centroid, self._site, self._obj = cell, mass, centroid, site, obj

    cell = property(lambda s: s._cell)
    mass = property(lambda s: s._mass)
    centroid = property(lambda s: s._centroid)
    site = property(lambda s: s._site)


class VoronoiCells(_TypedList):
    pass


class CudaUtils:
    """reference cuda_utils.h:45-124 static helpers (core_binds.h:388-393)."""

    _use = True

    @staticmethod
    def IsCudaAvailable() -> bool:
        return bool(_lib.load().ccb_cuda_available())

    @staticmethod
    def UseCuda() -> bool:
        return CudaUtils._use and CudaUtils.IsCudaAvailable()

    @staticmethod
    def SetUseCuda(use: bool) -> None:
        if not use:
            raise RuntimeError("coveragecontrol_b200 has no CPU path: SetUseCuda(False) is not supported")
        CudaUtils._use = True


# --------------------------------------------------------------------------------------------
# Parameters (parameters.h:49-175, parameters.cpp:36-261)
# --------------------------------------------------------------------------------------------
class Parameters:
    _DEFAULTS = dict(
        pNumRobots=32, pNumGaussianFeatures=32, pNumPolygons=0, pMaxVertices=10, pPolygonRadius=64.0,
        pPlotScale=1.0, pResolution=1.0, pWorldMapSize=1024, pRobotMapSize=1024, pLocalMapSize=256,
        pUpdateRobotMap=True, pUpdateExplorationMap=True, pUpdateSensorView=True, pUpdateSystemMap=True,
        pTruncationBND=2.0, pNorm=1.0, pMinSigma=40.0, pMaxSigma=50.0, pMinPeak=6.0, pMaxPeak=10.0,
        pUnknownImportance=0.5, pRobotMapUseUnknownImportance=False, pSensorSize=64,
        pCommunicationRange=128.0, pMaxRobotSpeed=5.0, pRobotInitDist=1024.0, pRobotPosHistorySize=20,
        pTimeStep=1.0, pAddNoisePositions=False, pPositionsNoiseSigma=0.0, pEpisodeSteps=2000,
        pCheckOscillations=True, pLloydMaxIterations=100, pLloydNumTries=10, pNumFrontiers=10,
    )
    # (toml path) -> attribute, exactly the keys parameters.cpp reads
    _TOML = {
        ("NumRobots",): "pNumRobots",
        ("IO", "PlotScale"): "pPlotScale",
        ("Environment", "Maps", "Resolution"): "pResolution",
        ("Environment", "Maps", "WorldMapSize"): "pWorldMapSize",
        ("Environment", "Maps", "RobotMapSize"): "pRobotMapSize",
        ("Environment", "Maps", "LocalMapSize"): "pLocalMapSize",
        ("Environment", "Maps", "UpdateSettings", "UpdateRobotMap"): "pUpdateRobotMap",
        ("Environment", "Maps", "UpdateSettings", "UpdateSensorView"): "pUpdateSensorView",
        ("Environment", "Maps", "UpdateSettings", "UpdateExplorationMap"): "pUpdateExplorationMap",
        ("Environment", "Maps", "UpdateSettings", "UpdateSystemMap"): "pUpdateSystemMap",
        ("Environment", "IDF", "NumGaussianFeatures"): "pNumGaussianFeatures",
        ("Environment", "IDF", "TruncationBND"): "pTruncationBND",
        ("Environment", "IDF", "Norm"): "pNorm",
        ("Environment", "IDF", "MinSigma"): "pMinSigma",
        ("Environment", "IDF", "MaxSigma"): "pMaxSigma",
        ("Environment", "IDF", "MinPeak"): "pMinPeak",
        ("Environment", "IDF", "MaxPeak"): "pMaxPeak",
        ("Environment", "IDF", "NumPolygons"): "pNumPolygons",
        ("Environment", "IDF", "MaxVertices"): "pMaxVertices",
        ("Environment", "IDF", "PolygonRadius"): "pPolygonRadius",
        ("Environment", "IDF", "UnknownImportance"): "pUnknownImportance",
        ("Environment", "IDF", "RobotMapUseUnknownImportance"): "pRobotMapUseUnknownImportance",
        ("RobotModel", "SensorSize"): "pSensorSize",
        ("RobotModel", "CommunicationRange"): "pCommunicationRange",
        ("RobotModel", "MaxRobotSpeed"): "pMaxRobotSpeed",
        ("RobotModel", "RobotInitDist"): "pRobotInitDist",
        ("RobotModel", "RobotPosHistorySize"): "pRobotPosHistorySize",
        ("RobotModel", "TimeStep"): "pTimeStep",
        ("RobotModel", "AddNoise", "AddNoisePositions"): "pAddNoisePositions",
        ("RobotModel", "AddNoise", "PositionsNoiseSigma"): "pPositionsNoiseSigma",
        ("Algorithm", "EpisodeSteps"): "pEpisodeSteps",
        ("Algorithm", "CheckOscillations"): "pCheckOscillations",
        ("Algorithm", "Global-CVT", "LloydMaxIterations"): "pLloydMaxIterations",
        ("Algorithm", "Global-CVT", "LloydNumTries"): "pLloydNumTries",
        ("Algorithm", "Exploration", "NumFrontiers"): "pNumFrontiers",
    }

    def __init__(self, config_file: str | None = None):
        for k, v in self._DEFAULTS.items():
            setattr(self, k, v)
        self.config_file_ = ""
        if config_file is not None:
            self.SetConfig(config_file)

    def SetConfig(self, config_file: str) -> None:
        self.config_file_ = str(config_file)
        self._parse()

    def _parse(self) -> None:
        print(f"Using config file: {self.config_file_}")
        if not os.path.exists(self.config_file_):
            print(f"Could not find config file {self.config_file_}", file=sys.stderr)
            raise RuntimeError("Could not open config file")
        try:
            with open(self.config_file_, "rb") as f:
                cfg = tomllib.load(f)
        except Exception as exc:  # parameters.cpp:47-52
            print(f"Error parsing config file: {exc}", file=sys.stderr)
            raise RuntimeError("Error parsing config file") from exc
        for path, attr in self._TOML.items():
            node = cfg
            for key in path:
                if not isinstance(node, dict) or key not in node:
                    node = None
                    break
                node = node[key]
            if node is None:
                continue
            default = self._DEFAULTS[attr]
            if isinstance(default, bool):
                if isinstance(node, bool):
                    setattr(self, attr, node)
            elif isinstance(default, int):
                if isinstance(node, int) and not isinstance(node, bool):
                    setattr(self, attr, int(node))
            else:
                if isinstance(node, (int, float)) and not isinstance(node, bool):
                    setattr(self, attr, float(node))
        if self.pNumRobots < 1:
            raise RuntimeError("NumRobots must be greater than 0")

    def PrintParameters(self) -> None:
        for k in self._DEFAULTS:
            print(f"{k[1:]}: {getattr(self, k)}")

    def copy(self) -> "Parameters":
        q = Parameters()
        for k in self._DEFAULTS:
            setattr(q, k, getattr(self, k))
        q.config_file_ = self.config_file_
        return q

    def _c(self) -> CParams:
        c = CParams()
        for name, ctype in CParams._fields_:
            v = getattr(self, name)
            setattr(c, name, int(v) if ctype is C.c_int else float(v))
        return c


# --------------------------------------------------------------------------------------------
# WorldIDF (world_idf.h:61-279)
# --------------------------------------------------------------------------------------------
class WorldIDF:
    def __init__(self, params: Parameters, arg=None, *, device: int = 0, stream: int = 0):
        lib = _lib.load()
        self._params = params.copy()
        self._device, self._stream = device, stream
        cp = self._params._c()
        self._h = lib.ccb_idf_create(C.byref(cp), device, stream)
        if not self._h:
            raise RuntimeError(_lib.last_error())
        self._bnds: list[BivariateNormalDistribution] = []
        self._polys: list[PolygonFeature] = []
        self._host = None        # cached host copy of the map
        self._mutable = None     # host array handed out by GetWorldMapMutable
        self._factor = None      # normalisation factor inherited by a deep copy
        if isinstance(arg, (str, os.PathLike)):
            self._load_features(os.fspath(arg))
            self.GenerateMap()
        elif arg is not None:
            m = np.ascontiguousarray(arg, np.float32)
            n = self._params.pWorldMapSize
            if m.shape != (n, n):
                print("Error: World map size does not match the specified size")
                raise ValueError("World map size does not match the specified size")
            _lib.check(lib.ccb_idf_set_map(self._h, _fp(m)))

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib is not None and getattr(_lib, "load", None):  # (module globals are gone at interpreter exit)
            _lib.load().ccb_idf_destroy(h)

    # world_idf.h:130-171
    def _load_features(self, file_name: str) -> None:
        if not os.path.exists(file_name):
            print(f"Error: Could not open file {file_name}")
            sys.exit(1)
        with open(file_name) as f:
            tok = f.read().split()
        i = 0
        while i < len(tok):
            t = tok[i]
            i += 1
            if t == "BND":
                x, y, sx, sy, rho, scale = (float(v) for v in tok[i:i + 6])
                i += 6
                self.AddNormalDistribution(BivariateNormalDistribution(Point2(x, y), Point2(sx, sy), rho, scale))
            elif t == "CircularBND":
                x, y, s, scale = (float(v) for v in tok[i:i + 4])
                i += 4
                self.AddNormalDistribution(BivariateNormalDistribution(Point2(x, y), s, scale))
            elif t == "Uniform":
                n, imp = int(tok[i]), float(tok[i + 1])
                i += 2
                pts = np.array([float(v) for v in tok[i:i + 2 * n]], np.float64).reshape(n, 2)
                i += 2 * n
                self.AddUniformDistributionPolygon(PolygonFeature(pts, imp))
            else:
                print(f"Error: Unknown feature type {t}")
                sys.exit(1)

    def _sync_mutable(self) -> None:
        if self._mutable is not None:
            _lib.check(_lib.load().ccb_idf_set_map(self._h, _fp(self._mutable)))
            self._host = None

    def AddUniformDistributionPolygon(self, poly_feature: PolygonFeature) -> None:
        pts = np.ascontiguousarray(PointVector(poly_feature.poly).to_numpy())
        _lib.check(_lib.load().ccb_idf_add_polygon(self._h, _dp(pts), len(pts), float(poly_feature.imp)), ValueError)
        self._polys.append(poly_feature)

    def AddNormalDistribution(self, dist) -> None:
        dists = dist if isinstance(dist, (list, tuple)) else [dist]
        lib = _lib.load()
        for d in dists:
            m, s = d.GetMean(), d.GetSigma()
            _lib.check(lib.ccb_idf_add_bnd(self._h, m[0], m[1], s[0], s[1], d.GetRho(), d.GetScale()))
            self._bnds.append(d)

    def GenerateMap(self, mode: int = 0) -> None:
        """world_idf.h:227-241.  mode 0 = the reference CUDA kernel's arithmetic; 1 = fp64-tabulated separable form."""
        _lib.check(_lib.load().ccb_idf_generate(self._h, mode))
        self._host = None
        self._factor = None
        if self._mutable is not None:
            self._mutable[...] = self._fetch()

    def GenerateMapCuda(self) -> None:
        self.GenerateMap(0)

    def GenerateMapCPU(self) -> None:
        """There is no CPU path here: the map is generated by the same kernel as GenerateMapCuda (mode 0, the
        reference CUDA kernel's arithmetic).  The reference's own CPU and CUDA maps differ by up to its test
        tolerance (atol 1e-2, python/tests/test_parity.py:46: glibc vs CUDA erfcf, strict vs inclusive polygon
        edges); callers that compare the two get identical maps."""
        self.GenerateMap(0)

    def _fetch(self) -> np.ndarray:
        n = self._params.pWorldMapSize
        out = np.empty((n, n), np.float32)
        _lib.check(_lib.load().ccb_idf_get_map(self._h, _fp(out)))
        return out

    def GetWorldMap(self) -> np.ndarray:
        if self._mutable is not None:
            v = self._mutable.view()
            v.flags.writeable = False
            return v
        if self._host is None:
            self._host = self._fetch()
            self._host.flags.writeable = False
        return self._host

    def GetWorldMapMutable(self) -> np.ndarray:
        """Writable host view; it is re-uploaded before every device use of this WorldIDF."""
        if self._mutable is None:
            self._mutable = self._fetch()
        return self._mutable

    def GetWorldMapDevicePtr(self) -> int:
        self._sync_mutable()
        return int(_lib.load().ccb_idf_map_dev(self._h))

    def GetNormalizationFactor(self) -> float:
        if self._factor is not None:
            return self._factor
        f = C.c_float(0)
        _lib.check(_lib.load().ccb_idf_norm_factor(self._h, C.byref(f)))
        return float(f.value)

    def GetNumFeatures(self) -> int:
        return len(self._bnds) + len(self._polys)

    def PrintMapSize(self) -> None:
        n = self._params.pWorldMapSize
        print(f"World map size: {n} {n}")

    def WriteWorldMap(self, file_name: str) -> int:
        np.savetxt(file_name, self.GetWorldMap(), fmt="%.9g")
        return 0

    def LoadMap(self, file_name: str) -> None:
        if not os.path.exists(file_name):
            print(f"Error: Could not open file {file_name}")
            sys.exit(1)
        n = self._params.pWorldMapSize
        m = np.loadtxt(file_name, dtype=np.float32).reshape(n, n)
        _lib.check(_lib.load().ccb_idf_set_map(self._h, _fp(np.ascontiguousarray(m))))
        self._host = None

    # world_idf.cpp:166-195
    def WriteDistributions(self, file_name: str) -> int:
        try:
            f = open(file_name, "w")
        except OSError:
            print(f"Could not open file: {file_name}", file=sys.stderr)
            return -1
        g = lambda v: format(float(v), f".{_MAX_PRECISION}g")  # noqa: E731
        with f:
            for d in self._bnds:
                m, s = d.GetMean(), d.GetSigma()
                if s[0] == s[1]:
                    f.write("CircularBND\n")
                    f.write(f"{g(m[0])} {g(m[1])} {g(s[0])} {g(d.GetScale())}\n")
                else:
                    f.write("BND\n")
                    f.write(f"{g(m[0])} {g(m[1])} {g(s[0])} {g(s[1])} {g(d.GetRho())} {g(d.GetScale())}\n")
            for p in self._polys:
                f.write("Uniform\n")
                f.write(f"{len(p.poly)} {g(p.imp)}\n")
                for pt in p.poly:
                    f.write(f"{g(pt[0])} {g(pt[1])}\n")
        return 0

    def _clone(self) -> "WorldIDF":
        """Deep copy (world_idf.h:173-196), map copied on the device."""
        self._sync_mutable()
        other = WorldIDF(self._params, device=self._device, stream=self._stream)
        for d in self._bnds:
            other.AddNormalDistribution(d)
        for p in self._polys:
            other.AddUniformDistributionPolygon(p)
        _lib.check(_lib.load().ccb_idf_copy_map(other._h, self._h))  # device to device, no host round trip
        other._factor = self._factor
        return other


# --------------------------------------------------------------------------------------------
# CoverageSystem (coverage_system.h:66-649)
# --------------------------------------------------------------------------------------------
_rng = np.random.default_rng()


def seed(value: int | None) -> None:
    """Seed the generator behind the random-environment constructors (the reference uses random_device)."""
    global _rng
    _rng = np.random.default_rng(value)


def _random_bnds(params: Parameters, n: int) -> BNDVector:
    """coverage_system.cpp:54-81."""
    hi = params.pWorldMapSize * params.pResolution - kLargeEps
    out = BNDVector()
    for _ in range(n):
        mean = Point2(_rng.uniform(kLargeEps, hi), _rng.uniform(kLargeEps, hi))
        sigma = params.pMinSigma if params.pMinSigma == params.pMaxSigma else _rng.uniform(params.pMinSigma, params.pMaxSigma)
        scale = params.pMinPeak if params.pMinPeak == params.pMaxPeak else _rng.uniform(params.pMinPeak, params.pMaxPeak)
        out.append(BivariateNormalDistribution(mean, sigma, scale))
    return out


def _random_polygons(params: Parameters, n: int) -> list[PolygonFeature]:
    """Stand-in for GenerateRandomPolygons (polygon_utils.cpp:70-103, CGAL random_polygon_2): a
    random star-shaped simple polygon of 4..pMaxVertices vertices inside a 2*radius box."""
    out = []
    world = params.pWorldMapSize * params.pResolution
    for _ in range(n):
        k = int(_rng.integers(4, max(5, params.pMaxVertices + 1)))
        ang = np.sort(_rng.uniform(0, 2 * math.pi, k))
        rad = _rng.uniform(0.3, 1.0, k) * params.pPolygonRadius
        pts = np.stack([rad * np.cos(ang), rad * np.sin(ang)], 1)
        pts -= pts.min(0)
        ext = pts.max(0)
        base = np.array([_rng.uniform(kLargeEps, world - kLargeEps - ext[0]),
                         _rng.uniform(kLargeEps, world - kLargeEps - ext[1])])
        imp = params.pMinPeak if params.pMinPeak == params.pMaxPeak else _rng.uniform(params.pMinPeak, params.pMaxPeak) * 0.00005
        out.append(PolygonFeature(pts + base, imp))
    return out


class _EnvBase:
    """Shared plumbing of CoverageSystem (E = 1) and BatchedCoverageSystem (E >= 1)."""

    def _create(self, params: Parameters, num_envs: int, num_robots: int, device: int, stream: int):
        lib = _lib.load()
        self._params = params.copy()
        self._E, self._R = int(num_envs), int(num_robots)
        self._device, self._stream = device, stream
        cp = self._params._c()
        self._h = lib.ccb_env_create(C.byref(cp), self._E, self._R, device, stream)
        if not self._h:
            raise RuntimeError(_lib.last_error())
        self._pos_cache = None
        self._ctrls = []
        _lib.apply_test_overrides(self._h)

    def set_option(self, name: str, value: int) -> None:
        """Test / A-B hook (include/ccb200_test.h ``ccb_test_set_option``): force a kernel choice on this batch."""
        _lib.check(_lib.load().ccb_test_set_option(self._h, name.encode(), int(value)))

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib is not None and getattr(_lib, "load", None):
            _lib.load().ccb_env_destroy(h)

    def _invalidate(self):
        self._pos_cache = None

    def _positions(self) -> np.ndarray:
        if self._pos_cache is None:
            out = np.empty((self._E, self._R, 2), np.float64)
            _lib.check(_lib.load().ccb_env_get_positions(self._h, _dp(out)))
            self._pos_cache = out
        return self._pos_cache

    def voronoi_labels(self, src: int = 0) -> np.ndarray:
        """Parity-test accessor: label of every cell as assigned by the fused kernel, [E, N, N] int32."""
        n = self._params.pWorldMapSize
        out = np.empty((self._E, n, n), np.int32)
        _lib.check(_lib.load().ccb_env_voronoi_labels(self._h, int(src), out.ctypes.data_as(C.POINTER(C.c_int32))))
        return out

    def _map(self, e: int, kind: int, robot: int = 0) -> np.ndarray:
        lib = _lib.load()
        dim = lib.ccb_env_map_dim(self._h, kind)
        out = np.empty((dim, dim), np.float32)
        st = lib.ccb_env_get_map(self._h, e, kind, robot, _fp(out))
        if st != 0:
            msg = _lib.last_error()
            raise (IndexError if "index" in msg.lower() else RuntimeError)(msg)
        out.flags.writeable = False
        return out


class CoverageSystem(_EnvBase):
    """Drop-in for the reference ``CoverageSystem`` binding (core_binds.h:269-384)."""

    def __init__(self, params: Parameters, a=None, b=None, c=None, *, device: int = 0, stream: int = 0):
        params = params.copy()
        if a is None:  # (params)  coverage_system.cpp:39-41
            a, b, c = params.pNumGaussianFeatures, params.pNumPolygons, params.pNumRobots
        if isinstance(a, (int, np.integer)):
            if c is None:  # (params, num_gaussians, num_robots)
                num_g, num_p, num_r = int(a), 0, int(b)
            else:
                num_g, num_p, num_r = int(a), int(b), int(c)
            idf = WorldIDF(params, device=device, stream=stream)
            idf.AddNormalDistribution(_random_bnds(params, num_g))
            for pf in _random_polygons(params, num_p):
                idf.AddUniformDistributionPolygon(pf)
            idf.GenerateMap()
            hi = params.pRobotInitDist - kLargeEps
            positions = _rng.uniform(kLargeEps, hi, size=(num_r, 2))
        elif isinstance(a, WorldIDF):
            idf = a._clone()  # SetWorldIDF deep-copies (coverage_system.h:261-264)
            if isinstance(b, (str, os.PathLike)):
                if not os.path.exists(b):
                    print(f"Error: Could not open file {b}")
                    sys.exit(1)
                positions = np.loadtxt(b, dtype=np.float64, ndmin=2).reshape(-1, 2)
                if params.pNumRobots != len(positions):  # coverage_system.cpp:131-137
                    print("Number of robots in the file does not match the number of robots in the parameters", file=sys.stderr)
                    sys.exit(1)
            else:
                positions = PointVector(b).to_numpy()
        else:  # (params, BNDVector, PointVector)  coverage_system.cpp:157-175
            idf = WorldIDF(params, device=device, stream=stream)
            idf.AddNormalDistribution(list(a))
            idf.GenerateMap()
            positions = PointVector(b).to_numpy()
        positions = np.ascontiguousarray(positions, np.float64).reshape(-1, 2)
        self._create(params, 1, len(positions), device, stream)
        self._world_idf = idf
        lib = _lib.load()
        _lib.check(lib.ccb_env_set_world_idf(self._h, 0, idf._h))
        _lib.check(lib.ccb_env_reset(self._h, _dp(positions)))
        self._world_mutable = None

    # ---- helpers ----
    def _check_robot(self, rid: int) -> None:
        if rid < 0 or rid >= self._R:
            raise IndexError("Robot index more than the number of robots")

    def _sync_world(self) -> None:
        if self._world_mutable is not None:
            _lib.check(_lib.load().ccb_env_set_world_host(self._h, 0, _fp(self._world_mutable)))

    # ---- world ----
    def GetWorldMap(self) -> np.ndarray:
        if self._world_mutable is not None:
            v = self._world_mutable.view()
            v.flags.writeable = False
            return v
        return self._map(0, MAP_WORLD)

    def GetWorldMapMutable(self) -> np.ndarray:
        """Writable host copy; re-uploaded before every device computation that reads the world map."""
        if self._world_mutable is None:
            m = self._map(0, MAP_WORLD)
            self._world_mutable = np.array(m, np.float32)
        return self._world_mutable

    def GetWorldIDFObject(self) -> WorldIDF:
        return self._world_idf

    GetWorldIDF = GetWorldIDFObject

    def GetNormalizationFactor(self) -> float:
        return self._world_idf.GetNormalizationFactor()

    # ---- stepping (coverage_system.h:279-333) ----
    def StepActions(self, actions) -> bool:
        if isinstance(actions, np.ndarray):
            a = np.ascontiguousarray(actions, np.float64)
        else:
            a = (actions if type(actions) is PointVector else PointVector(actions)).to_numpy()
        if a.shape != (self._R, 2):
            raise ValueError("The size of the actions don't match with the number of robots")
        self._sync_world()
        bad = C.c_int(-1)
        st = _lib.load().ccb_env_step_actions(self._h, _dp(a), C.byref(bad))
        self._invalidate()
        if st == 1 and bad.value >= 0:
            print("Control incorrect", file=sys.stderr)
            return True
        _lib.check(st)
        return False

    def StepAction(self, robot_id: int, action) -> bool:
        self._check_robot(robot_id)
        a = Point2(action)
        speed = float(np.hypot(a[0], a[1]))
        n = a / speed if speed > 0 else a
        return self.StepControl(robot_id, n, speed)

    def StepControl(self, robot_id: int, direction, speed: float) -> bool:
        self._check_robot(robot_id)
        d = Point2(direction)
        self._sync_world()
        st = _lib.load().ccb_env_step_control(self._h, 0, int(robot_id), d[0], d[1], float(speed))
        self._invalidate()
        if st == 2:  # robot_model.h:208-216 throws
            print("Speed needs to be non-negative\nZero-vector direction cannot be given in control", file=sys.stderr)
            raise RuntimeError("Speed needs to be non-negative")
        _lib.check(st)
        return False

    def _set_positions(self, robot: int, local: bool, pts: np.ndarray) -> None:
        self._sync_world()
        _lib.check(_lib.load().ccb_env_set_positions(self._h, 0, robot, int(local), _dp(pts)))
        self._invalidate()

    def SetLocalRobotPositions(self, relative_pos) -> None:
        self.SetRobotPositions(relative_pos)

    def SetRobotPositions(self, positions) -> None:
        p = np.ascontiguousarray(PointVector(positions).to_numpy())
        if len(p) != self._R:
            raise ValueError("The size of the positions don't match with the number of robots")
        self._set_positions(-1, True, p)

    def SetGlobalRobotPositions(self, positions) -> None:
        p = np.ascontiguousarray(PointVector(positions).to_numpy())
        if len(p) != self._R:
            raise ValueError("The size of the positions don't match with the number of robots")
        self._set_positions(-1, False, p)

    def SetLocalRobotPosition(self, robot_id: int, pos) -> None:
        self._check_robot(robot_id)
        self._set_positions(int(robot_id), True, np.ascontiguousarray(Point2(pos)))

    def SetGlobalRobotPosition(self, robot_id: int, pos) -> None:
        self._check_robot(robot_id)
        self._set_positions(int(robot_id), False, np.ascontiguousarray(Point2(pos)))

    # ---- positions / graph ----
    def _noise(self, pts: np.ndarray) -> np.ndarray:
        """coverage_system.cpp:340-366 AddNoise."""
        n = self._params.pWorldMapSize
        out = pts + _rng.normal(0.0, self._params.pPositionsNoiseSigma, size=pts.shape)
        return np.clip(out, kLargeEps, n - kLargeEps)

    def GetRobotPositions(self, force_no_noise: bool = False) -> PointVector:
        p = self._positions()[0]
        if self._params.pAddNoisePositions and not force_no_noise:
            p = self._noise(p)
        return PointVector(p)

    def GetRobotPosition(self, robot_id: int, force_no_noise: bool = False) -> np.ndarray:
        p = self._positions()[0][robot_id].copy()
        if self._params.pAddNoisePositions and not force_no_noise:
            p = self._noise(p)
        return p

    def GetNumRobots(self) -> int:
        return self._R

    def _neighbors(self, rid: int, pos: np.ndarray):
        """coverage_system.cpp:274-291 order: ascending neighbour index."""
        rel = []
        ids = []
        for j in range(self._R):
            if j == rid:
                continue
            d = (pos[j] - pos[rid]) if j > rid else -(pos[rid] - pos[j])
            if math.sqrt(d[0] * d[0] + d[1] * d[1]) < self._params.pCommunicationRange:
                rel.append(d)
                ids.append(j)
        return rel, ids

    def GetRelativePositonsNeighbors(self, robot_id: int) -> PointVector:  # [sic], core_binds.h:293
        self._check_robot(robot_id)
        pos = self._positions()[0]
        if self._params.pAddNoisePositions:  # coverage_system.cpp:626-645
            pos = self._noise(self._noise(pos))
        return PointVector(self._neighbors(robot_id, pos)[0])

    def GetNeighborIDs(self, robot_id: int):
        self._check_robot(robot_id)
        return intVector(self._neighbors(robot_id, self._positions()[0])[1])

    def GetRobotsInCommunication(self, robot_id: int) -> PointVector:
        """coverage_system.h:566-584: neighbours sorted by distance."""
        self._check_robot(robot_id)
        rel, _ = self._neighbors(robot_id, self._positions()[0])
        rel.sort(key=lambda d: math.sqrt(d[0] * d[0] + d[1] * d[1]))
        return PointVector(rel)

    # ---- maps ----
    def GetSystemMap(self):
        return self._map(0, MAP_SYSTEM)

    def GetSystemExploredIDFMap(self):
        return self._map(0, MAP_EXPLORED_IDF)

    def GetSystemExplorationMap(self):
        return self._map(0, MAP_EXPLORATION)

    def GetRobotMap(self, rid: int):
        self._check_robot(rid)
        return self._map(0, MAP_ROBOT, rid)

    def GetRobotLocalMap(self, rid: int):
        self._check_robot(rid)
        return self._map(0, MAP_LOCAL, rid)

    def GetRobotSensorView(self, rid: int):
        self._check_robot(rid)
        return self._map(0, MAP_SENSOR_VIEW, rid)

    def GetRobotObstacleMap(self, rid: int):
        self._check_robot(rid)
        m = self._map(0, MAP_OBSTACLE, rid)
        return np.array(m)  # return_value_policy::copy

    def GetRobotExplorationMap(self, rid: int):
        self._check_robot(rid)
        return self._map(0, MAP_ROBOT_EXPLORATION, rid)

    def GetRobotSystemMap(self, rid: int):
        self._check_robot(rid)
        return self._map(0, MAP_ROBOT_SYSTEM, rid)

    def GetRobotMapMutable(self, rid: int):
        raise NotImplementedError(
            "GetRobotMapMutable: robot maps are bit-mask views of the world map on the GPU (DESIGN.md, out of scope)")

    def GetSystemExploredIDFMapMutable(self):
        raise NotImplementedError(
            "GetSystemExploredIDFMapMutable: the explored-IDF map is a bit-mask view on the GPU (DESIGN.md, out of scope)")

    def GetCommunicationMaps(self, map_size: int) -> MapTypeVector:
        """coverage_system.h:587-596 (C++ float64 / round-half-away semantics): 2R maps."""
        out = np.zeros((self._R, 2, map_size, map_size), np.float32)
        _lib.check(_lib.load().ccb_env_comm_maps_cpp(self._h, 0, int(map_size), _fp(out)))
        return MapTypeVector(out.reshape(2 * self._R, map_size, map_size))

    def ClearRobotMaps(self) -> None:
        _lib.check(_lib.load().ccb_env_clear_robot_maps(self._h, 0))

    def ClearExploredIDF(self) -> None:
        _lib.check(_lib.load().ccb_env_clear_explored_idf(self._h, 0))

    # ---- Voronoi ----
    def ComputeVoronoiCells(self) -> None:
        self._sync_world()
        _lib.check(_lib.load().ccb_env_compute_voronoi(self._h))

    def GetVoronoiCells(self) -> VoronoiCells:
        self._sync_world()
        R, V = self._R, _lib.CCB_MAX_CELL_VERTS
        mass = np.empty(R)
        cen = np.empty((R, 2))
        site = np.empty((R, 2))
        obj = np.empty(R)
        nv = np.empty(R, np.int32)
        verts = np.empty((R, V, 2))
        _lib.check(_lib.load().ccb_env_get_cells(self._h, 0, _dp(mass), _dp(cen), _dp(site), _dp(obj),
                                                 nv.ctypes.data_as(c_int_p), _dp(verts)))
        return VoronoiCells(
            VoronoiCell(PointVector(verts[r, : nv[r]]), float(mass[r]), cen[r].copy(), site[r].copy(), float(obj[r]))
            for r in range(R))

    def GetVoronoiCell(self, robot_id: int) -> VoronoiCell:
        self._check_robot(robot_id)
        return self.GetVoronoiCells()[robot_id]

    def GetObjectiveValue(self) -> float:
        self._sync_world()
        out = np.empty(1)
        _lib.check(_lib.load().ccb_env_objective(self._h, _dp(out)))
        return float(out[0])

    def GetRobotVoronoiFeatures(self) -> DblVectorVector:
        out = np.empty((1, self._R, 7))
        _lib.check(_lib.load().ccb_env_voronoi_features(self._h, _dp(out)))
        return DblVectorVector(DblVector(row) for row in out[0])

    def GetLocalVoronoiFeatures(self, robot_id=None):
        """coverage_system.cpp:648-676 / coverage_system.h:624-633: one robot (DblVector) or all of them."""
        out = np.empty((1, self._R, 7))
        _lib.check(_lib.load().ccb_env_local_voronoi_features(self._h, _dp(out)))
        if robot_id is None:
            return DblVectorVector(DblVector(row) for row in out[0])
        self._check_robot(robot_id)
        return DblVector(out[0, robot_id])

    def GetRobotExplorationFeatures(self) -> DblVectorVector:
        """coverage_system.h:603-610 -> robot_model.h:347-403: per robot the pNumFrontiers best frontier
        points (worst first) as [dx, dy, dx, dy, ...] relative to the robot, zero-padded."""
        K = int(self._params.pNumFrontiers)
        out = np.zeros((1, self._R, 2 * K))
        _lib.check(_lib.load().ccb_env_exploration_features(self._h, K, _dp(out)))
        return DblVectorVector(DblVector(row) for row in out[0])

    def CheckOscillation(self, robot_id: int) -> bool:
        self._check_robot(robot_id)
        f = C.c_int(0)
        _lib.check(_lib.load().ccb_env_check_oscillation(self._h, 0, int(robot_id), C.byref(f)))
        return bool(f.value)

    def _ratios(self):
        out = np.empty((1, 2))
        _lib.check(_lib.load().ccb_env_exploration_ratios(self._h, _dp(out)))
        return out[0]

    def GetExplorationRatio(self) -> float:
        return float(self._ratios()[0])

    def GetWeightedExplorationRatio(self) -> float:
        return float(self._ratios()[1])

    # ---- I/O (coverage_system.cpp:368-403) ----
    def WriteRobotPositions(self, file_name: str, positions=None) -> int:
        pts = self._positions()[0] if positions is None else PointVector(positions).to_numpy()
        try:
            with open(file_name, "w") as f:
                for p in pts:
                    f.write(f"{format(p[0], f'.{_MAX_PRECISION}g')} {format(p[1], f'.{_MAX_PRECISION}g')}\n")
        except OSError:
            print(f"[Error] Could not open {file_name} for writing.", file=sys.stderr)
            return 1
        return 0

    def WriteEnvironment(self, pos_filename: str, env_filename: str) -> int:
        self.WriteRobotPositions(pos_filename)
        self._world_idf.WriteDistributions(env_filename)
        return 0

    # plotting is out of scope; keep the names so callers fail with a clear message
    _plot_warned = False

    def _no_plot(self, *a, **k):
        """The reference plots through gnuplot-iostream; plotting is outside the hot-path scope: no-ops (with
        one warning) so that evaluators run with their plot switches on."""
        if not CoverageSystem._plot_warned:
            CoverageSystem._plot_warned = True
            print("[coveragecontrol_b200] Plot*/RecordPlotData/RenderRecordedMap are no-ops (plotting is out of scope)", file=sys.stderr)
        return None

    PlotSystemMap = PlotWorldMap = PlotInitMap = PlotWorldMapRobots = PlotMapVoronoi = _no_plot
    PlotRobotLocalMap = PlotRobotExplorationMap = PlotRobotSensorView = PlotRobotSystemMap = _no_plot
    PlotRobotObstacleMap = PlotRobotCommunicationMaps = RecordPlotData = RenderRecordedMap = _no_plot

    # ---- device-side accessors (zero-copy hand-off to torch) ----
    def device_handle(self) -> int:
        return int(self._h)


# --------------------------------------------------------------------------------------------
# CVT controllers (clairvoyant_cvt.h, centralized_cvt.h, decentralized_cvt.h)
# --------------------------------------------------------------------------------------------
class _CVT:
    _KIND = 0

    def __init__(self, params: Parameters, a, b=None):
        env = b if b is not None else a  # (params, env) or (params, num_robots, env)
        if not isinstance(env, _EnvBase):
            raise TypeError("expected a CoverageSystem")
        self._env = env  # the reference holds a bare reference (no keep_alive); keeping it alive is harmless
        if isinstance(env, CoverageSystem):
            env._sync_world()
        self._h = _lib.load().ccb_ctrl_create(env._h, self._KIND)
        if not self._h:
            raise RuntimeError(_lib.last_error())

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and getattr(self._env, "_h", None) and _lib is not None and getattr(_lib, "load", None):
            _lib.load().ccb_ctrl_destroy(h)

    def ComputeActions(self) -> int:
        if isinstance(self._env, CoverageSystem):
            self._env._sync_world()
        _lib.check(_lib.load().ccb_ctrl_compute_actions(self._h))
        return 0

    def _fetch(self, fn) -> np.ndarray:
        out = np.empty((self._env._E, self._env._R, 2))
        _lib.check(fn(self._h, _dp(out)))
        return out

    def GetActions(self) -> PointVector:
        return PointVector(self._fetch(_lib.load().ccb_ctrl_get_actions)[0])

    def GetGoals(self) -> PointVector:
        return PointVector(self._fetch(_lib.load().ccb_ctrl_get_goals)[0])

    def IsConverged(self) -> bool:
        out = np.empty(self._env._E, np.int32)
        _lib.check(_lib.load().ccb_ctrl_is_converged(self._h, out.ctypes.data_as(c_int_p)))
        return bool(out[0])

    def GetVoronoi(self):
        raise NotImplementedError("the Voronoi class itself is not bound in the reference either (core_binds.h:187)")


class ClairvoyantCVT(_CVT):
    _KIND = 0


class CentralizedCVT(_CVT):
    _KIND = 1


class DecentralizedCVT(_CVT):
    _KIND = 2


class NearOptimalCVT:
    """Drop-in for the reference ``NearOptimalCVT`` (algorithms/near_optimal_cvt.h:60-118,
    near_optimal_cvt_algorithm.h:57-146): ``pLloydNumTries`` random site sets are relaxed with Lloyd's
    algorithm on the world map (a site moves to its cell's centroid while it is at least one resolution
    away, at most ``pLloydMaxIterations`` rounds), one relaxed set is chosen and the robots are assigned to
    its centroids with the Hungarian method (here: scipy's ``linear_sum_assignment``, same optimum).

    All tries advance together: they are the environments of one scratch batch over the same world map, so
    a Lloyd round is one batched whole-map Voronoi pass.  The reference chooses the try with the smallest
    ``GetSumIDFGoalDistSqr()``; that sum is never accumulated there (``GoalObjFunctional`` is commented out,
    voronoi.cpp:162), so every try scores 0 and the first one wins -- ``select="reference"`` (default)
    reproduces that, ``select="objective"`` takes the try with the smallest coverage objective instead.
    """

    def __init__(self, params: Parameters, a, b=None, *, select: str = "reference", initial_sites=None):
        env = b if b is not None else a  # (params, env) or (params, num_robots, env)
        if not isinstance(env, CoverageSystem):
            raise TypeError("expected a CoverageSystem")
        if select not in ("reference", "objective"):
            raise ValueError("select must be 'reference' or 'objective'")
        self._params = params.copy()
        self._env = env
        self._R = env._R if b is None else int(a)
        self._select = select
        self._initial_sites = None if initial_sites is None else np.ascontiguousarray(initial_sites, np.float64)
        self._positions = env._positions()[0].copy()
        self._actions = np.zeros((self._R, 2))
        self._goals = self._positions.copy()
        self._converged = False
        self._scratch = None
        self.lloyd_iterations = 0
        self.ComputeGoals()

    def _lloyd(self, sites: np.ndarray):
        """near_optimal_cvt_algorithm.h:75-101 for all tries at once; returns the centroids [tries, R, 2]
        of the relaxed sites and the coverage objective of every try."""
        p, lib = self._params, _lib.load()
        tries, R = sites.shape[0], sites.shape[1]
        if self._scratch is None or self._scratch._E != tries:
            sp = p.copy()
            sp.pNumRobots = R
            sp.pUpdateExplorationMap = False
            self._scratch = BatchedCoverageSystem(sp, tries, R, device=self._env._device)
        sc = self._scratch
        self._env._sync_world()
        for t in range(tries):
            _lib.check(lib.ccb_env_set_world_dev(sc._h, t, lib.ccb_env_world_dev(self._env._h)))
        active = np.ones(tries, bool)
        cen = np.empty((tries, R, 2))
        obj = np.empty((tries, R))
        sites = sites.copy()
        it = 0
        while True:
            sc.reset(sites)
            _lib.check(lib.ccb_env_get_cells_all(sc._h, None, _dp(cen), _dp(obj)))
            if it >= p.pLloydMaxIterations or not active.any():
                break
            far = np.linalg.norm(cen - sites, axis=2) >= p.pResolution   # sites that still move
            far &= active[:, None]
            active = far.any(axis=1)
            if not active.any():
                break
            sites[far] = cen[far]
            it += 1
        self.lloyd_iterations = it
        return cen, obj.sum(axis=1)

    def ComputeGoals(self) -> None:
        p = self._params
        R = self._R
        if self._initial_sites is not None:
            sites = self._initial_sites.reshape(-1, R, 2)
        else:
            hi = p.pWorldMapSize * p.pResolution - 0.001   # near_optimal_cvt_algorithm.h:72
            sites = _rng.uniform(0.001, hi, size=(max(1, p.pLloydNumTries), R, 2))
        cen, objective = self._lloyd(sites)
        best = 0 if self._select == "reference" else int(np.argmin(objective))
        self.best_try, self.try_objectives = best, objective
        centroids = cen[best]
        from scipy.optimize import linear_sum_assignment

        cost = np.linalg.norm(self._positions[:, None, :] - centroids[None, :, :], axis=2)  # :121-127
        rows, cols = linear_sum_assignment(cost)
        goals = np.empty((R, 2))
        goals[rows] = centroids[cols]
        self._goals = goals

    def ComputeActions(self) -> int:
        """near_optimal_cvt.h:93-113."""
        p = self._params
        self._converged = True
        self._positions = self._env._positions()[0].copy()
        for i in range(self._R):
            self._actions[i] = 0.0
            diff = self._goals[i] - self._positions[i]
            dist = float(np.sqrt(diff[0] * diff[0] + diff[1] * diff[1]))
            if dist < kEps:
                continue
            speed = min(p.pMaxRobotSpeed, dist / p.pTimeStep)
            self._actions[i] = speed * (diff / dist)
            self._converged = False
        return 0

    def GetActions(self) -> PointVector:
        return PointVector(self._actions)

    def GetGoals(self) -> PointVector:
        return PointVector(self._goals)

    def IsConverged(self) -> bool:
        return self._converged

    def GetVoronoi(self):
        raise NotImplementedError("the Voronoi class itself is not bound in the reference either (core_binds.h:187)")


# --------------------------------------------------------------------------------------------
# RobotModel (robot_model.h:61-432, core_binds.h:141-157): a single robot over a 1-robot environment
# --------------------------------------------------------------------------------------------
class RobotModel:
    def __init__(self, params: Parameters, global_start_position, world_idf: WorldIDF):
        p = params.copy()
        p.pNumRobots = 1
        self._start = Point2(global_start_position)
        self._env = CoverageSystem(p, world_idf, PointVector([self._start]))

    def StepControl(self, direction, speed: float) -> int:
        self._env.StepControl(0, direction, speed)
        return 0

    def SetRobotPosition(self, pos) -> None:
        self._env.SetLocalRobotPosition(0, pos)

    def SetGlobalRobotPosition(self, pos) -> None:
        self._env.SetGlobalRobotPosition(0, pos)

    def GetGlobalStartPosition(self):
        return self._start.copy()

    def GetGlobalCurrentPosition(self):
        return self._env.GetRobotPosition(0, True)

    def GetRobotLocalMap(self):
        return self._env.GetRobotLocalMap(0)

    def GetRobotSystemMap(self):
        return self._env.GetRobotSystemMap(0)

    def GetSensorView(self):
        return self._env.GetRobotSensorView(0)

    def GetExplorationMap(self):
        return self._env.GetRobotExplorationMap(0)

    def GetObstacleMap(self):
        return self._env.GetRobotObstacleMap(0)


# --------------------------------------------------------------------------------------------
# Batched environments (data generation): E independent CoverageSystems in one object
# --------------------------------------------------------------------------------------------
class BatchedCoverageSystem(_EnvBase):
    """E independent environments with identical Parameters, stepped together on one GPU.

    This is the device-resident form of the reference's sequential data-generation loop
    (python/data_generation/data_generation.py:217-259): no per-environment host round trip.
    """

    def __init__(self, params: Parameters, num_envs: int, num_robots: int | None = None, *, device: int = 0,
                 stream: int = 0):
        self._create(params, num_envs, num_robots or params.pNumRobots, device, stream)

    @property
    def num_envs(self) -> int:
        return self._E

    @property
    def num_robots(self) -> int:
        return self._R

    def set_world(self, e: int, world) -> None:
        lib = _lib.load()
        if isinstance(world, WorldIDF):
            world._sync_mutable()
            _lib.check(lib.ccb_env_set_world_idf(self._h, e, world._h))
        else:
            m = np.ascontiguousarray(world, np.float32)
            n = self._params.pWorldMapSize
            if m.shape != (n, n):
                raise ValueError("World map size does not match the specified size")
            _lib.check(lib.ccb_env_set_world_host(self._h, e, _fp(m)))

    def set_world_device(self, e: int, dev_ptr: int) -> None:
        _lib.check(_lib.load().ccb_env_set_world_dev(self._h, e, dev_ptr))

    def reset(self, positions) -> None:
        p = np.ascontiguousarray(positions, np.float64)
        if p.shape != (self._E, self._R, 2):
            raise ValueError(f"positions must have shape ({self._E}, {self._R}, 2)")
        _lib.check(_lib.load().ccb_env_reset(self._h, _dp(p)))
        self._invalidate()

    def reset_env(self, e: int, positions) -> None:
        """A fresh episode in slot ``e`` (set its world first with :meth:`set_world`): positions ``[R, 2]``."""
        p = np.ascontiguousarray(positions, np.float64)
        if p.shape != (self._R, 2):
            raise ValueError(f"positions must have shape ({self._R}, 2)")
        _lib.check(_lib.load().ccb_env_reset_env(self._h, int(e), _dp(p)))
        self._invalidate()

    def step_actions_device(self, actions_dev_ptr: int) -> None:
        """StepActions with the actions already on the device (``E x R x 2`` float64)."""
        _lib.check(_lib.load().ccb_env_step_actions_dev(self._h, actions_dev_ptr))
        self._invalidate()

    def step_actions(self, actions) -> None:
        a = np.ascontiguousarray(actions, np.float64)
        if a.shape != (self._E, self._R, 2):
            raise ValueError(f"actions must have shape ({self._E}, {self._R}, 2)")
        bad = C.c_int(-1)
        _lib.check(_lib.load().ccb_env_step_actions(self._h, _dp(a), C.byref(bad)))
        self._invalidate()

    def positions(self) -> np.ndarray:
        return self._positions().copy()

    def objective(self) -> np.ndarray:
        out = np.empty(self._E)
        _lib.check(_lib.load().ccb_env_objective(self._h, _dp(out)))
        return out

    def get_map(self, e: int, kind: int, robot: int = 0) -> np.ndarray:
        return self._map(e, kind, robot)

    def cells(self, e: int):
        R = self._R
        mass, cen, site, obj = np.empty(R), np.empty((R, 2)), np.empty((R, 2)), np.empty(R)
        _lib.check(_lib.load().ccb_env_get_cells(self._h, e, _dp(mass), _dp(cen), _dp(site), _dp(obj), None, None))
        return mass, cen, site, obj

    def voronoi_features(self) -> np.ndarray:
        out = np.empty((self._E, self._R, 7))
        _lib.check(_lib.load().ccb_env_voronoi_features(self._h, _dp(out)))
        return out

    def exploration_features(self, num_frontiers: int | None = None) -> np.ndarray:
        """GetRobotExplorationFeatures of every environment: [E, R, 2 * num_frontiers]."""
        K = int(self._params.pNumFrontiers if num_frontiers is None else num_frontiers)
        out = np.zeros((self._E, self._R, 2 * K))
        _lib.check(_lib.load().ccb_env_exploration_features(self._h, K, _dp(out)))
        return out

    def sync(self) -> None:
        _lib.check(_lib.load().ccb_env_sync(self._h))

    def world_device_ptr(self) -> int:
        return int(_lib.load().ccb_env_world_dev(self._h))

    def positions_device_ptr(self) -> int:
        return int(_lib.load().ccb_env_positions_dev(self._h))


class BatchedCVT(_CVT):
    """A CVT controller over every environment of a BatchedCoverageSystem."""

    def __init__(self, env: BatchedCoverageSystem, kind: int):
        self._KIND = int(kind)
        super().__init__(env._params, env)

    def compute_actions(self) -> None:
        _lib.check(_lib.load().ccb_ctrl_compute_actions(self._h))

    def actions(self) -> np.ndarray:
        return self._fetch(_lib.load().ccb_ctrl_get_actions)

    def goals(self) -> np.ndarray:
        return self._fetch(_lib.load().ccb_ctrl_get_goals)

    def converged(self) -> np.ndarray:
        out = np.empty(self._env._E, np.int32)
        _lib.check(_lib.load().ccb_ctrl_is_converged(self._h, out.ctypes.data_as(c_int_p)))
        return out.astype(bool)

    def rollout(self, steps: int, with_objective: bool = False, objective_dev_ptr: int = 0) -> None:
        """``steps`` x (ComputeActions; StepActions) on the device, no host round trip."""
        _lib.check(_lib.load().ccb_ctrl_rollout(self._h, int(steps), int(with_objective), objective_dev_ptr or None))
        self._env._invalidate()

    def rollout_obs(self, steps: int, with_objective: bool, objective_dev_ptr: int, cnn_size: int, use_comm: bool,
                    obs_dev_ptr: int, obs_step_stride: int = 0) -> None:
        """``rollout`` that also writes the LPAC observation (coverage_env_utils.py:227-277 get_maps) of every step
        to ``obs_dev_ptr + s * obs_step_stride`` floats; objective chain and observation kernels overlap."""
        _lib.check(_lib.load().ccb_ctrl_rollout_obs(self._h, int(steps), int(with_objective), objective_dev_ptr or None,
                                                    int(cnn_size), int(use_comm), obs_dev_ptr or None, int(obs_step_stride)))
        self._env._invalidate()

    def actions_device_ptr(self) -> int:
        return int(_lib.load().ccb_ctrl_actions_dev(self._h))
