"""ctypes front-end of the CPU oracle (oracle/cc_oracle.c).

TEST INFRASTRUCTURE ONLY.  Imported by ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``; never by
the product package ``coveragecontrol_b200``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle.so")
_REF = os.path.join(_HERE, "_ref", "libref_idf.so")


def build(verbose: bool = False) -> None:
    """Compile liboracle.so (and oracle/_ref when /root/reference is present)."""
    out = subprocess.run(["make", "-C", _HERE, "all"], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    if verbose:
        print(out.stdout)


class Params(C.Structure):
    """Mirror of ``orc_params`` (reference parameters.h:49-175)."""

    _fields_ = [
        ("pNumRobots", C.c_int),
        ("pResolution", C.c_double),
        ("pWorldMapSize", C.c_int),
        ("pRobotMapSize", C.c_int),
        ("pLocalMapSize", C.c_int),
        ("pUpdateRobotMap", C.c_int),
        ("pUpdateExplorationMap", C.c_int),
        ("pUpdateSensorView", C.c_int),
        ("pUpdateSystemMap", C.c_int),
        ("pTruncationBND", C.c_double),
        ("pNorm", C.c_double),
        ("pUnknownImportance", C.c_double),
        ("pRobotMapUseUnknownImportance", C.c_int),
        ("pSensorSize", C.c_int),
        ("pCommunicationRange", C.c_double),
        ("pMaxRobotSpeed", C.c_double),
        ("pRobotPosHistorySize", C.c_int),
        ("pTimeStep", C.c_double),
        ("pCheckOscillations", C.c_int),
    ]


class Bounds(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("left", "right", "bottom", "top", "width", "height")]


class Cell(C.Structure):
    _fields_ = [
        ("site", C.c_double * 2),
        ("mass", C.c_double),
        ("centroid", C.c_double * 2),
        ("sum_idf_site_dist", C.c_double),
        ("sum_idf_site_dist_sqr", C.c_double),
        ("nverts", C.c_int),
        ("verts", C.c_double * 128),
    ]

    def polygon(self) -> np.ndarray:
        return np.array(self.verts[: 2 * self.nverts], dtype=np.float64).reshape(-1, 2)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        _lib = C.CDLL(_LIB)
        L = _lib
        dp, fp, ip = C.POINTER(C.c_double), C.POINTER(C.c_float), C.POINTER(C.c_int)
        L.orc_params_default.argtypes = [C.POINTER(Params)]
        L.orc_idf_generate.restype = C.c_float
        L.orc_idf_generate.argtypes = [C.POINTER(Params), dp, C.c_int, dp, dp, ip, fp, C.c_int, C.c_int, fp]
        L.orc_idf_fp64.argtypes = [C.POINTER(Params), dp, C.c_int, C.c_double, dp, C.c_char_p]
        L.orc_point_in_monotone_polygon.argtypes = [fp, fp, C.c_int, C.c_float, C.c_float]
        L.orc_compute_offsets.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int, C.c_int,
                                          C.POINTER(Bounds), C.POINTER(Bounds)]
        L.orc_voronoi.argtypes = [dp, C.c_int, fp, C.c_int, C.c_double, C.c_double, C.c_double,
                                  C.c_int, C.c_int, C.POINTER(Cell)]
        L.orc_voronoi_bruteforce.argtypes = [dp, C.c_int, fp, C.c_int, C.c_int, C.c_int, C.c_double, ip, dp]
        L.orc_voronoi_labels_omp.argtypes = [dp, C.c_int, C.c_int, C.c_int, C.c_double, ip]
        L.orc_voronoi_labels_omp.restype = None
        L.orc_clip_cell.argtypes = [dp, C.c_int, C.c_int, C.c_double, C.c_double, dp, C.c_int]
        L.orc_env_create.restype = C.c_void_p
        L.orc_env_create.argtypes = [C.POINTER(Params), fp, dp, C.c_int]
        L.orc_env_destroy.argtypes = [C.c_void_p]
        L.orc_env_step_actions.argtypes = [C.c_void_p, dp]
        L.orc_env_set_global_positions.argtypes = [C.c_void_p, dp]
        L.orc_env_clear_robot_maps.argtypes = [C.c_void_p]
        L.orc_env_clear_robot_maps.restype = None
        L.orc_env_clear_explored_idf.argtypes = [C.c_void_p]
        L.orc_env_clear_explored_idf.restype = None
        L.orc_env_get_positions.argtypes = [C.c_void_p, dp]
        L.orc_env_objective.restype = C.c_double
        L.orc_env_objective.argtypes = [C.c_void_p]
        L.orc_env_compute_voronoi.argtypes = [C.c_void_p, C.POINTER(Cell)]
        for name in ("world_map", "system_map", "explored_map", "exploration_map"):
            f = getattr(L, "orc_env_" + name)
            f.restype = fp
            f.argtypes = [C.c_void_p]
        for name in ("robot_map", "sensor_view"):
            f = getattr(L, "orc_env_" + name)
            f.restype = fp
            f.argtypes = [C.c_void_p, C.c_int]
        L.orc_env_local_map.argtypes = [C.c_void_p, C.c_int, fp]
        L.orc_env_obstacle_map.argtypes = [C.c_void_p, C.c_int, fp]
        L.orc_env_comm_maps_cpp.argtypes = [C.c_void_p, C.c_int, C.c_int, fp, fp]
        L.orc_env_comm_maps_py.argtypes = [C.c_void_p, C.c_int, C.c_int, fp, fp]
        L.orc_env_neighbors.argtypes = [C.c_void_p, C.c_int, dp]
        L.orc_env_voronoi_features.argtypes = [C.c_void_p, C.c_int, dp]
        L.orc_env_exploration_features.argtypes = [C.c_void_p, C.c_int, C.c_int, dp]
        L.orc_env_exploration_features.restype = None
        L.orc_frontier_select.argtypes = [dp, C.c_int, C.c_int, C.POINTER(C.c_int)]
        L.orc_frontier_select.restype = C.c_int
        L.orc_env_check_oscillation.argtypes = [C.c_void_p, C.c_int]
        L.orc_ctrl_compute_actions.argtypes = [C.c_void_p, C.c_int, dp, dp]
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def default_params(**kw) -> Params:
    p = Params()
    lib().orc_params_default(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


def params_from(obj) -> Params:
    """Build oracle params from any object with the reference's ``p*`` attributes."""
    p = default_params()
    for name, _ in Params._fields_:
        if hasattr(obj, name):
            setattr(p, name, type(getattr(p, name))(getattr(obj, name)))
    return p


def _pack_polys(polys):
    """polys: list of (ndarray[n,2], imp)."""
    if not polys:
        z = np.zeros(1, np.float64)
        return z, z.copy(), np.zeros(1, np.int32), np.zeros(1, np.float32), 0
    px = np.concatenate([np.asarray(p, np.float64)[:, 0] for p, _ in polys])
    py = np.concatenate([np.asarray(p, np.float64)[:, 1] for p, _ in polys])
    sz = np.array([len(p) for p, _ in polys], np.int32)
    imp = np.array([i for _, i in polys], np.float32)
    return np.ascontiguousarray(px), np.ascontiguousarray(py), sz, imp, len(polys)


def idf_generate(p: Params, bnd: np.ndarray, polys=None, mode: int = 1):
    """bnd: [F,6] {mean_x, mean_y, sigma_x, sigma_y, scale, rho}.  Returns (map, factor)."""
    N = p.pWorldMapSize
    bnd = np.ascontiguousarray(bnd, np.float64).reshape(-1, 6)
    px, py, sz, imp, npoly = _pack_polys(polys or [])
    out = np.zeros((N, N), np.float32)
    f = lib().orc_idf_generate(C.byref(p), _dp(bnd), len(bnd), _dp(px), _dp(py), _ip(sz), _fp(imp),
                               npoly, mode, _fp(out))
    return out, float(f)


def idf_fp64(p: Params, bnd: np.ndarray, band: float = 0.0):
    N = p.pWorldMapSize
    bnd = np.ascontiguousarray(bnd, np.float64).reshape(-1, 6)
    out = np.zeros((N, N), np.float64)
    border = np.zeros((N, N), np.uint8)
    lib().orc_idf_fp64(C.byref(p), _dp(bnd), len(bnd), band, _dp(out), border.ctypes.data_as(C.c_char_p))
    return out, border


def compute_offsets(res, px, py, submap, map_size):
    ix, of = Bounds(), Bounds()
    lib().orc_compute_offsets(res, px, py, submap, map_size, C.byref(ix), C.byref(of))
    return ix, of


def voronoi(sites, map2d, res=1.0, single=False, robot_id=0, map_size=None):
    sites = np.ascontiguousarray(sites, np.float64).reshape(-1, 2)
    m = np.asarray(map2d, np.float32)
    assert m.strides[1] == 4
    w, h = map_size if map_size is not None else m.shape
    n = len(sites)
    cells = (Cell * (1 if single else n))()
    err = lib().orc_voronoi(_dp(sites), n, _fp(m), m.strides[0] // 4, float(w), float(h), res,
                            int(single), robot_id, cells)
    if err:
        raise RuntimeError(f"orc_voronoi failed ({err})")
    return list(cells)


def voronoi_bruteforce(sites, map2d, res=1.0):
    sites = np.ascontiguousarray(sites, np.float64).reshape(-1, 2)
    m = np.asarray(map2d, np.float32)
    W, H = m.shape
    labels = np.zeros((W, H), np.int32)
    sums = np.zeros((len(sites), 4), np.float64)
    lib().orc_voronoi_bruteforce(_dp(sites), len(sites), _fp(m), m.strides[0] // 4, W, H, res,
                                 _ip(labels), _dp(sums))
    return labels, sums


def voronoi_labels(sites, W: int, H: int, res=1.0):
    """Labels of a W x H grid under the stated rule, multi-threaded (large maps); == voronoi_bruteforce's labels."""
    sites = np.ascontiguousarray(sites, np.float64).reshape(-1, 2)
    labels = np.zeros((W, H), np.int32)
    lib().orc_voronoi_labels_omp(_dp(sites), len(sites), int(W), int(H), float(res), _ip(labels))
    return labels


def moments_from_labels(labels, map2d, sites, res=1.0, chunk_rows: int = 256):
    """[R, 4] = (sum v, sum cx v, sum cy v, sum |c - site|^2 v) per label, fp64, accumulated over chunks of rows
    (NumPy summation order inside a chunk; memory stays bounded for 16384^2 maps)."""
    sites = np.ascontiguousarray(sites, np.float64).reshape(-1, 2)
    R = len(sites)
    W, H = labels.shape
    out = np.zeros((R, 4))
    cyr = (np.arange(H) + 0.5) * res
    for r0 in range(0, W, chunk_rows):
        r1 = min(W, r0 + chunk_rows)
        lab = labels[r0:r1].ravel()
        v = np.asarray(map2d[r0:r1], np.float64).ravel()
        cx = np.repeat((np.arange(r0, r1) + 0.5) * res, H)
        cy = np.tile(cyr, r1 - r0)
        out[:, 0] += np.bincount(lab, v, R)
        out[:, 1] += np.bincount(lab, v * cx, R)
        out[:, 2] += np.bincount(lab, v * cy, R)
        d2 = (cx - sites[lab, 0]) ** 2 + (cy - sites[lab, 1]) ** 2
        out[:, 3] += np.bincount(lab, v * d2, R)
    return out


def clip_cell(sites, k, w, h):
    sites = np.ascontiguousarray(sites, np.float64).reshape(-1, 2)
    v = np.zeros(256, np.float64)
    n = lib().orc_clip_cell(_dp(sites), len(sites), k, float(w), float(h), _dp(v), 128)
    return v[: 2 * n].reshape(-1, 2).copy()


class Env:
    """Structure-faithful CPU environment (dense maps, as the reference keeps them)."""

    def __init__(self, p: Params, world_map: np.ndarray, positions: np.ndarray):
        self.p = p
        self.N, self.S, self.L = p.pWorldMapSize, p.pSensorSize, p.pLocalMapSize
        pos = np.ascontiguousarray(positions, np.float64).reshape(-1, 2)
        self.R = len(pos)
        wm = np.ascontiguousarray(world_map, np.float32)
        assert wm.shape == (self.N, self.N)
        self._h = lib().orc_env_create(C.byref(p), _fp(wm), _dp(pos), self.R)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_env_destroy(self._h)
            self._h = None

    def _view(self, ptr, shape):
        return np.ctypeslib.as_array(ptr, shape=shape)

    def step_actions(self, actions) -> int:
        a = np.ascontiguousarray(actions, np.float64).reshape(-1, 2)
        return lib().orc_env_step_actions(self._h, _dp(a))

    def set_global_positions(self, pos):
        a = np.ascontiguousarray(pos, np.float64).reshape(-1, 2)
        lib().orc_env_set_global_positions(self._h, _dp(a))

    def clear_robot_maps(self):
        lib().orc_env_clear_robot_maps(self._h)

    def clear_explored_idf(self):
        lib().orc_env_clear_explored_idf(self._h)

    def positions(self) -> np.ndarray:
        out = np.zeros((self.R, 2), np.float64)
        lib().orc_env_get_positions(self._h, _dp(out))
        return out

    def objective(self) -> float:
        return lib().orc_env_objective(self._h)

    def voronoi_cells(self):
        cells = (Cell * self.R)()
        lib().orc_env_compute_voronoi(self._h, cells)
        return list(cells)

    def world_map(self):
        return self._view(lib().orc_env_world_map(self._h), (self.N, self.N))

    def system_map(self):
        return self._view(lib().orc_env_system_map(self._h), (self.N, self.N))

    def explored_map(self):
        return self._view(lib().orc_env_explored_map(self._h), (self.N, self.N))

    def exploration_map(self):
        return self._view(lib().orc_env_exploration_map(self._h), (self.N, self.N))

    def robot_map(self, r):
        return self._view(lib().orc_env_robot_map(self._h, r), (self.N, self.N))

    def sensor_view(self, r):
        return self._view(lib().orc_env_sensor_view(self._h, r), (self.S, self.S))

    def local_map(self, r):
        out = np.zeros((self.L, self.L), np.float32)
        lib().orc_env_local_map(self._h, r, _fp(out))
        return out

    def obstacle_map(self, r):
        out = np.zeros((self.L, self.L), np.float32)
        lib().orc_env_obstacle_map(self._h, r, _fp(out))
        return out

    def comm_maps(self, r, size, semantics="py"):
        ox = np.zeros((size, size), np.float32)
        oy = np.zeros((size, size), np.float32)
        fn = lib().orc_env_comm_maps_py if semantics == "py" else lib().orc_env_comm_maps_cpp
        fn(self._h, r, size, _fp(ox), _fp(oy))
        return ox, oy

    def neighbors(self, r):
        rel = np.zeros((self.R, 2), np.float64)
        n = lib().orc_env_neighbors(self._h, r, _dp(rel))
        return rel[:n].copy()

    def exploration_features(self, r, num_frontiers):
        out = np.zeros(2 * num_frontiers)
        lib().orc_env_exploration_features(self._h, r, num_frontiers, _dp(out))
        return out

    def voronoi_features(self, r):
        out = np.zeros(7, np.float64)
        lib().orc_env_voronoi_features(self._h, r, _dp(out))
        return out

    def compute_actions(self, kind: int):
        """kind: 0 clairvoyant, 1 centralized, 2 decentralized -> (converged, actions, goals)."""
        a = np.zeros((self.R, 2), np.float64)
        g = np.zeros((self.R, 2), np.float64)
        c = lib().orc_ctrl_compute_actions(self._h, kind, _dp(a), _dp(g))
        if c < 0:
            raise RuntimeError("controller failed (goal out of bounds / degenerate cell)")
        return bool(c), a, g


# ---- the reference's own CUDA IDF kernel (oracle/_ref, GPU box only) ----------
_ref = None


def ref_idf_available() -> bool:
    return os.path.exists(_REF)


def ref_idf_generate(p: Params, bnd: np.ndarray, polys=None):
    """Run the REFERENCE generate_world_map_cuda() (needs a GPU)."""
    global _ref
    if _ref is None:
        _ref = C.CDLL(_REF)
        fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int)
        _ref.ref_idf_generate.argtypes = [fp, C.c_int, fp, fp, ip, fp, C.c_int, C.c_int, C.c_float,
                                          C.c_float, C.c_float, fp, fp]
    N = p.pWorldMapSize
    b = np.ascontiguousarray(bnd, np.float64).reshape(-1, 6).astype(np.float32)
    px, py, sz, imp, npoly = _pack_polys(polys or [])
    px, py = px.astype(np.float32), py.astype(np.float32)
    out = np.zeros((N, N), np.float32)
    f = C.c_float(0)
    err = _ref.ref_idf_generate(_fp(b), len(b), _fp(px), _fp(py), _ip(sz), _fp(imp), npoly, N,
                                float(p.pResolution), float(p.pTruncationBND), float(p.pNorm),
                                _fp(out), C.byref(f))
    if err:
        raise RuntimeError("reference generate_world_map_cuda failed")
    return out, float(f.value)
