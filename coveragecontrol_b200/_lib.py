"""ctypes binding of libccb200.so (the C ABI declared in include/ccb200.h).

There is no CPU implementation behind this module: if the library is missing it is built with
nvcc, and if that fails the import raises.  Every compute entry point fails loudly when no GPU is
present (``ccb_last_error`` -> RuntimeError).
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

c_double_p = C.POINTER(C.c_double)
c_float_p = C.POINTER(C.c_float)
c_int_p = C.POINTER(C.c_int)


class CParams(C.Structure):
    """``ccb_params`` (reference parameters.h:49-175)."""

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


# name -> (restype, argtypes).  Kept in one table so tests can check it against include/ccb200.h.
SIGNATURES = {
    "ccb_abi_version": (C.c_int, []),
    "ccb_last_error": (C.c_char_p, []),
    "ccb_cuda_available": (C.c_int, []),
    "ccb_device_count": (C.c_int, []),
    "ccb_params_default": (None, [C.POINTER(CParams)]),
    "ccb_idf_create": (C.c_void_p, [C.POINTER(CParams), C.c_int, C.c_void_p]),
    "ccb_idf_destroy": (None, [C.c_void_p]),
    "ccb_idf_add_bnd": (C.c_int, [C.c_void_p] + [C.c_double] * 6),
    "ccb_idf_add_polygon": (C.c_int, [C.c_void_p, c_double_p, C.c_int, C.c_float]),
    "ccb_idf_num_features": (C.c_int, [C.c_void_p]),
    "ccb_idf_generate": (C.c_int, [C.c_void_p, C.c_int]),
    "ccb_idf_get_map": (C.c_int, [C.c_void_p, c_float_p]),
    "ccb_idf_set_map": (C.c_int, [C.c_void_p, c_float_p]),
    "ccb_idf_map_dev": (C.c_void_p, [C.c_void_p]),
    "ccb_idf_copy_map": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ccb_idf_norm_factor": (C.c_int, [C.c_void_p, c_float_p]),
    "ccb_idf_generate_band": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "ccb_idf_band_max_dev": (C.c_void_p, [C.c_void_p]),
    "ccb_idf_normalize_band": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "ccb_idf_map_dev_mut": (C.c_void_p, [C.c_void_p]),
    "ccb_idf_get_bnds": (C.c_int, [C.c_void_p, c_double_p, C.c_int]),
    "ccb_idf_num_bnds": (C.c_int, [C.c_void_p]),
    "ccb_idf_num_polygons": (C.c_int, [C.c_void_p]),
    "ccb_idf_get_polygon": (C.c_int, [C.c_void_p, C.c_int, c_double_p, C.c_int, c_float_p]),
    "ccb_env_create": (C.c_void_p, [C.POINTER(CParams), C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "ccb_env_destroy": (None, [C.c_void_p]),
    "ccb_env_num_envs": (C.c_int, [C.c_void_p]),
    "ccb_env_num_robots": (C.c_int, [C.c_void_p]),
    "ccb_env_set_world_idf": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "ccb_env_set_world_host": (C.c_int, [C.c_void_p, C.c_int, c_float_p]),
    "ccb_env_set_world_dev": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "ccb_env_reset": (C.c_int, [C.c_void_p, c_double_p]),
    "ccb_env_reset_env": (C.c_int, [C.c_void_p, C.c_int, c_double_p]),
    "ccb_env_step_actions": (C.c_int, [C.c_void_p, c_double_p, c_int_p]),
    "ccb_env_step_control": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double]),
    "ccb_env_set_positions": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, c_double_p]),
    "ccb_env_get_positions": (C.c_int, [C.c_void_p, c_double_p]),
    "ccb_env_clear_robot_maps": (C.c_int, [C.c_void_p, C.c_int]),
    "ccb_env_clear_explored_idf": (C.c_int, [C.c_void_p, C.c_int]),
    "ccb_env_get_map": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, c_float_p]),
    "ccb_env_get_map_dev": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "ccb_env_map_dim": (C.c_int, [C.c_void_p, C.c_int]),
    "ccb_env_local_maps_dev": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ccb_env_obstacle_maps_dev": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ccb_env_comm_maps_dev": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "ccb_env_cnn_input_dev": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "ccb_env_graph_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "ccb_env_comm_maps_cpp": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_float_p]),
    "ccb_env_compute_voronoi": (C.c_int, [C.c_void_p]),
    "ccb_env_get_cells": (C.c_int, [C.c_void_p, C.c_int, c_double_p, c_double_p, c_double_p, c_double_p,
                                    c_int_p, c_double_p]),
    "ccb_env_get_cells_all": (C.c_int, [C.c_void_p, c_double_p, c_double_p, c_double_p]),
    "ccb_env_objective": (C.c_int, [C.c_void_p, c_double_p]),
    "ccb_env_objective_dev": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ccb_env_voronoi_labels": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int32)]),
    "ccb_env_voronoi_features": (C.c_int, [C.c_void_p, c_double_p]),
    "ccb_env_voronoi_features_dev": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ccb_env_local_voronoi_features": (C.c_int, [C.c_void_p, c_double_p]),
    "ccb_env_exploration_features": (C.c_int, [C.c_void_p, C.c_int, c_double_p]),
    "ccb_test_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int]),
    "ccb_selftest_triangulate": (C.c_int, [c_double_p, C.c_int, c_double_p, C.c_int]),
    "ccb_selftest_frontier_select": (C.c_int, [c_double_p, C.c_int, C.c_int, C.POINTER(C.c_int)]),
    "ccb_selftest_resize_weights": (C.c_int, [C.c_int, C.c_int, c_int_p, c_int_p, c_float_p, C.c_int]),
    "ccb_selftest_make_window": (C.c_int, [C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, c_int_p]),
    "ccb_selftest_vg_partition": (C.c_int, [c_double_p, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int,
                                            C.POINTER(C.c_uint32), c_int_p, c_int_p]),
    "ccb_selftest_vw_row_interval": (C.c_int, [C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_double_p, c_double_p,
                                               c_int_p, c_int_p, c_int_p, C.c_int, C.POINTER(C.c_uint32)]),
    "ccb_env_exploration_ratios": (C.c_int, [C.c_void_p, c_double_p]),
    "ccb_env_check_oscillation": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_int_p]),
    "ccb_env_sync": (C.c_int, [C.c_void_p]),
    "ccb_env_flagged_rows": (C.c_int, [C.c_void_p, c_int_p, c_int_p, C.c_int]),
    "ccb_env_world_dev": (C.c_void_p, [C.c_void_p]),
    "ccb_env_positions_dev": (C.c_void_p, [C.c_void_p]),
    "ccb_env_set_band_mode": (C.c_int, [C.c_void_p, C.c_int]),
    "ccb_env_voronoi_partials_dev": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "ccb_env_finalize_from_partials_dev": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ccb_env_step_actions_dev": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ccb_env_window_partials_dev": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "ccb_env_set_window_sums_dev": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ccb_peer_buffer_bytes": (C.c_size_t, [C.c_int, C.c_size_t]),
    "ccb_peer_allreduce_sum": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p), C.c_int,
                                         C.c_int, C.c_void_p]),
    "ccb_ctrl_create": (C.c_void_p, [C.c_void_p, C.c_int]),
    "ccb_ctrl_destroy": (None, [C.c_void_p]),
    "ccb_ctrl_compute_actions": (C.c_int, [C.c_void_p]),
    "ccb_ctrl_get_actions": (C.c_int, [C.c_void_p, c_double_p]),
    "ccb_ctrl_get_goals": (C.c_int, [C.c_void_p, c_double_p]),
    "ccb_ctrl_is_converged": (C.c_int, [C.c_void_p, c_int_p]),
    "ccb_ctrl_rollout": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "ccb_ctrl_rollout_obs": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t]),
    "ccb_ctrl_actions_dev": (C.c_void_p, [C.c_void_p]),
    "ccb_bench_kernel": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "ccb_bench_idf_kernel": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_float_p]),
    "ccb_bench_fp32_peak": (C.c_int, [C.c_int, c_double_p]),
    "ccb_test_count": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_uint64)]),
    "ccb_kernel_launch_count": (C.c_uint64, []),
}

CCB_MAX_CELL_VERTS = 64

_lib = None


def lib_path() -> str:
    return _build.LIB


def load():
    """Load (building first if needed) the shared library and declare every prototype."""
    global _lib
    if _lib is not None:
        return _lib
    if _build.needs_build():  # missing, or older than a source it is built from
        _build.build()        # (file lock + atomic replace: safe when every rank of a torchrun starts at once)
    lib = C.CDLL(_build.LIB)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    if lib.ccb_abi_version() != 1:
        raise RuntimeError("libccb200.so ABI version mismatch; rebuild with python -m coveragecontrol_b200.build --force")
    _lib = lib
    return lib


# Diagnostics: CCB_<OPTION>=<int> in the process environment forces a kernel choice on every environment batch created
# afterwards (tests and A/B timing; `ccb_test_set_option`, include/ccb200_test.h).  The library itself never reads
# the environment.
_OVERRIDES = {"CCB_VG_GENERIC": "vg_generic", "CCB_VG_PREFIX": "vg_prefix", "CCB_VG_FUSED": "vg_fused", "CCB_VG_BULK": "vg_bulk", "CCB_VW_GENERIC": "vw_generic",
              "CCB_CNN_GENERIC": "cnn_generic", "CCB_OBS_V1": "obs_v1", "CCB_OBS_SORTED": "obs_sorted", "CCB_OBS_PREFETCH": "obs_prefetch", "CCB_VG_PF": "vg_pf", "CCB_VOR_RPW": "vor_rpw", "CCB_VG_WARPS": "vg_warps",
              "CCB_VG_GRIDDIV": "vg_griddiv", "CCB_VG_PART_RPT": "vg_part_rpt", "CCB_NO_AUX": "no_aux",
              "CCB_NO_GRAPH": "no_graph", "CCB_NO_WFUSE": "no_wfuse"}


def apply_test_overrides(env_handle) -> None:
    for var, name in _OVERRIDES.items():
        v = os.environ.get(var)
        if v is not None and v.strip() != "":
            check(load().ccb_test_set_option(env_handle, name.encode(), int(v)))


def last_error() -> str:
    msg = load().ccb_last_error()
    return msg.decode() if msg else ""


def check(status: int, exc=RuntimeError) -> None:
    if status != 0:
        raise exc(last_error() or f"libccb200 call failed with status {status}")
