"""Frontier queue of GetExplorationFeatures (reference robot_model.h:383-401): the kernel's heap code
(csrc/maps.cuh expl_*, run on the host through ccb_selftest_frontier_select) against the real
std::priority_queue + FrontierCompare of the oracle (oracle/frontier_pq.cpp), with many ties -- which
candidates survive and the order they pop in are decided by the standard library's heap algorithms."""
import ctypes as C

import numpy as np
import pytest


def _both(values, K):
    from coveragecontrol_b200 import _lib
    from oracle import oracle as orc

    v = np.ascontiguousarray(values, np.float64)
    dp = v.ctypes.data_as(C.POINTER(C.c_double))
    a, b = np.full(K, -1, np.int32), np.full(K, -1, np.int32)
    na = _lib.load().ccb_selftest_frontier_select(dp, len(v), K, a.ctypes.data_as(C.POINTER(C.c_int)))
    nb = orc.lib().orc_frontier_select(dp, len(v), K, b.ctypes.data_as(C.POINTER(C.c_int)))
    return a[:na], b[:nb]


@pytest.mark.parametrize("K", [1, 2, 3, 7, 10, 16, 64])
@pytest.mark.parametrize("n", [0, 1, 5, 10, 11, 100, 1681])
@pytest.mark.parametrize("levels", [1, 2, 5, 0])
def test_heap_matches_std_priority_queue(K, n, levels):
    rng = np.random.default_rng(1000 * K + 10 * n + levels)
    # levels distinct values (0: all distinct): ties everywhere, as on an explored or an untouched map
    v = rng.random(n) if levels == 0 else rng.integers(0, levels, n).astype(np.float64) / 3.0
    a, b = _both(v, K)
    assert len(a) == len(b) == min(n, K)
    assert np.array_equal(a, b)
    if n:
        assert np.all(np.diff(v[b]) >= 0)  # popped worst first


def test_rejects_oversized_queue():
    from coveragecontrol_b200 import _lib

    v = np.zeros(4)
    out = np.zeros(128, np.int32)
    assert _lib.load().ccb_selftest_frontier_select(v.ctypes.data_as(C.POINTER(C.c_double)), 4, 65,
                                                    out.ctypes.data_as(C.POINTER(C.c_int))) == -1
