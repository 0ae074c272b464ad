"""CPU checks of the drop-in boundary: the C-ABI library loads here (no GPU) and exports every symbol
include/ccb200.h declares; the ctypes table covers them all; compute entry points fail loudly
without a GPU instead of falling back to anything."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT


def _declared(headers=("ccb200.h", "ccb200_test.h")):
    names = set()
    for h in headers:
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(ccb_[a-z0-9_]+)\s*\(", src))
    return sorted(names)


def test_public_header_has_no_test_hooks():
    """The drop-in boundary (include/ccb200.h) declares no self-test / profiling / option hooks: those live in
    include/ccb200_test.h."""
    public = _declared(("ccb200.h",))
    assert not [n for n in public if n.startswith(("ccb_selftest_", "ccb_test_", "ccb_bench_"))]
    hooks = set(_declared(("ccb200_test.h",))) - set(public)
    assert {"ccb_test_set_option", "ccb_bench_kernel", "ccb_selftest_vg_partition"} <= hooks


def test_header_declares_entry_points():
    names = _declared()
    assert len(names) > 60
    for must in ("ccb_idf_generate", "ccb_env_step_actions", "ccb_env_get_map", "ccb_env_get_cells",
                 "ccb_env_objective", "ccb_ctrl_compute_actions", "ccb_env_cnn_input_dev",
                 "ccb_env_voronoi_partials_dev"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from coveragecontrol_b200 import _lib

    lib = C.CDLL(_lib.lib_path())
    missing = [n for n in _declared() if not hasattr(lib, n)]
    assert not missing, f"declared in include/ccb200.h but not exported: {missing}"


def test_ctypes_table_matches_header():
    from coveragecontrol_b200 import _lib

    declared, bound = set(_declared()), set(_lib.SIGNATURES)
    assert declared == bound, (sorted(declared - bound), sorted(bound - declared))
    _lib.load()  # resolves every prototype
    assert _lib.load().ccb_abi_version() == 1


def test_no_cpu_fallback():
    """Without a GPU every handle constructor refuses; the product never routes through the oracle."""
    from coveragecontrol_b200 import _lib

    lib = _lib.load()
    if lib.ccb_cuda_available():
        pytest.skip("GPU present")
    p = _lib.CParams()
    lib.ccb_params_default(C.byref(p))
    assert not lib.ccb_idf_create(C.byref(p), 0, None)
    assert b"no CUDA device" in lib.ccb_last_error()
    assert not lib.ccb_env_create(C.byref(p), 1, 2, 0, None)
    import coverage_control as cc

    with pytest.raises(RuntimeError):
        cc.WorldIDF(cc.Parameters())
    pkg = os.path.join(ROOT, "coveragecontrol_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "liboracle" not in text, f
