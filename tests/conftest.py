import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu() -> bool:
    try:
        from coveragecontrol_b200 import _lib

        return bool(_lib.load().ccb_cuda_available())
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


# ---- shared fixtures: the reference's own KAT inputs (python/tests/test_coverage.py:31-43) ----
KAT_BND = np.array([[182, 438, 55, 55, 8, 0], [950, 97, 47, 47, 7, 0], [714, 920, 46, 46, 7, 0]], np.float64)
KAT_POS = np.array([[433, 154], [139, 319], [92, 112]], np.float64)
KAT_EXPECTED = {"clairvoyant": (98.90, 166), "centralized": (58.13, 104), "decentralized": (8.87, 17)}


def random_world(seed: int, N: int, F: int, R: int, sigma=(40.0, 60.0), peak=(6.0, 10.0)):
    """Synthetic inputs of SURVEY.md 8(d): circular Gaussians + robot positions."""
    rng = np.random.default_rng(seed)
    scale = N / 1024.0
    bnd = np.zeros((F, 6))
    bnd[:, 0:2] = rng.uniform(0, N, (F, 2))
    s = rng.uniform(sigma[0] * scale, sigma[1] * scale, F)
    bnd[:, 2] = s
    bnd[:, 3] = s
    bnd[:, 4] = rng.uniform(peak[0], peak[1], F)
    pos = rng.uniform(1e-4, N - 1e-4, (R, 2))
    return bnd, pos
