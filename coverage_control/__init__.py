"""Drop-in ``coverage_control`` package: the reference's import names (reference
python/coverage_control/__init__.py:11-14) resolved onto coveragecontrol_b200."""
from __future__ import annotations

from coveragecontrol_b200 import __version__  # noqa: F401
from . import _core  # noqa: F401
from .core import *  # noqa: F401,F403
from .io_utils import IOUtils  # noqa: F401


def __getattr__(name):
    if name == "CoverageEnvUtils":
        from coveragecontrol_b200.coverage_env_utils import CoverageEnvUtils

        return CoverageEnvUtils
    raise AttributeError(name)


__all__ = ["__version__", "core", "IOUtils", "CoverageEnvUtils"]
