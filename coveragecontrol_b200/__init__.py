"""coveragecontrol_b200 -- B200-native (hand-written sm_100a CUDA) implementation of the
CoverageControl grid hot path behind the reference's ``coverage_control._core`` API.

The compute lives in ``libccb200.so`` (C ABI, ``include/ccb200.h``); this package is the thin
Python host side.  There is no CPU fallback: without the library or without a GPU every
compute call raises.
"""
from __future__ import annotations

__version__ = "0.1.0"

from . import _lib  # noqa: F401
from .core import *  # noqa: F401,F403
from .core import seed  # noqa: F401


def __getattr__(name):
    if name == "CoverageEnvUtils":  # imports torch lazily
        from .coverage_env_utils import CoverageEnvUtils

        return CoverageEnvUtils
    raise AttributeError(name)
