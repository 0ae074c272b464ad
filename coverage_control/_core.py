"""Stands where the reference's pybind11 extension ``coverage_control._core`` stands
(reference cppsrc/python_bindings/python_binds.cpp:36)."""
from coveragecontrol_b200.core import *  # noqa: F401,F403
from coveragecontrol_b200.core import (CentralizedCVT, ClairvoyantCVT, DecentralizedCVT, NearOptimalCVT)  # noqa: F401
