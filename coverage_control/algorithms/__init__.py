"""reference python/coverage_control/algorithms/__init__.py (the four CVT controllers; ControllerNN needs the LPAC network, out of scope)."""
from .._core import CentralizedCVT, ClairvoyantCVT, DecentralizedCVT, NearOptimalCVT
from .controllers import ControllerCVT

__all__ = ["NearOptimalCVT", "ClairvoyantCVT", "CentralizedCVT", "DecentralizedCVT", "ControllerCVT"]
