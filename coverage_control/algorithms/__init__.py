"""reference python/coverage_control/algorithms/__init__.py (CVT trio; NearOptimalCVT is out of scope)."""
from .._core import CentralizedCVT, ClairvoyantCVT, DecentralizedCVT
from .controllers import ControllerCVT

__all__ = ["ClairvoyantCVT", "CentralizedCVT", "DecentralizedCVT", "ControllerCVT"]
