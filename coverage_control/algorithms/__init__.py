"""reference python/coverage_control/algorithms/__init__.py."""
from .._core import CentralizedCVT, ClairvoyantCVT, DecentralizedCVT, NearOptimalCVT
from .controllers import ControllerCVT, ControllerNN

__all__ = ["NearOptimalCVT", "ClairvoyantCVT", "CentralizedCVT", "DecentralizedCVT", "ControllerCVT", "ControllerNN"]
