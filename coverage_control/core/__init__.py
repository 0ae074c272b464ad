"""reference python/coverage_control/core/__init__.py:7-22."""
from .._core import (BivariateNormalDistribution, BNDVector, CoverageSystem, CudaUtils, DblVector,
                     DblVectorVector, Parameters, Point2, PointVector, PolygonFeature, RobotModel,
                     VoronoiCell, VoronoiCells, WorldIDF)

__all__ = ["Point2", "PointVector", "DblVector", "DblVectorVector", "PolygonFeature", "VoronoiCell",
           "VoronoiCells", "BivariateNormalDistribution", "BNDVector", "WorldIDF", "RobotModel",
           "CoverageSystem", "Parameters", "CudaUtils"]
