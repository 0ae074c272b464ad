"""ControllerCVT with the interface of reference python/coverage_control/algorithms/controllers.py:40-88."""
from __future__ import annotations

from .._core import CentralizedCVT, ClairvoyantCVT, CoverageSystem, DecentralizedCVT, NearOptimalCVT, Parameters


class ControllerCVT:
    def __init__(self, config: dict, params: Parameters, env: CoverageSystem):
        self.name = config["Name"]
        self.params = params
        algs = {"DecentralizedCVT": DecentralizedCVT, "ClairvoyantCVT": ClairvoyantCVT,
                "CentralizedCVT": CentralizedCVT, "NearOptimalCVT": NearOptimalCVT}
        if config["Algorithm"] not in algs:
            raise ValueError(f"Unknown controller type: {config['Algorithm']}")
        self.alg = algs[config["Algorithm"]](params, env)

    def step(self, env: CoverageSystem) -> (float, bool):
        self.alg.ComputeActions()
        actions = self.alg.GetActions()
        converged = self.alg.IsConverged()
        error_flag = env.StepActions(actions)
        if error_flag:
            raise ValueError("Error in step")
        return env.GetObjectiveValue(), converged
