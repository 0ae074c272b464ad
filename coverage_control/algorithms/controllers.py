"""ControllerCVT / ControllerNN with the interface of reference python/coverage_control/algorithms/controllers.py:40-160."""
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


class ControllerNN:
    """reference controllers.py:91-160: get the graph data, run the network, step the environment.

    The LPAC network itself (``coverage_control.nn``) is outside this package: pass ``model=`` (any ``torch.nn.Module``
    mapping the graph data to ``[R, 2]`` actions, with ``actions_mean`` / ``actions_std`` attributes, as ``cc_nn.LPAC``
    has), or a config with ``"ModelFile"`` (a pickled module, as the reference loads it).  The observation, the graph
    and the actions stay on the GPU: the reference's ``actions.cpu()`` -> ``PointVector`` -> ``StepActions`` round
    trip becomes ``ccb_env_step_actions_dev`` on the action tensor."""

    def __init__(self, config: dict, params: Parameters, env: CoverageSystem, model=None):
        import torch

        from ..io_utils import IOUtils

        self.config = config
        self.params = params
        self.name = self.config["Name"]
        self.use_cnn = self.config["UseCNN"]
        self.use_comm_map = self.config["UseCommMap"]
        self.cnn_map_size = self.config["CNNMapSize"]
        self.device = torch.device("cuda", env._device)
        if model is not None:
            self.model = model.to(self.device)
        elif "ModelFile" in self.config:
            self.model_file = IOUtils.sanitize_path(self.config["ModelFile"])
            self.model = torch.load(self.model_file, weights_only=False).to(self.device)
        else:
            raise ImportError("ControllerNN: building LPAC from LearningParams / ModelStateDict needs coverage_control.nn, "
                              "which is outside this package; pass model= or a ModelFile")
        self.actions_mean = self.model.actions_mean.to(self.device)
        self.actions_std = self.model.actions_std.to(self.device)
        self.model.eval()

    def step(self, env: CoverageSystem) -> (float, bool):
        import torch

        from coveragecontrol_b200 import _lib
        from coveragecontrol_b200.coverage_env_utils import CoverageEnvUtils

        pyg_data = CoverageEnvUtils.get_torch_geometric_data(env, self.params, True, self.use_comm_map, self.cnn_map_size)
        with torch.no_grad():
            actions = self.model(pyg_data)
        actions = actions * self.actions_std + self.actions_mean
        if actions.shape != (env._R, 2):
            raise ValueError("The size of the actions don't match with the number of robots")
        a64 = actions.detach().to(torch.float64).reshape(1, env._R, 2).contiguous()
        stream = torch.cuda.current_stream(a64.device)
        if env._stream != stream.cuda_stream:
            stream.synchronize()  # the environment's stream must see the finished action tensor
        env._sync_world()
        _lib.check(_lib.load().ccb_env_step_actions_dev(env._h, a64.data_ptr()))
        env._invalidate()
        converged = bool(torch.allclose(actions, torch.zeros_like(actions), atol=1e-5))
        return env.GetObjectiveValue(), converged
