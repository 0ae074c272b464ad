"""Mirror of reference ``python/coverage_control/coverage_env_utils.py`` (CoverageEnvUtils).

Same static-method names and tensor layouts; the per-robot Python loops and the CPU
``torchvision`` resize are replaced by batched CUDA kernels that write straight into a torch
tensor allocated on the GPU (zero-copy hand-off to LPAC).  Pass ``device="cpu"`` to get host
tensors as the reference returns.
"""
from __future__ import annotations

import math

import numpy
import torch

from . import _lib
from .core import (BatchedCoverageSystem, CoverageSystem, DblVector, DblVectorVector, Parameters,
                   PointVector, _EnvBase)


def _cuda_device(env: _EnvBase) -> torch.device:
    return torch.device("cuda", env._device)


def _finish(t: torch.Tensor, env: _EnvBase, device):
    """Make the kernel output visible to torch and move it if the caller asked for another device."""
    if env._stream != torch.cuda.current_stream(t.device).cuda_stream:
        _lib.check(_lib.load().ccb_env_sync(env._h))
    if isinstance(env, CoverageSystem):
        t = t[0]
    if device is not None and torch.device(device) != t.device:
        t = t.to(device)
    return t


def _prepare(env: _EnvBase) -> None:
    """Edits made through ``GetWorldMapMutable`` must be on the device before a kernel reads the world map."""
    if isinstance(env, CoverageSystem):
        env._sync_world()


def _noisy(env: _EnvBase) -> bool:
    """pAddNoisePositions: the reference builds the communication maps / edge weights from the NOISY getters
    (GetRobotPositions / GetRelativePositonsNeighbors, coverage_system.cpp:340-366,626-645); the device kernels
    read the true positions, so those quantities go through the host expressions instead."""
    return isinstance(env, CoverageSystem) and bool(env._params.pAddNoisePositions)


def _host_comm_maps(env, params: Parameters, map_size: int) -> torch.Tensor:
    """reference :143-191 verbatim semantics on the host (float32, torch.round, duplicates summed)."""
    R = env.GetNumRobots()
    out = torch.zeros((R, 2, map_size, map_size))
    for r in range(R):
        rel = CoverageEnvUtils.to_tensor(env.GetRelativePositonsNeighbors(r))
        if rel.numel() == 0:
            continue
        idx = torch.round(rel * map_size / (params.pCommunicationRange * params.pResolution * 2.0)
                          + (map_size / 2.0 - params.pResolution / 2.0)).t().long()
        val = rel / params.pCommunicationRange
        ok = ((idx >= 0) & (idx < map_size)).all(0)  # (an index past the map -- pResolution < 1 -- is dropped, as the kernel does)
        idx, val = idx[:, ok], val[ok]
        for ch in range(2):
            out[r][ch] = torch.sparse_coo_tensor(idx, val[:, ch], torch.Size([map_size, map_size])).to_dense()
    return out


class CoverageEnvUtils:
    """Utility functions for the coverage environment (reference coverage_env_utils.py:42)."""

    @staticmethod
    def to_tensor(data: object) -> torch.Tensor:
        """reference :46-95."""
        if isinstance(data, numpy.ndarray):
            return torch.from_numpy(numpy.copy(data.astype(numpy.float32)))
        if isinstance(data, PointVector):
            return torch.from_numpy(numpy.asarray(data.to_numpy(), numpy.float32).reshape(len(data), 2).copy())
        if isinstance(data, DblVectorVector):
            return torch.tensor([[float(v) for v in row] for row in data], dtype=torch.float32)
        if isinstance(data, DblVector):
            return torch.tensor([float(v) for v in data], dtype=torch.float32)
        raise ValueError(f"Unknown data type: {type(data)}")

    @staticmethod
    def get_raw_local_maps(env, params: Parameters, device=None) -> torch.Tensor:
        """reference :97-117 -> [R, L, L] (or [E, R, L, L] for a batched environment)."""
        L = params.pLocalMapSize
        _prepare(env)
        out = torch.empty((env._E, env._R, L, L), dtype=torch.float32, device=_cuda_device(env))
        _lib.check(_lib.load().ccb_env_local_maps_dev(env._h, out.data_ptr()))
        return _finish(out, env, device)

    @staticmethod
    def get_raw_obstacle_maps(env, params: Parameters, device=None) -> torch.Tensor:
        """reference :119-141."""
        L = params.pLocalMapSize
        out = torch.empty((env._E, env._R, L, L), dtype=torch.float32, device=_cuda_device(env))
        _lib.check(_lib.load().ccb_env_obstacle_maps_dev(env._h, out.data_ptr()))
        return _finish(out, env, device)

    @staticmethod
    def get_communication_maps(env, params: Parameters, map_size: int, device=None) -> torch.Tensor:
        """reference :143-191 (float32, round-half-even) -> [R, 2, C, C]."""
        if _noisy(env):
            t = _host_comm_maps(env, params, int(map_size))
            return t if device is not None and torch.device(device).type == "cpu" else t.to(device or _cuda_device(env))
        out = torch.empty((env._E, env._R, 2, map_size, map_size), dtype=torch.float32, device=_cuda_device(env))
        _lib.check(_lib.load().ccb_env_comm_maps_dev(env._h, int(map_size), out.data_ptr()))
        return _finish(out, env, device)

    @staticmethod
    def resize_maps(maps: torch.Tensor, resized_map_size: int) -> torch.Tensor:
        """reference :200-225; kept for callers that resize stored raw maps (torchvision, any device)."""
        import torchvision

        shape = maps.shape
        maps = maps.reshape(-1, maps.shape[-2], maps.shape[-1])
        maps = torchvision.transforms.functional.resize(
            maps, (resized_map_size, resized_map_size),
            interpolation=torchvision.transforms.InterpolationMode.BILINEAR, antialias=True)
        return maps.reshape(shape[:-2] + maps.shape[-2:])

    @staticmethod
    def get_maps(env, params: Parameters, resized_map_size: int, use_comm_map: bool, device=None) -> torch.Tensor:
        """reference :227-277 -> [R, 4 or 2, C, C]: crop, mask, antialias resize and the
        communication channels fused on the GPU."""
        C = int(resized_map_size)
        nch = 4 if use_comm_map else 2
        _prepare(env)
        out = torch.empty((env._E, env._R, nch, C, C), dtype=torch.float32, device=_cuda_device(env))
        noisy = bool(use_comm_map) and _noisy(env)
        _lib.check(_lib.load().ccb_env_cnn_input_dev(env._h, C, int(bool(use_comm_map) and not noisy), out.data_ptr()))
        out = _finish(out, env, None)
        if noisy:  # channels 1, 2 from the noisy neighbour positions, as the reference does
            out[:, 1:3] = _host_comm_maps(env, params, C).to(out.device)
        if device is not None and torch.device(device) != out.device:
            out = out.to(device)
        return out

    @staticmethod
    def get_voronoi_features(env, device=None) -> torch.Tensor:
        """reference :279-292 -> [R, 7]."""
        if isinstance(env, BatchedCoverageSystem):
            return torch.from_numpy(env.voronoi_features().astype(numpy.float32))
        feats = env.GetRobotVoronoiFeatures()
        return torch.tensor([[float(v) for v in row] for row in feats], dtype=torch.float32)

    @staticmethod
    def get_robot_positions(env) -> torch.Tensor:
        """reference :294-305."""
        if isinstance(env, BatchedCoverageSystem):
            return torch.from_numpy(env.positions().astype(numpy.float32))
        return CoverageEnvUtils.to_tensor(env.GetRobotPositions())

    @staticmethod
    def get_weights(env, params: Parameters, device=None) -> torch.Tensor:
        """reference :313-335 -> [R, R] ([E, R, R] batched).  ``device=None``: host tensor computed with the
        reference's own torch expressions; any ``device``: the ``k_graph`` kernel, no host round trip."""
        if device is not None and not _noisy(env):
            if float(params.pCommunicationRange) != float(env._params.pCommunicationRange):
                raise ValueError("get_weights(device=...): params.pCommunicationRange differs from the environment's")
            out = torch.empty((env._E, env._R, env._R), dtype=torch.float32, device=_cuda_device(env))
            _lib.check(_lib.load().ccb_env_graph_dev(env._h, out.data_ptr(), None))
            return _finish(out, env, device)
        onebyexp = 1.0 / math.exp(1.0)
        robot_positions = CoverageEnvUtils.get_robot_positions(env)
        pairwise_distances = torch.cdist(robot_positions, robot_positions, 2)
        edge_weights = torch.exp(-(pairwise_distances.square()) / (params.pCommunicationRange * params.pCommunicationRange))
        edge_weights.masked_fill_(edge_weights < onebyexp, 0)
        if edge_weights.dim() == 2:
            edge_weights.fill_diagonal_(0)
        else:
            edge_weights.diagonal(dim1=-2, dim2=-1).zero_()
        return edge_weights if device is None else edge_weights.to(device)

    @staticmethod
    def get_torch_geometric_data(env, params: Parameters, use_cnn: bool, use_comm_map: bool, map_size: int, device=None):
        """reference :337-377 for one CoverageSystem.  Everything is produced on the GPU (node features by
        ``get_maps`` / ``k_voronoi_features``, edge weights and normalised positions by ``k_graph``) and stays
        there unless ``device`` says otherwise: no host round trip between the simulator and the GNN.
        Returns ``torch_geometric.data.Data`` when torch_geometric is importable, else a ``GraphData`` with
        the same four fields (this image does not ship torch_geometric)."""
        if isinstance(env, BatchedCoverageSystem):
            raise ValueError("get_torch_geometric_data takes one CoverageSystem, as the reference; "
                             "use get_maps / get_weights(device=...) for a batch")
        if float(params.pCommunicationRange) != float(env._params.pCommunicationRange) or \
                int(params.pWorldMapSize) != int(env._params.pWorldMapSize):
            raise ValueError("get_torch_geometric_data: params differ from the environment's")
        dev = _cuda_device(env)
        _prepare(env)
        if use_cnn:
            features = CoverageEnvUtils.get_maps(env, params, map_size, use_comm_map)
        else:
            feats = torch.empty((env._E, env._R, 7), dtype=torch.float64, device=dev)
            _lib.check(_lib.load().ccb_env_voronoi_features_dev(env._h, feats.data_ptr()))
            features = _finish(feats, env, None).to(torch.float32)
        w = torch.empty((1, env._R, env._R), dtype=torch.float32, device=dev)
        pos = torch.empty((1, env._R, 2), dtype=torch.float32, device=dev)
        _lib.check(_lib.load().ccb_env_graph_dev(env._h, w.data_ptr(), pos.data_ptr()))
        w, pos = _finish(w, env, None), _finish(pos, env, None)
        if _noisy(env):  # reference :360-369: weights and positions from (separately drawn) noisy positions
            w = CoverageEnvUtils.get_weights(env, params).to(dev)
            half = float(params.pWorldMapSize) / 2.0
            pos = ((CoverageEnvUtils.get_robot_positions(env) + half) / float(params.pWorldMapSize)).to(dev)
        edge_index = w.nonzero().t().contiguous()          # row-major order, as to_sparse().coalesce().indices()
        edge_weight = w[edge_index[0], edge_index[1]]
        if device is not None:
            features, edge_index, edge_weight, pos = (t.to(device) for t in (features, edge_index, edge_weight, pos))
        try:
            import torch_geometric
        except ImportError:
            return GraphData(x=features, edge_index=edge_index, edge_weight=edge_weight, pos=pos)
        return torch_geometric.data.Data(x=features, edge_index=edge_index, edge_weight=edge_weight, pos=pos)


class GraphData:
    """The four fields of the ``torch_geometric.data.Data`` object the reference builds
    (coverage_env_utils.py:370-375), for installations without torch_geometric."""

    def __init__(self, x, edge_index, edge_weight, pos):
        self.x, self.edge_index, self.edge_weight, self.pos = x, edge_index, edge_weight, pos

    @property
    def num_nodes(self) -> int:
        return int(self.pos.shape[0])

    def to(self, device):
        return GraphData(self.x.to(device), self.edge_index.to(device), self.edge_weight.to(device), self.pos.to(device))
