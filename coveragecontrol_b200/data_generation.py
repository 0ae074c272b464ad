"""Batched dataset generation for LPAC (BASELINE config 3): the reference's
``python/data_generation/data_generation.py`` with the environments stepped together on the GPU.

The reference runs ONE environment at a time (``run_data_generation`` :217-259): an oracle controller
(ClairvoyantCVT, or CentralizedCVT) drives the robots; every ``EveryNumSteps``-th step ``step_with_save``
(:261-300) records the controller's actions, the robot positions, ``GetRobotVoronoiFeatures``, the local and
obstacle maps (resized to ``CNNMapSize``), the communication maps and the edge weights; one more sample
is taken when an episode ends (converged data).  Here ``NumEnvs`` environments advance in lockstep on one
GPU, every quantity of a sample is produced on the device for the whole batch by the kernels behind
``CoverageEnvUtils`` (one fused observation kernel instead of two raw maps + two CPU resizes per robot),
and a slot whose episode ends is re-initialised with a fresh random world.  The tensors have the
reference's names, shapes and dtypes (``save_dataset`` :412-470), so ``nn/data_loaders`` reads them as is;
only the order of the samples differs (they interleave environments).

    python -m coveragecontrol_b200.data_generation <data_params.toml> [--append-dir NAME] [--split true]
"""
from __future__ import annotations

import math
import os
import pathlib

import numpy as np
import torch

from . import _lib
from . import core as _core
from .core import BatchedCoverageSystem, BatchedCVT, Parameters, WorldIDF, kLargeEps
from .parallel import _wrap_dev

_KIND = {"ClairvoyantCVT": 0, "CentralizedCVT": 1, "DecentralizedCVT": 2}


def random_world(params: Parameters, device: int = 0) -> tuple[WorldIDF, np.ndarray]:
    """A random environment as ``CoverageSystem(params)`` builds it (coverage_system.cpp:39-112): random
    Gaussians and polygons, robots uniform in ``[0, pRobotInitDist]^2``."""
    idf = WorldIDF(params, device=device)
    idf.AddNormalDistribution(_core._random_bnds(params, params.pNumGaussianFeatures))
    for pf in _core._random_polygons(params, params.pNumPolygons):
        idf.AddUniformDistributionPolygon(pf)
    idf.GenerateMap()
    hi = params.pRobotInitDist - kLargeEps
    return idf, _core._rng.uniform(kLargeEps, hi, size=(params.pNumRobots, 2))


class BatchedDataGenerator:
    """``num_envs`` environments stepped together; ``generate(n)`` returns the dataset tensors on the CPU."""

    def __init__(self, params: Parameters, num_envs: int, *, cnn_map_size: int = 32, every_num_steps: int = 5,
                 converged_data_ratio: float = 0.02, algorithm: str = "ClairvoyantCVT", save_objective: bool = False,
                 device: int = 0, worlds=None):
        if algorithm not in _KIND:
            raise ValueError(f"unknown algorithm {algorithm}")
        self.params = params.copy()
        self.E, self.R, self.C = int(num_envs), params.pNumRobots, int(cnn_map_size)
        self.every = max(1, int(every_num_steps))
        self.ratio = float(converged_data_ratio)
        self.save_objective = bool(save_objective)
        self.device = device
        self.dev = torch.device("cuda", device)
        self.env = BatchedCoverageSystem(self.params, self.E, self.R, device=device)
        positions = np.zeros((self.E, self.R, 2))
        self._worlds = iter(worlds) if worlds is not None else None  # optional (world, positions) source (tests)
        for e in range(self.E):
            world, positions[e] = self._next_world()
            self.env.set_world(e, world)
        self.env.reset(positions)
        self.ctrl = BatchedCVT(self.env, _KIND[algorithm])
        self.steps = np.zeros(self.E, np.int64)   # steps taken in the current episode of every slot
        self.env_count = self.E
        self._obs = torch.empty((self.E, self.R, 4, self.C, self.C), dtype=torch.float32, device=self.dev)
        self._feat = torch.empty((self.E, self.R, 7), dtype=torch.float64, device=self.dev)

    def _next_world(self):
        if self._worlds is not None:
            return next(self._worlds)
        return random_world(self.params, self.device)

    # ---- one sample of every environment in `sel` (the state before the step, as step_with_save) ----------
    def _record(self, out: dict, start: int, sel: np.ndarray) -> None:
        lib = _lib.load()
        n = len(sel)
        idx = torch.as_tensor(sel, device=self.dev)
        act = _wrap_dev(self.ctrl.actions_device_ptr(), (self.E, self.R, 2), torch.float64, self.dev)
        pos = _wrap_dev(self.env.positions_device_ptr(), (self.E, self.R, 2), torch.float64, self.dev)
        _lib.check(lib.ccb_env_voronoi_features_dev(self.env._h, self._feat.data_ptr()))
        _lib.check(lib.ccb_env_cnn_input_dev(self.env._h, self.C, 1, self._obs.data_ptr()))
        posf = pos.to(torch.float32)
        dist = torch.cdist(posf, posf, 2)
        w = torch.exp(-(dist.square()) / (self.params.pCommunicationRange * self.params.pCommunicationRange))
        w.masked_fill_(w < 1.0 / math.e, 0)
        w.diagonal(dim1=-2, dim2=-1).zero_()
        sl = slice(start, start + n)
        out["actions"][sl].copy_(act[idx].to(torch.float32), non_blocking=True)
        out["robot_positions"][sl].copy_(posf[idx], non_blocking=True)
        out["coverage_features"][sl].copy_(self._feat[idx].to(torch.float32), non_blocking=True)
        out["local_maps"][sl].copy_(self._obs[idx, :, 0], non_blocking=True)
        out["comm_maps"][sl].copy_(self._obs[idx, :, 1:3], non_blocking=True)
        out["obstacle_maps"][sl].copy_(self._obs[idx, :, 3], non_blocking=True)
        out["edge_weights"][sl].copy_(w[idx], non_blocking=True)
        if self.save_objective:
            out["objectives"][sl] = torch.from_numpy(self.env.objective()[sel]).to(torch.float32)

    def generate(self, num_dataset: int) -> dict:
        n, R, C = int(num_dataset), self.R, self.C
        shapes = {"actions": (n, R, 2), "robot_positions": (n, R, 2), "local_maps": (n, R, C, C), "obstacle_maps": (n, R, C, C),
                  "comm_maps": (n, R, 2, C, C), "coverage_features": (n, R, 7), "edge_weights": (n, R, R)}
        # (pageable output: page-locking gigabytes costs more than it saves -- measured)
        out = {k: torch.empty(s) for k, s in shapes.items()}  # every row is written before generate() returns
        if self.save_objective:
            out["objectives"] = torch.zeros(n)
        num_conv = math.ceil(self.ratio * n)          # data_generation.py:104-107
        num_nonconv = n - num_conv
        count = nonconv = conv = 0
        episode = self.params.pEpisodeSteps
        while count < n:
            self.ctrl.compute_actions()
            converged = self.ctrl.converged()
            ended = converged | (self.steps >= episode)
            # which slots give a sample at this step (data_generation.py:239-257)
            want_nc = (~ended) & (self.steps % self.every == 0)
            sel_nc = np.flatnonzero(want_nc)[: max(0, num_nonconv - nonconv)]
            sel_c = np.flatnonzero(ended)[: max(0, num_conv - conv)]
            sel = np.concatenate([sel_nc, sel_c])[: n - count]
            if len(sel):
                self._record(out, count, sel)
                count += len(sel)
                nonconv += int(np.isin(sel, sel_nc).sum())
                conv += int(np.isin(sel, sel_c).sum())
            self.env.step_actions_device(self.ctrl.actions_device_ptr())
            self.steps += 1
            for e in np.flatnonzero(ended):  # a fresh episode in every slot that ended
                world, pos = self._next_world()
                self.env.set_world(int(e), world)
                self.env.reset_env(int(e), pos)
                self.steps[e] = 0
                self.env_count += 1
        torch.cuda.synchronize(self.dev)
        return out


def normalize_tensor(tensor: torch.Tensor, epsilon: float = 1e-6, zero_mean: bool = False, is_symmetric: bool = False):
    """data_generation.py:340-367."""
    dims = list(range(tensor.dim()))[:-1]
    # zero_mean: the reference returns torch.zeros_like(<the index tensor of the leading dimensions>), i.e. INTEGER zeros
    # of length dim() - 1 (which broadcasts for [n, R, 2] actions because dim() - 1 == 2); actions_mean.pt holds that
    mean = torch.zeros_like(torch.tensor(range(tensor.dim()))[:-1]) if zero_mean else tensor.mean(dim=dims)
    std = tensor.std(dim=dims)
    if is_symmetric:
        std = torch.ones_like(std) * std.mean()
    std = torch.where(std < epsilon, torch.full_like(std, epsilon), std)
    return (tensor - mean) / std, mean, std


def save_dataset(data: dict, dataset_dir, *, split=None, as_sparse: bool = False, normalize: bool = True) -> None:
    """The files of ``save_dataset`` (data_generation.py:412-470): ``split`` = (train, val) ratios or None."""
    root = pathlib.Path(dataset_dir)
    root.mkdir(parents=True, exist_ok=True)
    n = len(data["actions"])
    sparse_names = {"local_maps", "obstacle_maps", "edge_weights", "comm_maps"}

    def save(t: torch.Tensor, name: str):
        sp = as_sparse and name in sparse_names
        if split is None:
            torch.save(t.to_sparse() if sp else t, root / f"{name}.pt")
            return
        n_tr, n_va = int(n * split[0]), int(n * split[1])
        for sub, part in (("train", t[:n_tr]), ("val", t[n_tr:n_tr + n_va]), ("test", t[n_tr + n_va:])):
            (root / sub).mkdir(exist_ok=True)
            part = part.clone()
            torch.save(part.to_sparse() if sp else part, root / sub / f"{name}.pt")

    for name in ("robot_positions", "local_maps", "obstacle_maps", "edge_weights", "comm_maps", "actions", "coverage_features"):
        save(data[name], name)
    if "objectives" in data:
        save(data["objectives"], "objectives")
    if normalize:
        na, mean, std = normalize_tensor(data["actions"], is_symmetric=True, zero_mean=True)
        save(na, "normalized_actions")
        torch.save(mean, root / "actions_mean.pt")
        torch.save(std, root / "actions_std.pt")
        if "objectives" in data:
            no, mean, std = normalize_tensor(data["objectives"])
            save(no, "normalized_objectives")
            torch.save(mean, root / "objectives_mean.pt")
            torch.save(std, root / "objectives_std.pt")


def main(argv=None) -> None:
    import argparse
    import time
    import tomllib

    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("config_file")
    ap.add_argument("--append-dir", default="data")
    ap.add_argument("--split", default="true")
    ap.add_argument("--algorithm", default="ClairvoyantCVT")
    ap.add_argument("--num-envs", type=int, default=None, help="environments stepped together (config key NumEnvs, default 256)")
    a = ap.parse_args(argv)
    with open(a.config_file, "rb") as f:
        cfg = tomllib.load(f)
    params = Parameters(os.path.expandvars(os.path.expanduser(cfg["EnvironmentConfig"])))
    if "TimeStep" in cfg:
        params.pTimeStep = cfg["TimeStep"]
    gen = BatchedDataGenerator(params, a.num_envs or cfg.get("NumEnvs", 256), cnn_map_size=cfg["CNNMapSize"],
                               every_num_steps=cfg["EveryNumSteps"], converged_data_ratio=cfg["ConvergedDataRatio"],
                               algorithm=a.algorithm, save_objective=cfg.get("SaveObjective", False))
    t0 = time.perf_counter()
    data = gen.generate(cfg["NumDataset"])
    dt = time.perf_counter() - t0
    out = pathlib.Path(os.path.expandvars(os.path.expanduser(cfg["DataDir"]))) / a.append_dir
    split = None
    if a.split.lower() in ("1", "true", "yes"):
        split = (cfg["DataSetSplit"]["TrainRatio"], cfg["DataSetSplit"]["ValRatio"])
    save_dataset(data, out, split=split, as_sparse=cfg.get("SaveAsSparseQ", False), normalize=cfg.get("NormalizeQ", True))
    print(f"{cfg['NumDataset']} samples from {gen.env_count} environments in {dt:.1f} s -> {out}")


if __name__ == "__main__":
    main()
