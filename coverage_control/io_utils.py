"""IOUtils with the interface of reference python/coverage_control/io_utils.py."""
from __future__ import annotations

import os
import pathlib
import tomllib


class IOUtils:
    @staticmethod
    def sanitize_path(path_str: str) -> str:
        return os.path.normpath(os.path.expanduser(os.path.expandvars(path_str)))

    @staticmethod
    def load_toml(path: str) -> dict:
        path = IOUtils.sanitize_path(path)
        if not os.path.exists(path):
            raise FileNotFoundError(f"IOUtils::load_toml: File not found: {path}")
        with open(path, "rb") as f:
            return tomllib.load(f)

    @staticmethod
    def load_yaml(path: str) -> dict:
        """reference io_utils.py:93-117."""
        import yaml

        path = IOUtils.sanitize_path(path)
        if not os.path.exists(path):
            raise FileNotFoundError(f"IOUtils::load_yaml File not found: {path}")
        with open(path, "rb") as f:
            return yaml.load(f, Loader=yaml.FullLoader)

    @staticmethod
    def load_tensor(path: str):
        import torch

        if isinstance(path, pathlib.Path):
            path = str(path)
        path = IOUtils.sanitize_path(path)
        if not os.path.exists(path):
            raise FileNotFoundError(f"IOUtils::load_tensor: File not found: {path}")
        data = torch.load(path, weights_only=True)
        if isinstance(data, torch.jit.ScriptModule):
            return next(data.parameters())
        return data
