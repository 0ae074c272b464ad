"""Golden manifest of the LPAC dataset files written by the REFERENCE's own
DatasetGenerator.save_dataset / normalize_tensor (python/data_generation/data_generation.py:340-470), run in this
container on seeded synthetic tensors.  The reference module imports `coverage_control`, which resolves to this
repo's drop-in package (no GPU is needed to import it or to save tensors).  Output: dataset_manifest.json =
{case: {relative path: {dtype, shape, sparse, sha256 of the dense values}}}.

    python tests/golden/make_dataset_golden.py        (needs /root/reference)
"""
import hashlib
import importlib.util
import json
import os
import pathlib
import sys
import tempfile

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/python/data_generation/data_generation.py"

CASES = {
    "nosplit_dense": dict(split=None, sparse=False, normalize=True, objectives=True),
    "split_sparse": dict(split=(0.7, 0.2), sparse=True, normalize=True, objectives=True),
    "split_dense_noobj": dict(split=(0.5, 0.25), sparse=False, normalize=True, objectives=False),
    "nosplit_nonorm": dict(split=None, sparse=True, normalize=False, objectives=False),
}


def synthetic(n=23, R=4, C=8, seed=0):
    g = torch.Generator().manual_seed(seed)
    sparse_like = lambda *s: torch.rand(*s, generator=g) * (torch.rand(*s, generator=g) < 0.2)  # noqa: E731
    return {
        "robot_positions": torch.rand(n, R, 2, generator=g) * 100,
        "local_maps": sparse_like(n, R, C, C),
        "obstacle_maps": (torch.rand(n, R, C, C, generator=g) < 0.1).float(),
        "edge_weights": sparse_like(n, R, R),
        "comm_maps": sparse_like(n, R, 2, C, C) - 0.1 * sparse_like(n, R, 2, C, C),
        "actions": torch.randn(n, R, 2, generator=g) * 3,
        "coverage_features": torch.randn(n, R, 7, generator=g),
        "objectives": torch.rand(n, generator=g) * 1e4,
    }


def manifest(root):
    out = {}
    root = pathlib.Path(root)
    for f in sorted(root.rglob("*.pt")):
        t = torch.load(f, weights_only=False)
        sp = bool(t.is_sparse)
        d = t.to_dense() if sp else t
        out[str(f.relative_to(root))] = {
            "dtype": str(d.dtype), "shape": list(d.shape), "sparse": sp,
            "sha256": hashlib.sha256(d.contiguous().numpy().tobytes()).hexdigest()}
    return out


def reference_manifest(case):
    sys.path.insert(0, ROOT)
    spec = importlib.util.spec_from_file_location("ref_datagen", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    cfg = CASES[case]
    data = synthetic()
    gen = object.__new__(mod.DatasetGenerator)  # no constructor: it would read config files and build environments
    with tempfile.TemporaryDirectory() as tmp:
        gen.dataset_dir = gen.dataset_dir_path = pathlib.Path(tmp)
        gen.config = {"SaveAsSparseQ": cfg["sparse"], "NormalizeQ": cfg["normalize"],
                      "DataSetSplit": {"TrainRatio": (cfg["split"] or (0, 0))[0], "ValRatio": (cfg["split"] or (0, 0))[1]}}
        gen.split_dataset = cfg["split"] is not None
        gen.num_dataset = len(data["actions"])
        gen.save_objective = cfg["objectives"]
        for k, v in data.items():
            setattr(gen, k, v)
        gen.save_dataset()
        return manifest(tmp)


if __name__ == "__main__":
    out = {case: reference_manifest(case) for case in CASES}
    with open(os.path.join(HERE, "dataset_manifest.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print({k: len(v) for k, v in out.items()})
