"""The LPAC dataset files: coveragecontrol_b200.data_generation.save_dataset / normalize_tensor
against the REFERENCE's own DatasetGenerator.save_dataset (python/data_generation/data_generation.py:340-470) on the
same seeded tensors -- file names, train / val / test split, dtypes, shapes, sparse or dense storage and the values
(sha256 of the dense tensor).  tests/golden/dataset_manifest.json was written by running the reference's code here
(tests/golden/make_dataset_golden.py); where /root/reference is present the reference is also run live."""
import importlib.util
import json
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_dataset_golden as mk  # noqa: E402

with open(os.path.join(HERE, "golden", "dataset_manifest.json")) as f:
    GOLDEN = json.load(f)


def _ours(case, tmp_path):
    from coveragecontrol_b200.data_generation import save_dataset

    cfg = mk.CASES[case]
    data = mk.synthetic()
    if not cfg["objectives"]:
        del data["objectives"]
    save_dataset(data, tmp_path, split=cfg["split"], as_sparse=cfg["sparse"], normalize=cfg["normalize"])
    return mk.manifest(tmp_path)


@pytest.mark.parametrize("case", sorted(mk.CASES))
def test_files_equal_the_reference_manifest(case, tmp_path):
    got, ref = _ours(case, tmp_path), GOLDEN[case]
    assert sorted(got) == sorted(ref), "file names / folders differ"
    for name in ref:
        assert got[name] == ref[name], name


@pytest.mark.skipif(not os.path.exists(mk.REF), reason="reference tree absent (GPU box)")
@pytest.mark.parametrize("case", sorted(mk.CASES))
def test_manifest_is_what_the_reference_writes_now(case):
    assert mk.reference_manifest(case) == GOLDEN[case]


def test_normalize_tensor_equals_the_reference_formula():
    """data_generation.py:340-367, restated: mean / std over all but the last dimension, zero_mean -> integer zeros
    of length dim() - 1, is_symmetric -> one std for all components, std floored at epsilon"""
    import torch

    from coveragecontrol_b200.data_generation import normalize_tensor

    g = torch.Generator().manual_seed(1)
    t = torch.randn(50, 4, 2, generator=g) * torch.tensor([3.0, 0.5])
    out, mean, std = normalize_tensor(t, is_symmetric=True, zero_mean=True)
    s = t.std(dim=[0, 1])
    assert torch.equal(std, torch.ones_like(s) * s.mean()) and not mean.any() and mean.shape == (2,)
    assert torch.equal(out, (t - mean) / std)
    const = torch.ones(10, 3)
    const[:, 1] = torch.arange(10.0)
    out, mean, std = normalize_tensor(const)
    assert std[0] == pytest.approx(1e-6) and std[2] == pytest.approx(1e-6) and torch.isfinite(out).all()
    if os.path.exists(mk.REF):
        spec = importlib.util.spec_from_file_location("ref_datagen", mk.REF)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        for kw in ({}, {"zero_mean": True}, {"is_symmetric": True}, {"zero_mean": True, "is_symmetric": True}):
            a = normalize_tensor(t, **kw)
            b = mod.DatasetGenerator.normalize_tensor(None, t, **kw)
            for x, y in zip(a, b):
                assert x.dtype == y.dtype and x.shape == y.shape and torch.equal(x, y), kw
