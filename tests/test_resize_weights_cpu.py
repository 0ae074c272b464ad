"""The antialiased bilinear taps of the observation kernels (csrc/ccb200.cu resize_weights, uploaded to k_cnn_fused8 /
k_cnn_local_obstacle) on the HOST through ccb_selftest_resize_weights, against the INSTALLED torchvision -- the
operator the reference applies (coverage_env_utils.py:97-141 resize_maps) -- and against the committed fixture
tests/golden/resize_256_32_weights.npy (made by tests/golden/make_golden.py from the same call).  The reference's own
fixtures for this (data/coverage_env_utils/resized_local_maps.pt) are not in its tree."""
import ctypes as C
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _weights(L, Cc):
    from coveragecontrol_b200 import _lib

    lib = _lib.load()
    xmin, xsize = np.zeros(Cc, np.int32), np.zeros(Cc, np.int32)
    w = np.zeros(Cc * (2 * L + 8), np.float32)
    ip, fp = C.POINTER(C.c_int), C.POINTER(C.c_float)
    kmax = lib.ccb_selftest_resize_weights(L, Cc, xmin.ctypes.data_as(ip), xsize.ctypes.data_as(ip), w.ctypes.data_as(fp), len(w))
    assert kmax > 0
    w = w[: Cc * kmax].reshape(Cc, kmax)
    dense = np.zeros((Cc, L), np.float32)
    for o in range(Cc):
        assert 0 <= xmin[o] and xmin[o] + xsize[o] <= L and xsize[o] <= kmax
        dense[o, xmin[o]:xmin[o] + xsize[o]] = w[o, : xsize[o]]
        assert not w[o, xsize[o]:].any()
    return dense


def _torchvision_matrix(L, Cc):
    import torch
    import torchvision

    eye = torch.eye(L).unsqueeze(0)  # one image whose column k is the k-th unit vector
    out = torchvision.transforms.functional.resize(
        eye, (Cc, L), interpolation=torchvision.transforms.InterpolationMode.BILINEAR, antialias=True)[0]
    return out.numpy()


@pytest.mark.parametrize("L,Cc", [(256, 32), (128, 32), (64, 32), (100, 32), (48, 32), (32, 32), (256, 16), (96, 24), (200, 7)])
def test_taps_equal_torchvision(L, Cc):
    got, ref = _weights(L, Cc), _torchvision_matrix(L, Cc)
    assert np.array_equal(got != 0, np.abs(ref) > 1e-9), "tap ranges differ"
    assert np.abs(got - ref).max() <= 2e-7
    assert np.allclose(got.sum(1), 1.0, atol=1e-6)


def test_committed_fixture():
    ref = np.load(os.path.join(HERE, "golden", "resize_256_32_weights.npy"))
    got = _weights(256, 32)
    assert ref.shape == got.shape == (32, 256)
    assert np.abs(got - ref).max() <= 2e-7
    # interior output o reads inputs [8 o - 4, 8 o + 12) with weights (0.5 .. 7.5, 7.5 .. 0.5) / 64,
    # the border rows truncate and renormalise
    tri = np.concatenate([np.arange(0.5, 8), np.arange(7.5, 0, -1)]) / 64.0
    assert np.allclose(got[5, 36:52], tri, atol=1e-7) and not got[5, :36].any() and not got[5, 52:].any()
    assert np.allclose(got[0, :12], np.concatenate([np.arange(4.5, 8), np.arange(7.5, 0, -1)]) / 56.0, atol=1e-7)


def test_separable_resize_of_a_random_window_equals_torchvision():
    """width pass then height pass with these taps (the order the kernels and torch use) on a random 256 x 256 window"""
    import torch
    import torchvision

    rng = np.random.default_rng(0)
    x = rng.random((256, 256)).astype(np.float32)
    w = _weights(256, 32)
    tmp = (x[:, None, :] * w[None, :, :]).sum(-1, dtype=np.float32)       # [256, 32]: width
    out = (w[:, :, None] * tmp[None, :, :]).sum(1, dtype=np.float32)      # [32, 32]: height
    ref = torchvision.transforms.functional.resize(
        torch.from_numpy(x)[None], (32, 32), interpolation=torchvision.transforms.InterpolationMode.BILINEAR, antialias=True)[0].numpy()
    assert np.abs(out - ref).max() <= 1e-5  # fp32 summation order only
