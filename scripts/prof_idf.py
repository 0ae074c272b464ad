"""WorldIDF generation alone, for ncu (pipe utilisation of k_idf_generate) and for the IDF roofline line:

    python scripts/prof_idf.py [--world 1024 --features 32 --polygons 0 --mode 0 --iters 5]

Prints ms per GenerateMap (three kernels: generate, factor, normalise) and the algorithmic FLOP/s of
SURVEY.md 8(d): flops = 8 N^2 F + 16 N^2 F p_in + 2 N^2, p_in measured on the host from the features."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import coveragecontrol_b200 as ccb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--world", type=int, default=1024)
ap.add_argument("--features", type=int, default=32)
ap.add_argument("--polygons", type=int, default=0)
ap.add_argument("--mode", type=int, default=0)
ap.add_argument("--iters", type=int, default=5)
a = ap.parse_args()
N, F = a.world, a.features
p = ccb.Parameters()
p.pWorldMapSize = p.pRobotMapSize = N
rng = np.random.default_rng(0)
idf = ccb.WorldIDF(p)
sig = rng.uniform(40, 60, F)
for k in range(F):
    m = rng.uniform(0, N, 2)
    idf.AddNormalDistribution(ccb.BivariateNormalDistribution(ccb.Point2(m[0], m[1]), float(sig[k]), float(rng.uniform(6, 10))))
for _ in range(a.polygons):
    c = rng.uniform(64, N - 64, 2)
    nv = int(rng.integers(4, 11))
    ang = np.sort(rng.uniform(0, 2 * np.pi, nv))
    rad = rng.uniform(16, 64, nv)
    idf.AddUniformDistributionPolygon(ccb.PolygonFeature(ccb.PointVector(np.stack([c[0] + rad * np.cos(ang), c[1] + rad * np.sin(ang)], 1)),
                                                         float(5e-5 * rng.uniform(6, 10))))
for _ in range(2):
    idf.GenerateMap(a.mode)
torch.cuda.synchronize()
torch.cuda.profiler.start()
s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(a.iters):
    idf.GenerateMap(a.mode)
t.record()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
ms = s.elapsed_time(t) / a.iters
p_in = float(np.minimum(np.pi * (2 * sig) ** 2 / (N * N), 1.0).mean())  # truncation at 2 sigma
flops = 8.0 * N * N * F + 16.0 * N * N * F * p_in + 2.0 * N * N
print(f"IDF N={N} F={F} polygons={a.polygons} mode={a.mode}: {ms:.4f} ms per GenerateMap, p_in={p_in:.4f}, "
      f"algorithmic {flops / 1e9:.3f} GFLOP -> {flops / (ms * 1e-3) / 1e12:.2f} TFLOP/s, {4.0 * N * N / (ms * 1e-3) / 1e9:.0f} GB/s written")
