"""BASELINE config 5: world-size sweep -- IDF generation time and whole-map Voronoi mass/centroid
throughput against the HBM roofline, one environment per configuration.

    python scripts/sweep_c5.py > gpurun_out/sweep_c5.md

Columns: IDF generation (mode 0, the reference kernel's arithmetic) in ms and Gcell-feature/s; the
whole-map Voronoi label + reduce pass (ccb_bench_kernel which = 0: everything ComputeVoronoiCells launches
before the finalisation) in ms, algorithmic GB/s (4 N^2 bytes) and fraction of the measured HBM peak, for
the tile-ordered kernel and, where it applies (R <= 64), the site-ordered one."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import coveragecontrol_b200 as ccb  # noqa: E402
from coveragecontrol_b200 import BatchedCoverageSystem, _lib  # noqa: E402

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
lib = _lib.load()


def timed(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    t.record()
    torch.cuda.synchronize()
    return s.elapsed_time(t) / iters


def one(N, R, E=1):
    p = ccb.Parameters()
    p.pWorldMapSize = p.pRobotMapSize = N
    p.pNumRobots = R
    p.pUpdateExplorationMap = False
    rng = np.random.default_rng(N + R)
    scale = N / 1024.0
    idf = ccb.WorldIDF(p)
    F = R
    for _ in range(F):
        m = rng.uniform(0, N, 2)
        idf.AddNormalDistribution(ccb.BivariateNormalDistribution(ccb.Point2(m[0], m[1]), float(rng.uniform(40, 60) * min(scale, 4.0)),
                                                                  float(rng.uniform(6, 10))))
    idf_ms = timed(lambda: idf.GenerateMap(), iters=5, warm=2)
    row = {"N": N, "R": R, "F": F, "E": E, "idf_ms": idf_ms, "idf_gcf": N * N * F / idf_ms / 1e6}
    for mode, name in (("1", "tile"), ("0", "site")):
        if name == "site" and (R > 64 or N % 4):
            continue
        os.environ["CCB_VG_GENERIC"] = mode
        env = BatchedCoverageSystem(p, E, R)
        for e in range(E):
            env.set_world(e, idf)
        env.reset(rng.uniform(1e-4, N - 1e-4, (E, R, 2)))
        ms = timed(lambda: _lib.check(lib.ccb_bench_kernel(env._h, 0, 1, None)))
        gbs = 4.0 * N * N * E / (ms * 1e-3) / 1e9
        row[name] = (ms, gbs, gbs / PEAK)
        del env
    os.environ.pop("CCB_VG_GENERIC", None)
    return row


one(512, 16)  # warm-up (context, allocator pools): not reported
print(f"HBM peak used: {PEAK} GB/s\n")
print("| N | robots = features | envs | IDF ms | IDF Gcell-feature/s | tile-ordered ms | GB/s | of peak | site-ordered ms | GB/s | of peak |")
print("|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
for N, R, E in [(512, 16, 1), (1024, 32, 1), (2048, 64, 1), (4096, 64, 1), (4096, 128, 1), (8192, 64, 1), (8192, 256, 1), (16384, 64, 1),
                (16384, 1024, 1), (512, 16, 256), (1024, 32, 128), (2048, 64, 32)]:
    r = one(N, R, E)
    t = r.get("tile")
    s = r.get("site")
    f = lambda x: f"{x[0]:.4f} | {x[1]:.0f} | {100 * x[2]:.1f}%" if x else "- | - | -"  # noqa: E731
    print(f"| {N} | {R} | {E} | {r['idf_ms']:.3f} | {r['idf_gcf']:.1f} | {f(t)} | {f(s)} |", flush=True)
