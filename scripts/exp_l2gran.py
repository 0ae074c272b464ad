"""Experiment: does cudaLimitMaxL2FetchGranularity (32 / 64 / 128 bytes) change the sparse kernels' time?
    python scripts/exp_l2gran.py --gran 32 [--envs 512]
Sets the limit before the first kernel, then times the fused partition + moments kernel (7), the fused observation
kernel (8) and the streaming reduce for comparison."""
import argparse
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
import coveragecontrol_b200 as ccb  # noqa: E402
from coveragecontrol_b200 import BatchedCoverageSystem, BatchedCVT, _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--gran", type=int, default=0)
ap.add_argument("--envs", type=int, default=512)
args = ap.parse_args()
torch.cuda.init()
rt = C.CDLL("libcudart.so.12")
if args.gran:
    rc = rt.cudaDeviceSetLimit(C.c_int(5), C.c_size_t(args.gran))  # cudaLimitMaxL2FetchGranularity = 0x05
    v = C.c_size_t(0)
    rt.cudaDeviceGetLimit(C.byref(v), C.c_int(5))
    print("set rc", rc, "limit now", v.value)
lib = _lib.load()
p = bench.toml_params(ccb)
E = args.envs
batch = BatchedCoverageSystem(p, E, bench.N_ROBOTS)
pos = np.zeros((E, bench.N_ROBOTS, 2))
for e in range(E):
    bnd, pos[e] = bench.synth_inputs(e)
    idf = ccb.WorldIDF(p)
    for b in bnd:
        idf.AddNormalDistribution(ccb.BivariateNormalDistribution(ccb.Point2(b[0], b[1]), float(b[2]), float(b[4])))
    idf.GenerateMap()
    batch.set_world(e, idf)
batch.reset(pos)
ctrl = BatchedCVT(batch, 2)
ctrl.rollout(20)
obs = torch.empty((E, bench.N_ROBOTS, 4, 32, 32), device="cuda")
torch.cuda.synchronize()
for which in (7, 8, 4, 3):
    _lib.check(lib.ccb_bench_kernel(batch._h, 0, 1, obs.data_ptr()))
    _lib.check(lib.ccb_bench_kernel(batch._h, which, 3, obs.data_ptr()))
    torch.cuda.synchronize()
    s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    _lib.check(lib.ccb_bench_kernel(batch._h, which, 10, obs.data_ptr()))
    t.record()
    torch.cuda.synchronize()
    print(f"gran {args.gran} kernel {which}: {s.elapsed_time(t) / 10:.4f} ms per launch, E={E}")
