"""Small driver for ncu: builds a batch of C1/C2 environments and launches one kernel of the step
several times through the profiling hook (ccb_bench_kernel).

    python scripts/prof_kernels.py --which 0 --envs 64 --iters 5 [--steps 20]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
import coveragecontrol_b200 as ccb  # noqa: E402
from coveragecontrol_b200 import BatchedCoverageSystem, BatchedCVT, _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--which", type=int, default=0)
ap.add_argument("--envs", type=int, default=64)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--steps", type=int, default=20, help="decentralized roll-out steps before profiling")
args = ap.parse_args()

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
ctrl.rollout(args.steps)
obs = torch.empty((E, bench.N_ROBOTS, 4, 32, 32), device="cuda")
torch.cuda.synchronize()
s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
if args.which in (6, 9):  # the reduce / fix-up stages run on the tables of a full whole-map pass
    _lib.check(lib.ccb_bench_kernel(batch._h, 0, 1, obs.data_ptr()))
_lib.check(lib.ccb_bench_kernel(batch._h, args.which, 2, obs.data_ptr()))
torch.cuda.synchronize()
torch.cuda.profiler.start()  # ncu --profile-from-start off captures only the launches below
s.record()
_lib.check(lib.ccb_bench_kernel(batch._h, args.which, args.iters, obs.data_ptr()))
t.record()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(f"kernel {args.which}: {s.elapsed_time(t) / args.iters:.4f} ms per launch, E={E}")
