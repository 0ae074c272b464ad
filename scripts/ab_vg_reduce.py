"""A/B timing of the site-ordered whole-map Voronoi chain on the bench workload: k_vg_reduce alone, k_vg_partition
alone and the whole partition + items + reduce + fix-up chain (CUDA events, ccb_bench_kernel), for a list of
CCB_VG_* settings (CCB_VG_WARPS = warps per CTA, CCB_VG_GRIDDIV = grid divisor, CCB_VG_GENERIC = tile-ordered
kernel), plus the spread of the objectives between the settings.

    python scripts/ab_vg_reduce.py [--envs 1024] [--iters 20] [--steps 20] [--configs "WARPS=1;WARPS=4;GENERIC=1"]
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
import coveragecontrol_b200 as ccb  # noqa: E402
from coveragecontrol_b200 import BatchedCoverageSystem, BatchedCVT, _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=1024)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--steps", type=int, default=20, help="decentralized roll-out steps before timing")
ap.add_argument("--configs", default="WARPS=1;WARPS=4", help="';'-separated CCB_VG_* settings, e.g. 'WARPS=1,GRIDDIV=4'")
args = ap.parse_args()

lib = _lib.load()
p = bench.toml_params(ccb)
E = args.envs
worlds, pos = [], np.zeros((E, bench.N_ROBOTS, 2))
for e in range(E):
    bnd, pos[e] = bench.synth_inputs(e)
    idf = ccb.WorldIDF(p)
    for b in bnd:
        idf.AddNormalDistribution(ccb.BivariateNormalDistribution(ccb.Point2(b[0], b[1]), float(b[2]), float(b[4])))
    idf.GenerateMap()
    worlds.append(idf)


def timed(batch, which, iters, obs):
    s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    _lib.check(lib.ccb_bench_kernel(batch._h, which, 3, obs.data_ptr()))
    torch.cuda.synchronize()
    s.record()
    _lib.check(lib.ccb_bench_kernel(batch._h, which, iters, obs.data_ptr()))
    t.record()
    torch.cuda.synchronize()
    return s.elapsed_time(t) / iters


out, objs = {}, {}
for mode in args.configs.split(";"):
    for k in [k for k in os.environ if k.startswith("CCB_VG_")]:
        del os.environ[k]
    for kv in mode.split(","):
        if kv:
            os.environ["CCB_VG_" + kv.split("=")[0]] = kv.split("=")[1]
    batch = BatchedCoverageSystem(p, E, bench.N_ROBOTS)
    for e in range(E):
        batch.set_world(e, worlds[e])
    batch.reset(pos)
    ctrl = BatchedCVT(batch, 2)
    ctrl.rollout(args.steps)
    obs = torch.empty((E, bench.N_ROBOTS, 4, 32, 32), device="cuda")
    torch.cuda.synchronize()
    _lib.check(lib.ccb_bench_kernel(batch._h, 0, 1, obs.data_ptr()))
    r = {"reduce_ms": timed(batch, 6, args.iters, obs), "partition_ms": timed(batch, 7, args.iters, obs),
         "chain_ms": timed(batch, 0, args.iters, obs), "explored_chain_ms": timed(batch, 1, args.iters, obs)}
    r["reduce_gbs"] = 4.0 * 1024 * 1024 * E / (r["reduce_ms"] * 1e-3) / 1e9
    objs[mode] = np.array(batch.objective())
    out[mode] = r
    del ctrl, batch, obs
    torch.cuda.empty_cache()
base = objs[next(iter(objs))]
for mode, o in objs.items():
    out[mode]["objective_max_rel_diff_vs_first"] = float(np.max(np.abs(o - base) / np.abs(base)))
print(json.dumps({"envs": E, "iters": args.iters, "modes": out}))
