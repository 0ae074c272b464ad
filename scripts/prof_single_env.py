"""Single-environment latency (VERDICT r01 "next" 7): C1 (1024^2, 32 robots, 32 Gaussians) through the C-ABI, one
environment.  Prints the per-step time of (a) the API calls an evaluator makes, (b) roll-outs of 1 and 32 steps.
With --eager the kernels run outside CUDA graphs so an ncu launch list shows each of them."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--eager", action="store_true")
    ap.add_argument("--kind", type=int, default=0)
    ap.add_argument("--iters", type=int, default=200)
    a = ap.parse_args()
    import torch

    import coveragecontrol_b200 as ccb
    from bench import N_ROBOTS, synth_inputs
    from coveragecontrol_b200 import BatchedCoverageSystem, BatchedCVT, _lib

    lib = _lib.load()
    p = ccb.Parameters()
    one = BatchedCoverageSystem(p, 1, N_ROBOTS, device=0)
    if a.eager:
        one.set_option("no_graph", 1)
    bnd, pos = synth_inputs(0)
    idf = ccb.WorldIDF(p, device=0)
    for b in bnd:
        idf.AddNormalDistribution(ccb.BivariateNormalDistribution(ccb.Point2(b[0], b[1]), float(b[2]), float(b[4])))
    idf.GenerateMap()
    one.set_world(0, idf)
    one.reset(pos[None])
    c1 = BatchedCVT(one, a.kind)
    ha = np.empty((1, N_ROBOTS, 2))
    bad = C.c_int(-1)

    def api_step():
        c1.compute_actions()
        _lib.check(lib.ccb_ctrl_get_actions(c1._h, ha.ctypes.data_as(C.POINTER(C.c_double))))
        _lib.check(lib.ccb_env_step_actions(one._h, ha.ctypes.data_as(C.POINTER(C.c_double)), C.byref(bad)))

    def wall(fn, n):
        for _ in range(10):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / n * 1e6

    out = {"api_step_us": wall(api_step, a.iters)}
    out["rollout1_us"] = wall(lambda: c1.rollout(1), a.iters)
    out["rollout32_us_per_step"] = wall(lambda: c1.rollout(32), max(4, a.iters // 16)) / 32
    print(json.dumps(out))


if __name__ == "__main__":
    main()
