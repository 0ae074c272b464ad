"""C4 step on ONE GPU, stage by stage (CUDA events): the band partials for bands of N / G rows (G = 1, 2, 4, 8),
the finalisation, the controller and the move + marks.  Shows which part of the row-band step shards."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    import coveragecontrol_b200 as ccb
    from bench import c4_features, make_idf
    from coveragecontrol_b200 import BatchedCoverageSystem, BatchedCVT, _lib

    lib = _lib.load()
    n, r = 8192, 256
    p = ccb.Parameters()
    p.pWorldMapSize = p.pRobotMapSize = n
    p.pNumRobots = r
    p.pUpdateExplorationMap = False
    p.pTimeStep = 0.2
    bnd, polys = c4_features(n, 1024, 64)
    idf = make_idf(ccb, p, bnd, polys, 0)
    idf.GenerateMap()
    pos = np.random.default_rng(1).uniform(1e-4, n - 1e-4, (1, r, 2))
    stream = torch.cuda.Stream(device=0)
    env = BatchedCoverageSystem(p, 1, r, device=0, stream=stream.cuda_stream)
    env.set_world(0, idf)
    env.reset(pos)
    _lib.check(lib.ccb_env_set_band_mode(env._h, 1))
    sums = torch.zeros((1, r, 4), dtype=torch.float64, device="cuda")
    _lib.check(lib.ccb_env_voronoi_partials_dev(env._h, 0, 0, n, sums.data_ptr()))
    _lib.check(lib.ccb_env_finalize_from_partials_dev(env._h, sums.data_ptr()))
    ctrl = BatchedCVT(env, 0)
    s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn, iters=50):
        with torch.cuda.stream(stream):
            for _ in range(5):
                fn()
            s.record()
            for _ in range(iters):
                fn()
            t.record()
        torch.cuda.synchronize()
        return s.elapsed_time(t) / iters * 1e3

    out = {}
    for g in (1, 2, 4, 8):
        rows = n // g
        for which in (0, g - 1) if g > 1 else (0,):
            out[f"partials_rows_{rows}_band{which}_us"] = timed(lambda: _lib.check(lib.ccb_env_voronoi_partials_dev(env._h, 0, which * rows, (which + 1) * rows, sums.data_ptr())))
    _lib.check(lib.ccb_env_voronoi_partials_dev(env._h, 0, 0, n, sums.data_ptr()))
    out["finalize_us"] = timed(lambda: _lib.check(lib.ccb_env_finalize_from_partials_dev(env._h, sums.data_ptr())))
    out["compute_actions_us"] = timed(lambda: _lib.check(lib.ccb_ctrl_compute_actions(ctrl._h)))
    out["step_actions_us"] = timed(lambda: _lib.check(lib.ccb_env_step_actions_dev(env._h, lib.ccb_ctrl_actions_dev(ctrl._h))))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
