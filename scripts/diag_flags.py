import os, sys, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
import coveragecontrol_b200 as ccb
from coveragecontrol_b200 import BatchedCoverageSystem, BatchedCVT, _lib
lib = _lib.load(); p = bench.toml_params(ccb); E = 32
batch = BatchedCoverageSystem(p, E, bench.N_ROBOTS)
pos = np.zeros((E, bench.N_ROBOTS, 2))
for e in range(E):
    bnd, pos[e] = bench.synth_inputs(e)
    idf = ccb.WorldIDF(p)
    for b in bnd: idf.AddNormalDistribution(ccb.BivariateNormalDistribution(ccb.Point2(b[0], b[1]), float(b[2]), float(b[4])))
    idf.GenerateMap(); batch.set_world(e, idf)
batch.reset(pos)
ctrl = BatchedCVT(batch, 2)
for steps in (0, 20, 60, 100):
    if steps: ctrl.rollout(steps, False)
    obj = batch.objective()
    cnt = C.c_int(0); rows = np.zeros(4096, np.int32)
    _lib.check(lib.ccb_env_flagged_rows(batch._h, C.byref(cnt), rows.ctypes.data_as(C.POINTER(C.c_int)), 4096))
    n = cnt.value
    es = rows[:min(n, 4096)] // 1024
    print("after", steps, "more steps: flagged rows", n, "per env", np.bincount(es, minlength=E)[:16])
    if n:
        P = batch.positions()
        e0 = int(es[0]); r0 = int(rows[0] % 1024)
        np.save("gpurun_out/diag_pos.npy", P[e0]); print("env", e0, "row", r0)
        d = np.abs(P[e0][:, None, :] - P[e0][None, :, :]).sum(-1) + np.eye(32) * 1e9
        i, j = np.unravel_index(d.argmin(), d.shape); print("closest pair", i, j, P[e0][i], P[e0][j])
