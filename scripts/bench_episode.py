"""How the C2 step time evolves over an episode (the robots' seen regions grow, and with them the work of
the observation kernel): env-steps/s of bench.py's device step after 0, 100, 300, 600 steps."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
import coveragecontrol_b200 as ccb  # noqa: E402
from coveragecontrol_b200 import BatchedCoverageSystem, BatchedCVT  # noqa: E402

E = int(sys.argv[1]) if len(sys.argv) > 1 else 256
p = bench.toml_params(ccb)
env = BatchedCoverageSystem(p, E, bench.N_ROBOTS)
pos = np.zeros((E, bench.N_ROBOTS, 2))
for e in range(E):
    bnd, pos[e] = bench.synth_inputs(e)
    idf = ccb.WorldIDF(p)
    for b in bnd:
        idf.AddNormalDistribution(ccb.BivariateNormalDistribution(ccb.Point2(b[0], b[1]), float(b[2]), float(b[4])))
    idf.GenerateMap()
    env.set_world(e, idf)
env.reset(pos)
ctrl = BatchedCVT(env, 2)
obs = torch.empty((E, bench.N_ROBOTS, 4, 32, 32), dtype=torch.float32, device="cuda")
obj = torch.empty((1, E), dtype=torch.float64, device="cuda")
done, rows = 0, []
for target in (0, 100, 300, 600):
    while done < target:
        ctrl.rollout_obs(1, True, obj.data_ptr(), 32, True, obs.data_ptr(), 0)
        done += 1
    for _ in range(3):
        ctrl.rollout_obs(1, True, obj.data_ptr(), 32, True, obs.data_ptr(), 0)
    torch.cuda.synchronize()
    s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        ctrl.rollout_obs(1, True, obj.data_ptr(), 32, True, obs.data_ptr(), 0)
    t.record()
    torch.cuda.synchronize()
    done += 13
    ms = s.elapsed_time(t) / 10
    rows.append({"after_steps": target, "ms_per_step": ms, "env_steps_per_s": E / (ms * 1e-3), "converged_envs": int(ctrl.converged().sum())})
print(json.dumps({"envs": E, "rows": rows}))
