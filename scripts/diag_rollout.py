import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import coveragecontrol_b200 as cc
from coveragecontrol_b200 import BatchedCoverageSystem, BatchedCVT, _lib
N, R, E, steps, C = 512, 10, 4, 7, 32
p = cc.Parameters(); p.pWorldMapSize = p.pRobotMapSize = N; p.pNumRobots = R
rng = np.random.default_rng(5)
worlds = [rng.random((N, N)).astype(np.float32) for _ in range(E)]
pos = rng.uniform(1, N - 1, (E, R, 2))
def run(fused, noaux, obs_on=True, objective=True):
    os.environ["CCB_NO_AUX"] = "1" if noaux else "0"
    env = BatchedCoverageSystem(p, E, R)
    for e in range(E): env.set_world(e, worlds[e])
    env.reset(pos)
    ctrl = BatchedCVT(env, 2)
    obj = torch.zeros((steps, E), dtype=torch.float64, device="cuda")
    obs = torch.zeros((steps, E, R, 4, C, C), dtype=torch.float32, device="cuda")
    traj = []
    for s in range(steps):
        if fused:
            ctrl.rollout_obs(1, objective, obj[s].data_ptr(), C, True, obs[s].data_ptr() if obs_on else 0, 0)
        else:
            ctrl.rollout(1, objective, obj[s].data_ptr())
            if obs_on: _lib.check(_lib.load().ccb_env_cnn_input_dev(env._h, C, 1, obs[s].data_ptr()))
        traj.append(env.positions())
    torch.cuda.synchronize()
    return np.array(traj), obj.cpu().numpy(), obs.cpu().numpy()
base = run(False, True)
for name, args in [("sep,noaux again", (False, True)), ("sep,aux", (False, False)), ("fused,noaux", (True, True)), ("fused,aux", (True, False)),
                   ("fused,aux,no-obs", (True, False, False)), ("fused,aux,no-objective", (True, False, True, False))]:
    r = run(*args)
    d = np.abs(r[0] - base[0]).reshape(steps, -1).max(1)
    print(name, "pos maxdiff per step", d, "obj maxrel", np.abs(r[1] / np.where(base[1] == 0, 1, base[1]) - 1).max() if args[-1] is not False or len(args) < 4 else "-", "obs maxdiff", np.abs(r[2] - base[2]).max())
