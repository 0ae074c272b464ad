"""Golden LPAC observation (BASELINE config 2: local / comm_x / comm_y / obstacle channels, 32 x 32) produced by the
REFERENCE's own python/coverage_control/coverage_env_utils.py -- get_maps, get_communication_maps, resize_maps,
get_weights, run unmodified in this container -- over a stand-in environment that hands it the oracle's raw maps
(oracle/cc_oracle.c; the CUDA raw maps are held bit-equal to those by tests/test_gpu_observation.py).
The world is a closed-form pattern and the robots follow stored actions from stored start positions, so the GPU test
can rebuild the identical state through the drop-in API (StepActions).  Output: lpac_obs_golden.npz.

    python tests/golden/make_lpac_obs_golden.py        (needs /root/reference; torch_geometric is stubbed)
"""
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/python/coverage_control/coverage_env_utils.py"

N, R, S, L, C, COMM = 384, 6, 64, 256, 32, 128.0


def world_pattern(n=N):
    """closed form, exactly representable: both sides build the same float32 map"""
    i = np.arange(n, dtype=np.int64)
    v = ((i[:, None] * 37 + i[None, :] * 91 + (i[:, None] * i[None, :]) % 17) % 251).astype(np.float32) / np.float32(251)
    v[(i[:, None] // 48 + i[None, :] // 48) % 3 == 0] = 0.0  # patches of exact zeros
    return v


def start_and_actions(steps=9, seed=0):
    """start positions and per-step actions (speed <= 4.5 < pMaxRobotSpeed): both sides step with StepActions"""
    rng = np.random.default_rng(seed)
    pos = rng.uniform(20, N - 20, (R, 2))
    pos[0] = [3.5, 200.25]          # windows clipped by the borders
    pos[1] = [N - 2.0, N - 7.5]
    pos[2] = pos[3] + [8.0, 0.0]    # comm-map index on an exact half (round-half-even vs half-away)
    ang = rng.uniform(0, 2 * np.pi, (steps, R))
    speed = rng.uniform(0.5, 4.5, (steps, R))
    act = np.stack([np.cos(ang), np.sin(ang)], -1) * speed[..., None]
    act[:, 2] = act[:, 3]           # the pair keeps its offset of 8 cells
    act[:, 0, 0] = -np.abs(act[:, 0, 0])  # robot 0 runs into the border clamp
    return pos, act


def params(cc_or_orc):
    from oracle import oracle as orc

    return orc.default_params(pWorldMapSize=N, pRobotMapSize=N, pSensorSize=S, pLocalMapSize=L, pNumRobots=R,
                              pCommunicationRange=COMM)


def main():
    import coverage_control as cc
    from oracle import oracle as orc

    tg = types.ModuleType("torch_geometric")
    tg.data = types.ModuleType("torch_geometric.data")
    tg.data.Data = lambda **kw: types.SimpleNamespace(**kw)
    sys.modules.setdefault("torch_geometric", tg)
    sys.modules.setdefault("torch_geometric.data", tg.data)
    spec = importlib.util.spec_from_file_location("coverage_control._reference_env_utils", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    ref = mod.CoverageEnvUtils

    p = params(None)
    start, actions = start_and_actions()
    oenv = orc.Env(p, world_pattern(), start)
    for a in actions:
        assert oenv.step_actions(a) == 0

    class StandIn:
        def GetNumRobots(self):
            return R

        def GetRobotLocalMap(self, r):
            return oenv.local_map(r)

        def GetRobotObstacleMap(self, r):
            return oenv.obstacle_map(r)

        def GetCommunicationMaps(self, size):
            return None  # computed and discarded by the reference (coverage_env_utils.py:259-263)

        def GetRelativePositonsNeighbors(self, r):
            return cc.PointVector(oenv.neighbors(r))

        def GetRobotPositions(self):
            return cc.PointVector(oenv.positions())

    cp = cc.Parameters()
    cp.pWorldMapSize = cp.pRobotMapSize = N
    cp.pSensorSize, cp.pLocalMapSize, cp.pNumRobots, cp.pCommunicationRange = S, L, R, COMM
    env = StandIn()
    maps = ref.get_maps(env, cp, C, True).numpy()
    weights = ref.get_weights(env, cp).numpy()
    # get_torch_geometric_data (coverage_env_utils.py:337-377) with the stub Data container: x is `maps` again
    data = ref.get_torch_geometric_data(env, cp, True, True, C)
    assert np.array_equal(data.x.numpy(), maps)
    np.savez_compressed(os.path.join(HERE, "lpac_obs_golden.npz"), start=start, actions=actions, maps=maps, weights=weights,
                        final_positions=oenv.positions(), edge_index=data.edge_index.numpy(),
                        edge_weight=data.edge_weight.numpy(), pos=data.pos.numpy())
    print(maps.shape, float(np.abs(maps).sum()), weights.shape)


if __name__ == "__main__":
    main()
