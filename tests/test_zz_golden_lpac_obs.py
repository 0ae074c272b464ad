"""LPAC observation against tests/golden/lpac_obs_golden.npz: the output of the REFERENCE's own
coverage_env_utils.py (get_maps incl. get_communication_maps and resize_maps, get_weights), run unmodified by
tests/golden/make_lpac_obs_golden.py over the oracle's raw maps.  CPU: the fixture is reproduced from the oracle plus
the installed torchvision (so it is not stale).  GPU: the CUDA path (k_cnn_fused8 / k_comm_maps / k_graph) through
the drop-in API on the same world and robot trajectory.  Tolerance: fp32 summation order of the 16 x 16 taps,
rtol 1e-5 (torch.allclose's default, as the reference's test_coverage_env_utils.py) + atol 1e-6.
(The file name sorts last on purpose: the GPU half was written after round 1's GPU minutes were spent.)"""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_lpac_obs_golden as mk  # noqa: E402

GOLD = np.load(os.path.join(HERE, "golden", "lpac_obs_golden.npz"))


def test_fixture_is_reproduced_by_oracle_and_torchvision():
    import torch
    import torchvision

    from oracle import oracle as orc

    oenv = orc.Env(mk.params(None), mk.world_pattern(), GOLD["start"])
    for a in GOLD["actions"]:
        assert oenv.step_actions(a) == 0
    assert np.array_equal(oenv.positions(), GOLD["final_positions"])
    assert oenv.positions()[0, 0] == pytest.approx(1e-4)  # robot 0 ended on the border clamp
    rs = lambda t: torchvision.transforms.functional.resize(  # noqa: E731
        t, (mk.C, mk.C), interpolation=torchvision.transforms.InterpolationMode.BILINEAR, antialias=True).numpy()
    local = rs(torch.from_numpy(np.stack([oenv.local_map(r) for r in range(mk.R)])))
    obst = rs(torch.from_numpy(np.stack([oenv.obstacle_map(r) for r in range(mk.R)])))
    maps = GOLD["maps"]
    assert np.allclose(local, maps[:, 0], rtol=1e-5, atol=1e-6) and np.allclose(obst, maps[:, 3], rtol=1e-5, atol=1e-6)
    for r in range(mk.R):
        ox, oy = oenv.comm_maps(r, mk.C, "py")
        assert np.array_equal(ox, maps[r, 1]) and np.array_equal(oy, maps[r, 2])
    # the pair 8 cells apart lands on an exact half: the Python semantics (half to even) are the golden ones
    assert maps[3, 1].any() and maps[2, 1].any()
    assert (GOLD["weights"] > 0).sum() == 6
    # the graph fields follow from the weights and the positions (coverage_env_utils.py:362-369)
    w = torch.from_numpy(GOLD["weights"]).to_sparse().coalesce()
    assert np.array_equal(GOLD["edge_index"], w.indices().numpy()) and np.array_equal(GOLD["edge_weight"], w.values().numpy())
    assert np.allclose(GOLD["pos"], (GOLD["final_positions"].astype(np.float32) + mk.N / 2.0) / mk.N, atol=1e-7)


@pytest.mark.gpu
def test_cuda_observation_equals_the_reference_output():
    import torch

    import coverage_control as cc

    p = cc.Parameters()
    p.pWorldMapSize = p.pRobotMapSize = mk.N
    p.pSensorSize, p.pLocalMapSize, p.pNumRobots, p.pCommunicationRange = mk.S, mk.L, mk.R, mk.COMM
    env = cc.CoverageSystem(p, cc.WorldIDF(p, mk.world_pattern()), cc.PointVector(GOLD["start"]))
    for a in GOLD["actions"]:
        assert not env.StepActions(cc.PointVector(a))
    assert np.allclose(np.array(env.GetRobotPositions()), GOLD["final_positions"], rtol=0, atol=1e-9)
    U = cc.CoverageEnvUtils
    maps = U.get_maps(env, p, mk.C, True).cpu()
    gold = torch.from_numpy(GOLD["maps"])
    assert maps.shape == gold.shape
    assert torch.allclose(maps[:, 0], gold[:, 0], rtol=1e-5, atol=1e-6), "local channel"
    assert torch.allclose(maps[:, 3], gold[:, 3], rtol=1e-5, atol=1e-6), "obstacle channel"
    assert torch.allclose(maps[:, 1:3], gold[:, 1:3], rtol=1e-6, atol=1e-7), "communication channels"
    assert torch.equal(maps[:, 1:3] != 0, gold[:, 1:3] != 0)
    w_host = U.get_weights(env, p)
    w_dev = U.get_weights(env, p, device="cpu")
    gw = torch.from_numpy(GOLD["weights"])
    assert torch.allclose(w_host, gw, rtol=0, atol=1e-6)
    assert torch.allclose(w_dev, gw, rtol=0, atol=1e-5) and torch.equal(w_dev != 0, gw != 0)
    # the graph the GNN receives (coverage_env_utils.py:337-377): same edges in the same order, weights, positions
    g = U.get_torch_geometric_data(env, p, True, True, mk.C)
    assert torch.equal(g.edge_index.cpu(), torch.from_numpy(GOLD["edge_index"]))
    assert torch.allclose(g.edge_weight.cpu(), torch.from_numpy(GOLD["edge_weight"]), rtol=0, atol=1e-5)
    assert torch.allclose(g.pos.cpu(), torch.from_numpy(GOLD["pos"]), rtol=0, atol=1e-6)
    assert torch.allclose(g.x.cpu(), maps)
