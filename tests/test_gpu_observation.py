"""LPAC observation tensors (coverage_env_utils.py) from the CUDA path vs oracle + torchvision."""
import numpy as np
import pytest

from conftest import random_world

pytestmark = pytest.mark.gpu


def _setup(N=1024, R=8, L=256, S=64, seed=21, steps=6):
    import coverage_control as cc
    from coverage_control.algorithms import DecentralizedCVT
    from oracle import oracle as orc

    p = cc.Parameters()
    p.pWorldMapSize = p.pRobotMapSize = N
    p.pNumRobots = R
    p.pLocalMapSize = L
    p.pSensorSize = S
    bnd, pos = random_world(seed, N, 10, R)
    pos[0] = [20.0, 30.0]
    pos[1] = [N - 40.0, 60.0]
    pos[2] = pos[3] + [8.0, -8.0]  # exact .5 index in the comm map: round-half-even vs half-away
    op = orc.params_from(p)
    wm, _ = orc.idf_generate(op, bnd, mode=1)
    oenv = orc.Env(op, wm, pos)
    env = cc.CoverageSystem(p, cc.WorldIDF(p, wm), cc.PointVector(pos))
    ctrl = DecentralizedCVT(p, env)
    for _ in range(steps):
        ctrl.ComputeActions()
        a = ctrl.GetActions()
        env.StepActions(a)
        oenv.step_actions(np.array(a))
    return cc, orc, p, env, oenv


def test_raw_local_obstacle_and_comm_maps():
    import torch

    cc, orc, p, env, oenv = _setup()
    U = cc.CoverageEnvUtils
    R = env.GetNumRobots()
    local = U.get_raw_local_maps(env, p).cpu().numpy()
    obst = U.get_raw_obstacle_maps(env, p).cpu().numpy()
    comm = U.get_communication_maps(env, p, 32).cpu().numpy()
    assert local.shape == (R, 256, 256) and comm.shape == (R, 2, 32, 32)
    for r in range(R):
        assert np.array_equal(local[r], oenv.local_map(r))
        assert np.array_equal(obst[r], oenv.obstacle_map(r))
        ox, oy = oenv.comm_maps(r, 32, "py")
        assert np.array_equal(comm[r, 0], ox) and np.array_equal(comm[r, 1], oy)
    cpp = env.GetCommunicationMaps(32)
    for r in range(R):
        ox, oy = oenv.comm_maps(r, 32, "cpp")
        assert np.array_equal(cpp[2 * r], ox) and np.array_equal(cpp[2 * r + 1], oy)
    # python semantics, restated with torch exactly as coverage_env_utils.py:164-186 does
    for r in range(R):
        nb = torch.from_numpy(oenv.neighbors(r).astype(np.float32))
        if len(nb) == 0:
            assert not comm[r].any()
            continue
        idx = torch.round(nb * 32 / (p.pCommunicationRange * p.pResolution * 2.0) + (32 / 2.0 - p.pResolution / 2.0))
        idx = torch.transpose(idx, 1, 0).long()
        val = nb / p.pCommunicationRange
        for ch in range(2):
            dense = torch.sparse_coo_tensor(idx, val[:, ch], torch.Size([32, 32])).to_dense().numpy()
            assert np.allclose(comm[r, ch], dense, rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("C", [32, 64])
def test_cnn_input_matches_torchvision_resize(C):
    import torch
    import torchvision

    cc, orc, p, env, oenv = _setup()
    U = cc.CoverageEnvUtils
    R = env.GetNumRobots()
    maps = U.get_maps(env, p, C, True).cpu()
    assert maps.shape == (R, 4, C, C)
    raw_local = torch.from_numpy(np.stack([oenv.local_map(r) for r in range(R)]))
    raw_obst = torch.from_numpy(np.stack([oenv.obstacle_map(r) for r in range(R)]))
    rs = lambda t: torchvision.transforms.functional.resize(  # noqa: E731
        t, (C, C), interpolation=torchvision.transforms.InterpolationMode.BILINEAR, antialias=True)
    # torch.allclose defaults, as the reference's test_coverage_env_utils.py uses
    assert torch.allclose(maps[:, 0], rs(raw_local), rtol=1e-5, atol=1e-8)
    assert torch.allclose(maps[:, 3], rs(raw_obst), rtol=1e-5, atol=1e-8)
    comm = U.get_communication_maps(env, p, C).cpu()
    assert torch.equal(maps[:, 1:3], comm)
    two = U.get_maps(env, p, C, False).cpu()
    assert torch.equal(two[:, 0], maps[:, 0]) and torch.equal(two[:, 1], maps[:, 3])


def test_weights_and_positions():
    import torch

    cc, orc, p, env, oenv = _setup(steps=2)
    U = cc.CoverageEnvUtils
    pos = U.get_robot_positions(env)
    assert torch.allclose(pos, torch.from_numpy(oenv.positions().astype(np.float32)))
    w = U.get_weights(env, p)
    d = torch.cdist(pos, pos, 2)
    ref = torch.exp(-(d.square()) / (p.pCommunicationRange ** 2))
    ref[ref < 1.0 / np.e] = 0
    ref.fill_diagonal_(0)
    assert torch.allclose(w, ref)


def test_graph_data_on_device():
    """get_weights(device=...) and get_torch_geometric_data (reference coverage_env_utils.py:313-377): the k_graph
    kernel against the reference's own torch expressions on the host; tolerance 1e-5 on float32 weights (torch's
    cdist goes through a matrix product for more than 25 points), edges compared away from the 1/e threshold."""
    import torch

    cc, orc, p, env, oenv = _setup(steps=3)
    U = cc.CoverageEnvUtils
    host = U.get_weights(env, p)
    dev = U.get_weights(env, p, device="cuda")
    assert dev.is_cuda and dev.shape == host.shape and dev.dtype == torch.float32
    pos = U.get_robot_positions(env).double()
    d2 = torch.cdist(pos, pos, 2).square()
    exact = torch.exp(-d2 / p.pCommunicationRange ** 2)
    sure = (exact - 1.0 / np.e).abs() > 1e-4  # pairs whose side of the threshold does not hang on float32 rounding
    assert torch.allclose(dev.cpu()[sure], host[sure], rtol=0, atol=1e-5)
    assert torch.all(dev.cpu().diagonal() == 0)
    for use_cnn in (True, False):
        g = U.get_torch_geometric_data(env, p, use_cnn, True, 32)
        assert g.x.is_cuda and g.edge_index.is_cuda and g.edge_weight.is_cuda and g.pos.is_cuda
        want_x = U.get_maps(env, p, 32, True) if use_cnn else U.get_voronoi_features(env)
        assert torch.allclose(g.x.cpu(), want_x.cpu())
        ref_pos = (U.get_robot_positions(env) + p.pWorldMapSize / 2.0) / p.pWorldMapSize
        assert torch.allclose(g.pos.cpu(), ref_pos, rtol=0, atol=1e-6)
        sp = host.to_sparse().coalesce()
        if bool(sure.all()):
            assert torch.equal(g.edge_index.cpu(), sp.indices().long())
            assert torch.allclose(g.edge_weight.cpu(), sp.values().float(), rtol=0, atol=1e-5)
        assert g.edge_index.dtype == torch.long and g.edge_index.shape[0] == 2
        assert g.edge_weight.shape[0] == g.edge_index.shape[1] and torch.all(g.edge_weight >= 1.0 / np.e - 1e-6)
        assert g.num_nodes == p.pNumRobots
    with pytest.raises(ValueError):
        q = p.copy() if hasattr(p, "copy") else p
        q.pCommunicationRange = p.pCommunicationRange + 1.0
        U.get_weights(env, q, device="cuda")


@pytest.mark.parametrize("N,fill", [(1000, False), (1024, True), (600, False)])
def test_cnn_input_every_window_alignment(N, fill):
    """Window origins with every column misalignment (oy & 3), windows overhanging all four map edges,
    optional unknown-importance fill of unseen cells (robot_model.h:180-187)."""
    import torch
    import torchvision

    import coverage_control as cc
    from oracle import oracle as orc

    R = 12
    p = cc.Parameters()
    p.pWorldMapSize = p.pRobotMapSize = N
    p.pNumRobots = R
    p.pRobotMapUseUnknownImportance = fill
    bnd, pos = random_world(33, N, 10, R)
    for k in range(8):  # columns 8 apart cover every oy & 3 and oy & 31 phase, near and across the edges
        pos[k] = [40.0 + 97.0 * k, 3.0 + 1.0 * k + (N - 20.0 if k % 2 else 0.0)]
    pos[8] = [N - 5.0, N / 2 + 1.0]
    pos[9] = [4.0, N / 2 + 2.0]
    op = orc.params_from(p)
    wm, _ = orc.idf_generate(op, bnd, mode=1)
    oenv = orc.Env(op, wm, pos)
    env = cc.CoverageSystem(p, cc.WorldIDF(p, wm), cc.PointVector(pos))
    maps = cc.CoverageEnvUtils.get_maps(env, p, 32, True).cpu()
    raw_local = torch.from_numpy(np.stack([oenv.local_map(r) for r in range(R)]))
    raw_obst = torch.from_numpy(np.stack([oenv.obstacle_map(r) for r in range(R)]))
    rs = lambda t: torchvision.transforms.functional.resize(  # noqa: E731
        t, (32, 32), interpolation=torchvision.transforms.InterpolationMode.BILINEAR, antialias=True)
    assert torch.allclose(maps[:, 0], rs(raw_local), rtol=1e-5, atol=1e-8)
    assert torch.allclose(maps[:, 3], rs(raw_obst), rtol=1e-5, atol=1e-8)


def _net():
    import torch

    class Net(torch.nn.Module):
        """Stand-in for the LPAC network: any function of the graph data with the fields ControllerNN reads."""

        def __init__(self):
            super().__init__()
            self.register_buffer("actions_mean", torch.tensor([0.25, -0.5]))
            self.register_buffer("actions_std", torch.tensor([2.0, 1.5]))

        def forward(self, data):
            R = data.pos.shape[0]
            deg = torch.zeros(R, device=data.pos.device).index_add_(0, data.edge_index[0], data.edge_weight)
            feat = data.x.reshape(R, -1).mean(dim=1)
            return torch.stack([0.5 - data.pos[:, 0] + 0.1 * deg, 0.5 - data.pos[:, 1] + feat], dim=1)

    return Net()


def test_controller_nn_steps_on_the_device():
    """reference algorithms/controllers.py:91-160: ControllerNN.step with the observation, the graph and the actions left
    on the GPU against the reference's own sequence (actions.cpu() -> PointVector -> StepActions) on a twin environment."""
    import torch

    from coverage_control.algorithms import ControllerNN

    cc, orc, p, env, _ = _setup(steps=2)
    _, _, _, twin, _ = _setup(steps=2)
    U = cc.CoverageEnvUtils
    cfg = {"Name": "nn", "UseCNN": True, "UseCommMap": True, "CNNMapSize": 32}
    ctrl = ControllerNN(cfg, p, env, model=_net())
    assert ctrl.name == "nn" and next(ctrl.model.buffers()).is_cuda
    for _ in range(3):
        obj, converged = ctrl.step(env)
        data = U.get_torch_geometric_data(twin, p, True, True, 32)
        with torch.no_grad():
            a = ctrl.model(data) * ctrl.actions_std + ctrl.actions_mean
        assert not twin.StepActions(cc.PointVector(a.cpu().numpy()))
        # (the toy network's index_add_ is an atomic float32 sum: two evaluations may differ in the last bit)
        assert np.allclose(np.array(env.GetRobotPositions()), np.array(twin.GetRobotPositions()), rtol=0, atol=1e-5)
        assert obj == pytest.approx(twin.GetObjectiveValue(), rel=1e-7)
        assert converged is False
    with pytest.raises(ImportError):
        ControllerNN(cfg, p, env)
