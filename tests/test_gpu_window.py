"""DecentralizedCVT's per-robot window Voronoi (csrc/vwindow.cuh): the row-interval kernel against the
generic region kernel and against the oracle, on geometry chosen to hit the exact-rule branches
(bisectors through cell centres, along rows, along columns), clipped windows and the unknown-fill mode."""
import os

import numpy as np
import pytest

from conftest import random_world

pytestmark = pytest.mark.gpu


def _goals(p, wm, pos, steps, generic):
    from coveragecontrol_b200 import BatchedCoverageSystem, BatchedCVT

    old = os.environ.get("CCB_VW_GENERIC")
    os.environ["CCB_VW_GENERIC"] = "1" if generic else "0"
    try:
        env = BatchedCoverageSystem(p, 1, len(pos))
    finally:
        if old is None:
            os.environ.pop("CCB_VW_GENERIC", None)
        else:
            os.environ["CCB_VW_GENERIC"] = old
    env.set_world(0, wm)
    env.reset(pos[None])
    ctrl = BatchedCVT(env, 2)
    out = []
    for _ in range(steps):
        ctrl.compute_actions()
        out.append((ctrl.goals()[0].copy(), ctrl.actions()[0].copy()))
        env.step_actions(ctrl.actions())
    return out


@pytest.mark.parametrize("N,L,S,fill,ties", [(512, 256, 64, False, True), (512, 128, 32, True, True),
                                              (512, 128, 32, True, False), (256, 64, 16, False, False)])
def test_window_kernel_matches_generic_kernel_and_oracle(N, L, S, fill, ties):
    """ties=True puts bisectors exactly through cell centres: there the build's stated rule (ties to the
    lowest site index = the robot itself in its own list) and the reference's scan-line rule
    (voronoi.cpp:137-141, ties to the geometrically lower cell) differ, so only the two kernels are
    compared; ties=False shifts the sites off the lattice and also compares with the oracle."""
    import coverage_control as cc
    from oracle import oracle as orc

    R = 14
    p = cc.Parameters()
    p.pWorldMapSize = p.pRobotMapSize = N
    p.pNumRobots = R
    p.pLocalMapSize = L
    p.pSensorSize = S
    p.pCommunicationRange = L / 2.0
    p.pRobotMapUseUnknownImportance = fill
    p.pTimeStep = 0.2
    bnd, pos = random_world(77, N, 14, R, sigma=(60.0, 90.0))
    c = N / 2.0
    q = 0.0 if ties else 0.3
    pos[0] = [c, c]                # bisector with robot 1 along a row, through cell centres (x = c + 5.5)
    pos[1] = [c + 11.0 + q, c]
    pos[2] = [c, c + 11.0 + q]     # bisector with robot 0 along a column, through cell centres
    pos[3] = [c + 7.0 + q, c + 7.0]  # diagonal bisector through cell centres
    pos[4] = [c + 1e-9, c - 13.0 - q]  # nearly flat bisector slope
    pos[5] = [2.5, 3.5]        # clipped windows
    pos[6] = [N - 1.5, N - 9.0]
    pos[7] = [N - 30.0, 12.0]
    op = orc.params_from(p)
    wm, _ = orc.idf_generate(op, bnd, mode=1)
    steps = 8
    new = _goals(p, wm, pos, steps, generic=False)
    gen = _goals(p, wm, pos, steps, generic=True)
    oenv = orc.Env(op, wm, pos)
    for s in range(steps):
        _, oa, og = oenv.compute_actions(2)
        assert np.allclose(new[s][0], gen[s][0], rtol=1e-11, atol=1e-9), f"goals: window kernel vs generic, step {s}"
        if not ties:
            assert np.allclose(new[s][0], og, rtol=1e-4, atol=1e-6), f"goals vs oracle, step {s}"
            assert np.allclose(new[s][1], oa, rtol=1e-4, atol=1e-6), f"actions vs oracle, step {s}"
            oenv.step_actions(oa)


@pytest.mark.parametrize("fill", [False, True])
def test_observation_kernel_supplies_the_next_window_moments(monkeypatch, fill):
    """With a DecentralizedCVT attached, the CNN-input kernel also reduces the robots' window Voronoi
    (obs.cuh, VOR variant) and the next ComputeActions uses those moments: same goals / trajectory as with
    the two kernels kept apart (CCB_NO_WFUSE), observation bit-identical."""
    import torch

    import coverage_control as cc
    from coveragecontrol_b200 import BatchedCoverageSystem, BatchedCVT, _lib

    N, R, E, C = 512, 14, 3, 32
    p = cc.Parameters()
    p.pWorldMapSize = p.pRobotMapSize = N
    p.pNumRobots = R
    p.pRobotMapUseUnknownImportance = fill
    p.pTimeStep = 0.2
    rng = np.random.default_rng(23)
    worlds = [rng.random((N, N)).astype(np.float32) for _ in range(E)]
    pos = rng.uniform(1, N - 1, (E, R, 2))
    c = N / 2.0
    pos[0, 0], pos[0, 1], pos[0, 2] = [c, c], [c + 11.0, c], [c, c + 11.0]   # exact ties (rows / columns)
    pos[1, 0], pos[1, 1] = [2.5, 3.5], [N - 1.5, N - 9.0]                     # clipped windows
    runs = []
    for nofuse in ("1", "0"):
        monkeypatch.setenv("CCB_NO_WFUSE", nofuse)
        monkeypatch.setenv("CCB_NO_GRAPH", "1")
        env = BatchedCoverageSystem(p, E, R)
        for e in range(E):
            env.set_world(e, worlds[e])
        env.reset(pos)
        ctrl = BatchedCVT(env, 2)
        obs = torch.zeros((E, R, 4, C, C), dtype=torch.float32, device="cuda")
        trace = []
        for s in range(6):
            ctrl.compute_actions()
            trace.append(ctrl.goals().copy())
            env.step_actions(ctrl.actions())
            _lib.check(_lib.load().ccb_env_cnn_input_dev(env._h, C, 1, obs.data_ptr()))
            trace.append(obs.cpu().numpy().copy())
        runs.append(trace)
    for s, (a, b) in enumerate(zip(*runs)):
        if s % 2 == 0:
            assert np.allclose(a, b, rtol=1e-11, atol=1e-9), f"goals differ at step {s // 2}"
        else:
            assert np.array_equal(a, b), f"observation differs at step {s // 2}"


def test_local_voronoi_features_match_oracle():
    """CoverageSystem::GetLocalVoronoiFeatures (coverage_system.cpp:648-676): the robot's own cell among itself and
    its neighbours over the clipped local window of its own map -> [cx, cy, mass, sum d^2 v, 0, sum d v, 0]."""
    import coverage_control as cc
    from coverage_control.algorithms import DecentralizedCVT
    from oracle import oracle as orc

    N, R, L = 512, 10, 128
    p = cc.Parameters()
    p.pWorldMapSize = p.pRobotMapSize = N
    p.pNumRobots = R
    p.pLocalMapSize = L
    p.pSensorSize = 48
    p.pCommunicationRange = 100
    bnd, pos = random_world(91, N, 14, R, sigma=(60.0, 90.0))
    pos[0], pos[1] = [3.2, 5.7], [N - 2.3, N - 7.9]  # clipped windows
    pos[2] = pos[3] + [20.3, -11.2]                   # close neighbours
    op = orc.params_from(p)
    wm, _ = orc.idf_generate(op, bnd, mode=1)
    env = cc.CoverageSystem(p, cc.WorldIDF(p, wm), cc.PointVector(pos))
    oenv = orc.Env(op, wm, pos)
    ctrl = DecentralizedCVT(p, env)
    for _ in range(5):
        ctrl.ComputeActions()
        a = ctrl.GetActions()
        env.StepActions(a)
        oenv.step_actions(np.array(a))
    feats = np.array([list(f) for f in env.GetLocalVoronoiFeatures()])
    P = oenv.positions()
    for r in range(R):
        ix, of = orc.compute_offsets(p.pResolution, P[r, 0], P[r, 1], L, N)
        x0, y0 = ix.left + of.left, ix.bottom + of.bottom
        trimmed = np.ascontiguousarray(oenv.robot_map(r)[x0:x0 + of.width, y0:y0 + of.height])
        shift = np.array([x0, y0]) * p.pResolution
        nb = oenv.neighbors(r).reshape(-1, 2) + P[r]          # neighbours' global positions
        sites = np.concatenate([[P[r] - shift], nb - shift]) if len(nb) else np.array([P[r] - shift])
        cell = orc.voronoi(sites, trimmed, p.pResolution, single=True, robot_id=0, map_size=(of.width, of.height))[0]
        want = [cell.centroid[0], cell.centroid[1], cell.mass, cell.sum_idf_site_dist_sqr, 0.0, cell.sum_idf_site_dist, 0.0]
        assert np.allclose(feats[r], want, rtol=1e-6, atol=1e-6), (r, feats[r], want)
        assert np.allclose(list(env.GetLocalVoronoiFeatures(r)), feats[r])




def _clear_setup(R, seed):
    import coverage_control as cc
    from oracle import oracle as orc

    N = 512
    p = cc.Parameters()
    p.pWorldMapSize = p.pRobotMapSize = N
    p.pNumRobots = R
    p.pLocalMapSize = 256
    p.pSensorSize = 32
    p.pCommunicationRange = 128.0
    bnd, pos = random_world(seed, N, 9, R, sigma=(50.0, 80.0))
    op = orc.params_from(p)
    wm, _ = orc.idf_generate(op, bnd, mode=1)
    return cc, orc, p, op, wm, pos


def test_clear_robot_maps_invalidates_the_fused_window_moments():
    """the observation kernel's VOR variant caches DecentralizedCVT's window moments for the
    current positions; ClearRobotMaps / ClearExploredIDF (coverage_system.h:399-408, called mid-run by the
    reference's evaluators) change the robot maps without moving a robot, so the cache must not survive them.
    Step by step against the dense oracle environment, clearing after get_maps and before ComputeActions."""
    from coverage_control.algorithms import DecentralizedCVT

    R = 10
    cc, orc, p, op, wm, pos = _clear_setup(R, 21)
    env = cc.CoverageSystem(p, cc.WorldIDF(p, wm), cc.PointVector(pos))
    oenv = orc.Env(op, wm, pos)
    ctrl = DecentralizedCVT(p, env)
    for s in range(14):
        cc.CoverageEnvUtils.get_maps(env, p, 32, True)  # leaves the window moments of these positions behind
        if s in (6, 9):
            env.ClearRobotMaps(), oenv.clear_robot_maps()
            if s == 9:
                env.ClearExploredIDF(), oenv.clear_explored_idf()
            for r in (0, R - 1):
                assert np.array_equal(env.GetRobotMap(r), oenv.robot_map(r))
        ctrl.ComputeActions()
        _, oa, og = oenv.compute_actions(2)
        assert np.allclose(np.array(ctrl.GetGoals()), og, rtol=1e-9, atol=1e-7), f"goals, step {s}"
        env.StepActions(ctrl.GetActions())
        oenv.step_actions(oa)
    assert np.array_equal(env.GetSystemMap(), oenv.system_map())
    assert np.array_equal(env.GetRobotMap(3), oenv.robot_map(3))


def test_clear_robot_maps_between_graph_replays():
    """The same through the device roll-out, whose second identical call is captured into a CUDA graph and later
    ones replay it: a clear between two replays must reach the replayed kernels (the graph is keyed on which cached
    results are valid on entry)."""
    import torch

    from coveragecontrol_b200 import BatchedCoverageSystem, BatchedCVT, _lib

    R, E = 8, 3
    cc, orc, p, op, wm, _ = _clear_setup(R, 22)
    pos = np.random.default_rng(5).uniform(1, 511, (E, R, 2))
    env = BatchedCoverageSystem(p, E, R)
    for e in range(E):
        env.set_world(e, wm)
    env.reset(pos)
    ctrl = BatchedCVT(env, 2)
    oenvs = [orc.Env(op, wm, pos[e]) for e in range(E)]
    obs = torch.zeros((E, R, 4, 32, 32), device="cuda")
    obj = torch.zeros((1, E), dtype=torch.float64, device="cuda")
    lib = _lib.load()
    for s in range(12):
        ctrl.rollout_obs(1, True, obj.data_ptr(), 32, True, obs.data_ptr(), 0)
        for oe in oenvs:
            _, oa, _ = oe.compute_actions(2)
            oe.step_actions(oa)
        if s in (5, 8):  # after the graph exists (calls 0 and 1 are eager / capture)
            _lib.check(lib.ccb_env_clear_robot_maps(env._h, 1))
            oenvs[1].clear_robot_maps()
        got = env.positions()
        for e in range(E):
            assert np.allclose(got[e], oenvs[e].positions(), atol=1e-7), f"positions, step {s}, env {e}"
    torch.cuda.synchronize()
    for e in range(E):
        assert float(obj[0, e]) == pytest.approx(oenvs[e].objective(), rel=1e-9)


@pytest.mark.parametrize("fill", [False, True])
def test_item_ordered_observation_kernel_equals_the_row_ordered_one(monkeypatch, fill):
    """k_obs_fused (items = non-zero width-pass outputs dealt to the lanes; the product path) against round 1's
    k_cnn_fused8 (whole window rows per warp; CCB_OBS_V1=1): the same 16-tap order, so the local channel is bit-identical;
    the obstacle channel is closed-form in the new kernel (different summation order, 1e-6); window moments -> goals
    equal to 1e-11; 40 steps so that seen regions grow past the 128-column mark, clipped windows, exact ties."""
    import torch

    import coverage_control as cc
    from coveragecontrol_b200 import BatchedCoverageSystem, BatchedCVT, _lib

    N, R, E, C = 512, 12, 4, 32
    p = cc.Parameters()
    p.pWorldMapSize = p.pRobotMapSize = N
    p.pNumRobots = R
    p.pRobotMapUseUnknownImportance = fill
    rng = np.random.default_rng(31)
    worlds = [rng.random((N, N)).astype(np.float32) for _ in range(E)]
    pos = rng.uniform(1, N - 1, (E, R, 2))
    c = N / 2.0
    pos[0, 0], pos[0, 1], pos[0, 2] = [c, c], [c + 11.0, c], [c, c + 11.0]
    pos[1, 0], pos[1, 1], pos[1, 2] = [2.5, 3.5], [N - 1.5, N - 9.0], [N - 0.7, 0.4]
    runs = []
    for v1 in ("1", "0"):
        monkeypatch.setenv("CCB_OBS_V1", v1)
        monkeypatch.setenv("CCB_NO_GRAPH", "1")
        env = BatchedCoverageSystem(p, E, R)
        for e in range(E):
            env.set_world(e, worlds[e])
        env.reset(pos)
        ctrl = BatchedCVT(env, 2)
        obs = torch.zeros((E, R, 4, C, C), dtype=torch.float32, device="cuda")
        trace = []
        for s in range(40):
            ctrl.compute_actions()
            trace.append(ctrl.goals().copy())
            env.step_actions(ctrl.actions())
            _lib.check(_lib.load().ccb_env_cnn_input_dev(env._h, C, 1, obs.data_ptr()))
            if s % 4 == 3:
                trace.append(obs.cpu().numpy().copy())
        runs.append(trace)
    for a, b in zip(*runs):
        if a.ndim == 3:
            assert np.allclose(a, b, rtol=1e-11, atol=1e-9), "goals differ"
        else:
            assert np.array_equal(a[:, :, 0:3], b[:, :, 0:3]), "local / communication channels differ"
            assert np.allclose(a[:, :, 3], b[:, :, 3], rtol=1e-5, atol=1e-7), "obstacle channel differs"
