"""CoverageSystem.GetRobotExplorationFeatures (reference coverage_system.h:603-610 ->
robot_model.h:347-403) through the drop-in package against the oracle's restatement (whose frontier queue
is the real std::priority_queue): bit-exact, including which of many equally valued candidates survive."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(N, R, S, steps, seed, *, expl=True, unknown=False, K=10, speed=5.0, corner=False):
    import coverage_control as cc
    from coverage_control.algorithms import ClairvoyantCVT
    from oracle import oracle as orc

    rng = np.random.default_rng(seed)
    p = cc.Parameters()
    p.pWorldMapSize = p.pRobotMapSize = N
    p.pNumRobots = R
    p.pLocalMapSize = 2 * S
    p.pSensorSize = S
    p.pCommunicationRange = 2.0 * S
    p.pUpdateExplorationMap = expl
    p.pRobotMapUseUnknownImportance = unknown
    p.pNumFrontiers = K
    p.pMaxRobotSpeed = speed
    wm = rng.random((N, N)).astype(np.float32)
    wm[rng.random((N, N)) < 0.3] = 0.0  # zero cells: block().count() differs from the window area
    pos = rng.uniform(1.0, N - 1.0, (R, 2))
    if corner:  # windows and candidate lattices clipped at every border, robots on the limits
        pos[0], pos[1], pos[2] = (1e-4, 1e-4), (N - 1e-4, N - 1e-4), (0.4, N - 0.6)
    env = cc.CoverageSystem(p, cc.WorldIDF(p, wm), cc.PointVector(pos))
    oenv = orc.Env(orc.params_from(p), wm, pos)
    ctrl = ClairvoyantCVT(p, env)
    for step in range(steps + 1):
        got = np.array([list(f) for f in env.GetRobotExplorationFeatures()])
        ref = np.array([oenv.exploration_features(r, K) for r in range(R)])
        assert got.shape == (R, 2 * K)
        assert np.array_equal(got, ref), f"step {step}: robots {np.argwhere((got != ref).any(axis=1)).ravel().tolist()} differ"
        if step < steps:
            ctrl.ComputeActions()
            a = np.array(ctrl.GetActions())
            env.StepActions(ctrl.GetActions())
            oenv.step_actions(a)
    return got


@pytest.mark.parametrize("N,R,S,steps,seed", [(256, 6, 32, 6, 0), (128, 4, 16, 4, 1), (512, 8, 64, 3, 2)])
def test_exploration_features_match_oracle(N, R, S, steps, seed):
    got = _run(N, R, S, steps, seed)
    assert np.any(got != 0.0)


def test_exploration_features_at_the_borders():
    _run(128, 5, 32, 3, 3, corner=True)


def test_exploration_features_with_unknown_importance_fill():
    _run(128, 4, 16, 3, 4, unknown=True)


def test_exploration_features_without_exploration_map():
    """pUpdateExplorationMap = false: the exploration map is all zeros, every candidate scores 0 and the
    queue's tie behaviour alone picks the points."""
    _run(128, 3, 16, 2, 5, expl=False)


def test_exploration_features_queue_sizes_and_speed():
    _run(128, 3, 16, 1, 6, K=1)
    _run(128, 3, 16, 1, 7, K=37, speed=2.5)


def test_exploration_features_batched_matches_single():
    import coverage_control as cc
    from coveragecontrol_b200 import BatchedCoverageSystem

    N, R, E = 128, 4, 3
    rng = np.random.default_rng(8)
    p = cc.Parameters()
    p.pWorldMapSize = p.pRobotMapSize = N
    p.pNumRobots = R
    p.pLocalMapSize, p.pSensorSize, p.pCommunicationRange = 32, 16, 32.0
    worlds = [rng.random((N, N)).astype(np.float32) for _ in range(E)]
    pos = rng.uniform(1.0, N - 1.0, (E, R, 2))
    batch = BatchedCoverageSystem(p, E, R)
    for e in range(E):
        batch.set_world(e, worlds[e])
    batch.reset(pos)
    got = batch.exploration_features()
    for e in range(E):
        env = cc.CoverageSystem(p, cc.WorldIDF(p, worlds[e]), cc.PointVector(pos[e]))
        assert np.array_equal(got[e], np.array([list(f) for f in env.GetRobotExplorationFeatures()]))
