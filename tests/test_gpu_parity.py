"""GPU parity tests: the CUDA path (through the C ABI / the _core mirror) against the CPU oracle."""
import numpy as np
import pytest

from conftest import KAT_BND, KAT_EXPECTED, KAT_POS, random_world

pytestmark = pytest.mark.gpu


def _cc():
    import coverage_control as cc

    return cc


def _orc():
    from oracle import oracle as orc

    return orc


def _make_idf(cc, params, bnd, polys=()):
    idf = cc.WorldIDF(params)
    for b in bnd:
        idf.AddNormalDistribution(cc.BivariateNormalDistribution(cc.Point2(b[0], b[1]), float(b[2]), float(b[4])))
    for pts, imp in polys:
        idf.AddUniformDistributionPolygon(cc.PolygonFeature(cc.PointVector(pts), imp))
    return idf


# ---------------------------------------------------------------------------------------------
# IDF
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N,F,seed", [(1024, 32, 0), (512, 7, 1), (192, 5, 2)])
def test_idf_mode0_matches_reference_cuda_kernel_bit_for_bit(N, F, seed):
    """Mode 0 vs the reference's own generate_world_map.cu compiled in place (oracle/_ref).  Sizes are
    multiples of 32: the reference kernel leaves the cells beyond floor(N/32)*32 UNINITIALISED and takes
    its normalisation maximum over them too (generate_world_map.cu:238-256), so for ragged sizes its
    output depends on what the allocation held before; ragged sizes are checked against the CPU
    restatement instead (test_idf_polygons_and_ragged_size)."""
    cc, orc = _cc(), _orc()
    if not orc.ref_idf_available():
        pytest.skip("oracle/_ref/libref_idf.so not built")
    p = cc.Parameters()
    p.pWorldMapSize = p.pRobotMapSize = N
    bnd, _ = random_world(seed, N, F, 4)
    idf = _make_idf(cc, p, bnd)
    idf.GenerateMap()
    got = idf.GetWorldMap()
    ref, ref_f = orc.ref_idf_generate(orc.params_from(p), bnd)
    M = (N // 32) * 32  # the reference leaves rows/cols beyond floor(N/32)*32 uninitialised
    assert np.array_equal(got[:M, :M], ref[:M, :M]), f"max abs diff {np.abs(got[:M,:M] - ref[:M,:M]).max()}"
    if M == N:
        assert idf.GetNormalizationFactor() == ref_f


def test_idf_mode0_close_to_cpu_restatement():
    """Same tolerance the reference accepts between its CPU and CUDA paths (test_parity.py:46)."""
    cc, orc = _cc(), _orc()
    p = cc.Parameters()
    bnd, _ = random_world(3, 1024, 32, 4)
    idf = _make_idf(cc, p, bnd)
    idf.GenerateMap()
    got = idf.GetWorldMap()
    for mode in (0, 1):
        ref, f = orc.idf_generate(orc.params_from(p), bnd, mode=mode)
        assert np.allclose(got, ref, atol=1e-2)
        assert abs(idf.GetNormalizationFactor() - f) <= 1e-3 * f


def test_idf_mode1_within_1e5_of_fp64_truth():
    """north_star tolerance: relative 1e-5 on IDF (of the map maximum), fast separable mode."""
    cc, orc = _cc(), _orc()
    p = cc.Parameters()
    bnd, _ = random_world(4, 1024, 32, 4)
    idf = _make_idf(cc, p, bnd)
    idf.GenerateMap(1)
    got = idf.GetWorldMap().astype(np.float64)
    truth, border = orc.idf_fp64(orc.params_from(p), bnd, band=1e-3)
    truth = truth * (p.pNorm / truth.max())
    ok = border == 0  # cells whose truncation test is within 1e-3 of the threshold are reported separately
    assert np.abs(got - truth)[ok].max() <= 1e-5 * truth.max()
    assert (~ok).mean() < 1e-3
    # flipped borderline cells change by at most the truncated tail (~e^-2 of one peak)
    assert np.abs(got - truth).max() < 0.2


def test_idf_polygons_and_ragged_size():
    cc, orc = _cc(), _orc()
    p = cc.Parameters()
    p.pWorldMapSize = p.pRobotMapSize = 250  # not a multiple of 32 (nor of 4: 250 % 4 = 2)
    bnd, _ = random_world(5, 250, 3, 4, sigma=(20, 30))
    tri = [(np.array([[30.2, 40.1], [120.7, 55.3], [60.4, 150.9]]), 0.002)]
    quad = [(np.array([[150.3, 150.2], [220.6, 160.1], [230.2, 230.7], [140.9, 210.3]]), 0.003)]
    # y-monotone pieces the library makes (ear-clipping triangles); same union as the polygons
    pieces = tri + [(quad[0][0][[0, 1, 2]], 0.003), (quad[0][0][[0, 2, 3]], 0.003)]
    # polygons only: no erfc in play, so the CUDA result must equal the restatement bit for bit
    only = _make_idf(cc, p, bnd[:0], tri + quad)
    only.GenerateMap()
    got = only.GetWorldMap()
    ref, ref_f = orc.idf_generate(orc.params_from(p), bnd[:0], polys=pieces, mode=1)
    assert got.shape == (250, 250)
    assert np.array_equal(got, ref)
    assert only.GetNormalizationFactor() == ref_f
    assert np.count_nonzero(got) > 5000
    # Gaussians + polygons: glibc vs CUDA erfcf differ by ULPs that the 4-corner cancellation
    # amplifies, hence the reference's own tolerance class (atol)
    idf = _make_idf(cc, p, bnd, tri + quad)
    idf.GenerateMap()
    got = idf.GetWorldMap()
    ref, _ = orc.idf_generate(orc.params_from(p), bnd, polys=pieces, mode=1)
    assert np.allclose(got, ref, atol=1e-4)
    # empty world: factor = pNorm, map all zero (generate_world_map.cu:252-256)
    empty = cc.WorldIDF(p)
    empty.GenerateMap()
    assert not empty.GetWorldMap().any()
    assert empty.GetNormalizationFactor() == pytest.approx(p.pNorm)


def test_idf_polygons_match_reference_cuda_kernel():
    """Polygons alone and Gaussians + polygons vs the reference's own kernel, bit for bit, on a size the
    reference kernel covers completely (a multiple of 32; see the ragged-size note above)."""
    cc, orc = _cc(), _orc()
    if not orc.ref_idf_available():
        pytest.skip("oracle/_ref/libref_idf.so not built")
    p = cc.Parameters()
    p.pWorldMapSize = p.pRobotMapSize = 256
    bnd, _ = random_world(5, 256, 3, 4, sigma=(20, 30))
    tri = [(np.array([[30.2, 40.1], [120.7, 55.3], [60.4, 150.9]]), 0.002)]
    quad = [(np.array([[150.3, 150.2], [220.6, 160.1], [230.2, 230.7], [140.9, 210.3]]), 0.003)]
    pieces = tri + [(quad[0][0][[0, 1, 2]], 0.003), (quad[0][0][[0, 2, 3]], 0.003)]
    for b in (bnd[:0], bnd):
        idf = _make_idf(cc, p, b, tri + quad)
        idf.GenerateMap()
        rmap, rf = orc.ref_idf_generate(orc.params_from(p), b, polys=pieces)
        assert np.array_equal(idf.GetWorldMap(), rmap)
        assert idf.GetNormalizationFactor() == rf


# ---------------------------------------------------------------------------------------------
# Voronoi labels and reductions
# ---------------------------------------------------------------------------------------------
def _env_from(cc, p, wm, pos):
    idf = cc.WorldIDF(p, wm)
    return cc.CoverageSystem(p, idf, cc.PointVector(pos))


@pytest.mark.parametrize("N,R,seed", [(1024, 32, 0), (1024, 3, 1), (512, 64, 2), (256, 1, 3), (300, 17, 4)])
def test_voronoi_labels_bit_exact_and_moments(N, R, seed):
    cc, orc = _cc(), _orc()
    p = cc.Parameters()
    p.pWorldMapSize = p.pRobotMapSize = N
    p.pNumRobots = R
    bnd, pos = random_world(seed, N, 16, R)
    wm, _ = orc.idf_generate(orc.params_from(p), bnd, mode=1)
    env = _env_from(cc, p, wm, pos)
    labels = env.voronoi_labels(0)[0]
    ref_labels, ref_sums = orc.voronoi_bruteforce(pos, wm)
    assert np.array_equal(labels, ref_labels)  # bit-exact, lowest-index tie rule
    cells = env.GetVoronoiCells()
    mass = np.array([c.mass for c in cells])
    assert np.allclose(mass, ref_sums[:, 0], rtol=1e-9, atol=1e-12)
    obj = env.GetObjectiveValue()
    assert obj == pytest.approx(ref_sums[:, 3].sum(), rel=1e-9)
    # reference-faithful scan-line oracle (voronoi.cpp:55-150): centroid incl. vertex-mean fallback
    ocells = orc.voronoi(pos, wm)
    for c, o in zip(cells, ocells):
        assert c.mass == pytest.approx(o.mass, rel=1e-4, abs=1e-9)
        tol = 1e-4 * N if o.mass >= 1e-10 else 1e-9 * N
        assert np.allclose(c.centroid, list(o.centroid), rtol=1e-4 if o.mass >= 1e-10 else 0, atol=tol if o.mass < 1e-10 else 1e-6)
        # same polygon vertex set
        a = np.array(sorted(map(tuple, np.round(np.array(c.cell), 6))))
        b = np.array(sorted(map(tuple, np.round(o.polygon(), 6))))
        assert a.shape == b.shape and np.allclose(a, b, atol=1e-5)


def test_voronoi_ties_go_to_lowest_index():
    """Sites placed symmetrically about cell centres: exact bisector hits, coincident robots."""
    cc, orc = _cc(), _orc()
    N = 256
    p = cc.Parameters()
    p.pWorldMapSize = p.pRobotMapSize = N
    pos = np.array([[100.5, 64.5], [100.5, 192.5], [28.5, 128.5], [172.5, 128.5], [100.5, 64.5]], np.float64)
    p.pNumRobots = len(pos)
    wm = np.random.default_rng(0).random((N, N)).astype(np.float32)
    env = _env_from(cc, p, wm, pos)
    labels = env.voronoi_labels(0)[0]
    ref_labels, ref_sums = orc.voronoi_bruteforce(pos, wm)
    assert np.array_equal(labels, ref_labels)
    assert not (labels == 4).any()  # the duplicate of site 0 never wins
    mass = np.array([c.mass for c in env.GetVoronoiCells()])
    assert np.allclose(mass, ref_sums[:, 0], rtol=1e-9)


def test_voronoi_many_sites_slow_paths():
    """More sites than lanes (CTA-level cull) and dense clusters (> 4 candidates per strip)."""
    cc, orc = _cc(), _orc()
    N, R = 256, 100
    p = cc.Parameters()
    p.pWorldMapSize = p.pRobotMapSize = N
    p.pNumRobots = R
    rng = np.random.default_rng(7)
    pos = np.concatenate([rng.uniform(1, N - 1, (60, 2)), rng.normal(128, 6, (40, 2)).clip(1, N - 1)])
    wm = rng.random((N, N)).astype(np.float32)
    env = _env_from(cc, p, wm, pos)
    labels = env.voronoi_labels(0)[0]
    ref_labels, ref_sums = orc.voronoi_bruteforce(pos, wm)
    assert np.array_equal(labels, ref_labels)
    cells = env.GetVoronoiCells()
    assert np.allclose([c.mass for c in cells], ref_sums[:, 0], rtol=1e-9)
    assert env.GetObjectiveValue() == pytest.approx(ref_sums[:, 3].sum(), rel=1e-9)


# ---------------------------------------------------------------------------------------------
# the reference's own known answers through the drop-in API (python/tests/test_coverage.py)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["clairvoyant", "centralized", "decentralized"])
def test_reference_known_answers(name):
    cc = _cc()
    from coverage_control.algorithms import CentralizedCVT, ClairvoyantCVT, DecentralizedCVT

    params = cc.Parameters()
    params.pNumGaussianFeatures = 3
    params.pNumRobots = 3
    bnd_list = cc.BNDVector()
    for b in KAT_BND:
        bnd_list.append(cc.BivariateNormalDistribution(cc.Point2(b[0], b[1]), b[2], b[4]))
    robot_pos_list = cc.PointVector()
    for q in KAT_POS:
        robot_pos_list.append(cc.Point2(q[0], q[1]))
    env = cc.CoverageSystem(params, bnd_list, robot_pos_list)
    init_cost = env.GetObjectiveValue()
    controller = {"clairvoyant": ClairvoyantCVT, "centralized": CentralizedCVT,
                  "decentralized": DecentralizedCVT}[name](params, env)
    step_num = None
    for i in range(0, params.pEpisodeSteps):
        controller.ComputeActions()
        actions = controller.GetActions()
        if env.StepActions(actions):
            break
        if controller.IsConverged():
            break
        step_num = i
    current_cost = env.GetObjectiveValue()
    improvement = 100 * (init_cost - current_cost) / init_cost
    assert round(improvement, 2) == KAT_EXPECTED[name][0]
    assert step_num == KAT_EXPECTED[name][1]


# ---------------------------------------------------------------------------------------------
# trajectories and maps against the structure-faithful CPU environment
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind,name", [(0, "ClairvoyantCVT"), (1, "CentralizedCVT"), (2, "DecentralizedCVT")])
def test_rollout_matches_oracle(kind, name):
    cc, orc = _cc(), _orc()
    import coverage_control.algorithms as algs

    N, R, steps = 512, 12, 25
    p = cc.Parameters()
    p.pWorldMapSize = p.pRobotMapSize = N
    p.pNumRobots = R
    p.pLocalMapSize = 128
    p.pSensorSize = 32
    p.pCommunicationRange = 96
    p.pTimeStep = 0.2
    bnd, pos = random_world(11 + kind, N, 12, R)
    pos[0] = [3.0, 5.0]          # windows clipped at the world edge
    pos[1] = [N - 2.0, N - 7.0]
    op = orc.params_from(p)
    wm, _ = orc.idf_generate(op, bnd, mode=1)
    oenv = orc.Env(op, wm, pos)
    env = _env_from(cc, p, wm, pos)
    ctrl = getattr(algs, name)(p, env)
    for s in range(steps):
        conv, oa, og = oenv.compute_actions(kind)
        ctrl.ComputeActions()
        a = np.array(ctrl.GetActions())
        g = np.array(ctrl.GetGoals())
        assert np.allclose(g, og, rtol=1e-4, atol=1e-6), f"goals differ at step {s}"
        assert np.allclose(a, oa, rtol=1e-4, atol=1e-6), f"actions differ at step {s}"
        assert ctrl.IsConverged() == conv
        assert not env.StepActions(ctrl.GetActions())
        assert oenv.step_actions(oa) == 0
        assert np.allclose(np.array(env.GetRobotPositions()), oenv.positions(), rtol=0, atol=1e-7)
        assert env.GetObjectiveValue() == pytest.approx(oenv.objective(), rel=1e-4)
    # every derived map is bit-identical to the dense maps the reference keeps
    assert np.array_equal(env.GetWorldMap(), oenv.world_map())
    assert np.array_equal(env.GetSystemExploredIDFMap(), oenv.explored_map())
    assert np.array_equal(env.GetSystemExplorationMap(), oenv.exploration_map())
    assert np.array_equal(env.GetSystemMap(), oenv.system_map())
    for r in (0, 1, R - 1):
        assert np.array_equal(env.GetRobotMap(r), oenv.robot_map(r))
        assert np.array_equal(env.GetRobotSensorView(r), oenv.sensor_view(r))
        assert np.array_equal(env.GetRobotLocalMap(r), oenv.local_map(r))
        assert np.array_equal(env.GetRobotObstacleMap(r), oenv.obstacle_map(r))
    feats = np.array([list(f) for f in env.GetRobotVoronoiFeatures()])
    ofeats = np.array([oenv.voronoi_features(r) for r in range(R)])
    assert np.allclose(feats, ofeats, rtol=1e-6, atol=1e-6)
    for r in range(R):
        assert np.allclose(np.array(env.GetRelativePositonsNeighbors(r)).reshape(-1, 2), oenv.neighbors(r))


def test_batched_rollout_equals_single_environments():
    """E environments stepped on the device == each one stepped through the single-env API."""
    cc, orc = _cc(), _orc()
    from coveragecontrol_b200 import BatchedCoverageSystem, BatchedCVT

    N, R, E, steps = 256, 6, 5, 12
    p = cc.Parameters()
    p.pWorldMapSize = p.pRobotMapSize = N
    p.pNumRobots = R
    p.pLocalMapSize = 64
    p.pSensorSize = 16
    p.pCommunicationRange = 48
    op = orc.params_from(p)
    worlds, poss = [], []
    for e in range(E):
        bnd, pos = random_world(100 + e, N, 6, R)
        wm, _ = orc.idf_generate(op, bnd, mode=1)
        worlds.append(wm)
        poss.append(pos)
    for kind in (0, 1, 2):
        batch = BatchedCoverageSystem(p, E, R)
        for e in range(E):
            batch.set_world(e, worlds[e])
        batch.reset(np.stack(poss))
        ctrl = BatchedCVT(batch, kind)
        ctrl.rollout(steps)
        got = batch.positions()
        obj = batch.objective()
        for e in range(E):
            oenv = orc.Env(op, worlds[e], poss[e])
            for _ in range(steps):
                _, a, _ = oenv.compute_actions(kind)
                oenv.step_actions(a)
            assert np.allclose(got[e], oenv.positions(), atol=1e-6), f"kind {kind} env {e}"
            assert obj[e] == pytest.approx(oenv.objective(), rel=1e-4)


def test_fused_rollout_with_observation_equals_separate_calls():
    """ccb_ctrl_rollout_obs (objective chain and observation kernels on two streams) == rollout + get_maps."""
    import torch

    cc, orc = _cc(), _orc()
    from coveragecontrol_b200 import BatchedCoverageSystem, BatchedCVT, _lib

    N, R, E, steps, C = 512, 10, 4, 7, 32
    p = cc.Parameters()
    p.pWorldMapSize = p.pRobotMapSize = N
    p.pNumRobots = R
    rng = np.random.default_rng(5)
    worlds = [rng.random((N, N)).astype(np.float32) for _ in range(E)]
    pos = rng.uniform(1, N - 1, (E, R, 2))
    out = []
    for fused in (False, True):
        env = BatchedCoverageSystem(p, E, R)
        for e in range(E):
            env.set_world(e, worlds[e])
        env.reset(pos)
        ctrl = BatchedCVT(env, 2)
        obj = torch.zeros((steps, E), dtype=torch.float64, device="cuda")
        obs = torch.zeros((steps, E, R, 4, C, C), dtype=torch.float32, device="cuda")
        if fused:
            ctrl.rollout_obs(steps, True, obj.data_ptr(), C, True, obs.data_ptr(), obs[0].numel())
        else:
            for s in range(steps):
                ctrl.rollout(1, True, obj[s].data_ptr())
                _lib.check(_lib.load().ccb_env_cnn_input_dev(env._h, C, 1, obs[s].data_ptr()))
        torch.cuda.synchronize()
        out.append((env.positions(), obj.cpu().numpy(), obs.cpu().numpy(), env.objective()))
    assert np.allclose(out[0][0], out[1][0], rtol=0, atol=1e-9)
    assert np.allclose(out[0][1], out[1][1], rtol=1e-12)
    assert np.allclose(out[0][2], out[1][2], rtol=1e-6, atol=1e-9)
    assert np.allclose(out[0][3], out[1][3], rtol=1e-12)
    assert np.allclose(out[1][1][-1], out[1][3], rtol=1e-12)


@pytest.mark.parametrize("kind", [0, 1, 2])
def test_rollout_graph_replay_equals_eager(monkeypatch, kind):
    """Identical roll-out calls are captured into a CUDA graph on the second call and replayed afterwards
    (csrc/ccb200.cu ccb_ctrl_rollout_obs); the trajectory, objective and observation must be those of the
    eager launches, and the host-side bookkeeping must stay consistent with calls made in between."""
    import torch

    cc, orc = _cc(), _orc()
    from coveragecontrol_b200 import BatchedCoverageSystem, BatchedCVT, _lib

    N, R, E, C = 512, 9, 3, 32
    p = cc.Parameters()
    p.pWorldMapSize = p.pRobotMapSize = N
    p.pNumRobots = R
    rng = np.random.default_rng(17)
    worlds = [rng.random((N, N)).astype(np.float32) for _ in range(E)]
    pos = rng.uniform(1, N - 1, (E, R, 2))
    res = []
    for nograph in ("1", "0"):
        monkeypatch.setenv("CCB_NO_GRAPH", nograph)
        env = BatchedCoverageSystem(p, E, R)
        for e in range(E):
            env.set_world(e, worlds[e])
        env.reset(pos)
        ctrl = BatchedCVT(env, kind)
        obj = torch.zeros((2, E), dtype=torch.float64, device="cuda")
        obs = torch.zeros((E, R, 4, C, C), dtype=torch.float32, device="cuda")
        launches0 = _lib.load().ccb_kernel_launch_count()
        trace = []
        for it in range(6):
            ctrl.rollout_obs(2, True, obj.data_ptr(), C, True, obs.data_ptr(), 0)
            torch.cuda.synchronize()
            trace.append((env.positions(), obj.cpu().numpy().copy(), obs.cpu().numpy().copy()))
            if it == 3:  # calls in between: host-returning getters and a plain step through the host API
                trace.append((env.objective(), np.array(env.cells(0)[0])))
                ctrl.compute_actions()
                env.step_actions(ctrl.actions())
        trace.append(env.objective())
        res.append((trace, _lib.load().ccb_kernel_launch_count() - launches0))
    (eager, n_eager), (graph, n_graph) = res
    assert n_graph == n_eager > 0  # replays are counted like the launches they stand for
    for a, b in zip(eager, graph):
        if isinstance(a, tuple):
            for x, y in zip(a, b):
                assert np.allclose(x, y, rtol=1e-9, atol=1e-9)
        else:
            assert np.allclose(a, b, rtol=1e-9)


def test_bench_configuration_against_oracle():
    """The configuration bench.py measures (C2: 1024^2 world, 32 robots, TOML parameters, DecentralizedCVT,
    objective + LPAC observation), batched wide enough that every default fast path is taken -- site-ordered
    whole-map Voronoi, observation kernel fused with the window Voronoi, roll-outs replayed from a CUDA
    graph -- against the CPU oracle stepping single environments."""
    import torch
    import torchvision

    import bench
    cc, orc = _cc(), _orc()
    from coveragecontrol_b200 import BatchedCoverageSystem, BatchedCVT

    E, C, calls, per_call = 36, 32, 4, 2
    p = bench.toml_params(cc)
    op = orc.params_from(p)
    env = BatchedCoverageSystem(p, E, bench.N_ROBOTS)
    pos = np.zeros((E, bench.N_ROBOTS, 2))
    worlds = {}
    for e in range(E):
        bnd, pos[e] = bench.synth_inputs(1000 + e)
        idf = cc.WorldIDF(p)
        for b in bnd:
            idf.AddNormalDistribution(cc.BivariateNormalDistribution(cc.Point2(b[0], b[1]), float(b[2]), float(b[4])))
        idf.GenerateMap()
        env.set_world(e, idf)
        if e in (0, 17, 35):
            worlds[e] = np.array(idf.GetWorldMap())
    env.reset(pos)
    ctrl = BatchedCVT(env, 2)
    obj = torch.zeros((per_call, E), dtype=torch.float64, device="cuda")
    obs = torch.zeros((E, bench.N_ROBOTS, 4, C, C), dtype=torch.float32, device="cuda")
    for _ in range(calls):
        ctrl.rollout_obs(per_call, True, obj.data_ptr(), C, True, obs.data_ptr(), 0)
    torch.cuda.synchronize()
    got_pos, got_obj, got_obs = env.positions(), obj.cpu().numpy()[-1], obs.cpu()
    rs = lambda t: torchvision.transforms.functional.resize(  # noqa: E731
        t, (C, C), interpolation=torchvision.transforms.InterpolationMode.BILINEAR, antialias=True)
    for e, wm in worlds.items():
        oenv = orc.Env(op, wm, pos[e])
        for _ in range(calls * per_call):
            _, a, _ = oenv.compute_actions(2)
            oenv.step_actions(a)
        assert np.allclose(got_pos[e], oenv.positions(), rtol=0, atol=1e-6), f"env {e}"
        assert got_obj[e] == pytest.approx(oenv.objective(), rel=1e-4)
        local = torch.from_numpy(np.stack([oenv.local_map(r) for r in range(bench.N_ROBOTS)]))
        obst = torch.from_numpy(np.stack([oenv.obstacle_map(r) for r in range(bench.N_ROBOTS)]))
        assert torch.allclose(got_obs[e, :, 0], rs(local), rtol=1e-5, atol=1e-8)
        assert torch.allclose(got_obs[e, :, 3], rs(obst), rtol=1e-5, atol=1e-8)
        comm = np.stack([np.stack(oenv.comm_maps(r, C, "py")) for r in range(bench.N_ROBOTS)])
        assert np.array_equal(got_obs[e, :, 1:3].numpy(), comm)
