"""Batched dataset generation (coveragecontrol_b200/data_generation.py, BASELINE config 3) against the
reference's per-environment pipeline (python/data_generation/data_generation.py:217-300) replayed here with
single CoverageSystem objects and CoverageEnvUtils: same schedule, same tensors."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_batched_generator_equals_per_environment_pipeline(tmp_path):
    import torch

    import coverage_control as cc
    from coverage_control.algorithms import ClairvoyantCVT
    from coveragecontrol_b200.data_generation import BatchedDataGenerator, save_dataset

    N, R, E, C, every, n = 512, 6, 3, 32, 2, 16
    p = cc.Parameters()
    p.pWorldMapSize = p.pRobotMapSize = N
    p.pNumRobots = R
    p.pEpisodeSteps = 5
    p.pTimeStep = 0.2
    rng = np.random.default_rng(42)
    worlds = [(rng.random((N, N)).astype(np.float32) * (rng.random((N, N)) < 0.4), rng.uniform(1, N - 1, (R, 2))) for _ in range(12)]
    gen = BatchedDataGenerator(p, E, cnn_map_size=C, every_num_steps=every, converged_data_ratio=0.25, save_objective=True,
                               worlds=iter(worlds))
    data = gen.generate(n)

    # ---- the reference's loop, one CoverageSystem per slot, same schedule --------------------------
    U = cc.CoverageEnvUtils
    src = iter(worlds)

    def fresh():
        wm, pos = next(src)
        env = cc.CoverageSystem(p, cc.WorldIDF(p, wm), cc.PointVector(pos))
        return env, ClairvoyantCVT(p, env)

    slots = [fresh() for _ in range(E)]
    steps = [0] * E
    num_conv = math.ceil(0.25 * n)
    num_nonconv = n - num_conv
    ref = {k: [] for k in data}
    count = nonconv = conv = 0
    while count < n:
        acts, ended = [], []
        for env, alg in slots:
            alg.ComputeActions()
            acts.append(alg.GetActions())
        ended = [slots[e][1].IsConverged() or steps[e] >= p.pEpisodeSteps for e in range(E)]
        sel_nc = [e for e in range(E) if not ended[e] and steps[e] % every == 0][: max(0, num_nonconv - nonconv)]
        sel_c = [e for e in range(E) if ended[e]][: max(0, num_conv - conv)]
        sel = (sel_nc + sel_c)[: n - count]
        for e in sel:
            env = slots[e][0]
            ref["actions"].append(U.to_tensor(acts[e]))
            ref["robot_positions"].append(U.get_robot_positions(env))
            ref["coverage_features"].append(U.get_voronoi_features(env))
            maps = U.get_maps(env, p, C, True).cpu()
            ref["local_maps"].append(maps[:, 0])
            ref["comm_maps"].append(maps[:, 1:3])
            ref["obstacle_maps"].append(maps[:, 3])
            ref["edge_weights"].append(U.get_weights(env, p))
            ref["objectives"].append(torch.tensor(env.GetObjectiveValue(), dtype=torch.float32))
        count += len(sel)
        nonconv += len([e for e in sel if e in sel_nc])
        conv += len([e for e in sel if e in sel_c])
        for e in range(E):
            slots[e][0].StepActions(acts[e])
            steps[e] += 1
        for e in range(E):
            if ended[e]:
                slots[e] = fresh()
                steps[e] = 0
    for k in data:
        want = torch.stack(ref[k]).to(torch.float32)
        assert data[k].shape == want.shape, k
        tol = dict(rtol=1e-5, atol=1e-6) if k != "objectives" else dict(rtol=1e-5, atol=0)
        assert torch.allclose(data[k], want, **tol), k
    assert gen.env_count > E  # episodes ended and slots were re-initialised

    # ---- files as save_dataset writes them (data_generation.py:412-470) ---------------------------------
    save_dataset(data, tmp_path / "ds", split=(0.7, 0.2))
    tr = torch.load(tmp_path / "ds" / "train" / "local_maps.pt")
    assert tr.shape == (int(n * 0.7), R, C, C)
    assert (tmp_path / "ds" / "actions_mean.pt").exists() and (tmp_path / "ds" / "test" / "normalized_actions.pt").exists()
    total = sum(len(torch.load(tmp_path / "ds" / s / "actions.pt")) for s in ("train", "val", "test"))
    assert total == n
