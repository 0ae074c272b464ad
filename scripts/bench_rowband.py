"""BASELINE config 4: one large world tiled in row bands over G GPUs (NCCL all-reduce of the per-robot
moment partials).  Run on 1 GPU directly or under torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port 29520 \
        scripts/bench_rowband.py [--world 8192 --robots 256 --features 1024 --polygons 64 --steps 20]

Per step: ClairvoyantCVT ComputeActions + StepActions (band partials -> all-reduce -> finalise -> actions
-> move + sensor-window marks).  Prints one JSON line on rank 0 (steps/s, ms/step max over ranks, the
IDF generation time of the banded world)."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--world", type=int, default=8192)
    ap.add_argument("--robots", type=int, default=256)
    ap.add_argument("--features", type=int, default=1024)
    ap.add_argument("--polygons", type=int, default=64)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--kind", type=int, default=0, help="0 ClairvoyantCVT, 1 CentralizedCVT, 2 DecentralizedCVT")
    ap.add_argument("--eager", action="store_true", help="no CUDA graph (the ncu launch list shows every kernel)")
    ap.add_argument("--collective", default="auto", choices=["auto", "peer", "nccl"])
    a = ap.parse_args()
    import torch
    import torch.distributed as dist

    import coveragecontrol_b200 as ccb
    from coveragecontrol_b200 import BatchedCoverageSystem, BatchedCVT
    from coveragecontrol_b200.parallel import BandedCoverageSystem, BandedWorldIDF

    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    N, R = a.world, a.robots
    p = ccb.Parameters()
    p.pWorldMapSize = p.pRobotMapSize = N
    p.pNumRobots = R
    p.pUpdateExplorationMap = False
    p.pTimeStep = 0.2
    rng = np.random.default_rng(0)
    idf = ccb.WorldIDF(p, device=local)
    for _ in range(a.features):
        m = rng.uniform(0, N, 2)
        idf.AddNormalDistribution(ccb.BivariateNormalDistribution(ccb.Point2(m[0], m[1]), float(rng.uniform(40, 60)),
                                                                  float(rng.uniform(6, 10))))
    for _ in range(a.polygons):  # star-shaped simple polygons, 4-10 vertices, radius 64 (coverage_system.cpp:83-98)
        c = rng.uniform(64, N - 64, 2)
        nv = int(rng.integers(4, 11))
        ang = np.sort(rng.uniform(0, 2 * np.pi, nv))
        rad = rng.uniform(16, 64, nv)
        pts = np.stack([c[0] + rad * np.cos(ang), c[1] + rad * np.sin(ang)], 1)
        idf.AddUniformDistributionPolygon(ccb.PolygonFeature(ccb.PointVector(pts), float(5e-5 * rng.uniform(6, 10))))
    pos = rng.uniform(1e-4, N - 1e-4, (1, R, 2))

    s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s.record()
    BandedWorldIDF(idf)
    t.record()
    torch.cuda.synchronize()
    idf_ms = s.elapsed_time(t)

    stream = torch.cuda.Stream(device=local)
    env = BatchedCoverageSystem(p, 1, R, device=local, stream=stream.cuda_stream)
    env.set_world(0, idf)
    env.reset(pos)
    banded = BandedCoverageSystem(env, 0, stream=stream, graph=not a.eager, collective=a.collective)
    banded.compute_cells()
    ctrl = BatchedCVT(env, a.kind)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(a.warmup):
        banded.step(ctrl)
    barrier()
    with torch.cuda.stream(stream):
        s.record()
        for _ in range(a.steps):
            banded.step(ctrl)
        t.record()
    torch.cuda.synchronize()
    ms = torch.tensor([s.elapsed_time(t)], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    obj = banded.step(ctrl, with_objective=True)
    barrier()
    obj = [float(obj[0])]
    banded.close()
    if rank == 0:
        per = float(ms.item()) / a.steps
        print(json.dumps({"workload": f"C4 row bands: {N}x{N} world, {R} robots, {a.features} Gaussians + {a.polygons} polygons, ClairvoyantCVT",
                          "n_gpus": world, "collective": banded.collective, "rows_per_gpu": banded.rows[1] - banded.rows[0], "ms_per_step": per,
                          "env_steps_per_s": 1e3 / per, "algorithmic_GBps_per_gpu": 4.0 * N * (banded.rows[1] - banded.rows[0]) / (per * 1e-3) / 1e9,
                          "idf_gen_ms_banded": idf_ms, "objective": float(obj[0])}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
