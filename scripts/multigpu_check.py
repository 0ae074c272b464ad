"""Row-band tiling over G GPUs (NCCL): run with
    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port 29511 scripts/multigpu_check.py
Every rank generates its band of the IDF, the bands are normalised with the all-reduced maximum and
all-gathered; then a ClairvoyantCVT roll-out on the banded world (partials -> ncclAllReduce ->
finalise) is compared, on every rank, with the same roll-out done on rank-local whole-map kernels."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist

    import coveragecontrol_b200 as ccb
    from conftest import random_world
    from coveragecontrol_b200 import BatchedCoverageSystem, BatchedCVT
    from coveragecontrol_b200.parallel import BandedCoverageSystem, BandedWorldIDF

    rank, local = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    N, R, F = int(os.environ.get("CHECK_N", 2048)), 64, 48
    p = ccb.Parameters()
    p.pWorldMapSize = p.pRobotMapSize = N
    p.pNumRobots = R
    bnd, pos = random_world(1, N, F, R)

    def make_idf():
        idf = ccb.WorldIDF(p, device=local)
        for b in bnd:
            idf.AddNormalDistribution(ccb.BivariateNormalDistribution(ccb.Point2(b[0], b[1]), float(b[2]), float(b[4])))
        return idf

    whole = make_idf()
    whole.GenerateMap()
    banded_idf = make_idf()
    BandedWorldIDF(banded_idf)
    a, b = np.array(whole.GetWorldMap()), np.array(banded_idf.GetWorldMap())
    assert np.array_equal(a.view(np.int32), b.view(np.int32)), "banded IDF differs from the whole-map IDF"

    # the peer-memory all-reduce alone: random vectors, bit-exact against the rank-ordered sum of an all-gather,
    # launch by launch and replayed from a CUDA graph; a deliberately late rank must not change a bit
    from coveragecontrol_b200.parallel import PeerExchange

    world = dist.get_world_size()
    n = 4 * R
    px = PeerExchange.create(n, local)
    collective = "nccl" if world > 1 else "none"
    if px is not None:
        collective = "peer"
        st = torch.cuda.Stream(device=local)
        gen = torch.Generator(device="cuda").manual_seed(1234 + rank)
        for it in range(40):
            v = (torch.rand(n, dtype=torch.float64, device="cuda", generator=gen) - 0.5) * 10.0 ** (it % 7)
            parts = [torch.empty_like(v) for _ in range(world)]
            dist.all_gather(parts, v)
            want = torch.zeros_like(v)
            for q in parts:
                want = want + q  # rank order
            if it % 5 == rank % 5:
                torch.cuda._sleep(2_000_000)  # ~1 ms late
            st.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(st):
                px.allreduce_sum_(v, st)
            st.synchronize()
            assert torch.equal(v, want), ("peer all-reduce differs from the rank-ordered sum", it)
        v = torch.zeros(n, dtype=torch.float64, device="cuda")
        src = torch.rand(n, dtype=torch.float64, device="cuda", generator=gen)
        g = torch.cuda.CUDAGraph()
        st.synchronize()
        with torch.cuda.graph(g, stream=st):
            v.copy_(src)
            px.allreduce_sum_(v, st)
        parts = [torch.empty_like(src) for _ in range(world)]
        dist.all_gather(parts, src)
        want = torch.zeros_like(src)
        for q in parts:
            want = want + q
        for it in range(25):
            with torch.cuda.stream(st):
                g.replay()
            st.synchronize()
            assert torch.equal(v, want), ("graph replay of the peer all-reduce", it)
        del g
        px.check()
    # every controller family on the tiled world against the same roll-out on rank-local whole-map kernels:
    # eager launches first, then the captured graph (partials -> ncclAllReduce -> finalise -> actions -> move, one graph)
    stream = torch.cuda.Stream(device=local)
    for kind, name in ((0, "ClairvoyantCVT"), (1, "CentralizedCVT"), (2, "DecentralizedCVT")):
        for graph in (False, True):
            plain = BatchedCoverageSystem(p, 1, R, device=local)
            plain.set_world(0, whole)
            plain.reset(pos[None])
            benv = BatchedCoverageSystem(p, 1, R, device=local, stream=stream.cuda_stream)
            benv.set_world(0, banded_idf)
            benv.reset(pos[None])
            banded = BandedCoverageSystem(benv, 0, stream=stream, graph=graph)
            assert banded.collective == collective, (banded.collective, collective)
            banded.compute_cells()
            c0, c1 = BatchedCVT(plain, kind), BatchedCVT(benv, kind)
            for it in range(10):
                c0.rollout(1)
                obj = banded.step(c1, with_objective=(it % 2 == 1))
                stream.synchronize()
                assert np.allclose(plain.positions(), benv.positions(), rtol=0, atol=1e-9), (name, graph, it)
                if obj is not None:
                    o = float(obj[0])
                    assert abs(o - plain.objective()[0]) <= 1e-10 * abs(o), (name, graph, it)
            banded.close()  # graphs with a captured all-reduce must go before the process group does
    # replicas stay identical across ranks
    mine = torch.from_numpy(benv.positions()).cuda()
    ref = mine.clone()
    dist.broadcast(ref, src=0)
    assert torch.equal(mine, ref), "replicated robot state diverged between ranks"
    dist.barrier()
    if rank == 0:
        print(f"multigpu_check ok: world_size={dist.get_world_size()} N={N} R={R} rows/rank={banded.rows} collective={collective}")
    torch.cuda.synchronize()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
