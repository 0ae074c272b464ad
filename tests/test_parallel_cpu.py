"""Host-side logic of the multi-GPU paths on CPU: partitioning helpers, and world_size-2 gloo runs of
the row-band reduce (band partials from the oracle -> all_reduce(SUM) == full-map moments) and of the
IDF band normalisation (band maxima -> all_reduce(MAX) == global maximum)."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT, random_world


def test_band_rows_partition():
    from coveragecontrol_b200.parallel import band_rows

    for n, g in [(1024, 1), (1024, 2), (8192, 8), (1000, 3), (300, 4), (65, 4), (3, 8)]:
        bands = [band_rows(n, g, r) for r in range(g)]
        assert bands[0][0] == 0 and bands[-1][1] == n
        for (a0, a1), (b0, b1) in zip(bands, bands[1:]):
            assert a1 == b0 and a0 <= a1
        if n % 64 == 0 and n // 64 >= g:
            assert all(a % 64 == 0 and b % 64 == 0 for a, b in bands)
    with pytest.raises(ValueError):
        band_rows(10, 2, 2)


def test_shard_envs():
    from coveragecontrol_b200.parallel import shard_envs

    for total, g in [(1024, 8), (10, 4), (3, 8), (0, 2)]:
        parts = [list(shard_envs(total, g, r)) for r in range(g)]
        assert sum(parts, []) == list(range(total))
        assert max(map(len, parts)) - min(map(len, parts)) <= 1


def test_ordered_float_encoding_is_monotone():
    from coveragecontrol_b200.parallel import float_to_ordered, ordered_to_float

    x = np.array([-np.inf, -3.5, -1e-30, -0.0, 0.0, 1e-38, 1.0, 7.25, np.inf], np.float32)
    o = float_to_ordered(x)
    assert np.all(np.diff(o.astype(np.int64)) >= 0)
    assert np.array_equal(ordered_to_float(o).view(np.int32), x.view(np.int32))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist

    from coveragecontrol_b200.parallel import (allreduce_max_, allreduce_sum_, band_rows, float_to_ordered,
                                               ordered_to_float)
    from oracle import oracle as orc

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        N, R = 256, 9
        bnd, pos = random_world(5, N, 6, R)
        p = orc.default_params(pWorldMapSize=N, pRobotMapSize=N, pNumRobots=R)
        raw, _ = orc.idf_generate(p, bnd, mode=1)
        r0, r1 = band_rows(N, world, rank)
        # IDF: the band maximum (as the kernel's ordered int) all-reduced with MAX is the map maximum
        mx = torch.from_numpy(float_to_ordered(np.array([raw[r0:r1].max() if r1 > r0 else -np.inf], np.float32)).copy())
        allreduce_max_(mx)
        assert ordered_to_float(mx.numpy())[0] == raw.max()
        # Voronoi: moments of the band's rows, all-reduced, are the full-map moments
        labels, full = orc.voronoi_bruteforce(pos, raw)
        ii, jj = np.meshgrid(np.arange(N) + 0.5, np.arange(N) + 0.5, indexing="ij")
        part = np.zeros((1, R, 4))
        for k in range(R):
            m = labels[r0:r1] == k
            v = raw[r0:r1][m].astype(np.float64)
            x, y = ii[r0:r1][m], jj[r0:r1][m]
            part[0, k] = [v.sum(), (x * v).sum(), (y * v).sum(), ((x * x + y * y) * v).sum()]
        t = torch.from_numpy(part)
        allreduce_sum_(t)
        tot = t.numpy()[0]
        assert np.allclose(tot[:, 0], full[:, 0], rtol=1e-12, atol=1e-12)
        # objective from the all-reduced moments == the oracle's direct sum
        obj = tot[:, 3] - 2 * (pos[:, 0] * tot[:, 1] + pos[:, 1] * tot[:, 2]) + (pos ** 2).sum(1) * tot[:, 0]
        assert np.allclose(obj, full[:, 3], rtol=1e-9, atol=1e-6)
        # the peer-memory exchange needs NCCL ranks with mapped device buffers: on a gloo group every rank gets
        # None alike and the all-reduce above stays the collective
        from coveragecontrol_b200.parallel import PeerExchange

        assert PeerExchange.create(4 * R, 0) is None
        q.put((rank, "ok"))
    except Exception as exc:  # surface the failure in the parent
        q.put((rank, repr(exc)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_row_band_reduce_world_size_2_gloo():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = sorted(q.get(timeout=240) for _ in procs)
    for pr in procs:
        pr.join(timeout=60)
    assert res == [(0, "ok"), (1, "ok")], res


def test_peer_exchange_buffer_layout():
    """csrc/peer.cuh: [2 parities][G ranks][count] doubles, flags [2][G] u64 after them on a 128-byte boundary."""
    from coveragecontrol_b200 import _lib

    lib = _lib.load()
    for world, count in ((2, 1024), (8, 1024), (8, 3), (1, 1)):
        data = 2 * world * count * 8
        flags = (data + 127) // 128 * 128
        assert lib.ccb_peer_buffer_bytes(world, count) == flags + 2 * world * 8
