"""CPU cross-check of the oracle's map state (oracle/cc_oracle.c env_*) against an
independent NumPy model written from the reference: a cell of a robot's map holds the world value once it has
been inside the robot's sensor window (robot_model.h:98-107, 220-256), the system map is the world value where
any robot has seen the cell and -1 elsewhere (coverage_system.h:104-120), local / obstacle maps are zero- / one-
padded windows around round(p / res) (robot_model.h:293-345, map_utils.h:51-106), neighbours are the robots
strictly within pCommunicationRange (coverage_system.cpp:274-291) and the communication maps are the reference's
torch expression (coverage_env_utils.py:164-189), evaluated here with torch on the CPU."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc


def _rnd(v):  # std::round
    return int(np.floor(abs(v) + 0.5) * np.sign(v))


def _window(p, size, N, res):
    """rows / columns of the map inside the size-window around p and where they land in the window"""
    cx, cy = _rnd(p[0] / res), _rnd(p[1] / res)
    ox, oy = cx - size // 2, cy - size // 2
    x0, x1 = max(ox, 0), min(cx + size // 2, N)
    y0, y1 = max(oy, 0), min(cy + size // 2, N)
    return ox, oy, x0, max(x1, x0), y0, max(y1, y0)


def _outside(p, size, N, res):  # map_utils.h:140-155
    lim = N * res + size * res / 2.0
    return p[0] <= -size * res / 2.0 or p[1] <= -size * res / 2.0 or p[0] >= lim or p[1] >= lim


class Model:
    def __init__(self, world, pos, S, L, res):
        self.w, self.N, self.S, self.L, self.res = world, world.shape[0], S, L, res
        self.R = len(pos)
        self.seen = np.zeros((self.R, self.N, self.N), bool)
        self.visit(pos)

    def visit(self, pos):
        self.pos = np.array(pos)
        for r, p in enumerate(self.pos):
            if _outside(p, self.S, self.N, self.res):
                continue
            _, _, x0, x1, y0, y1 = _window(p, self.S, self.N, self.res)
            self.seen[r, x0:x1, y0:y1] = True

    def robot_map(self, r):
        return np.where(self.seen[r], self.w, np.float32(0))

    def padded(self, r, size, src, pad):
        out = np.full((size, size), pad, np.float32)
        p = self.pos[r]
        if _outside(p, size, self.N, self.res):
            return out
        ox, oy, x0, x1, y0, y1 = _window(p, size, self.N, self.res)
        out[x0 - ox:x1 - ox, y0 - oy:y1 - oy] = src[x0:x1, y0:y1]
        return out

    def system_map(self):
        any_seen = self.seen.any(0)
        return np.where(any_seen, self.w, np.float32(0)) - np.where(any_seen, np.float32(0), np.float32(1))


def _comm_maps_torch(rel, C, comm_range, res):
    """coverage_env_utils.py:164-189 for one robot; rel = relative neighbour positions [K, 2] (float64)"""
    maps = torch.zeros((2, C, C))
    if len(rel) == 0:
        return maps.numpy()
    neighbors_pos = torch.from_numpy(np.asarray(rel)).float()
    scaled = torch.round(neighbors_pos * C / (comm_range * res * 2.0) + (C / 2.0 - res / 2.0))
    idx = torch.transpose(scaled, 1, 0).long()
    val = neighbors_pos / comm_range
    # pResolution < 1 pushes the index of a far neighbour past the map; torch's unchecked sparse -> dense conversion
    # then writes at the wrapped linear offset (or past the buffer): undefined behaviour the restatement and the CUDA
    # path do not imitate -- such neighbours are dropped
    ok = ((idx >= 0) & (idx < C)).all(0)
    idx, val = idx[:, ok], val[ok]
    maps[0] = torch.sparse_coo_tensor(idx, val[:, 0], torch.Size([C, C])).to_dense()
    maps[1] = torch.sparse_coo_tensor(idx, val[:, 1], torch.Size([C, C])).to_dense()
    return maps.numpy()


@pytest.mark.parametrize("N,S,L,R,res,seed", [(256, 32, 96, 7, 1.0, 0), (192, 16, 64, 5, 1.0, 1), (128, 24, 128, 4, 0.5, 2)])
def test_map_state_after_a_random_walk(N, S, L, R, res, seed):
    rng = np.random.default_rng(seed)
    world = rng.random((N, N)).astype(np.float32)
    p = orc.default_params(pWorldMapSize=N, pRobotMapSize=N, pSensorSize=S, pLocalMapSize=L, pNumRobots=R, pResolution=res,
                           pCommunicationRange=0.35 * N * res, pMaxRobotSpeed=5.0 * res)
    ext = N * res
    pos = rng.uniform(0, ext, (R, 2))
    pos[0] = [1e-3, 0.4 * ext]            # windows clipped on every side
    pos[1] = [ext - 1e-3, ext - 1e-3]
    pos[2] = [0.5 * ext, 2.5 * res]
    env = orc.Env(p, world, pos)
    model = Model(world, env.positions(), S, L, res)
    for step in range(12):
        ang = rng.uniform(0, 2 * np.pi, R)
        speed = rng.uniform(0, 5.0 * res, R)
        speed[3 % R] = 0.0  # a robot that stays where it is
        act = np.stack([np.cos(ang), np.sin(ang)], 1) * speed[:, None]
        assert env.step_actions(act) == 0
        model.visit(env.positions())
        if step % 4 != 3:
            continue
        assert np.array_equal(env.system_map(), model.system_map())
        assert np.array_equal(env.explored_map(), np.where(model.seen.any(0), world, np.float32(0)))
        assert np.array_equal(env.exploration_map(), np.where(model.seen.any(0), np.float32(0), np.float32(1)))
        for r in range(R):
            assert np.array_equal(env.robot_map(r), model.robot_map(r)), f"robot map {r}"
            assert np.array_equal(env.sensor_view(r), model.padded(r, S, world, 0.0)), f"sensor view {r}"
            assert np.array_equal(env.local_map(r), model.padded(r, L, model.robot_map(r), 0.0)), f"local map {r}"
            assert np.array_equal(env.obstacle_map(r), model.padded(r, L, np.zeros((N, N), np.float32), 1.0)), f"obstacle map {r}"
            # neighbours: strictly inside the communication range, in ascending robot index, relative positions
            cur = model.pos
            d = cur - cur[r]
            nb = [j for j in range(R) if j != r and np.sqrt(d[j, 0] ** 2 + d[j, 1] ** 2) < p.pCommunicationRange]
            rel = env.neighbors(r)
            assert np.array_equal(rel, d[nb]), f"neighbours of {r}"
            ox, oy = env.comm_maps(r, 32, "py")
            ref = _comm_maps_torch(rel, 32, p.pCommunicationRange, res)
            assert np.allclose(ox, ref[0], atol=1e-6) and np.allclose(oy, ref[1], atol=1e-6), f"comm maps {r}"


def test_comm_map_rounding_differs_between_cpp_and_python_on_exact_halves():
    """C++ rounds half away from zero in fp64, the Python the network consumes rounds half to even
    in fp32 -- e.g. dx = 8 with range 128, C = 32 -> index 16.5 -> 17 (C++) / 16 (Python)"""
    N = 256
    p = orc.default_params(pWorldMapSize=N, pRobotMapSize=N, pSensorSize=16, pLocalMapSize=64, pNumRobots=2,
                           pCommunicationRange=128.0)
    world = np.zeros((N, N), np.float32)
    env = orc.Env(p, world, np.array([[100.0, 100.0], [108.0, 100.0]]))
    px, _ = env.comm_maps(0, 32, "py")
    cx, _ = env.comm_maps(0, 32, "cpp")
    assert np.argwhere(px != 0).tolist() == [[16, 16]]  # y index: 0 * .. + 15.5 -> 16 (half to even)
    assert np.argwhere(cx != 0).tolist() == [[17, 16]]
    ref = _comm_maps_torch(env.neighbors(0), 32, 128.0, 1.0)
    assert np.array_equal(px, ref[0])


def test_voronoi_features_against_numpy():
    """robot_model.h:408-430 GetVoronoiFeatures: the single-site "Voronoi" over the clipped L-window of the robot's
    own map.  Cell centres are taken relative to the robot (origin_shift = -site, voronoi.h:171), and the distance
    terms subtract the site once more (voronoi.h:89-95 -- the reference's double shift, kept as it is):
    [cx, cy, mass, sum |c - 2 s|^2 v, 0, sum |c - 2 s| v, 0] with c the centre and s the robot, both relative to the
    window origin; zero mass -> the rectangle's centroid relative to the robot (voronoi.cpp:41-53)."""
    rng = np.random.default_rng(4)
    N, S, L, R = 192, 32, 64, 5
    world = rng.random((N, N)).astype(np.float32)
    p = orc.default_params(pWorldMapSize=N, pRobotMapSize=N, pSensorSize=S, pLocalMapSize=L, pNumRobots=R)
    pos = rng.uniform(20, N - 20, (R, 2))
    pos[0] = [3.2, 100.7]       # window clipped by the border
    pos[1] = [N - 1.5, N - 2.5]
    env = orc.Env(p, world, pos)
    model = Model(world, env.positions(), S, L, 1.0)
    for _ in range(6):
        ang = rng.uniform(0, 2 * np.pi, R)
        env.step_actions(np.stack([np.cos(ang), np.sin(ang)], 1) * 4.0)
        model.visit(env.positions())
    for r in range(R):
        _, _, x0, x1, y0, y1 = _window(model.pos[r], L, N, 1.0)
        v = model.robot_map(r)[x0:x1, y0:y1].astype(np.float64)
        s = model.pos[r] - [x0, y0]
        cx = np.arange(x1 - x0) + 0.5
        cy = np.arange(y1 - y0) + 0.5
        px, py = cx[:, None] - s[0], cy[None, :] - s[1]          # relative to the robot
        mass = v.sum()
        d2 = (px - s[0]) ** 2 + (py - s[1]) ** 2                   # the site subtracted a second time
        ref = [(px * v).sum() / mass, (py * v).sum() / mass, mass, (d2 * v).sum(), 0.0, (np.sqrt(d2) * v).sum(), 0.0]
        assert np.allclose(env.voronoi_features(r), ref, rtol=1e-10, atol=1e-9), f"robot {r}"
    # nothing seen inside the window: the vertex mean of the window rectangle, relative to the robot
    blank = orc.Env(p, np.zeros((N, N), np.float32), pos)
    for r in range(R):
        _, _, x0, x1, y0, y1 = _window(pos[r], L, N, 1.0)
        s = blank.positions()[r] - [x0, y0]
        f = blank.voronoi_features(r)
        assert np.allclose(f[:2], [(x1 - x0) / 2.0 - s[0], (y1 - y0) / 2.0 - s[1]], atol=1e-12) and not f[2:].any()
