"""CPU cross-check of the oracle's ClairvoyantCVT loop (oracle/cc_oracle.c orc_ctrl_compute_actions +
orc_env_step_actions) against an independent NumPy model written from the reference: goals = centroids of the
nearest-site cells over the world map (clairvoyant_cvt.h:89-99), action = min(vmax, dist / dt) * unit(diff), zero
when dist < 1e-10 or the robot oscillates (clairvoyant_cvt.h:101-127; CheckOscillation coverage_system.h:341-363:
more than 2 of the last min(6, len - 1) history entries within 1e-4 of the last one), kinematics with the speed
clamp and the [1e-4, N res - 1e-4] position clamp (robot_model.h:197-256), and the position history that pops
WITHOUT pushing when it is full (coverage_system.cpp:263-271)."""
from collections import deque

import numpy as np
import pytest

from oracle import oracle as orc


class Model:
    def __init__(self, world, pos, res, vmax, dt, hist_size):
        self.w = world.astype(np.float64)
        self.N = world.shape[0]
        self.res, self.vmax, self.dt, self.H = res, vmax, dt, hist_size
        self.start = np.array(pos, np.float64)
        self.local = np.zeros_like(self.start)
        self.pos = self.start.copy()
        self.hist = [deque() for _ in pos]
        self._post_step()

    def _post_step(self):
        for r, h in enumerate(self.hist):
            if len(h) > 0 and len(h) == self.H:
                h.popleft()
            else:
                h.append(self.pos[r].copy())

    def oscillates(self, r):
        h = self.hist[r]
        if len(h) < 3:
            return False
        last = h[-1]
        look = list(h)[::-1][: min(6, len(h) - 1)]
        return sum(np.hypot(*(p - last)) < 1e-4 for p in look) > 2

    def goals(self):
        c = (np.arange(self.N) + 0.5) * self.res
        dx = c[:, None, None] - self.pos[None, None, :, 0]
        dy = c[None, :, None] - self.pos[None, None, :, 1]
        lab = np.argmin(dx * dx + dy * dy, axis=2)
        g = np.zeros_like(self.pos)
        for k in range(len(self.pos)):
            sel = lab == k
            m = self.w[sel].sum()
            assert m > 1e-10  # the worlds of this test have mass everywhere
            g[k] = [(self.w * c[:, None])[sel].sum() / m, (self.w * c[None, :])[sel].sum() / m]
        return g

    def actions(self):
        g = self.goals()
        a = np.zeros_like(self.pos)
        for r in range(len(self.pos)):
            diff = g[r] - self.pos[r]
            dist = np.hypot(*diff)
            if dist < 1e-10 or self.oscillates(r):
                continue
            a[r] = min(self.vmax, dist / self.dt) * (diff / dist)
        return a, g

    def step(self, a):
        ext = self.N * self.res
        for r in range(len(self.pos)):
            sp = np.hypot(*a[r])
            d = a[r] / sp if sp > 0 else a[r]
            sp = min(sp, self.vmax)
            n = np.hypot(*d)
            d = d / n if n > 0 else d
            new_local = self.local[r] + d * sp * self.dt if sp > 1e-10 else self.local[r]
            gpos = new_local + self.start[r]
            gpos = np.where(gpos <= 0, 1e-4, gpos)
            gpos = np.where(gpos >= ext, ext - 1e-4, gpos)
            self.local[r] = gpos - self.start[r]
            self.pos[r] = gpos
        self._post_step()


@pytest.mark.parametrize("N,R,H,steps,seed", [(96, 4, 20, 60, 0), (64, 3, 5, 80, 1), (80, 6, 4, 50, 2)])
def test_clairvoyant_rollout_against_numpy(N, R, H, steps, seed):
    rng = np.random.default_rng(seed)
    i = np.arange(N) + 0.5
    world = (0.05 + np.exp(-((i[:, None] - 0.3 * N) ** 2 + (i[None, :] - 0.6 * N) ** 2) / (2 * (0.2 * N) ** 2))
             + 0.5 * rng.random((N, N))).astype(np.float32)
    p = orc.default_params(pWorldMapSize=N, pRobotMapSize=N, pSensorSize=16, pLocalMapSize=32, pNumRobots=R,
                           pRobotPosHistorySize=H, pMaxRobotSpeed=3.0)
    pos = rng.uniform(2, N - 2, (R, 2))
    env = orc.Env(p, world, pos)
    model = Model(world, pos, 1.0, 3.0, p.pTimeStep, H)
    stopped = 0
    for step in range(steps):
        conv, a, g = env.compute_actions(0)
        ma, mg = model.actions()
        assert np.allclose(g, mg, rtol=0, atol=1e-8), f"goals differ at step {step}"
        assert np.allclose(a, ma, rtol=0, atol=1e-7), f"actions differ at step {step}"
        assert conv == (not ma.any())
        stopped += int(np.sum(~ma.any(axis=1)))
        assert env.step_actions(a) == 0
        model.step(ma)
        assert np.allclose(env.positions(), model.pos, rtol=0, atol=1e-7), f"positions differ at step {step}"
    assert stopped > 0  # robots did reach their centroid / the oscillation rule did fire during the roll-out


@pytest.mark.parametrize("N,R,S,L,comm,seed", [(128, 5, 24, 64, 50.0, 3), (96, 4, 16, 48, 200.0, 4)])
def test_decentralized_rollout_against_numpy(N, R, S, L, comm, seed):
    """decentralized_cvt.h:85-173: every robot reduces ITS OWN map (the cells its sensor has covered so far) over the
    clipped L-window around itself, among itself and the robots strictly within pCommunicationRange; the goal is the
    centroid of its own cell in world coordinates"""
    from test_oracle_maps_cpu import Model as MapModel, _window

    rng = np.random.default_rng(seed)
    world = (0.1 + rng.random((N, N))).astype(np.float32)
    p = orc.default_params(pWorldMapSize=N, pRobotMapSize=N, pSensorSize=S, pLocalMapSize=L, pNumRobots=R,
                           pCommunicationRange=comm, pMaxRobotSpeed=3.0)
    pos = rng.uniform(2, N - 2, (R, 2))
    pos[0] = [4.0, 5.5]  # local window clipped by two borders
    env = orc.Env(p, world, pos)
    kin = Model(world, pos, 1.0, 3.0, p.pTimeStep, p.pRobotPosHistorySize)
    maps = MapModel(world, env.positions(), S, L, 1.0)
    c = np.arange(N) + 0.5
    for step in range(25):
        conv, a, g = env.compute_actions(2)
        cur = kin.pos
        mg = np.zeros_like(cur)
        for r in range(R):
            _, _, x0, x1, y0, y1 = _window(cur[r], L, N, 1.0)
            d = cur - cur[r]
            nb = [j for j in range(R) if j != r and np.hypot(*d[j]) < comm]
            sites = np.vstack([cur[r:r + 1], cur[nb]])
            cx, cy = c[x0:x1], c[y0:y1]
            d2 = (cx[:, None, None] - sites[None, None, :, 0]) ** 2 + (cy[None, :, None] - sites[None, None, :, 1]) ** 2
            own = np.argmin(d2, axis=2) == 0
            v = maps.robot_map(r)[x0:x1, y0:y1].astype(np.float64) * own
            m = v.sum()
            assert m > 1e-10
            mg[r] = [(v * cx[:, None]).sum() / m, (v * cy[None, :]).sum() / m]
        assert np.allclose(g, mg, rtol=0, atol=1e-8), f"goals differ at step {step}"
        ma = np.zeros_like(cur)
        for r in range(R):
            diff = mg[r] - cur[r]
            dist = np.hypot(*diff)
            if dist >= 1e-10 and not kin.oscillates(r):
                ma[r] = min(3.0, dist / p.pTimeStep) * (diff / dist)
        assert np.allclose(a, ma, rtol=0, atol=1e-7), f"actions differ at step {step}"
        assert env.step_actions(a) == 0
        kin.step(ma)
        maps.visit(env.positions())
        assert np.allclose(env.positions(), kin.pos, rtol=0, atol=1e-7)


def test_centralized_rollout_against_numpy():
    """centralized_cvt.h:90-130: global cells over the system's explored-IDF map (the world where ANY robot's sensor
    has been, 0 elsewhere)"""
    from test_oracle_maps_cpu import Model as MapModel

    rng = np.random.default_rng(6)
    N, R, S = 112, 5, 24
    world = (0.1 + rng.random((N, N))).astype(np.float32)
    p = orc.default_params(pWorldMapSize=N, pRobotMapSize=N, pSensorSize=S, pLocalMapSize=64, pNumRobots=R, pMaxRobotSpeed=3.0)
    pos = rng.uniform(2, N - 2, (R, 2))
    env = orc.Env(p, world, pos)
    kin = Model(world, pos, 1.0, 3.0, p.pTimeStep, p.pRobotPosHistorySize)
    maps = MapModel(world, env.positions(), S, 64, 1.0)
    for step in range(25):
        kin.w = np.where(maps.seen.any(0), world, np.float32(0)).astype(np.float64)
        conv, a, g = env.compute_actions(1)
        ma, mg = kin.actions()
        assert np.allclose(g, mg, rtol=0, atol=1e-8), f"goals differ at step {step}"
        assert np.allclose(a, ma, rtol=0, atol=1e-7), f"actions differ at step {step}"
        assert env.step_actions(a) == 0
        kin.step(ma)
        maps.visit(env.positions())
        assert np.allclose(env.positions(), kin.pos, rtol=0, atol=1e-7)
