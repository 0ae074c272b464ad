"""The oracle's GetExplorationFeatures restatement (oracle/cc_oracle.c orc_env_exploration_features, reference
robot_model.h:347-403) against an independent NumPy evaluation of the same scores: the selected candidate points
lie on the reference's lattice, come out worst first, and none of the points left out scores higher than the worst
one kept (which of several equally scored points are kept is the priority queue's business:
tests/test_frontier_heap_cpu.py)."""
import numpy as np
import pytest


def _scores(p, world, seen, robot_map, pos):
    """benefit / distance of every candidate of robot_model.h:356-381, {(kx, ky): (qx, qy, value)}."""
    N, S, res = p.pWorldMapSize, p.pSensorSize, p.pResolution
    tsd = p.pMaxRobotSpeed * p.pTimeStep * res
    out = {}
    xs, ys = [], []
    for arr, c in ((xs, pos[0]), (ys, pos[1])):
        v, lim = max(c - 4 * tsd, 1e-4), min(c + 4 * tsd, float(N))
        while v < lim:
            arr.append(v)
            v += res
    unexplored_map = (~seen).astype(np.int64) if p.pUpdateExplorationMap else np.zeros((N, N), np.int64)
    nonzero = (robot_map != 0).astype(np.int64)
    for kx, qx in enumerate(xs):
        for ky, qy in enumerate(ys):
            dist = max(float(np.hypot(qx - pos[0], qy - pos[1])), tsd)
            # numpy's round is half-to-even, std::round half away from zero: positions are positive here
            cx, cy = int(np.floor(qx / res + 0.5)), int(np.floor(qy / res + 0.5))
            x0, x1 = max(cx - S // 2, 0), min(cx + S // 2, N)
            y0, y1 = max(cy - S // 2, 0), min(cy + S // 2, N)
            un = int(unexplored_map[x0:x1, y0:y1].sum())
            benefit = float(un)
            if un > 0.1 * S * S:
                benefit += 2.0 * int(nonzero[x0:x1, y0:y1].sum())
            out[(kx, ky)] = (qx, qy, benefit / dist)
    return out


@pytest.mark.parametrize("seed,expl,K", [(0, True, 10), (1, True, 4), (2, False, 10), (3, True, 25)])
def test_oracle_exploration_features_against_numpy(seed, expl, K):
    from oracle import oracle as orc

    rng = np.random.default_rng(seed)
    N, R, S = 96, 3, 16
    p = orc.default_params(pNumRobots=R)
    p.pWorldMapSize = p.pRobotMapSize = N
    p.pSensorSize, p.pLocalMapSize, p.pCommunicationRange = S, 2 * S, 2.0 * S
    p.pUpdateExplorationMap = int(expl)
    world = (rng.random((N, N)) + 0.5).astype(np.float32)  # strictly positive: seen <=> robot map non-zero
    world[rng.random((N, N)) < 0.25] = 0.0                  # ... except these, which only the count() sees
    pos = rng.uniform(2.0, N - 2.0, (R, 2))
    pos[0] = [1.5, N - 2.5]                                 # clipped windows and a clipped lattice
    env = orc.Env(p, world, pos)
    seen = np.zeros((R, N, N), bool)

    def mark():
        P = env.positions()
        for r in range(R):
            ix, of = orc.compute_offsets(p.pResolution, P[r, 0], P[r, 1], S, N)
            x0, y0 = ix.left + of.left, ix.bottom + of.bottom
            seen[r, x0:x0 + of.width, y0:y0 + of.height] = True

    mark()
    for step in range(4):
        P = env.positions()
        for r in range(R):
            feats = env.exploration_features(r, K).reshape(K, 2)
            sc = _scores(p, world, seen[r], np.asarray(env.robot_map(r)), P[r])
            by_point = {(round(qx - P[r, 0], 9), round(qy - P[r, 1], 9)): v for (qx, qy, v) in sc.values()}
            n = min(K, len(sc))
            vals = []
            for k in range(n):
                key = (round(feats[k, 0], 9), round(feats[k, 1], 9))
                assert key in by_point, f"robot {r}: {feats[k]} is not a candidate point"
                vals.append(by_point[key])
            assert np.all(feats[n:] == 0.0)
            assert len({tuple(f) for f in feats[:n]}) == n, "a candidate was returned twice"
            assert np.all(np.diff(vals) >= -1e-12), "not popped worst first"
            rest = sorted(by_point.values())
            assert vals[0] >= rest[-n] - 1e-12 if n < len(rest) else True, "a better candidate was left out"
            assert np.allclose(sorted(vals), rest[-n:], rtol=1e-12, atol=1e-12)
        env.step_actions(rng.uniform(-3.0, 3.0, (R, 2)))
        mark()
