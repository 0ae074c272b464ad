"""csrc/maps.cuh clip_cell_warp skips, 32 sites at a time, every site farther from the cell's own site than twice the
polygon's circumradius about it plus a margin, instead of testing its bisector against the vertices.  The claim: a
skipped site's half-plane test would have said "does not reach the cell", so the vertices stay bit for bit those of
the plain sequential clip (clip_cell).  Restated here operation for operation (the library is built with -fmad=false:
no contraction, Python floats reproduce every rounding) and held to exactly that over random, clustered and lattice
site sets, map sizes up to 16384 and sites on / outside the box."""
import math

import numpy as np


def clip_plain(sites, k, w, h, prune):
    """clip_cell (prune=False) / the order of operations of clip_cell_warp (prune=True); returns the vertex list and
    how many half-planes went through the vertex test."""
    kx, ky = sites[k]
    poly = [(0.0, 0.0), (w, 0.0), (w, h), (0.0, h)]
    margin = 1e-6 * (w + h)
    tested = 0

    def reach2():
        r2 = 0.0
        for vx, vy in poly:
            r2 = max(r2, (vx - kx) * (vx - kx) + (vy - ky) * (vy - ky))
        rr = 2.0 * math.sqrt(r2) + margin
        return rr * rr

    n = len(sites)
    for base in range(0, n, 32):
        if not poly:
            break
        chunk = list(range(base, min(base + 32, n)))
        d2 = {j: (sites[j][0] - kx) * (sites[j][0] - kx) + (sites[j][1] - ky) * (sites[j][1] - ky) for j in chunk}
        bound = reach2() if prune else math.inf
        todo = [j for j in chunk if j != k and d2[j] <= bound]
        while todo and poly:
            j = todo.pop(0)
            jx, jy = sites[j]
            nx, ny = jx - kx, jy - ky
            if nx == 0 and ny == 0:
                continue
            c = ((jx * jx + jy * jy) - (kx * kx + ky * ky)) / 2.0
            tested += 1
            dv = [nx * vx + ny * vy - c for vx, vy in poly]
            if not any(d > 0 for d in dv):
                continue  # the bisector does not reach the cell
            tmp = []
            na = len(poly)
            for v in range(na):
                u = 0 if v + 1 == na else v + 1
                (vx, vy), (ux, uy) = poly[v], poly[u]
                vin, uin = dv[v] <= 0, dv[u] <= 0
                if vin:
                    tmp.append((vx, vy))
                if vin != uin:
                    t = dv[v] / (dv[v] - dv[u])
                    tmp.append((vx + t * (ux - vx), vy + t * (uy - vy)))
            nb = len(tmp)
            poly = [tmp[v] for v in range(nb) if not (nb > 1 and tmp[v] == tmp[v - 1])]
            if prune:
                bound = reach2()
                todo = [q for q in todo if d2[q] <= bound]
    return poly, tested


def site_sets():
    rng = np.random.default_rng(0)
    for n, size in ((32, 1024.0), (64, 1024.0), (256, 8192.0), (300, 16384.0), (7, 64.0), (100, 512.0)):
        yield rng.uniform(0.0, size, (n, 2)), size, size
        yield np.concatenate([rng.uniform(0, size, (n // 2, 2)), rng.normal(size / 2, size * 0.004, (n - n // 2, 2))]), size, size
    g = np.arange(0.0, 1024.0, 128.0) + 64.0
    yield np.stack(np.meshgrid(g, g), -1).reshape(-1, 2), 1024.0, 1024.0                      # lattice: many exact ties
    yield np.array([[0.0, 0.0], [1024.0, 1024.0], [512.0, 0.0], [-30.0, 400.0], [1100.0, 90.0], [512.0, 512.0]]), 1024.0, 1024.0
    yield rng.uniform(0.0, 300.0, (40, 2)), 300.0, 120.0                                   # a window rectangle


def test_pruned_clip_equals_plain_clip_bit_for_bit():
    saved = total = 0
    for sites, w, h in site_sets():
        pts = [(float(x), float(y)) for x, y in sites]
        for k in range(len(pts)):
            a, ta = clip_plain(pts, k, w, h, False)
            b, tb = clip_plain(pts, k, w, h, True)
            assert a == b, (len(pts), k)
            assert tb <= ta
            saved += ta - tb
            total += ta
    assert saved > 0.5 * total  # the point of it: most half-plane tests are skipped
