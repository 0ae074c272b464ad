"""The arithmetic of the block-prefix walk (csrc/vglobal.cuh: vg_add4 / vg_add8 / vg_pfx_step / dd_add / dd_diff and the
carried P(T) of k_vg_partition<1, true>) restated operation for operation on the CPU -- every fp64 operation rounded
once, fma through exact rationals -- and held against exact rational sums.  It pins the accuracy argument of
DESIGN.md 3.1a without a GPU: an interval sum is off by at most a few ulp of the PARTIAL BLOCKS at its two boundaries
(the <= 7 cells between a block edge and the boundary, added in plain fp64) plus ~1e-30 of the row's running sums
(the double-double entries) -- never by an ulp of the running sums themselves; zero runs are exactly 0; a cell of mass
1e-9 behind a prefix of 1e3 is intact.  (tests/test_gpu_vglobal.py holds the kernel itself against the oracle.)"""
from fractions import Fraction

import numpy as np

PB = 8


def rn(x: Fraction) -> float:
    return float(x)  # Fraction -> float rounds to nearest, ties to even: one IEEE rounding


def fma(a: float, b: float, c: float) -> float:
    return rn(Fraction(a) * Fraction(b) + Fraction(c))


def add(a: float, b: float) -> float:
    return a + b


def two_sum(a, b):
    s = a + b
    bb = s - a
    return s, (a - (s - bb)) + (b - bb)


def quick(a, b):
    s = a + b
    return s, b - (s - a)


def dd_add(a, b):
    s = two_sum(a[0], b[0])
    t = two_sum(a[1], b[1])
    s = quick(s[0], s[1] + t[0])
    return quick(s[0], s[1] + t[1])


def dd_diff(a, b):
    d = two_sum(a[0], -b[0])
    return d[0] + (d[1] + (a[1] - b[1]))


def add4(x, y0, acc):
    v0, v1, v2, v3 = (float(v) for v in x)
    s = (v0 + v1) + (v2 + v3)
    t = fma(3.0, v3, fma(2.0, v2, v1))
    u = fma(9.0, v3, fma(4.0, v2, v1))
    acc[0] = acc[0] + s
    acc[1] = acc[1] + fma(y0, s, t)
    acc[2] = acc[2] + fma(y0 * y0, s, fma(y0 + y0, t, u))


def add8(cells, keep, y0):
    x = [float(c) if (keep >> i) & 1 else 0.0 for i, c in enumerate(cells)]
    acc = [0.0, 0.0, 0.0]
    add4(x[:4], y0, acc)
    add4(x[4:], y0 + 4.0, acc)
    return acc


def build_entries(row):
    """k_vg_prefix: E[b + 1] = E[b] (+) blocksum(b), double-double, one entry per 8 columns."""
    n = len(row)
    e = [((0.0, 0.0), (0.0, 0.0), (0.0, 0.0))]
    for b in range(n // PB):
        bs = add8(row[PB * b:PB * b + PB], 0xFF, PB * b + 0.5)
        e.append(tuple(dd_add(e[-1][k], (bs[k], 0.0)) for k in range(3)))
    return e


def prefix_at(row, entries, T):
    """P(T) of the walk: the entry of T's block (+) the cells from the block edge to T."""
    b, f = T >> 3, T & (PB - 1)
    p = entries[b]
    if f:
        h = add8(row[PB * b:PB * b + PB], (1 << f) - 1, PB * b + 0.5)
        p = tuple(dd_add(p[k], (h[k], 0.0)) for k in range(3))
    return p


def exact(row, lo, hi):
    s = sy = syy = Fraction(0)
    for c in range(lo, hi):
        v, y = Fraction(float(row[c])), Fraction(2 * c + 1, 2)
        s, sy, syy = s + v, sy + y * v, syy + y * y * v
    return s, sy, syy


def check_row(row, cuts, rel):
    entries = build_entries(row)
    total = exact(row, 0, len(row))
    prev = prefix_at(row, entries, 0)
    assert prev == ((0.0, 0.0),) * 3
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        cur = prefix_at(row, entries, hi)
        got = [dd_diff(cur[k], prev[k]) for k in range(3)]
        want = exact(row, lo, hi)
        part_lo, part_hi = exact(row, lo & ~(PB - 1), lo), exact(row, hi & ~(PB - 1), hi)
        for k in range(3):
            err = abs(Fraction(got[k]) - want[k])
            # the final rounding of the interval's own sum, a few ulp of the two partial blocks, ~1e-30 of the running sums
            bound = (abs(want[k]) * Fraction(2) ** -52 + (part_lo[k] + part_hi[k]) * Fraction(2) ** -50
                     + abs(total[k]) * Fraction(rel))
            assert err <= bound, (lo, hi, k, float(err), float(bound))
        prev = cur


def test_interval_sums_from_carried_prefixes_are_exact_to_double_double():
    rng = np.random.default_rng(0)
    for n in (64, 256, 1024):
        row = (rng.random(n) ** 8).astype(np.float32) * np.float32(3.0)
        cuts = sorted({0, n, *rng.integers(0, n + 1, 9).tolist()})
        check_row(row, cuts, 1e-30)


def test_low_mass_cell_behind_a_heavy_prefix_keeps_its_digits():
    n = 1024
    row = np.full(n, 2.0, np.float32)  # running sums ~2e3 (mass), ~1e6, ~7e8 (moments)
    row[-40:] = 0.0
    row[-1] = np.float32(1e-9)
    entries = build_entries(row)
    lo = n - 40
    got = [dd_diff(prefix_at(row, entries, n)[k], prefix_at(row, entries, lo)[k]) for k in range(3)]
    want = exact(row, lo, n)
    for k in range(3):
        assert abs(Fraction(got[k]) - want[k]) <= abs(want[k]) * Fraction(2) ** -50, k
    assert got[0] > 1e-10  # (the `mass < 1e-10` fallback of voronoi.cpp:41-53 must not fire: 1e-9 > 1e-10)


def test_light_cell_next_to_heavy_cells_of_the_same_block():
    """The corner the partial blocks leave: cells outside the interval but inside its boundary block are added and
    subtracted in fp64, so the error is an ulp of THEM: 1e-9 next to 2.0 keeps ~6 digits, enough for the 1e-4 gate."""
    n = 64
    row = np.zeros(n, np.float32)
    row[40:43] = 2.0
    row[43] = np.float32(1e-9)
    entries = build_entries(row)
    got = [dd_diff(prefix_at(row, entries, 44)[k], prefix_at(row, entries, 43)[k]) for k in range(3)]
    want = exact(row, 43, 44)
    for k in range(3):
        assert abs(Fraction(got[k]) - want[k]) <= abs(want[k]) * Fraction(1, 10 ** 6), k


def test_a_run_of_zero_cells_sums_to_exactly_zero():
    rng = np.random.default_rng(1)
    n = 512
    row = rng.random(n).astype(np.float32)
    row[100:300] = 0.0
    entries = build_entries(row)
    for lo, hi in ((100, 300), (101, 299), (104, 296), (107, 108), (150, 151), (100, 101)):
        a, b = prefix_at(row, entries, lo), prefix_at(row, entries, hi)
        assert a == b, (lo, hi)  # bitwise equal double-doubles
        assert [dd_diff(b[k], a[k]) for k in range(3)] == [0.0, 0.0, 0.0]
