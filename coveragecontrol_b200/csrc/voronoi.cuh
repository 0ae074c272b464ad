// voronoi.cuh -- discrete Voronoi assignment fused with the per-cell mass / moment reductions.
//
// Replaces reference cppsrc/core/src/voronoi.cpp:55-327 (CGAL arrangement + scan-line
// rasterisation + fp64 accumulation, one OpenMP thread per site) and the per-robot windowed
// variant of algorithms/decentralized_cvt.h:85-143.
//
// Semantics: every cell centre ((i+.5)res, (j+.5)res) is labelled with its nearest site; the
// distance is d2 = dx*dx + dy*dy evaluated in fp64 without FMA contraction and ties go to the
// lowest site index (the build's stated rule).  The labels are never stored: each
// warp streams a 64-column x <= 64-row region of the map once and accumulates, per label,
//   S0 = sum v, Sx = sum (i+.5) v, Sy = sum (j+.5) v, S2 = sum ((i+.5)^2 + (j+.5)^2) v
// in fp64 (cell units; a 24-bit v times a <= 15-bit half-integer is exact in fp64), from which
// mass, centroid and the objective sum |c - p|^2 v follow in the finalize kernel.
//
// How the exact labelling stays off the critical path:
//   * h_k(c) = d2_k(c) - d2_0(c) is AFFINE in the cell centre c.  A site that loses to another
//     site at all four corner cells of a region loses everywhere inside it: each warp culls the
//     site list to the 1-3 sites that can own a cell of its region (one site per lane, fp64, once
//     per region); regions with more owners are split in row halves and culled again;
//   * two owners (the common boundary case): along a column h changes sign at most once, so each
//     lane converts the bisector into one integer row threshold per column (fp64, once per region,
//     the two cells next to the threshold re-checked with the exact rule when they are within
//     rounding of the bisector) and the per-cell label is a single integer compare;
//   * three owners: h_k evaluated incrementally in fp32 per cell, with the fp64 rule only when
//     the best two are closer than a rigorous bound of the fp32 error;
//   * accumulation is predicated per surviving label: no per-cell atomics, no label runs; one
//     warp reduction and one set of fp64 atomics per (region, label).
//
// Memory: a lane owns 4 consecutive columns (one 16-byte load per row), lanes 0-15 take one row and
// lanes 16-31 the next, so one warp load is two full 256-byte row segments.  Loads are issued one
// group of passes ahead of their use, straight into registers: with ~24 warps resident per SM that
// keeps > 40 KB in flight per SM, more than HBM latency x bandwidth needs, without any shared-memory
// staging.  Masked sources (a robot's own map, the explored map) are BIT PLANES (common.cuh): the
// warp first loads the <= 128 mask words of its region; empty regions end there, empty row pairs
// are skipped, and world values are only requested for row pairs with a visible cell.
#pragma once

#include "common.cuh"

namespace ccb {

constexpr int VOR_THREADS = 256;
constexpr int VOR_WARPS = VOR_THREADS / 32;
constexpr int VOR_WX = 2, VOR_WY = 4;  // warp grid of a CTA tile (global mode)
constexpr int VOR_TW = 64;       // region width: 16 lanes x 4 cells; lanes 16-31 take the next row
constexpr int VOR_MAXRPW = 64;   // rows per warp region (4 mask words per lane cover 64 rows x 2 words)
constexpr int VOR_MAXFAST = 4;   // regions with more competing sites are split, then take the generic path
constexpr int VOR_MINSPLIT = 4;  // do not split below this many rows
constexpr int VOR_PAD_ROWS = 8;  // rows the streaming loop may read past a region (see stream_region)

enum { SRC_WORLD = 0, SRC_EXPLORED = 1, SRC_ROBOT = 2 };

struct VorArgs {
  EnvDev d;
  int rpw;      // rows per warp region (even, <= VOR_MAXRPW)
  int tiles_x, tiles_y;
  double *sums; // [E][R][4]
  int *labels;  // optional [E][N*N] (global mode only): the label of every cell, for parity tests
};

__device__ __forceinline__ double cell_centre(int i, double res) {
  return __dmul_rn(__dadd_rn((double)i, 0.5), res);
}
__device__ __forceinline__ double d2_exact(double cx, double cy, double px, double py) {
  double dx = __dsub_rn(cx, px), dy = __dsub_rn(cy, py);
  return __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
}

struct Acc6 {
  double c0, c1, c2, c3, sx, sxx;
  __device__ __forceinline__ void zero() { c0 = c1 = c2 = c3 = sx = sxx = 0.0; }
};

// What every warp of a CTA needs to stream its regions: one copy per CTA in shared memory (fields are
// read where they are used instead of living in registers across the whole kernel).
struct VorCta {
  int N, R, NW;        // map size, robots per environment, mask words per row
  double res;
  float fill;          // value of not-yet-seen cells of a robot map (0 unless pRobotMapUseUnknownImportance)
  int use_sensor;
  const float *world;     // this environment's world map
  const uint32_t *mask;   // bit plane of the selected map (robot plane / explored plane); unused for SRC_WORLD
  int e, robot;        // robot: output slot in window mode
  const double *s_px, *s_py;  // site coordinates (shared memory)
  const int *s_out;           // site -> output slot
  double *sums;
  int *labels;                // optional label plane of this environment
};
// One warp's region, by value (registers).
struct StripCtx {
  int r0, r1;          // region rows [r0, r1), r1 - r0 <= VOR_MAXRPW
  int c0, c1;          // valid columns [c0, c1)
  int col_base;        // absolute column of lane 0's first cell (multiple of VOR_TW)
};

// ---- reduction / emission ---------------------------------------------------------------------
struct Sum4 {
  double m, sx, sy, s2;
};
__device__ __forceinline__ Sum4 fold_and_reduce(const Acc6 &a, int col) {
  // fold the per-column sums with this lane's column coordinates (cell units), then warp-reduce
  const double y0 = (double)col + 0.5, y1 = y0 + 1.0, y2 = y0 + 2.0, y3 = y0 + 3.0;
  Sum4 r;
  r.m = (a.c0 + a.c1) + (a.c2 + a.c3);
  r.sy = fma(y0, a.c0, fma(y1, a.c1, fma(y2, a.c2, y3 * a.c3)));
  r.s2 = a.sxx + fma(y0 * y0, a.c0, fma(y1 * y1, a.c1, fma(y2 * y2, a.c2, (y3 * y3) * a.c3)));
  r.sx = a.sx;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    r.m += __shfl_xor_sync(0xffffffffu, r.m, o);
    r.sx += __shfl_xor_sync(0xffffffffu, r.sx, o);
    r.sy += __shfl_xor_sync(0xffffffffu, r.sy, o);
    r.s2 += __shfl_xor_sync(0xffffffffu, r.s2, o);
  }
  return r;
}

__device__ __forceinline__ void emit(double *sums, int e, int R, int slot, double m, double sx, double sy, double s2) {
  if (m == 0.0 && sx == 0.0 && sy == 0.0 && s2 == 0.0) return;
  // the pointer comes out of a shared-memory struct: tell the compiler it is global memory, otherwise
  // every atomicAdd carries a shared-memory CAS-loop alternative
  double *dst = reinterpret_cast<double *>(__cvta_generic_to_global(sums)) + ((size_t)e * R + slot) * 4;
  atomicAdd(dst + 0, m);
  atomicAdd(dst + 1, sx);
  atomicAdd(dst + 2, sy);
  atomicAdd(dst + 3, s2);
}

// ---- loads ---------------------------------------------------------------------------------------
// this lane's four cells of one row; aligned = rows are 16-byte aligned (N % 4 == 0) and the four
// columns are inside the map; otherwise element loads of the columns flagged in inmap
__device__ __forceinline__ float4 load4(const float *p, bool vec, unsigned inmap) {
  if (vec) return __ldg(reinterpret_cast<const float4 *>(p));
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (inmap & 1u) v.x = __ldg(p + 0);
  if (inmap & 2u) v.y = __ldg(p + 1);
  if (inmap & 4u) v.z = __ldg(p + 2);
  if (inmap & 8u) v.w = __ldg(p + 3);
  return v;
}

// mask words of a region: lane l holds, for q = 0..3, the word of row r0 + 16 q + (l >> 1) and
// half l & 1 (columns col_base + 32 half ..), already clipped to the valid columns [c0, c1).
struct RegionMask {
  uint32_t w[4];
};
__device__ __forceinline__ bool load_region_mask(const VorCta &g, const StripCtx &c, int lane, RegionMask &m) {
  const int half = lane & 1;
  const int wd = (c.col_base >> 5) + half;
  const int cb = c.col_base + 32 * half;  // column of bit 0
  // columns [c0, c1) -> bits [lo, hi) of this word
  const int lo = min(max(c.c0 - cb, 0), 32), hi = min(max(c.c1 - cb, 0), 32);
  uint32_t keep = 0u;
  if (hi > lo) keep = ((hi - lo >= 32) ? 0xffffffffu : ((1u << (hi - lo)) - 1u)) << lo;
  uint32_t any = 0u;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int row = c.r0 + 16 * q + (lane >> 1);
    uint32_t v = 0u;
    if (row < c.r1 && wd < g.NW && keep) v = __ldg(g.mask + (size_t)row * g.NW + wd) & keep;
    m.w[q] = v;
    any |= v;
  }
  return __any_sync(0xffffffffu, any != 0u);
}

// ---- consumers: what a warp does with the four cells a lane holds in one pass ----------------------
// One owner: every cell of the region belongs to site k0.  Column sums c0..c3 plus the row-weighted
// sums of the lane's row total: 11 fp64 operations per pass.
template <bool LABELS>
struct ConsOne {
  Acc6 a;
  int k0_slot;  // output slot written as label
  __device__ __forceinline__ void init() { a.zero(); }
  __device__ __forceinline__ void pass(const VorCta &g, float v0, float v1, float v2, float v3, int u, double xd, int ra,
                                       int rb, int col, unsigned in4) {
    if (LABELS) {
      if (g.labels && ra + u < rb) {
        int *lp_out = g.labels + (size_t)(ra + u) * g.N + col;
        if (in4 & 1u) lp_out[0] = k0_slot;
        if (in4 & 2u) lp_out[1] = k0_slot;
        if (in4 & 4u) lp_out[2] = k0_slot;
        if (in4 & 8u) lp_out[3] = k0_slot;
      }
    }
    const double dv0 = (double)v0, dv1 = (double)v1, dv2 = (double)v2, dv3 = (double)v3;
    a.c0 += dv0, a.c1 += dv1, a.c2 += dv2, a.c3 += dv3;
    const double s = (dv0 + dv1) + (dv2 + dv3);
    a.sx = fma(xd, s, a.sx);
    a.sxx = fma(xd * xd, s, a.sxx);
  }
};

// NCAND (2..4) owners: per-cell labels from the fp32 affine differences h_i = d2_i - d2_0 (the exact
// fp64 rule for a lane's four cells whenever the best two are within a rigorous bound of the fp32
// error), and per-COLUMN running sums (v, x v, x^2 v) that are flushed into the lane's per-owner
// accumulators (shared memory, private to the lane) whenever the column's label changes -- a Voronoi
// cell is convex, so a column changes owner at most NCAND - 1 times per region.  14 fp64 operations
// per pass whatever the number of owners; no thresholds, no per-owner predicated sets.  The cold part
// (exact labels, flushes) runs through a shared-memory scratch copy of the running sums with rolled
// loops, so that it costs little code and no registers in the streaming loop.
constexpr int RUNS_LANE_DOUBLES = VOR_MAXFAST * 4 + 12;  // per lane: owner accumulators + scratch

template <int NCAND, bool SINGLE, bool LABELS>
struct ConsRuns {
  float h0[NCAND - 1], ha[NCAND - 1], hb[NCAND - 1];  // h_i at (row ra, this lane's first column); per row; per column
  float margin;
  unsigned cur;                    // owner (candidate index) of each column's running sums, one per byte; 0xff: none yet
  double c[4], cx[4], cxx[4];
  double *lacc;                    // this lane's shared memory: lacc[(owner * 4 + v) * 32], v = m, sx, sy, s2;
                                   // scratch at lacc[(VOR_MAXFAST * 4 + k) * 32], k = 0..11
  const int *cand;                 // NCAND site indices, ascending (shared memory)

  __device__ __forceinline__ void init(const VorCta &g, int ra, int rows, int col, const int *cand_, double *lacc_) {
    cand = cand_, lacc = lacc_;
    const double res = g.res;
    const double cx0 = cell_centre(ra, res), cy0 = cell_centre(col, res);
    const int k0 = cand[0];
    const double p0x = g.s_px[k0], p0y = g.s_py[k0];
    const double f0 = d2_exact(cx0, cy0, p0x, p0y);
    double hmax = 0.0;
#pragma unroll
    for (int i = 1; i < NCAND; ++i) {
      const int k = cand[i];
      const double px = g.s_px[k], py = g.s_py[k];
      const double H0 = d2_exact(cx0, cy0, px, py) - f0;
      const double HA = 2.0 * res * (p0x - px), HB = 2.0 * res * (p0y - py);
      hmax = fmax(hmax, fabs(H0) + rows * fabs(HA) + 4.0 * fabs(HB));
      h0[i - 1] = (float)H0, ha[i - 1] = (float)HA, hb[i - 1] = (float)HB;
    }
    // fp32 evaluation error of h_i is < 2^-21 hmax; differences of two h's double it
    margin = (float)(hmax * (1.0 / 262144.0));
    cur = 0xffffffffu;
#pragma unroll
    for (int j = 0; j < 4; ++j) c[j] = cx[j] = cxx[j] = 0.0;
#pragma unroll
    for (int t = 0; t < NCAND * 4; ++t) lacc[t * 32] = 0.0;
  }

  // Cold: exact labels where the fp32 ones are in doubt, then every column whose label changed moves
  // its running sums to the previous owner's accumulators.  Returns the final labels.
  __device__ __forceinline__ unsigned change(const VorCta &g, unsigned labs, bool amb, int row, int col) {
    double *scr = lacc + VOR_MAXFAST * 4 * 32;
#pragma unroll
    for (int j = 0; j < 4; ++j) scr[j * 32] = c[j], scr[(4 + j) * 32] = cx[j], scr[(8 + j) * 32] = cxx[j];
    if (amb) {
      const double res = g.res;
      const double cxr = cell_centre(row, res);
      labs = 0;
#pragma unroll 1
      for (int j = 0; j < 4; ++j) {
        const double cy = cell_centre(col + j, res);
        int best = 0;
        double bd = d2_exact(cxr, cy, g.s_px[cand[0]], g.s_py[cand[0]]);
#pragma unroll 1
        for (int i = 1; i < NCAND; ++i) {
          const double dd = d2_exact(cxr, cy, g.s_px[cand[i]], g.s_py[cand[i]]);
          if (dd < bd) bd = dd, best = i;
        }
        labs |= (unsigned)best << (8 * j);
      }
    }
#pragma unroll 1
    for (int j = 0; j < 4; ++j) {
      const unsigned lj = (labs >> (8 * j)) & 0xffu, cj = (cur >> (8 * j)) & 0xffu;
      if (lj == cj) continue;
      if (cj != 0xffu && (!SINGLE || cj == 0u)) {
        const double y = (double)(col + j) + 0.5;
        const double cc = scr[j * 32];
        double *dst = lacc + cj * 128;
        dst[0] += cc;
        dst[32] += scr[(4 + j) * 32];
        dst[64] = fma(y, cc, dst[64]);
        dst[96] += fma(y * y, cc, scr[(8 + j) * 32]);
      }
      scr[j * 32] = 0.0, scr[(4 + j) * 32] = 0.0, scr[(8 + j) * 32] = 0.0;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) c[j] = scr[j * 32], cx[j] = scr[(4 + j) * 32], cxx[j] = scr[(8 + j) * 32];
    cur = labs;
    return labs;
  }

  __device__ __forceinline__ void pass(const VorCta &g, float v0, float v1, float v2, float v3, int u, double xd, int ra,
                                       int rb, int col, unsigned in4) {
    unsigned labs = 0;
    bool amb = false;
    {
      const float uf = (float)u;
      float hr[NCAND - 1];
#pragma unroll
      for (int i = 0; i < NCAND - 1; ++i) hr[i] = fmaf(uf, ha[i], h0[i]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float best = 0.f, second = 3.0e38f;  // candidate 0 has h = 0
        unsigned bl = 0;
#pragma unroll
        for (int i = 0; i < NCAND - 1; ++i) {
          const float gi = fmaf((float)j, hb[i], hr[i]);
          if (gi < best) second = best, best = gi, bl = (unsigned)(i + 1) << (8 * j);
          else second = fminf(second, gi);
        }
        amb |= (second - best < margin);
        labs |= bl;
      }
    }
    if (amb | (labs != cur)) labs = change(g, labs, amb, ra + u, col);
    if (LABELS && !SINGLE) {
      if (g.labels && ra + u < rb) {
        int *lp_out = g.labels + (size_t)(ra + u) * g.N + col;
        if (in4 & 1u) lp_out[0] = g.s_out[cand[labs & 0xffu]];
        if (in4 & 2u) lp_out[1] = g.s_out[cand[(labs >> 8) & 0xffu]];
        if (in4 & 4u) lp_out[2] = g.s_out[cand[(labs >> 16) & 0xffu]];
        if (in4 & 8u) lp_out[3] = g.s_out[cand[labs >> 24]];
      }
    }
    const double dv0 = (double)v0, dv1 = (double)v1, dv2 = (double)v2, dv3 = (double)v3;
    const double xd2 = xd * xd;
    c[0] += dv0, c[1] += dv1, c[2] += dv2, c[3] += dv3;
    cx[0] = fma(xd, dv0, cx[0]), cx[1] = fma(xd, dv1, cx[1]), cx[2] = fma(xd, dv2, cx[2]), cx[3] = fma(xd, dv3, cx[3]);
    cxx[0] = fma(xd2, dv0, cxx[0]), cxx[1] = fma(xd2, dv1, cxx[1]);
    cxx[2] = fma(xd2, dv2, cxx[2]), cxx[3] = fma(xd2, dv3, cxx[3]);
  }

  __device__ __forceinline__ void finish(const VorCta &g, int col) { change(g, 0xfefefefeu, false, 0, col); }
};

// ---- the streaming skeleton ------------------------------------------------------------------------
// Rows [ra, rb) of region c (ra - c.r0 even) go through cons.pass() once, a lane holding four
// consecutive cells of one row per pass (lanes 16-31: the next row).
//   plain (unmasked source, every lane inside the map and the region, even row count, 16-byte rows):
//     vector loads through a register ring that keeps U passes (2 KB per warp) in flight;
//   otherwise: per pass the visibility nibble comes from the region's mask words (one shuffle), row
//     pairs with nothing visible are skipped, cells are selected against the mask.
template <int SRC, class Cons>
__device__ __forceinline__ void stream_region(const VorCta &g, const StripCtx &c, int ra, int rb, uint32_t m0, uint32_t m1,
                                              uint32_t m2, uint32_t m3, Cons &cons) {
  constexpr int U = 4;
  const int lane = threadIdx.x & 31;
  const int hp = lane >> 4, c16 = lane & 15;
  const int col = c.col_base + 4 * c16;
  const int rows = rb - ra;
  const int npass = (rows + 1) >> 1;
  const int N = g.N;
  unsigned in4 = 0;    // inside the valid columns of the region
  unsigned inmap = 0;  // inside the map (loadable)
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (col + i >= c.c0 && col + i < c.c1) in4 |= 1u << i;
    if (col + i < N) inmap |= 1u << i;
  }
  const bool vec = ((N & 3) == 0) && inmap == 0xfu;
  const float *wp = g.world + (size_t)(ra + hp) * N + col;  // this lane's row of pass 0
  double xd = (double)(ra + hp) + 0.5;                      // its row coordinate (cell units)
  const bool plain = (SRC == SRC_WORLD) && ((rows & 1) == 0) && __all_sync(0xffffffffu, in4 == 0xfu && vec);
  if (plain) {
    const float4 *p4 = reinterpret_cast<const float4 *>(wp);
    const size_t s4 = (size_t)N >> 1;  // two rows, in float4 units
    int t = 0;
    if (npass >= U) {
      float4 ring[U];
#pragma unroll
      for (int k = 0; k < U; ++k) ring[k] = __ldg(p4 + k * s4);
      const float4 *pn = p4 + (size_t)U * s4;  // pass t + U
      // consume pass t + k, refill its slot with pass t + U + k.  The last iteration refills with rows
      // past the region (never used): the world buffer is padded by VOR_PAD_ROWS rows for that.
#pragma unroll 1
      for (; t + U <= npass; t += U) {
#pragma unroll
        for (int k = 0; k < U; ++k) {
          cons.pass(g, ring[k].x, ring[k].y, ring[k].z, ring[k].w, 2 * (t + k) + hp, xd + (double)(2 * k), ra, rb, col, in4);
          ring[k] = __ldg(pn + k * s4);
        }
        xd += 2.0 * U;
        pn += (size_t)U * s4;
      }
    }
#pragma unroll 1
    for (; t < npass; ++t) {  // fewer than U passes left
      const float4 w = __ldg(p4 + (size_t)t * s4);
      cons.pass(g, w.x, w.y, w.z, w.w, 2 * t + hp, xd, ra, rb, col, in4);
      xd += 2.0;
    }
  } else {
    const float fill = (SRC == SRC_ROBOT) ? g.fill : 0.f;
    const bool use_sensor = (SRC == SRC_WORLD) || g.use_sensor;
    const size_t pstride = (size_t)2 * N;
    const int um0 = ra - c.r0;  // even: pass t sits in mask word group (um0 + 2 t) >> 4
#pragma unroll 1
    for (int t = 0; t < npass; ++t, xd += 2.0) {
      const int u = 2 * t + hp;
      const unsigned inside = (ra + u < rb) ? in4 : 0u;
      unsigned vis = inside;
      if (SRC != SRC_WORLD) {
        const int um = um0 + 2 * t;  // + hp stays inside the same 16-row word group
        const int q = um >> 4;
        const uint32_t sel = q == 0 ? m0 : (q == 1 ? m1 : (q == 2 ? m2 : m3));
        const uint32_t word = __shfl_sync(0xffffffffu, sel, (((um + hp) & 15) << 1) | (c16 >> 3));
        vis &= (word >> (4 * (c16 & 7))) & 0xfu;
        if (!__any_sync(0xffffffffu, (fill != 0.f ? inside : vis) != 0u)) continue;  // nothing in this row pair
      }
      float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
      if (vis != 0u && use_sensor) w = load4(wp + (size_t)t * pstride, vec, inmap);
      float f0 = 0.f, f1 = 0.f, f2 = 0.f, f3 = 0.f;
      if (SRC == SRC_ROBOT && fill != 0.f) {
        f0 = (inside & 1u) ? fill : 0.f, f1 = (inside & 2u) ? fill : 0.f;
        f2 = (inside & 4u) ? fill : 0.f, f3 = (inside & 8u) ? fill : 0.f;
      }
      const float v0 = (vis & 1u) ? w.x : f0, v1 = (vis & 2u) ? w.y : f1;
      const float v2 = (vis & 4u) ? w.z : f2, v3 = (vis & 8u) ? w.w : f3;
      cons.pass(g, v0, v1, v2, v3, u, xd, ra, rb, col, in4);
    }
  }
}

// One owner (site index k0).  Not inlined: every streaming variant gets its own register allocation and
// the caller's state stays out of the streaming loops.
template <int SRC, bool SINGLE, bool LABELS>
__device__ __noinline__ void region_one(const VorCta *gp, StripCtx c, int ra, int rb, int k0, uint32_t m0, uint32_t m1,
                                        uint32_t m2, uint32_t m3) {
  const VorCta &g = *gp;
  const int lane = threadIdx.x & 31;
  ConsOne<LABELS && !SINGLE> cons;
  cons.init();
  cons.k0_slot = g.s_out[k0];
  stream_region<SRC>(g, c, ra, rb, m0, m1, m2, m3, cons);
  const Sum4 tot = fold_and_reduce(cons.a, c.col_base + 4 * (lane & 15));
  if (lane == 0) emit(g.sums, g.e, g.R, SINGLE ? g.robot : g.s_out[k0], tot.m, tot.sx, tot.sy, tot.s2);
}

// NCAND owners; cand (shared memory) holds their site indices, ascending.  lacc: this warp's per-lane
// owner accumulators (shared memory, [4][4][32] doubles).
template <int NCAND, int SRC, bool SINGLE, bool LABELS>
__device__ __noinline__ void region_runs(const VorCta *gp, StripCtx c, int ra, int rb, const int *cand, double *lacc,
                                         uint32_t m0, uint32_t m1, uint32_t m2, uint32_t m3) {
  const VorCta &g = *gp;
  const int lane = threadIdx.x & 31;
  const int col = c.col_base + 4 * (lane & 15);
  ConsRuns<NCAND, SINGLE, LABELS> cons;
  cons.init(g, ra, rb - ra, col, cand, lacc + lane);
  stream_region<SRC>(g, c, ra, rb, m0, m1, m2, m3, cons);
  cons.finish(g, col);
  // per owner: reduce the lanes' accumulators and emit
#pragma unroll 1
  for (int s = 0; s < (SINGLE ? 1 : NCAND); ++s) {
    const double *src = lacc + lane + s * 128;
    double m = src[0], sx = src[32], sy = src[64], s2 = src[96];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      m += __shfl_xor_sync(0xffffffffu, m, o);
      sx += __shfl_xor_sync(0xffffffffu, sx, o);
      sy += __shfl_xor_sync(0xffffffffu, sy, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if (lane == 0) emit(g.sums, g.e, g.R, SINGLE ? g.robot : g.s_out[cand[s]], m, sx, sy, s2);
  }
}

// direct loader for the generic path
template <int SRC>
__device__ __forceinline__ void load_cells_direct(const VorCta &g, const StripCtx &c, int row, int col, float v[4]) {
  const size_t off = (size_t)row * g.N + col;
  const float fill = (SRC == SRC_ROBOT) ? g.fill : 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[i] = 0.f;
    if (col + i < c.c0 || col + i >= c.c1) continue;
    bool visible = true;
    if (SRC != SRC_WORLD) visible = plane_bit(g.mask, g.NW, row, col + i);
    if (visible) v[i] = (SRC == SRC_WORLD || g.use_sensor) ? __ldg(g.world + off + i) : 0.f;
    else v[i] = fill;
  }
}

// Generic path: any number of candidates, fp64 rule per cell, per-lane label runs flushed with
// global atomics.  Only small regions that still contain more than VOR_MAXFAST owners come here.
template <int SRC>
__device__ __noinline__ void region_slow(const VorCta *gp, int c0, int c1, int col_base, int ra, int rb,
                                         const int *cand, int nc, bool single, int lane) {
  const VorCta &g = *gp;
  StripCtx c;
  c.r0 = ra, c.r1 = rb, c.c0 = c0, c.c1 = c1, c.col_base = col_base;
  const double res = g.res;
  const int col = col_base + 4 * (lane & 15);
  int cur = -1;
  double m = 0, sx = 0, sy = 0, s2 = 0;
  for (int row = ra + (lane >> 4); row < rb; row += 2) {
    float v[4];
    load_cells_direct<SRC>(g, c, row, col, v);
    const double cx = cell_centre(row, res);
    const double xd = (double)row + 0.5;
#pragma unroll 1
    for (int i = 0; i < 4; ++i) {
      if (v[i] == 0.f && !g.labels) continue;
      if (col + i < c0 || col + i >= c1) continue;
      const double cy = cell_centre(col + i, res);
      int best = 0;
      double bd = d2_exact(cx, cy, g.s_px[cand[0]], g.s_py[cand[0]]);
      for (int k = 1; k < nc; ++k) {
        double dd = d2_exact(cx, cy, g.s_px[cand[k]], g.s_py[cand[k]]);
        if (dd < bd) bd = dd, best = k;
      }
      if (single && best != 0) continue;
      const int slot = single ? g.robot : g.s_out[cand[best]];
      if (!single && g.labels) g.labels[(size_t)row * g.N + col + i] = slot;
      if (v[i] == 0.f) continue;
      if (slot != cur) {
        if (cur >= 0) {
          emit(g.sums, g.e, g.R, cur, m, sx, sy, s2);
        }
        cur = slot, m = sx = sy = s2 = 0;
      }
      const double dv = (double)v[i], yd = (double)(col + i) + 0.5;
      m += dv;
      sx = fma(xd, dv, sx);
      sy = fma(yd, dv, sy);
      s2 = fma(xd * xd + yd * yd, dv, s2);
    }
  }
  if (cur >= 0) {
    emit(g.sums, g.e, g.R, cur, m, sx, sy, s2);
  }
}

// Warp-level cull of the candidate list (lane k < nc holds candidate k) for the rectangle
// rows [r0, r1) x columns [c0, c1).  Two exact steps, both resting on h_k - h_j being affine:
//   1. j* = winner at the rectangle's centre cell; drop every k that loses to j* at all 4 corners;
//   2. among the survivors drop every k that loses to one and the same survivor j at all 4 corners.
// Returns the ballot of surviving lanes.
__device__ __noinline__ unsigned cull_rect(double res, int c0, int c1, int r0, int r1, bool has, double kx, double ky,
                                          int lane) {
  double fcen = INFINITY;
  if (has) fcen = d2_exact(cell_centre((r0 + r1 - 1) >> 1, res), cell_centre((c0 + c1 - 1) >> 1, res), kx, ky);
  double best = fcen;
  int bl = lane;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double ob = __shfl_xor_sync(0xffffffffu, best, o);
    int ol = __shfl_xor_sync(0xffffffffu, bl, o);
    if (ob < best || (ob == best && ol < bl)) best = ob, bl = ol;
  }
  // corner distances share their one-dimensional squares
  const double dx0 = __dsub_rn(cell_centre(r0, res), kx), dx1 = __dsub_rn(cell_centre(r1 - 1, res), kx);
  const double dy0 = __dsub_rn(cell_centre(c0, res), ky), dy1 = __dsub_rn(cell_centre(c1 - 1, res), ky);
  const double qx0 = __dmul_rn(dx0, dx0), qx1 = __dmul_rn(dx1, dx1);
  const double qy0 = __dmul_rn(dy0, dy0), qy1 = __dmul_rn(dy1, dy1);
  double f[4];
  f[0] = __dadd_rn(qx0, qy0), f[1] = __dadd_rn(qx1, qy0);
  f[2] = __dadd_rn(qx0, qy1), f[3] = __dadd_rn(qx1, qy1);
  bool keep = false;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const double fj = __shfl_sync(0xffffffffu, f[q], bl);
    if (has && f[q] - fj <= 1e-7 * (1.0 + fj)) keep = true;
  }
  unsigned mask = __ballot_sync(0xffffffffu, keep);
  if (__popc(mask) > 1) {
    // pairwise domination among the survivors (a dominated site may still be used as dominator)
    unsigned todo = mask;
    while (todo) {
      const int j = __ffs(todo) - 1;
      todo &= todo - 1;
      bool dominated = true;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const double fj = __shfl_sync(0xffffffffu, f[q], j);
        dominated = dominated && (f[q] - fj > 1e-7 * (1.0 + fj));
      }
      if (lane != j && dominated) keep = false;
    }
    mask = __ballot_sync(0xffffffffu, keep);
  }
  return mask;
}

// One warp, one region.  cand[0..nc) are site indices (ascending) valid for the whole CTA rectangle.
// Masked sources first load the region's mask words: an empty region ends there.  The region is
// culled; if more than VOR_MAXFAST sites can own cells it is split in row halves and each half is
// culled again (fewer rows, fewer owners) before falling back to the generic path.
template <int SRC, bool SINGLE, bool LABELS>
__device__ __forceinline__ void process_region(const VorCta &g, const StripCtx &c, const int *s_cand, int nc, int *wcand,
                                               double *lacc, int lane) {
  RegionMask rm;
  rm.w[0] = rm.w[1] = rm.w[2] = rm.w[3] = 0u;
  if (SRC != SRC_WORLD) {
    const bool any = load_region_mask(g, c, lane, rm);
    if (!any && !(SRC == SRC_ROBOT && g.fill != 0.f)) return;  // nothing visible: all values are zero
  }
  if (nc > 32) {
    region_slow<SRC>(&g, c.c0, c.c1, c.col_base, c.r0, c.r1, s_cand, nc, SINGLE, lane);
    return;
  }
  const bool has = lane < nc;
  const int my_site = has ? s_cand[lane] : 0;
  const double kx = g.s_px[my_site], ky = g.s_py[my_site];
  int r = c.r0;
  int len = c.r1 - r;
  while (r < c.r1) {
    if (len > c.r1 - r) len = c.r1 - r;
    const unsigned wmask = cull_rect(g.res, c.c0, c.c1, r, r + len, has, kx, ky, lane);
    const int wnc = __popc(wmask);
    if (wnc > VOR_MAXFAST && len > VOR_MINSPLIT) {
      len = ((len >> 1) + 1) & ~1;  // split (keep it even): retry the first half
      continue;
    }
    const int ra = r, rb = r + len;
    r += len;
    if (SINGLE && !(wmask & 1u)) continue;  // the robot owns no cell of these rows
    // the surviving candidates, ascending, into this warp's list
    {
      const bool keep = (wmask >> lane) & 1u;
      if (keep) wcand[__popc(wmask & ((1u << lane) - 1u))] = my_site;
      __syncwarp();
    }
    if (wnc == 1) {
      region_one<SRC, SINGLE, LABELS>(&g, c, ra, rb, wcand[0], rm.w[0], rm.w[1], rm.w[2], rm.w[3]);
    } else if (wnc == 2) {
      region_runs<2, SRC, SINGLE, LABELS>(&g, c, ra, rb, wcand, lacc, rm.w[0], rm.w[1], rm.w[2], rm.w[3]);
    } else if (wnc == 3) {
      region_runs<3, SRC, SINGLE, LABELS>(&g, c, ra, rb, wcand, lacc, rm.w[0], rm.w[1], rm.w[2], rm.w[3]);
    } else if (wnc == 4) {
      region_runs<4, SRC, SINGLE, LABELS>(&g, c, ra, rb, wcand, lacc, rm.w[0], rm.w[1], rm.w[2], rm.w[3]);
    } else {
      region_slow<SRC>(&g, c.c0, c.c1, c.col_base, ra, rb, wcand, wnc, SINGLE, lane);
    }
    __syncwarp();
  }
}

// dynamic shared memory: double px[R], py[R]; int out[R]; int cand[R]; int wcand[VOR_WARPS][32];
// then (16-byte aligned) per warp the lanes' owner accumulators and scratch, double [RUNS_LANE_DOUBLES][32 lanes].
//
// SINGLE = false (global): grid = (tiles_x * tiles_y, 1, E); a CTA owns a (2 x 64)-column x
//   (4 x rpw)-row tile of the map, one region per warp; every robot is a site and every site's sums
//   are produced.  SRC = SRC_WORLD or SRC_EXPLORED.
// SINGLE = true (window):  grid = (1, R, E); a CTA owns one robot's whole clipped L-window and its
//   warps loop over the window's regions; sites = the robot + its neighbours within communication
//   range (decentralized_cvt.h:100-120), only the robot's own cell is reduced.  SRC = SRC_ROBOT.
__host__ __device__ inline size_t vor_smem_bytes(int R) {
  return (((size_t)R * 24 + VOR_WARPS * 32 * sizeof(int) + 15) & ~(size_t)15) + (size_t)VOR_WARPS * RUNS_LANE_DOUBLES * 32 * 8;
}

template <int SRC, bool SINGLE, bool LABELS>
__global__ void __launch_bounds__(VOR_THREADS, 3) k_voronoi(VorArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const EnvDev &d = a.d;
  const int ns_cap = d.R;
  double *s_px = reinterpret_cast<double *>(smem_raw);
  double *s_py = s_px + ns_cap;
  int *s_out = reinterpret_cast<int *>(s_py + ns_cap);
  int *s_cand = s_out + ns_cap;
  int *s_wcand = s_cand + ns_cap;
  const size_t acc_off = ((size_t)ns_cap * 24 + VOR_WARPS * 32 * sizeof(int) + 15) & ~(size_t)15;
  __shared__ VorCta s_cta;
  __shared__ int s_ns, s_nc;
  __shared__ double s_best[VOR_WARPS];
  __shared__ int s_besti[VOR_WARPS];
  __shared__ int s_woff[VOR_WARPS + 1];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int e = blockIdx.z;
  const int robot = SINGLE ? blockIdx.y : 0;
  const double res = d.res;
  const double *pos = d.pos_g + (size_t)e * d.R * 2;

  // ---- CTA rectangle: rows [tr0, tr1), columns [tc0, tc1), column tiles start at col_base0 -----
  int tr0, tr1, tc0, tc1, col_base0;
  if (SINGLE) {
    // decentralized_cvt.h:89-95: clipped L-window around the robot
    Window w = make_window(res, pos[2 * robot], pos[2 * robot + 1], d.L, d.N);
    tr0 = w.x0, tr1 = w.x1, tc0 = w.y0, tc1 = w.y1;
    col_base0 = (tc0 / VOR_TW) * VOR_TW;
  } else {
    const int th = VOR_WY * a.rpw, tw = VOR_WX * VOR_TW;
    const int tx = blockIdx.x / a.tiles_y, ty = blockIdx.x - tx * a.tiles_y;
    tr0 = tx * th, tr1 = min(tr0 + th, d.N);
    col_base0 = ty * tw;
    tc0 = col_base0, tc1 = min(col_base0 + tw, d.N);
  }
  if (tr0 >= tr1 || tc0 >= tc1) return;

  // ---- site list ---------------------------------------------------------------------------
  if (!SINGLE) {
    for (int k = tid; k < d.R; k += VOR_THREADS) {
      s_px[k] = pos[2 * k], s_py[k] = pos[2 * k + 1], s_out[k] = k;
    }
    if (tid == 0) s_ns = d.R;
  } else if (warp == 0) {
    const double sx = pos[2 * robot], sy = pos[2 * robot + 1];
    if (lane == 0) s_px[0] = sx, s_py[0] = sy, s_out[0] = robot;
    int n = 1;
    for (int base = 0; base < d.R; base += 32) {
      int j = base + lane;
      bool keep = false;
      double jx = 0, jy = 0;
      if (j < d.R && j != robot) {
        jx = pos[2 * j], jy = pos[2 * j + 1];
        double dx = __dsub_rn(jx, sx), dy = __dsub_rn(jy, sy);
        keep = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy))) < d.comm_range;
      }
      unsigned m = __ballot_sync(0xffffffffu, keep);
      if (keep) {
        int slot = n + __popc(m & ((1u << lane) - 1u));
        s_px[slot] = jx, s_py[slot] = jy, s_out[slot] = j;
      }
      n += __popc(m);
    }
    if (lane == 0) s_ns = n;
  }
  __syncthreads();
  const int ns = s_ns;

  // ---- CTA-level cull (only when there are more sites than lanes) ----------------------------
  int nc = ns;
  if (ns > 32) {
    const double cx = cell_centre((tr0 + tr1 - 1) >> 1, res), cy = cell_centre((tc0 + tc1 - 1) >> 1, res);
    double best = INFINITY;
    int bi = 0x7fffffff;
    for (int k = tid; k < ns; k += VOR_THREADS) {
      double f = d2_exact(cx, cy, s_px[k], s_py[k]);
      if (f < best) best = f, bi = k;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      double ob = __shfl_xor_sync(0xffffffffu, best, o);
      int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob < best || (ob == best && oi < bi)) best = ob, bi = oi;
    }
    if (lane == 0) s_best[warp] = best, s_besti[warp] = bi;
    __syncthreads();
    best = s_best[0], bi = s_besti[0];
    for (int w = 1; w < VOR_WARPS; ++w)
      if (s_best[w] < best || (s_best[w] == best && s_besti[w] < bi)) best = s_best[w], bi = s_besti[w];
    const double jx = s_px[bi], jy = s_py[bi];
    const double ccx[2] = {cell_centre(tr0, res), cell_centre(tr1 - 1, res)};
    const double ccy[2] = {cell_centre(tc0, res), cell_centre(tc1 - 1, res)};
    int running = 0;
    for (int base = 0; base < ns; base += VOR_THREADS) {
      int k = base + tid;
      bool keep = false;
      if (k < ns) {
        double kx = s_px[k], ky = s_py[k];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          double fk = d2_exact(ccx[q & 1], ccy[q >> 1], kx, ky);
          double fj = d2_exact(ccx[q & 1], ccy[q >> 1], jx, jy);
          if (fk - fj <= 1e-7 * (1.0 + fj)) keep = true;
        }
      }
      unsigned m = __ballot_sync(0xffffffffu, keep);
      if (lane == 0) s_woff[warp] = __popc(m);
      __syncthreads();
      int off = running;
      for (int w = 0; w < warp; ++w) off += s_woff[w];
      if (keep) s_cand[off + __popc(m & ((1u << lane) - 1u))] = k;
      int tot = 0;
      for (int w = 0; w < VOR_WARPS; ++w) tot += s_woff[w];
      running += tot;
      __syncthreads();
    }
    if (tid == 0) s_nc = running;
    __syncthreads();
    nc = s_nc;
  } else {
    if (tid < ns) s_cand[tid] = tid;
    __syncthreads();
  }
  if (SINGLE && s_cand[0] != 0) return;  // the robot itself was culled: it owns no cell here

  // ---- regions: (row chunk, 64-column tile) pairs dealt round-robin to the warps ---------------
  if (tid == 0) {
    const size_t nn = (size_t)d.N * d.N, pw = (size_t)d.N * d.NW;
    VorCta &g = s_cta;
    g.N = d.N, g.R = d.R, g.NW = d.NW, g.res = d.res, g.fill = d.fill, g.use_sensor = d.use_sensor;
    g.world = d.world + (size_t)e * nn;
    g.mask = nullptr;
    if (SRC == SRC_ROBOT) g.mask = d.seenb + ((size_t)e * d.R + robot) * pw;
    if (SRC == SRC_EXPLORED) g.mask = d.sysb + ((size_t)e * 2) * pw;
    g.e = e, g.robot = robot;
    g.s_px = s_px, g.s_py = s_py, g.s_out = s_out, g.sums = a.sums;
    g.labels = (LABELS && a.labels && !SINGLE) ? a.labels + (size_t)e * nn : nullptr;
  }
  __syncthreads();
  const int rpw = a.rpw;
  const int nrow = (tr1 - tr0 + rpw - 1) / rpw;
  const int ncol = (tc1 - 1 - col_base0) / VOR_TW + 1;
  for (int sidx = warp; sidx < nrow * ncol; sidx += VOR_WARPS) {
    const int ri = sidx / ncol, ci = sidx - ri * ncol;
    StripCtx c;
    c.r0 = tr0 + ri * rpw;
    c.r1 = min(c.r0 + rpw, tr1);
    c.col_base = col_base0 + ci * VOR_TW;
    c.c0 = max(c.col_base, tc0);
    c.c1 = min(c.col_base + VOR_TW, tc1);
    if (c.c0 >= c.c1) continue;
    process_region<SRC, SINGLE, LABELS>(s_cta, c, s_cand, nc, s_wcand + warp * 32,
                                        reinterpret_cast<double *>(smem_raw + acc_off) + warp * (RUNS_LANE_DOUBLES * 32), lane);
    __syncwarp();
  }
}

}  // namespace ccb
