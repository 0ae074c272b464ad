// voronoi.cuh -- discrete Voronoi assignment fused with the per-cell mass / moment reductions: the
// TILE-ORDERED, fully general kernel (any number of sites, any map size, masked sources, robot windows).
// The batched product path uses the site-ordered kernels of vglobal.cuh (whole map) and vwindow.cuh (robot
// windows) where they apply; this kernel covers the rest and large single worlds, where it is the faster
// one.  It also provides the shared helpers (cell_centre, d2_exact, reduce4, emit).
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
constexpr int VOR_MAXFAST = 3;   // regions with more competing sites are split, then take the generic path
constexpr int VOR_MINSPLIT = 4;  // do not split below this many rows
constexpr int VOR_PAD_ROWS = 8;  // rows the streaming loop may read past a region (see stream_region)

enum { SRC_WORLD = 0, SRC_EXPLORED = 1, SRC_ROBOT = 2 };

struct VorArgs {
  EnvDev d;
  int rpw;      // rows per warp region (even, <= VOR_MAXRPW)
  int tiles_x, tiles_y;
  int row0, row1;  // global mode: only rows [row0, row1) are labelled and reduced (row band of a multi-GPU world)
  double *sums; // [E][R][4]
  int *labels;  // optional [E][N*N] (global mode only): the label of every cell, for parity tests
};

// (the host bodies serve the CPU self-test of the partition sweep, ccb_selftest_vg_partition: x86-64 without
// -mfma rounds every operation separately, like the __d*_rn intrinsics)
__host__ __device__ __forceinline__ double cell_centre(int i, double res) {
#ifdef __CUDA_ARCH__
  return __dmul_rn(__dadd_rn((double)i, 0.5), res);
#else
  volatile double c = (double)i + 0.5;
  return c * res;
#endif
}
__host__ __device__ __forceinline__ double d2_exact(double cx, double cy, double px, double py) {
#ifdef __CUDA_ARCH__
  double dx = __dsub_rn(cx, px), dy = __dsub_rn(cy, py);
  return __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
#else
  volatile double dx = cx - px, dy = cy - py;
  volatile double xx = dx * dx, yy = dy * dy;
  return xx + yy;
#endif
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
__device__ __forceinline__ double shfl_xor_d(double v, int o) { return __shfl_xor_sync(0xffffffffu, v, o); }
// Warp sum of four values per lane with 6 shuffles instead of 20: the first two butterfly steps halve
// the number of values a lane carries (lanes with the bit set keep the upper half), the remaining
// three steps reduce one value.  Lane l ends up with the total of value ((l >> 4) & 1) * 2 + ((l >> 3) & 1);
// the result is returned to lane 0 .. through shared exchange of the four holders.
__device__ __forceinline__ Sum4 reduce4(double m, double sx, double sy, double s2) {
  const int lane = threadIdx.x & 31;
  // step 1 (xor 16): lanes 0-15 keep (m, sx), lanes 16-31 keep (sy, s2)
  const bool up1 = lane & 16;
  double a0 = up1 ? sy : m, a1 = up1 ? s2 : sx;      // kept
  double b0 = up1 ? m : sy, b1 = up1 ? sx : s2;      // sent
  a0 += shfl_xor_d(b0, 16), a1 += shfl_xor_d(b1, 16);
  // step 2 (xor 8): lanes with bit 3 clear keep a0, the others a1
  const bool up2 = lane & 8;
  double c0 = up2 ? a1 : a0, d0 = up2 ? a0 : a1;
  c0 += shfl_xor_d(d0, 8);
  c0 += shfl_xor_d(c0, 4);
  c0 += shfl_xor_d(c0, 2);
  c0 += shfl_xor_d(c0, 1);
  // holders: lane 0 -> m, lane 8 -> sx, lane 16 -> sy, lane 24 -> s2
  Sum4 r;
  r.m = __shfl_sync(0xffffffffu, c0, 0);
  r.sx = __shfl_sync(0xffffffffu, c0, 8);
  r.sy = __shfl_sync(0xffffffffu, c0, 16);
  r.s2 = __shfl_sync(0xffffffffu, c0, 24);
  return r;
}
__device__ __forceinline__ Sum4 fold_and_reduce(const Acc6 &a, int col) {
  // fold the per-column sums with this lane's column coordinates (cell units), then warp-reduce
  const double y0 = (double)col + 0.5, y1 = y0 + 1.0, y2 = y0 + 2.0, y3 = y0 + 3.0;
  const double m = (a.c0 + a.c1) + (a.c2 + a.c3);
  const double sy = fma(y0, a.c0, fma(y1, a.c1, fma(y2, a.c2, y3 * a.c3)));
  const double s2 = a.sxx + fma(y0 * y0, a.c0, fma(y1 * y1, a.c1, fma(y2 * y2, a.c2, (y3 * y3) * a.c3)));
  return reduce4(m, a.sx, sy, s2);
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

// the stated rule between two candidates k0 < k1 (site indices) at one cell: 1 if k1 is strictly nearer
__device__ __noinline__ int exact_label2(const double *s_px, const double *s_py, int k0, int k1, int row, int col,
                                         double res) {
  const double cx = cell_centre(row, res), cy = cell_centre(col, res);
  return d2_exact(cx, cy, s_px[k1], s_py[k1]) < d2_exact(cx, cy, s_px[k0], s_py[k0]) ? 1 : 0;
}

// Two owners, careful version (cold): returns T in [0, rows], packed with two flags: bit 30 =
// candidate 1 owns the rows u < T (else the rows u >= T); bit 29 = every cell of the column needs
// the exact rule.  h(u) = hc + u HA is d2_1 - d2_0 at row r0 + u of column `col`.
__device__ __noinline__ int column_threshold(const double *s_px, const double *s_py, int k0, int k1, int r0, int col,
                                             int rows, double hc, double HA, double eps, double res) {
  if (fabs(HA) * rows <= eps) {
    // the bisector runs (numerically) along this column: one label for all rows, unless it is
    // within rounding of the cell centres, in which case every cell takes the exact rule
    int T = (hc < 0.0) ? rows : 0;
    return T | (1 << 30) | ((fabs(hc) <= eps) ? (1 << 29) : 0);
  }
  const double x = -hc / HA;       // h(x) = 0
  const bool first1 = HA > 0.0;    // h increasing: candidate 1 (h < 0) owns the rows below x
  double tf = first1 ? ceil(x) : floor(x) + 1.0;
  tf = fmin(fmax(tf, 0.0), (double)rows);
  int T = (int)tf;
  // re-check the two cells next to the threshold with the exact rule when they are within
  // rounding of the bisector (|h| <= eps there)
  if (T - 1 >= 0 && T - 1 < rows && fabs(hc + (T - 1) * HA) <= eps) {
    const int lab = exact_label2(s_px, s_py, k0, k1, r0 + T - 1, col, res);
    if ((lab == 1) != first1) T -= 1;
  }
  if (T >= 0 && T < rows && fabs(hc + T * HA) <= eps) {
    const int lab = exact_label2(s_px, s_py, k0, k1, r0 + T, col, res);
    if ((lab == 1) == first1) T += 1;
  }
  return T | (first1 ? (1 << 30) : 0);
}

// 1.0 where p, else 0.0: fma(v, one_if(p), acc) is a predicated fp64 add in one instruction (the
// C++ `if (p) acc += v` compiles to an add plus two 32-bit selects)
__device__ __forceinline__ double one_if(bool p) { return __hiloint2double(p ? 0x3ff00000 : 0, 0); }

// Two or three owners k0 < k1 (< k2), branch-free: set 0 takes every cell, set i the cells of owner i
// through 0/1 masks, owner 0 gets the difference at the end.  A region's boundary crosses a different
// column in almost every pass, so anything that branches per label change diverges on every pass.
//   NC == 2: along a column h = d2_1 - d2_0 is affine and changes sign once: an integer row threshold
//            per column (fp64, once per region; the cells next to it re-checked with the exact rule
//            when they are within rounding of the bisector), one integer compare per cell;
//   NC == 3: fp32 affine differences per cell, the exact rule when the best two are closer than a
//            rigorous bound of the fp32 error.
template <int NC, bool SINGLE, bool LABELS>
struct ConsMask {
  static constexpr int NSETS = SINGLE ? 1 : NC;
  Acc6 acc[NSETS];
  int T0, T1, T2, T3;   // NC == 2: cell (u, j) belongs to candidate 1 iff (first1 ? u : -u) < T_j
  bool first1;
  float h0a, haa, hba, h0b, hab, hbb, margin;  // NC == 3
  int k0, k1, k2;

  // returns false when whole columns need the exact rule (a bisector along a column through the cell
  // centres): the caller takes the generic path
  __device__ __forceinline__ bool init(const VorCta &g, int ra, int rows, int col, int k0_, int k1_, int k2_) {
    k0 = k0_, k1 = k1_, k2 = k2_;
#pragma unroll
    for (int s = 0; s < NSETS; ++s) acc[s].zero();
    T0 = T1 = T2 = T3 = 0, first1 = true;
    h0a = haa = hba = h0b = hab = hbb = margin = 0.f;
    const double res = g.res;
    const double cx0 = cell_centre(ra, res), cy0 = cell_centre(col, res);
    const double p0x = g.s_px[k0], p0y = g.s_py[k0];
    const double f0 = d2_exact(cx0, cy0, p0x, p0y);
    const double p1x = g.s_px[k1], p1y = g.s_py[k1];
    const double H0 = d2_exact(cx0, cy0, p1x, p1y) - f0;
    const double HA = 2.0 * res * (p0x - p1x), HB = 2.0 * res * (p0y - p1y);
    double hmax = fabs(H0) + rows * fabs(HA) + 4.0 * fabs(HB);
    unsigned exact_cols = 0;
    if (NC == 2) {
      const double eps = 1e-9 * (1.0 + f0 + hmax);  // >> the fp64 rounding of d2 and of the affine form
      int tt[4] = {0, 0, 0, 0};
      bool careful = fabs(HA) * rows <= eps;
      if (!careful) {
        // h(u) = hc + u HA changes sign at x = -hc / HA: one reciprocal per lane, one multiply per
        // column; a threshold whose neighbouring cells are within eps of the bisector is redone by
        // the careful version (which applies the exact rule there)
        first1 = HA > 0.0;
        const double ninv = -1.0 / HA;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const double hc = fma((double)j, HB, H0);
          const double x = hc * ninv;
          double tf = first1 ? ceil(x) : floor(x) + 1.0;
          tf = fmin(fmax(tf, 0.0), (double)rows);
          tt[j] = (int)tf;
          const double hl = fma(tf - 1.0, HA, hc), hr = fma(tf, HA, hc);
          careful |= (fabs(hl) <= eps) | (fabs(hr) <= eps);
        }
      }
      if (careful) {  // rare, per lane
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
          const int t = column_threshold(g.s_px, g.s_py, k0, k1, ra, col + j, rows, fma((double)j, HB, H0), HA, eps, res);
          first1 = (t >> 30) & 1;  // depends on the sign of HA only (or is 1 for every flat column)
          tt[j] = t & 0xffff;
          exact_cols |= ((t >> 29) & 1u) << j;
        }
      }
      // pass() tests (first1 ? u : -u) < Ts_j: u < T_j, resp. u >= T_j <=> -u < 1 - T_j
      if (!first1) tt[0] = 1 - tt[0], tt[1] = 1 - tt[1], tt[2] = 1 - tt[2], tt[3] = 1 - tt[3];
      T0 = tt[0], T1 = tt[1], T2 = tt[2], T3 = tt[3];
    } else {
      const double p2x = g.s_px[k2], p2y = g.s_py[k2];
      const double G0 = d2_exact(cx0, cy0, p2x, p2y) - f0;
      const double GA = 2.0 * res * (p0x - p2x), GB = 2.0 * res * (p0y - p2y);
      hmax = fmax(hmax, fabs(G0) + rows * fabs(GA) + 4.0 * fabs(GB));
      h0a = (float)H0, haa = (float)HA, hba = (float)HB;
      h0b = (float)G0, hab = (float)GA, hbb = (float)GB;
      // fp32 evaluation error of h_k is < 2^-21 hmax; differences of two h's double it
      margin = (float)(hmax * (1.0 / 262144.0));
    }
    return !__any_sync(0xffffffffu, exact_cols != 0u);
  }

  __device__ __forceinline__ void pass(const VorCta &g, float v0, float v1, float v2, float v3, int u, double xd, int ra,
                                       int rb, int col, unsigned in4) {
    // q1_j / q2_j: cell j belongs to candidate 1 / 2
    bool q10 = false, q11 = false, q12 = false, q13 = false, q20 = false, q21 = false, q22 = false, q23 = false;
    if (NC == 2) {
      const int us = first1 ? u : -u;
      q10 = us < T0, q11 = us < T1, q12 = us < T2, q13 = us < T3;
    } else {
      const float uf = (float)u;
      const float hr0 = fmaf(uf, haa, h0a), hr1 = fmaf(uf, hab, h0b);
      bool amb = false;
      unsigned labs = 0;  // one label per byte
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float g1 = fmaf((float)i, hba, hr0), g2 = fmaf((float)i, hbb, hr1);
        float best = 0.f, second;  // winner among {0: 0, 1: g1, 2: g2}
        unsigned bl = 0;
        if (g1 < best) second = best, best = g1, bl = 1u << (8 * i);
        else second = g1;
        if (g2 < best) second = best, best = g2, bl = 2u << (8 * i);
        else second = fminf(second, g2);
        amb |= (second - best < margin);
        labs |= bl;
      }
      if (amb) {  // rare: a cell of this lane's four is within the fp32 error of a bisector -> exact rule
        const double res = g.res;
        const double cxr = cell_centre(ra + u, res);
        const double ax = g.s_px[k0], ay = g.s_py[k0], bx = g.s_px[k1], by = g.s_py[k1], ex = g.s_px[k2], ey = g.s_py[k2];
        labs = 0;
#pragma unroll 1
        for (int i = 0; i < 4; ++i) {
          const double cy = cell_centre(col + i, res);
          unsigned best = 0;
          double bd = d2_exact(cxr, cy, ax, ay);
          double dd = d2_exact(cxr, cy, bx, by);
          if (dd < bd) bd = dd, best = 1;
          dd = d2_exact(cxr, cy, ex, ey);
          if (dd < bd) bd = dd, best = 2;
          labs |= best << (8 * i);
        }
      }
      q10 = (labs & 0xffu) == 1u, q11 = ((labs >> 8) & 0xffu) == 1u, q12 = ((labs >> 16) & 0xffu) == 1u, q13 = (labs >> 24) == 1u;
      q20 = (labs & 0xffu) == 2u, q21 = ((labs >> 8) & 0xffu) == 2u, q22 = ((labs >> 16) & 0xffu) == 2u, q23 = (labs >> 24) == 2u;
    }
    if (LABELS && !SINGLE) {
      if (g.labels && ra + u < rb) {
        int *lp_out = g.labels + (size_t)(ra + u) * g.N + col;
        if (in4 & 1u) lp_out[0] = g.s_out[q10 ? k1 : (q20 ? k2 : k0)];
        if (in4 & 2u) lp_out[1] = g.s_out[q11 ? k1 : (q21 ? k2 : k0)];
        if (in4 & 4u) lp_out[2] = g.s_out[q12 ? k1 : (q22 ? k2 : k0)];
        if (in4 & 8u) lp_out[3] = g.s_out[q13 ? k1 : (q23 ? k2 : k0)];
      }
    }
    if (SINGLE) {  // only the robot's own cells (candidate 0) count
      v0 = (q10 || q20) ? 0.f : v0, v1 = (q11 || q21) ? 0.f : v1;
      v2 = (q12 || q22) ? 0.f : v2, v3 = (q13 || q23) ? 0.f : v3;
    }
    const double dv0 = (double)v0, dv1 = (double)v1, dv2 = (double)v2, dv3 = (double)v3;
    const double xd2 = xd * xd;
    {
      Acc6 &t0 = acc[0];
      t0.c0 += dv0, t0.c1 += dv1, t0.c2 += dv2, t0.c3 += dv3;
      const double s = (dv0 + dv1) + (dv2 + dv3);
      t0.sx = fma(xd, s, t0.sx);
      t0.sxx = fma(xd2, s, t0.sxx);
    }
    if constexpr (!SINGLE) {
      Acc6 &a = acc[1];
      const double b0 = one_if(q10), b1 = one_if(q11), b2 = one_if(q12), b3 = one_if(q13);
      a.c0 = fma(dv0, b0, a.c0), a.c1 = fma(dv1, b1, a.c1), a.c2 = fma(dv2, b2, a.c2), a.c3 = fma(dv3, b3, a.c3);
      const double s_k = fma(dv0, b0, dv1 * b1) + fma(dv2, b2, dv3 * b3);
      a.sx = fma(xd, s_k, a.sx);
      a.sxx = fma(xd2, s_k, a.sxx);
    }
    if constexpr (!SINGLE && NC >= 3) {
      Acc6 &a = acc[2];
      const double b0 = one_if(q20), b1 = one_if(q21), b2 = one_if(q22), b3 = one_if(q23);
      a.c0 = fma(dv0, b0, a.c0), a.c1 = fma(dv1, b1, a.c1), a.c2 = fma(dv2, b2, a.c2), a.c3 = fma(dv3, b3, a.c3);
      const double s_k = fma(dv0, b0, dv1 * b1) + fma(dv2, b2, dv3 * b3);
      a.sx = fma(xd, s_k, a.sx);
      a.sxx = fma(xd2, s_k, a.sxx);
    }
  }
};

// ---- the streaming skeleton ------------------------------------------------------------------------
// Rows [ra, rb) of region c (ra - c.r0 even) go through cons.pass() once, a lane holding four
// consecutive cells of one row per pass (lanes 16-31: the next row).
//   plain (unmasked source, every lane inside the map and the region, even row count, 16-byte rows):
//     vector loads through a register ring that keeps U passes (2 KB per warp) in flight;
//   otherwise: per pass the visibility nibble comes from the region's mask words (one shuffle), row
//     pairs with nothing visible are skipped, cells are selected against the mask.
template <int SRC, int U, class Cons>
__device__ __forceinline__ void stream_region(const VorCta &g, const StripCtx &c, int ra, int rb, uint32_t m0, uint32_t m1,
                                              uint32_t m2, uint32_t m3, Cons &cons) {
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
    // visibility nibble of pass t for this lane (0 past the last pass) and, where something is visible,
    // the request for its world values.  t is warp-uniform, so the shuffle is executed by every lane.
    auto fetch = [&](int t, float4 &w, unsigned &vis) {
      vis = 0u;
      w = make_float4(0.f, 0.f, 0.f, 0.f);
      if (t >= npass) return;
      const int u = 2 * t + hp;
      vis = (ra + u < rb) ? in4 : 0u;
      if (SRC != SRC_WORLD) {
        const int um = um0 + 2 * t;  // + hp stays inside the same 16-row word group
        const int q = um >> 4;
        const uint32_t sel = q == 0 ? m0 : (q == 1 ? m1 : (q == 2 ? m2 : m3));
        const uint32_t word = __shfl_sync(0xffffffffu, sel, (((um + hp) & 15) << 1) | (c16 >> 3));
        vis &= (word >> (4 * (c16 & 7))) & 0xfu;
      }
      if (vis != 0u && use_sensor) w = load4(wp + (size_t)t * pstride, vec, inmap);
    };
    // a ring of UM passes in flight: the loads of pass t + UM are issued before pass t is consumed
    constexpr int UM = 4;
    float4 ring[UM];
    unsigned visr[UM];
#pragma unroll
    for (int k = 0; k < UM; ++k) fetch(k, ring[k], visr[k]);
#pragma unroll 1
    for (int t0 = 0; t0 < npass; t0 += UM) {
#pragma unroll
      for (int k = 0; k < UM; ++k) {
        const int t = t0 + k;
        if (t < npass) {  // warp-uniform
          const int u = 2 * t + hp;
          const unsigned inside = (ra + u < rb) ? in4 : 0u;
          const unsigned vis = visr[k];
          const float4 w = ring[k];
          // row pairs with nothing visible contribute nothing (unless unseen cells carry a fill value)
          const bool live = (SRC == SRC_WORLD) || __any_sync(0xffffffffu, (fill != 0.f ? inside : vis) != 0u);
          if (live) {
            float f0 = 0.f, f1 = 0.f, f2 = 0.f, f3 = 0.f;
            if (SRC == SRC_ROBOT && fill != 0.f) {
              f0 = (inside & 1u) ? fill : 0.f, f1 = (inside & 2u) ? fill : 0.f;
              f2 = (inside & 4u) ? fill : 0.f, f3 = (inside & 8u) ? fill : 0.f;
            }
            const float v0 = (vis & 1u) ? w.x : f0, v1 = (vis & 2u) ? w.y : f1;
            const float v2 = (vis & 4u) ? w.z : f2, v3 = (vis & 8u) ? w.w : f3;
            cons.pass(g, v0, v1, v2, v3, u, xd + (double)(2 * t), ra, rb, col, in4);
          }
        }
        fetch(t + UM, ring[k], visr[k]);
      }
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
  stream_region<SRC, 4>(g, c, ra, rb, m0, m1, m2, m3, cons);
  const Sum4 tot = fold_and_reduce(cons.a, c.col_base + 4 * (lane & 15));
  if (lane == 0) emit(g.sums, g.e, g.R, SINGLE ? g.robot : g.s_out[k0], tot.m, tot.sx, tot.sy, tot.s2);
}

// Two or three owners.  Returns false (nothing accumulated) when the caller has to take the generic path.
template <int NC, int SRC, bool SINGLE, bool LABELS>
__device__ __noinline__ bool region_mask(const VorCta *gp, StripCtx c, int ra, int rb, int k0, int k1, int k2, uint32_t m0,
                                         uint32_t m1, uint32_t m2, uint32_t m3) {
  const VorCta &g = *gp;
  const int lane = threadIdx.x & 31;
  const int col = c.col_base + 4 * (lane & 15);
  ConsMask<NC, SINGLE, LABELS> cons;
  if (!cons.init(g, ra, rb - ra, col, k0, k1, k2)) return false;
  stream_region<SRC, (NC == 2 ? 4 : 2)>(g, c, ra, rb, m0, m1, m2, m3, cons);
  Sum4 tot = fold_and_reduce(cons.acc[0], col);
  if (SINGLE) {
    if (lane == 0) emit(g.sums, g.e, g.R, g.robot, tot.m, tot.sx, tot.sy, tot.s2);
    return true;
  }
  const int ksel[3] = {k0, k1, k2};
#pragma unroll
  for (int k = 1; k < ConsMask<NC, SINGLE, LABELS>::NSETS; ++k) {
    const Sum4 part = fold_and_reduce(cons.acc[k], col);
    if (lane == 0) emit(g.sums, g.e, g.R, g.s_out[ksel[k]], part.m, part.sx, part.sy, part.s2);
    tot.m -= part.m, tot.sx -= part.sx, tot.sy -= part.sy, tot.s2 -= part.s2;
  }
  if (lane == 0) emit(g.sums, g.e, g.R, g.s_out[k0], tot.m, tot.sx, tot.sy, tot.s2);
  return true;
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
  // reference site j*: the winner at the centre cell, found on fp32 keys (any site is a valid
  // reference for the domination test below; a near-tie at the centre only changes how much is culled)
  float fcen = 3.0e38f;
  if (has) fcen = (float)d2_exact(cell_centre((r0 + r1 - 1) >> 1, res), cell_centre((c0 + c1 - 1) >> 1, res), kx, ky);
  const unsigned kmin = __reduce_min_sync(0xffffffffu, __float_as_uint(fcen));  // non-negative floats order as integers
  const int bl = __ffs(__ballot_sync(0xffffffffu, __float_as_uint(fcen) == kmin)) - 1;
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
  if (__popc(mask) > 2) {  // (with two survivors the other one is j* itself: nothing to gain)
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
                                               int lane) {
  RegionMask rm;
  rm.w[0] = rm.w[1] = rm.w[2] = rm.w[3] = 0u;
  if (SRC == SRC_WORLD) {
    // Ask for the whole region (<= 64 rows x 256 bytes = 128 lines, four per lane) to be brought into
    // L2 now: the requests are all in flight while the warp culls and plans, and the streaming loop's
    // loads then see L2 latency instead of HBM latency.
    const int pcol = c.col_base + 32 * (lane & 1);
    if (pcol < g.N) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int row = c.r0 + 16 * q + (lane >> 1);
        if (row < c.r1) asm volatile("prefetch.global.L2 [%0];" ::"l"(g.world + (size_t)row * g.N + pcol));
      }
    }
  } else {
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
    bool done = false;
    if (wnc == 1) {
      region_one<SRC, SINGLE, LABELS>(&g, c, ra, rb, wcand[0], rm.w[0], rm.w[1], rm.w[2], rm.w[3]);
      done = true;
    } else if (wnc == 2) {
      done = region_mask<2, SRC, SINGLE, LABELS>(&g, c, ra, rb, wcand[0], wcand[1], wcand[1], rm.w[0], rm.w[1], rm.w[2], rm.w[3]);
    } else if (wnc == 3) {
      done = region_mask<3, SRC, SINGLE, LABELS>(&g, c, ra, rb, wcand[0], wcand[1], wcand[2], rm.w[0], rm.w[1], rm.w[2], rm.w[3]);
    }
    if (!done) region_slow<SRC>(&g, c.c0, c.c1, c.col_base, ra, rb, wcand, wnc, SINGLE, lane);
    __syncwarp();
  }
}

// dynamic shared memory: double px[R], py[R]; int out[R]; int cand[R]; int wcand[VOR_WARPS][32].
//
// SINGLE = false (global): grid = (tiles_x * tiles_y, 1, E); a CTA owns a (2 x 64)-column x
//   (4 x rpw)-row tile of the map, one region per warp; every robot is a site and every site's sums
//   are produced.  SRC = SRC_WORLD or SRC_EXPLORED.
// SINGLE = true (window):  grid = (1, R, E); a CTA owns one robot's whole clipped L-window and its
//   warps loop over the window's regions; sites = the robot + its neighbours within communication
//   range (decentralized_cvt.h:100-120), only the robot's own cell is reduced.  SRC = SRC_ROBOT.
__host__ __device__ inline size_t vor_smem_bytes(int R) {
  return ((size_t)R * 24 + VOR_WARPS * 32 * sizeof(int) + 15) & ~(size_t)15;
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
    tr0 = a.row0 + tx * th, tr1 = min(tr0 + th, a.row1);
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
    process_region<SRC, SINGLE, LABELS>(s_cta, c, s_cand, nc, s_wcand + warp * 32, lane);
    __syncwarp();
  }
}

}  // namespace ccb
