// voronoi.cuh -- discrete Voronoi assignment fused with the per-cell mass / moment reductions.
//
// Replaces reference cppsrc/core/src/voronoi.cpp:55-327 (CGAL arrangement + scan-line
// rasterisation + fp64 accumulation, one OpenMP thread per site) and the per-robot windowed
// variant of algorithms/decentralized_cvt.h:85-143.
//
// Semantics: every cell centre ((i+.5)res, (j+.5)res) is labelled with its nearest site; the
// distance is d2 = dx*dx + dy*dy evaluated in fp64 without FMA contraction and ties go to the
// lowest site index (the build's stated rule).  The labels are never stored: each
// warp streams a 128-column x rpw-row strip of the map once and accumulates, per label,
//   S0 = sum v, Sx = sum (i+.5) v, Sy = sum (j+.5) v, S2 = sum ((i+.5)^2 + (j+.5)^2) v
// in fp64 (cell units; 24-bit v times a <= 15-bit half-integer is exact in fp64), from which mass,
// centroid and the objective sum |c - p|^2 v follow in the finalize kernel.
//
// How the exact labelling stays off the critical path:
//   * d2_k(c) - d2_j(c) is affine in c, so a site that loses to the strip-centre winner at all four
//     corner cells of a strip loses everywhere inside it: each warp culls the site list to the
//     1-4 sites that can own a cell of its strip (one site per lane, fp64, once per strip);
//   * per cell the survivors are compared in fp32 strip-local coordinates; only when the best two
//     are closer than a rigorous fp32 error bound is the fp64 rule evaluated (a few cells per map);
//   * accumulation is predicated per surviving label, no per-cell atomics and no label runs.
#pragma once

#include "common.cuh"

namespace ccb {

constexpr int VOR_THREADS = 256;
constexpr int VOR_WARPS = VOR_THREADS / 32;
constexpr int VOR_TY = 128;  // 32 lanes x 4 cells

enum { SRC_WORLD = 0, SRC_EXPLORED = 1, SRC_ROBOT = 2 };

struct VorArgs {
  EnvDev d;
  int single;   // 0: all robots are sites, sums for every site. 1: per-robot L-window, own cell only
  int src;      // SRC_*
  int rpw;      // rows per warp
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

struct StripCtx {
  const EnvDev *d;
  int e, robot;        // robot: mask bit for SRC_ROBOT / output slot in single mode
  int src;
  int r0, r1;          // strip rows [r0, r1)
  int c0, c1;          // valid columns [c0, c1)
  int col_base;        // absolute column of lane 0's first cell (multiple of 128)
  const double *s_px, *s_py;  // site coordinates (shared memory)
  const int *s_out;           // site -> output slot
  double *sums;
  int *labels;                // optional label plane of this environment
};

// ---- staging: the whole strip goes global -> shared with cp.async ---------------------------------
// Every lane copies exactly the 16 bytes per row it later consumes, so a per-thread
// cp.async.wait_group is the only synchronisation.  All rows of the strip are in flight at once:
// memory-level parallelism comes from shared memory, not from registers.
constexpr int VOR_MAXROWS = 16;  // rows per strip kept in shared memory (rpw <= VOR_MAXROWS)

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, bool valid) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  const int sz = valid ? 16 : 0;  // src-size 0: the 16 destination bytes are zero-filled, no global read
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(sa), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async4(void *smem, const void *gmem, bool valid) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  const int sz = valid ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(sa), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

// Stage the strip.  wbuf: [rows][32] float4 (this warp's slice), mbuf: same size, raw masks.
// Returns per-row 4-bit visibility of this lane's cells packed into a 64-bit word (row u -> bits
// 4u..4u+3): inside the region AND visible in the selected map.  Cells that are not visible are
// never read from the world map.
__device__ __forceinline__ unsigned long long stage_strip(const StripCtx &c, int col, int lane, float4 *wbuf,
                                                          uint4 *mbuf) {
  const EnvDev &d = *c.d;
  const size_t nn = (size_t)d.N * d.N;
  const int rows = c.r1 - c.r0;
  const bool aligned = (d.N & 3) == 0;
  unsigned in4 = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (col + i >= c.c0 && col + i < c.c1) in4 |= 1u << i;
  unsigned long long vis = 0;
  const float *world = d.world + (size_t)c.e * nn;
  if (c.src == SRC_WORLD) {
    for (int u = 0; u < rows; ++u) vis |= (unsigned long long)in4 << (4 * u);
  } else if (in4) {
    // phase A: masks
    if (c.src == SRC_EXPLORED) {
      const uint8_t *sys = d.sys + (size_t)c.e * nn;
      if (aligned && in4 == 0xfu) {
        for (int u = 0; u < rows; ++u)
          cp_async4(reinterpret_cast<uint32_t *>(mbuf) + u * 32 + lane, sys + (size_t)(c.r0 + u) * d.N + col, true);
        cp_async_wait_all();
        for (int u = 0; u < rows; ++u) {
          const uint32_t q = reinterpret_cast<const uint32_t *>(mbuf)[u * 32 + lane];
          const unsigned m = (q & 1u) | ((q >> 7) & 2u) | ((q >> 14) & 4u) | ((q >> 21) & 8u);
          vis |= (unsigned long long)m << (4 * u);
        }
      } else {
        for (int u = 0; u < rows; ++u) {
          unsigned m = 0;
          for (int i = 0; i < 4; ++i)
            if (in4 & (1u << i)) m |= (unsigned)(__ldg(sys + (size_t)(c.r0 + u) * d.N + col + i) & 1u) << i;
          vis |= (unsigned long long)m << (4 * u);
        }
      }
    } else {
      const uint32_t *seen = d.seen + ((size_t)c.e * d.W + (c.robot >> 5)) * nn;
      const int sh = c.robot & 31;
      if (aligned && in4 == 0xfu) {
        for (int u = 0; u < rows; ++u) cp_async16(mbuf + u * 32 + lane, seen + (size_t)(c.r0 + u) * d.N + col, true);
        cp_async_wait_all();
        for (int u = 0; u < rows; ++u) {
          const uint4 q = mbuf[u * 32 + lane];
          const unsigned m = ((q.x >> sh) & 1u) | (((q.y >> sh) & 1u) << 1) | (((q.z >> sh) & 1u) << 2) |
                             (((q.w >> sh) & 1u) << 3);
          vis |= (unsigned long long)m << (4 * u);
        }
      } else {
        for (int u = 0; u < rows; ++u) {
          unsigned m = 0;
          for (int i = 0; i < 4; ++i)
            if (in4 & (1u << i)) m |= ((__ldg(seen + (size_t)(c.r0 + u) * d.N + col + i) >> sh) & 1u) << i;
          vis |= (unsigned long long)m << (4 * u);
        }
      }
    }
  }
  // phase B: world values, only where something is visible
  if (!d.use_sensor) {
    for (int u = 0; u < rows; ++u) wbuf[u * 32 + lane] = make_float4(0.f, 0.f, 0.f, 0.f);
  } else if (aligned && (in4 == 0xfu || in4 == 0)) {
    for (int u = 0; u < rows; ++u) {
      const bool need = ((vis >> (4 * u)) & 0xfull) != 0;
      cp_async16(wbuf + u * 32 + lane, need ? world + (size_t)(c.r0 + u) * d.N + col : world, need);
    }
    cp_async_wait_all();
  } else {
    for (int u = 0; u < rows; ++u) {
      float t[4] = {0.f, 0.f, 0.f, 0.f};
      for (int i = 0; i < 4; ++i)
        if ((vis >> (4 * u + i)) & 1ull) t[i] = __ldg(world + (size_t)(c.r0 + u) * d.N + col + i);
      wbuf[u * 32 + lane] = make_float4(t[0], t[1], t[2], t[3]);
    }
  }
  return vis;
}

__device__ __forceinline__ void reduce_and_emit(const StripCtx &c, const Acc6 &a, int col, int lane,
                                               double out[4]) {
  // fold the per-column sums with this lane's column coordinates (cell units)
  double y0 = (double)col + 0.5, y1 = y0 + 1.0, y2 = y0 + 2.0, y3 = y0 + 3.0;
  double m = (a.c0 + a.c1) + (a.c2 + a.c3);
  double sy = fma(y0, a.c0, fma(y1, a.c1, fma(y2, a.c2, y3 * a.c3)));
  double syy = fma(y0 * y0, a.c0, fma(y1 * y1, a.c1, fma(y2 * y2, a.c2, (y3 * y3) * a.c3)));
  double s2 = a.sxx + syy;
  double sx = a.sx;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    m += __shfl_xor_sync(0xffffffffu, m, o);
    sx += __shfl_xor_sync(0xffffffffu, sx, o);
    sy += __shfl_xor_sync(0xffffffffu, sy, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  out[0] = m, out[1] = sx, out[2] = sy, out[3] = s2;
  (void)lane;
  (void)c;
}

__device__ __forceinline__ void emit(double *sums, int e, int R, int slot, const double v[4]) {
  if (v[0] == 0.0 && v[1] == 0.0 && v[2] == 0.0 && v[3] == 0.0) return;
  double *dst = sums + ((size_t)e * R + slot) * 4;
  atomicAdd(dst + 0, v[0]);
  atomicAdd(dst + 1, v[1]);
  atomicAdd(dst + 2, v[2]);
  atomicAdd(dst + 3, v[3]);
}

// exact rule among `n` candidates (ascending site order): returns the index of the winner
template <int NC>
__device__ __noinline__ int exact_label(const StripCtx &c, const int (&site)[NC], int row, int col, double res) {
  double cx = cell_centre(row, res), cy = cell_centre(col, res);
  int best = 0;
  double bd = d2_exact(cx, cy, c.s_px[site[0]], c.s_py[site[0]]);
#pragma unroll
  for (int k = 1; k < NC; ++k) {
    double dd = d2_exact(cx, cy, c.s_px[site[k]], c.s_py[site[k]]);
    if (dd < bd) bd = dd, best = k;
  }
  return best;
}

// Fast path: NC (1..4) candidate sites can own cells of this strip.
// SINGLE: accumulate only cells owned by candidate 0 (the robot itself).
template <int NC, bool SINGLE>
__device__ void strip_fast(const StripCtx &c, const int (&site)[NC], float margin, int lane, float4 *wbuf,
                           uint4 *mbuf) {
  const EnvDev &d = *c.d;
  const double res = d.res;
  const float resf = (float)res;
  float qx[NC], qy[NC];
#pragma unroll
  for (int k = 0; k < NC; ++k) {
    qx[k] = (float)(c.s_px[site[k]] - (double)c.r0 * res);
    qy[k] = (float)(c.s_py[site[k]] - (double)c.col_base * res);
  }
  constexpr int NSETS = SINGLE ? 1 : NC;
  Acc6 acc[NSETS];
#pragma unroll
  for (int s = 0; s < NSETS; ++s) acc[s].zero();
  const int col = c.col_base + 4 * lane;
  float yl[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) yl[i] = resf * ((float)(4 * lane + i) + 0.5f);

  const unsigned long long vis = stage_strip(c, col, lane, wbuf, mbuf);
  const float fill = (c.src == SRC_ROBOT) ? d.fill : 0.f;
  {
    for (int row = c.r0; row < c.r1; ++row) {
      const int u = row - c.r0;
      const float4 w4 = wbuf[u * 32 + lane];
      const unsigned vb = (unsigned)(vis >> (4 * u)) & 0xfu;
      float v[4];
      v[0] = (vb & 1u) ? w4.x : 0.f, v[1] = (vb & 2u) ? w4.y : 0.f;
      v[2] = (vb & 4u) ? w4.z : 0.f, v[3] = (vb & 8u) ? w4.w : 0.f;
      if (fill != 0.f) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (!(vb & (1u << i)) && col + i >= c.c0 && col + i < c.c1) v[i] = fill;
      }
      int lab[4] = {0, 0, 0, 0};
      if (NC > 1) {
        const float xl = resf * ((float)(row - c.r0) + 0.5f);
        float dx2[NC];
  #pragma unroll
        for (int k = 0; k < NC; ++k) {
          float dx = xl - qx[k];
          dx2[k] = dx * dx;
        }
  #pragma unroll
        for (int i = 0; i < 4; ++i) {
          float best, second;
          int bl = 0;
          {
            float dy = yl[i] - qy[0];
            best = fmaf(dy, dy, dx2[0]);
            second = INFINITY;
          }
  #pragma unroll
          for (int k = 1; k < NC; ++k) {
            float dy = yl[i] - qy[k];
            float dd = fmaf(dy, dy, dx2[k]);
            if (dd < best) {
              second = best;
              best = dd;
              bl = k;
            } else {
              second = fminf(second, dd);
            }
          }
          if (second - best < margin) bl = exact_label<NC>(c, site, row, col + i, res);
          lab[i] = bl;
        }
      }
      if (!SINGLE && c.labels) {
  #pragma unroll
        for (int i = 0; i < 4; ++i)
          if (col + i >= c.c0 && col + i < c.c1) c.labels[(size_t)row * d.N + col + i] = c.s_out[site[lab[i]]];
      }
      const double dv0 = (double)v[0], dv1 = (double)v[1], dv2 = (double)v[2], dv3 = (double)v[3];
      const double xd = (double)row + 0.5;
      const double xd2 = xd * xd;
      if (!SINGLE) {
        // set 0: every cell of the strip
        Acc6 &t = acc[0];
        t.c0 += dv0, t.c1 += dv1, t.c2 += dv2, t.c3 += dv3;
        double s = (dv0 + dv1) + (dv2 + dv3);
        t.sx = fma(xd, s, t.sx);
        t.sxx = fma(xd2, s, t.sxx);
  #pragma unroll
        for (int k = 1; k < NC; ++k) {
          Acc6 &a = acc[k];
          double s_k = 0.0;
          if (lab[0] == k) a.c0 += dv0, s_k += dv0;
          if (lab[1] == k) a.c1 += dv1, s_k += dv1;
          if (lab[2] == k) a.c2 += dv2, s_k += dv2;
          if (lab[3] == k) a.c3 += dv3, s_k += dv3;
          a.sx = fma(xd, s_k, a.sx);
          a.sxx = fma(xd2, s_k, a.sxx);
        }
      } else {
        Acc6 &a = acc[0];
        double s_k = 0.0;
        if (lab[0] == 0) a.c0 += dv0, s_k += dv0;
        if (lab[1] == 0) a.c1 += dv1, s_k += dv1;
        if (lab[2] == 0) a.c2 += dv2, s_k += dv2;
        if (lab[3] == 0) a.c3 += dv3, s_k += dv3;
        a.sx = fma(xd, s_k, a.sx);
        a.sxx = fma(xd2, s_k, a.sxx);
      }
    }
  }

  double tot[4];
  reduce_and_emit(c, acc[0], col, lane, tot);
  if (SINGLE) {
    if (lane == 0) emit(c.sums, c.e, d.R, c.robot, tot);
    return;
  }
  double rest[4] = {tot[0], tot[1], tot[2], tot[3]};
#pragma unroll
  for (int k = 1; k < NC; ++k) {
    double part[4];
    reduce_and_emit(c, acc[k], col, lane, part);
    if (lane == 0) emit(c.sums, c.e, d.R, c.s_out[site[k]], part);
#pragma unroll
    for (int q = 0; q < 4; ++q) rest[q] -= part[q];
  }
  if (lane == 0) emit(c.sums, c.e, d.R, c.s_out[site[0]], rest);
}

// direct (unstaged) loader for the generic path
__device__ inline void load_cells_direct(const StripCtx &c, int row, int col, float v[4]) {
  const EnvDev &d = *c.d;
  const size_t nn = (size_t)d.N * d.N;
  const size_t off = (size_t)row * d.N + col;
  const float fill = (c.src == SRC_ROBOT) ? d.fill : 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[i] = 0.f;
    if (col + i < c.c0 || col + i >= c.c1) continue;
    bool visible = true;
    if (c.src == SRC_EXPLORED) visible = __ldg(d.sys + (size_t)c.e * nn + off + i) & 1;
    else if (c.src == SRC_ROBOT)
      visible = (__ldg(d.seen + ((size_t)c.e * d.W + (c.robot >> 5)) * nn + off + i) >> (c.robot & 31)) & 1u;
    v[i] = visible ? (d.use_sensor ? __ldg(d.world + (size_t)c.e * nn + off + i) : 0.f) : fill;
  }
}

// Generic path: any number of candidates, fp64 rule per cell, per-lane label runs flushed with
// global atomics.  Only strips that contain > 4 competing sites come here.
__device__ void strip_slow(const StripCtx &c, const int *cand, int nc, bool single, int lane) {
  const EnvDev &d = *c.d;
  const double res = d.res;
  const int col = c.col_base + 4 * lane;
  int cur = -1;
  double m = 0, sx = 0, sy = 0, s2 = 0;
  for (int row = c.r0; row < c.r1; ++row) {
    float v[4];
    load_cells_direct(c, row, col, v);
    const double cx = cell_centre(row, res);
    const double xd = (double)row + 0.5;
#pragma unroll 1
    for (int i = 0; i < 4; ++i) {
      if (v[i] == 0.f && !c.labels) continue;
      if (col + i < c.c0 || col + i >= c.c1) continue;
      const double cy = cell_centre(col + i, res);
      int best = 0;
      double bd = d2_exact(cx, cy, c.s_px[cand[0]], c.s_py[cand[0]]);
      for (int k = 1; k < nc; ++k) {
        double dd = d2_exact(cx, cy, c.s_px[cand[k]], c.s_py[cand[k]]);
        if (dd < bd) bd = dd, best = k;
      }
      if (single && best != 0) continue;
      const int slot = single ? c.robot : c.s_out[cand[best]];
      if (!single && c.labels) c.labels[(size_t)row * d.N + col + i] = slot;
      if (v[i] == 0.f) continue;
      if (slot != cur) {
        if (cur >= 0) {
          double o[4] = {m, sx, sy, s2};
          emit(c.sums, c.e, d.R, cur, o);
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
    double o[4] = {m, sx, sy, s2};
    emit(c.sums, c.e, d.R, cur, o);
  }
}

// One warp, one strip: cull the candidate list to the sites that can own a cell of the strip,
// then stream it.  cand[0..nc) are site indices (ascending) valid for the whole CTA rectangle.
__device__ void process_strip(StripCtx &c, const int *s_cand, int nc, bool single, int rpw, int *wcand, int lane,
                              float4 *wbuf, uint4 *mbuf) {
  const EnvDev &d = *c.d;
  const double res = d.res;
  if (nc > 32) {
    strip_slow(c, s_cand, nc, single, lane);
    return;
  }
  // warp-level cull: lane k < nc holds candidate k
  const bool has = lane < nc;
  const int my_site = has ? s_cand[lane] : 0;
  const double kx = c.s_px[my_site], ky = c.s_py[my_site];
  double fcen = INFINITY;
  if (has) fcen = d2_exact(cell_centre((c.r0 + c.r1 - 1) >> 1, res), cell_centre((c.c0 + c.c1 - 1) >> 1, res), kx, ky);
  double best = fcen;
  int bl = lane;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double ob = __shfl_xor_sync(0xffffffffu, best, o);
    int ol = __shfl_xor_sync(0xffffffffu, bl, o);
    if (ob < best || (ob == best && ol < bl)) best = ob, bl = ol;
  }
  const double jx = __shfl_sync(0xffffffffu, kx, bl), jy = __shfl_sync(0xffffffffu, ky, bl);
  bool keep = false;
  double dmax = 0.0;
  if (has) {
    const double ccx[2] = {cell_centre(c.r0, res), cell_centre(c.r1 - 1, res)};
    const double ccy[2] = {cell_centre(c.c0, res), cell_centre(c.c1 - 1, res)};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      double fk = d2_exact(ccx[q & 1], ccy[q >> 1], kx, ky);
      double fj = d2_exact(ccx[q & 1], ccy[q >> 1], jx, jy);
      if (fk - fj <= 1e-7 * (1.0 + fj)) keep = true;
      dmax = fmax(dmax, fk);
    }
  }
  const unsigned wmask = __ballot_sync(0xffffffffu, keep);
  const int wnc = __popc(wmask);
  if (single && !(wmask & 1u)) return;  // the robot owns no cell of this strip
  if (!keep) dmax = 0.0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) dmax = fmax(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
  // fp32 error bound of (d2_a - d2_b) in strip-local coordinates: < 2^-19.6 (dmax + span^2), DESIGN.md
  const double span = res * (double)(VOR_TY + rpw);
  const float margin = (float)((dmax + span * span) * (1.0 / 131072.0));

  // gather the kept candidates' site indices (ascending) into registers of every lane
  int ksite[4] = {0, 0, 0, 0};
  {
    unsigned mm = wmask;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (mm) {
        int src_lane = __ffs(mm) - 1;
        mm &= mm - 1;
        ksite[q] = __shfl_sync(0xffffffffu, my_site, src_lane);
      }
    }
  }
  if (wnc == 1) {
    int s1[1] = {ksite[0]};
    if (single) strip_fast<1, true>(c, s1, margin, lane, wbuf, mbuf);
    else strip_fast<1, false>(c, s1, margin, lane, wbuf, mbuf);
  } else if (wnc == 2) {
    int s2[2] = {ksite[0], ksite[1]};
    if (single) strip_fast<2, true>(c, s2, margin, lane, wbuf, mbuf);
    else strip_fast<2, false>(c, s2, margin, lane, wbuf, mbuf);
  } else if (wnc == 3) {
    int s3[3] = {ksite[0], ksite[1], ksite[2]};
    if (single) strip_fast<3, true>(c, s3, margin, lane, wbuf, mbuf);
    else strip_fast<3, false>(c, s3, margin, lane, wbuf, mbuf);
  } else if (wnc == 4) {
    if (single) strip_fast<4, true>(c, ksite, margin, lane, wbuf, mbuf);
    else strip_fast<4, false>(c, ksite, margin, lane, wbuf, mbuf);
  } else {
    if (keep) wcand[__popc(wmask & ((1u << lane) - 1u))] = my_site;
    __syncwarp();
    strip_slow(c, wcand, wnc, single, lane);
    __syncwarp();
  }
}

// dynamic shared memory: double px[ns_cap], py[ns_cap]; int out[ns_cap]; int cand[ns_cap];
// int wcand[VOR_WARPS][32]
//
// Global mode: grid = (tiles_x * tiles_y, 1, E); a CTA owns a 128-column x (8 rpw)-row tile, one
// strip per warp.  Window mode (single): grid = (1, R, E); a CTA owns the robot's whole clipped
// L-window and its warps loop over the window's strips, so the neighbour list is built once.
__global__ void __launch_bounds__(VOR_THREADS) k_voronoi(VorArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const EnvDev &d = a.d;
  const int ns_cap = d.R;
  double *s_px = reinterpret_cast<double *>(smem_raw);
  double *s_py = s_px + ns_cap;
  int *s_out = reinterpret_cast<int *>(s_py + ns_cap);
  int *s_cand = s_out + ns_cap;
  int *s_wcand = s_cand + ns_cap;
  // per-warp staging buffers (16-byte aligned): world tile, then mask tile
  size_t stage_off = ((size_t)ns_cap * 24 + VOR_WARPS * 32 * sizeof(int) + 15) & ~(size_t)15;
  float4 *s_wbuf = reinterpret_cast<float4 *>(smem_raw + stage_off);
  uint4 *s_mbuf = reinterpret_cast<uint4 *>(s_wbuf + (size_t)VOR_WARPS * a.rpw * 32);
  __shared__ int s_ns, s_nc;
  __shared__ double s_best[VOR_WARPS];
  __shared__ int s_besti[VOR_WARPS];
  __shared__ int s_woff[VOR_WARPS + 1];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int e = blockIdx.z;
  const int robot = a.single ? blockIdx.y : 0;
  const double res = d.res;
  const double *pos = d.pos_g + (size_t)e * d.R * 2;

  // ---- CTA rectangle: rows [tr0, tr1), columns [tc0, tc1), column tiles start at col_base0 -----
  int tr0, tr1, tc0, tc1, col_base0;
  if (a.single) {
    // decentralized_cvt.h:89-95: clipped L-window around the robot
    Window w = make_window(res, pos[2 * robot], pos[2 * robot + 1], d.L, d.N);
    tr0 = w.x0, tr1 = w.x1, tc0 = w.y0, tc1 = w.y1;
    col_base0 = (tc0 / VOR_TY) * VOR_TY;
  } else {
    const int th = VOR_WARPS * a.rpw;
    const int tx = blockIdx.x / a.tiles_y, ty = blockIdx.x - tx * a.tiles_y;
    tr0 = tx * th, tr1 = min(tr0 + th, d.N);
    col_base0 = ty * VOR_TY;
    tc0 = col_base0, tc1 = min(col_base0 + VOR_TY, d.N);
  }
  if (tr0 >= tr1 || tc0 >= tc1) return;

  // ---- site list ---------------------------------------------------------------------------
  if (!a.single) {
    for (int k = tid; k < d.R; k += VOR_THREADS) {
      s_px[k] = pos[2 * k], s_py[k] = pos[2 * k + 1], s_out[k] = k;
    }
    if (tid == 0) s_ns = d.R;
  } else if (warp == 0) {
    // decentralized_cvt.h:100-120: self first, then neighbours within range in index order
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
  if (a.single && s_cand[0] != 0) return;  // the robot itself was culled: it owns no cell here

  // ---- strips: (row chunk, column tile) pairs dealt round-robin to the warps -------------------
  StripCtx c;
  c.d = &d, c.e = e, c.robot = robot, c.src = a.src;
  c.s_px = s_px, c.s_py = s_py, c.s_out = s_out, c.sums = a.sums;
  c.labels = (a.labels && !a.single) ? a.labels + (size_t)e * d.N * d.N : nullptr;
  const int nrow = (tr1 - tr0 + a.rpw - 1) / a.rpw;
  const int ncol = (tc1 - 1 - col_base0) / VOR_TY + 1;
  for (int sidx = warp; sidx < nrow * ncol; sidx += VOR_WARPS) {
    const int ri = sidx / ncol, ci = sidx - ri * ncol;
    c.r0 = tr0 + ri * a.rpw;
    c.r1 = min(c.r0 + a.rpw, tr1);
    c.col_base = col_base0 + ci * VOR_TY;
    c.c0 = max(c.col_base, tc0);
    c.c1 = min(c.col_base + VOR_TY, tc1);
    if (c.c0 >= c.c1) continue;
    process_strip(c, s_cand, nc, a.single != 0, a.rpw, s_wcand + warp * 32, lane,
                  s_wbuf + (size_t)warp * a.rpw * 32, s_mbuf + (size_t)warp * a.rpw * 32);
    __syncwarp();
  }
}

}  // namespace ccb
