// vwindow.cuh -- DecentralizedCVT's per-robot Voronoi (decentralized_cvt.h:85-143) as one pass over
// the SEEN part of the robot's local window.
//
// Reference: every robot copies its N x N map, crops the L-window around itself, builds a CGAL Voronoi
// of {self} + {neighbours within pCommunicationRange} and sums its own cell (voronoi.cpp:55-150).
// Here: one CTA per (robot, environment).  Only the robot's own cell is reduced, and that cell is
// convex, so on every window row it is ONE column interval [lo, hi):
//   1. a thread per window row intersects the half-lines h_k(j) = d2_k - d2_self >= 0 of the K
//      neighbours (h_k is affine in the column; the crossing comes from one fp64 division and the
//      two cells next to it are re-checked with the stated fp64 rule when they are within rounding of
//      the bisector; a bisector that runs along the row through cell centres flags the row for the
//      per-cell rule);
//   2. the robot's bit plane, clipped to the window and ANDed with the row intervals, is staged in
//      shared memory (9 words per row) and the rows that still have a bit are compacted into a list;
//   3. the warps deal the listed rows between them in batches whose world values are all requested
//      before the first is used; a lane owns 8 consecutive columns (one mask byte, 16-byte loads from
//      the aligned column below the window origin, misalignment S uniform over the CTA) and keeps
//      fp64 column sums plus row-weighted sums -- the same exact cell-unit moments as voronoi.cuh;
//   4. one deterministic CTA reduction writes sums[e][r][0..3] (no atomics, no memset).
// Rule: a cell belongs to the robot unless some neighbour is STRICTLY nearer (d2 in fp64 without FMA
// contraction; self is site 0 of the local list, so ties go to the robot), as in voronoi.cuh.
//
// Restrictions (the caller falls back to k_voronoi<SRC_ROBOT, true> otherwise): L <= 256, N % 4 == 0.
#pragma once

#include "common.cuh"
#include "voronoi.cuh"

namespace ccb {

constexpr int VW_THREADS = 256;
constexpr int VW_WARPS = VW_THREADS / 32;
constexpr int VW_MAXL = 256;
constexpr int VW_WORDS = VW_MAXL / 32 + 1;  // mask words a window row can touch
constexpr int VW_BATCH = 3;

struct VwShared {
  uint32_t mask[VW_MAXL * VW_WORDS + 1];  // own & seen
  uint32_t own[VW_MAXL * VW_WORDS + 1];   // own cell (only read when unseen cells carry a fill value)
  int lo[VW_MAXL], hi[VW_MAXL];           // own-cell interval of every window row, map columns
  uint32_t flags[VW_MAXL / 32];           // row has a cell to add
  uint32_t exact[VW_MAXL / 32];           // row needs the per-cell rule
  int act[VW_MAXL];
  int nact, K;
  double red[VW_WARPS][4];
};
// dynamic shared memory: VwShared followed by px[R], py[R] (self first, then the neighbours)
__host__ __device__ inline size_t vw_smem_bytes(int R) { return sizeof(VwShared) + sizeof(double) * 2 * (size_t)R; }

// 1 if neighbour (kx, ky) is strictly nearer to cell (row, col) than the robot (sx, sy)
// (vw_beats / vw_row_interval / vw_range_bits also compile for the host: ccb_selftest_vw_row_interval, CPU tests)
__host__ __device__ __forceinline__ bool vw_beats(double res, int row, int col, double kx, double ky, double sx, double sy) {
  const double cx = cell_centre(row, res), cy = cell_centre(col, res);
  return d2_exact(cx, cy, kx, ky) < d2_exact(cx, cy, sx, sy);
}

// own-cell interval of one window row (columns y0 + lo .. y0 + hi - 1); returns true when the row needs
// the per-cell rule (then [lo, hi) only holds the unambiguous constraints)
__host__ __device__ __noinline__ bool vw_row_interval(double res, int row, int y0, int ncols, int K, const double *px,
                                                      const double *py, int &lo_out, int &hi_out) {
  const double cx = cell_centre(row, res), cy0 = cell_centre(y0, res);
  const double sx = px[0], sy = py[0];
  const double f0 = d2_exact(cx, cy0, sx, sy);
  int lo = 0, hi = ncols;
  bool exact = false;
  for (int k = 1; k <= K; ++k) {
    const double kx = px[k], ky = py[k];
    const double hc = d2_exact(cx, cy0, kx, ky) - f0;  // h at the first column
    const double B = 2.0 * res * (sy - ky);             // h(j + 1) - h(j): centres are res apart
    const double eps = 1e-9 * (1.0 + f0 + fabs(hc) + ncols * fabs(B));
    if (fabs(B) * ncols <= eps) {
      // the bisector runs (numerically) along this row
      if (fabs(hc) <= eps) exact = true;
      else if (hc < 0.0) hi = 0;
      continue;
    }
    const double x = -hc / B;  // h(x) = 0, in columns from y0
    if (B > 0.0) {             // h increasing: the robot owns the columns >= x
      double tf = fmin(fmax(ceil(x), 0.0), (double)ncols);
      int T = (int)tf;
      if (T - 1 >= 0 && T - 1 < ncols && fabs(hc + (T - 1) * B) <= eps && !vw_beats(res, row, y0 + T - 1, kx, ky, sx, sy)) T -= 1;
      if (T >= 0 && T < ncols && fabs(hc + T * B) <= eps && vw_beats(res, row, y0 + T, kx, ky, sx, sy)) T += 1;
      lo = max(lo, T);
    } else {                   // h decreasing: the robot owns the columns <= x
      double tf = fmin(fmax(floor(x) + 1.0, 0.0), (double)ncols);
      int T = (int)tf;
      if (T - 1 >= 0 && T - 1 < ncols && fabs(hc + (T - 1) * B) <= eps && vw_beats(res, row, y0 + T - 1, kx, ky, sx, sy)) T -= 1;
      if (T >= 0 && T < ncols && fabs(hc + T * B) <= eps && !vw_beats(res, row, y0 + T, kx, ky, sx, sy)) T += 1;
      hi = min(hi, T);
    }
  }
  lo_out = lo, hi_out = hi;
  return exact;
}

// bits [a, b) of a 32-bit word whose bit 0 is column cb
__host__ __device__ __forceinline__ uint32_t vw_range_bits(int a, int b, int cb) {
  const int lo = min(max(a - cb, 0), 32), hi = min(max(b - cb, 0), 32);
  if (hi <= lo) return 0u;
  return ((hi - lo >= 32) ? 0xffffffffu : ((1u << (hi - lo)) - 1u)) << lo;
}

struct VwAcc {
  double cs[8], sx, sxx;
};

template <int S>
__device__ __forceinline__ void vw_accumulate(const EnvDev &d, VwShared &sh, const float *__restrict__ world, const Window &w,
                                              int wbase, VwAcc &acc) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nact = sh.nact;
  const int gy0 = w.oy + 8 * lane;
  const int gyb = gy0 - S;
  constexpr int NV = S == 0 ? 2 : 3;
  bool grp_in[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) grp_in[j] = gyb + 4 * j >= 0 && gyb + 4 * j + 3 < d.N;
  const int off = gy0 - 32 * wbase;
  const int wi = min(off >> 5, VW_WORDS - 1), shf = off & 31;  // lanes past the window read zeros
  const bool lane_in = 8 * lane < d.L;
  const float fill = d.fill;
  const bool dense = fill != 0.f;
#pragma unroll 1
  for (int base = warp * VW_BATCH; base < nact; base += VW_WARPS * VW_BATCH) {
    float4 v[VW_BATCH][NV];
    unsigned vis[VW_BATCH], own[VW_BATCH];
#pragma unroll
    for (int b = 0; b < VW_BATCH; ++b) {
      vis[b] = own[b] = 0u;
#pragma unroll
      for (int j = 0; j < NV; ++j) v[b][j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (base + b < nact && lane_in) {
        const int a = sh.act[base + b];
        vis[b] = __funnelshift_r(sh.mask[a * VW_WORDS + wi], sh.mask[a * VW_WORDS + wi + 1], shf) & 0xffu;
        if (dense) own[b] = __funnelshift_r(sh.own[a * VW_WORDS + wi], sh.own[a * VW_WORDS + wi + 1], shf) & 0xffu;
        if (vis[b] != 0u && d.use_sensor) {
          const float4 *src = reinterpret_cast<const float4 *>(world + (size_t)(w.ox + a) * d.N + gyb);
#pragma unroll
          for (int j = 0; j < NV; ++j)
            if (grp_in[j]) v[b][j] = __ldcs(src + j);  // read once per CTA: streaming load, 3-5 % over __ldg
        }
      }
    }
#pragma unroll
    for (int b = 0; b < VW_BATCH; ++b) {
      if (base + b >= nact) break;  // warp-uniform
      float x[4 * NV];
#pragma unroll
      for (int j = 0; j < NV; ++j) x[4 * j] = v[b][j].x, x[4 * j + 1] = v[b][j].y, x[4 * j + 2] = v[b][j].z, x[4 * j + 3] = v[b][j].w;
      const double xd = (double)(w.ox + sh.act[base + b]) + 0.5;
      double s = 0.0;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float unseen = (dense && ((own[b] >> i) & 1u)) ? fill : 0.f;
        const float c = ((vis[b] >> i) & 1u) ? x[S + i] : unseen;
        const double dv = (double)c;
        acc.cs[i] += dv;
        s += dv;
      }
      acc.sx = fma(xd, s, acc.sx);
      acc.sxx = fma(xd * xd, s, acc.sxx);
    }
  }
}

// robot0: first robot of the launch (a shard of the robots when the windows are dealt over several GPUs)
__global__ void __launch_bounds__(VW_THREADS, 3) k_voronoi_window(EnvDev d, double *__restrict__ sums, int robot0) {
  extern __shared__ __align__(16) unsigned char vw_smem[];
  VwShared &sh = *reinterpret_cast<VwShared *>(vw_smem);
  double *s_px = reinterpret_cast<double *>(vw_smem + sizeof(VwShared));
  double *s_py = s_px + d.R;
  const int r = robot0 + blockIdx.x, e = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t er = (size_t)e * d.R + r;
  const double res = d.res;
  const double *pos = d.pos_g + (size_t)e * d.R * 2;
  const double sx = pos[2 * r], sy = pos[2 * r + 1];
  // decentralized_cvt.h:89-95: clipped L-window around the robot
  const Window w = make_window(res, sx, sy, d.L, d.N);
  const int ncols = w.y1 - w.y0;
  if (tid < VW_MAXL / 32) sh.flags[tid] = 0u, sh.exact[tid] = 0u;
  if (tid == 0) sh.nact = 0, sh.mask[VW_MAXL * VW_WORDS] = 0u, sh.own[VW_MAXL * VW_WORDS] = 0u;

  // ---- sites: self, then the neighbours within communication range in ascending index
  //      (decentralized_cvt.h:100-120) ------------------------------------------------------------
  if (warp == 0) {
    if (lane == 0) s_px[0] = sx, s_py[0] = sy;
    int n = 1;
    for (int base = 0; base < d.R; base += 32) {
      const int j = base + lane;
      bool keep = false;
      double jx = 0, jy = 0;
      if (j < d.R && j != r) {
        jx = pos[2 * j], jy = pos[2 * j + 1];
        const double dx = __dsub_rn(jx, sx), dy = __dsub_rn(jy, sy);
        keep = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy))) < d.comm_range;
      }
      const unsigned m = __ballot_sync(0xffffffffu, keep);
      if (keep) {
        const int slot = n + __popc(m & ((1u << lane) - 1u));
        s_px[slot] = jx, s_py[slot] = jy;
      }
      n += __popc(m);
    }
    if (lane == 0) sh.K = n - 1;
  }
  __syncthreads();
  const int K = sh.K;

  // window rows that can hold a cell to add: inside the clipped window and (unless unseen cells carry a value)
  // inside the rows of the robot's seen box (EnvDev::seen_box, kept by k_mark_windows)
  const bool dense = d.fill != 0.f;
  int alo = w.x0 - w.ox, ahi = w.x1 - w.ox;
  if (!dense) alo = max(alo, d.seen_box[4 * er] - w.ox), ahi = min(ahi, d.seen_box[4 * er + 1] - w.ox);
  ahi = min(ahi, VW_MAXL);

  // ---- 1. own-cell interval of every such row (thread a <-> window row a = map row w.ox + a) ----
  {
    const int a = tid;  // VW_THREADS == VW_MAXL
    const int row = w.ox + a;
    int lo = 0, hi = 0;
    if (a >= alo && a < ahi && ncols > 0) {
      if (K == 0) lo = 0, hi = ncols;
      else if (vw_row_interval(res, row, w.y0, ncols, K, s_px, s_py, lo, hi)) atomicOr(&sh.exact[a >> 5], 1u << (a & 31));
    }
    sh.lo[a] = w.y0 + lo, sh.hi[a] = w.y0 + max(hi, lo);
  }
  __syncthreads();

  // ---- 2. mask words: seen & window & own interval, batches of VW_STAGE loads issued before use ------
  const uint32_t *plane = robot_plane(d, d.seenb, e, r);
  const int wbase = w.oy >> 5;  // arithmetic shift: floor for negative origins
  const int nstage = max(ahi - alo, 0) * VW_WORDS;
  constexpr int VW_STAGE = 4;
#pragma unroll 1
  for (int t0 = 0; t0 < nstage; t0 += VW_STAGE * VW_THREADS) {
    uint32_t mw[VW_STAGE], ow[VW_STAGE];
#pragma unroll
    for (int it = 0; it < VW_STAGE; ++it) {
      const int t = t0 + tid + it * VW_THREADS;
      const int ar = t / VW_WORDS, j = t - ar * VW_WORDS;
      const int a = alo + ar, wd = wbase + j;
      uint32_t m = 0u, o = 0u;
      if (t < nstage && wd >= 0 && wd < d.NW) {
        o = vw_range_bits(sh.lo[a], sh.hi[a], wd << 5);
        if (o) m = __ldg(plane + (size_t)(w.ox + a) * d.NW + wd) & o;
      }
      mw[it] = m, ow[it] = o;
    }
#pragma unroll
    for (int it = 0; it < VW_STAGE; ++it) {
      const int t = t0 + tid + it * VW_THREADS;
      if (t >= nstage) continue;
      const int ar = t / VW_WORDS, j = t - ar * VW_WORDS;
      const int a = alo + ar;
      uint32_t m = mw[it], o = ow[it];
      if ((dense ? o : m) && ((sh.exact[a >> 5] >> (a & 31)) & 1u)) {
        // per-cell rule on this row: drop the cells where some neighbour is strictly nearer
        uint32_t todo = dense ? o : m;
        while (todo) {
          const int bit = __ffs(todo) - 1;
          todo &= todo - 1;
          const int col = ((wbase + j) << 5) + bit;
          bool lost = false;
          for (int k = 1; k <= K && !lost; ++k) lost = vw_beats(res, w.ox + a, col, s_px[k], s_py[k], sx, sy);
          if (lost) m &= ~(1u << bit), o &= ~(1u << bit);
        }
      }
      sh.mask[a * VW_WORDS + j] = m, sh.own[a * VW_WORDS + j] = o;
      if (dense ? o : m) atomicOr(&sh.flags[a >> 5], 1u << (a & 31));
    }
  }
  __syncthreads();
  {
    // rows with work, in ascending order (a fixed order makes the fp64 sums reproducible run to run)
    const int a = tid;
    int before = 0;
#pragma unroll
    for (int wq = 0; wq < VW_MAXL / 32; ++wq) {
      const uint32_t f = sh.flags[wq];
      before += wq < (a >> 5) ? __popc(f) : (wq == (a >> 5) ? __popc(f & ((1u << (a & 31)) - 1u)) : 0);
    }
    if ((sh.flags[a >> 5] >> (a & 31)) & 1u) sh.act[before] = a;
    if (a == VW_MAXL - 1) sh.nact = before + ((sh.flags[a >> 5] >> (a & 31)) & 1u);
  }
  __syncthreads();

  // ---- 3. accumulate the listed rows ---------------------------------------------------------------
  VwAcc acc;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc.cs[i] = 0.0;
  acc.sx = acc.sxx = 0.0;
  const float *world = d.world + (size_t)e * cell_count(d);
  switch (w.oy & 3) {
    case 0: vw_accumulate<0>(d, sh, world, w, wbase, acc); break;
    case 1: vw_accumulate<1>(d, sh, world, w, wbase, acc); break;
    case 2: vw_accumulate<2>(d, sh, world, w, wbase, acc); break;
    default: vw_accumulate<3>(d, sh, world, w, wbase, acc); break;
  }

  // ---- 4. fold with the column coordinates, reduce over the CTA in a fixed order -------------------
  double m = 0.0, smy = 0.0, s2 = acc.sxx;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const double y = (double)(w.oy + 8 * lane + i) + 0.5;
    m += acc.cs[i];
    smy = fma(y, acc.cs[i], smy);
    s2 = fma(y * y, acc.cs[i], s2);
  }
  const Sum4 tot = reduce4(m, acc.sx, smy, s2);
  if (lane == 0) sh.red[warp][0] = tot.m, sh.red[warp][1] = tot.sx, sh.red[warp][2] = tot.sy, sh.red[warp][3] = tot.s2;
  __syncthreads();
  if (tid < 4) {
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < VW_WARPS; ++k) t += sh.red[k][tid];
    sums[er * 4 + tid] = t;
  }
}

}  // namespace ccb
