// voronoi.cuh -- discrete Voronoi assignment fused with the per-cell mass / moment reductions.
//
// Replaces reference cppsrc/core/src/voronoi.cpp:55-327 (CGAL arrangement + scan-line
// rasterisation + fp64 accumulation, one OpenMP thread per site) and the per-robot windowed
// variant of algorithms/decentralized_cvt.h:85-143.
//
// Semantics: every cell centre ((i+.5)res, (j+.5)res) is labelled with its nearest site; the
// distance is d2 = dx*dx + dy*dy evaluated in fp64 without FMA contraction and ties go to the
// lowest site index (the build's stated rule).  The labels are never stored: each
// warp streams a 64-column x rpw-row region of the map once and accumulates, per label,
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
//     warp reduction and one set of fp64 atomics per (region, label);
//   * the region is staged global -> shared with cp.async in double-buffered 4 KB chunks (each lane
//     copies the 16 bytes per row it later consumes, so a per-thread wait is the only sync):
//     memory-level parallelism comes from shared memory, not from registers.
#pragma once

#include "common.cuh"

namespace ccb {

constexpr int VOR_THREADS = 256;
constexpr int VOR_WARPS = VOR_THREADS / 32;
constexpr int VOR_WX = 2, VOR_WY = 4;  // warp grid of a CTA tile (global mode)
constexpr int VOR_TW = 64;       // region width: 16 lanes x 4 cells; lanes 16-31 take the next row
constexpr int VOR_CH = 8;        // passes (2 rows each) per staging chunk: 8 x 512 B = 4 KB
constexpr int VOR_NBUF = 3;      // ring of staging chunks per warp (prefetch distance VOR_NBUF - 1)
constexpr int VOR_MAXRPW = 128;  // rows per warp region
constexpr int VOR_MAXFAST = 3;   // regions with more competing sites are split, then take the generic path
constexpr int VOR_MINSPLIT = 4;
// passes of the streaming loop interleaved for ILP (2 warps per scheduler cannot hide the
// LDS -> F2F -> DADD chain of a single pass)
constexpr int VOR_UNROLL = 1;  // do not split below this many rows

enum { SRC_WORLD = 0, SRC_EXPLORED = 1, SRC_ROBOT = 2 };

struct VorArgs {
  EnvDev d;
  int src;      // SRC_*
  int rpw;      // rows per warp region (even)
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

// Everything a warp needs to stream one region, by value: nothing here points at the kernel
// parameter block (taking its address would force a copy of it into local memory).
struct StripCtx {
  int N, R;            // map size, robots per environment
  double res;
  float fill;          // value of not-yet-seen cells of a robot map (0 unless pRobotMapUseUnknownImportance)
  int use_sensor;
  const float *world;     // this environment's world map
  const uint32_t *seen;   // this environment's seen plane holding `robot`'s bit (SRC_ROBOT)
  const uint8_t *sys;     // this environment's system mask (SRC_EXPLORED)
  int e, robot;        // robot: mask bit for SRC_ROBOT / output slot in window mode
  int src;
  int r0, r1;          // region rows [r0, r1)
  int c0, c1;          // valid columns [c0, c1)
  int col_base;        // absolute column of lane 0's first cell (multiple of VOR_TW)
  const double *s_px, *s_py;  // site coordinates (shared memory)
  const int *s_out;           // site -> output slot
  double *sums;
  int *labels;                // optional label plane of this environment
};

// ---- staging ----------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(unsigned smem_addr, const void *gmem, int src_bytes) {
  // src_bytes 0: the 16 destination bytes are zero-filled and global memory is not read
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_addr), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async4(unsigned smem_addr, const void *gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(smem_addr), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// ---- bulk (TMA 1-D) copies completing on an mbarrier: one 256-byte row per lane, one warp
//      instruction per chunk (SASS: UBLKCP + SYNCS) ------------------------------------------------
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  unsigned done = 0;
  int spins = 0;
  while (!done) {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done && ++spins > (1 << 22)) __trap();  // a lost copy must fail loudly, never hang the GPU
  }
}

// ---- staging of one chunk ------------------------------------------------------------------------
// Pass t of a region covers rows r0 + 2t (lanes 0-15) and r0 + 2t + 1 (lanes 16-31); a chunk is up
// to VOR_CH passes; a lane's slot for pass j of a chunk is wslot[j * 32] (16 bytes).
// Masked sources (cold helper, scalars by value so that nothing is forced into local memory):
// stage the visibility masks into the destination slots themselves (a lane only ever touches its own
// 16 bytes per pass), wait, fold them into the per-pass 4-bit visibility word, then issue the world
// copies for the passes where something is visible; cells that are not visible are never read.
__device__ __noinline__ unsigned stage_masked(const uint8_t *sys, const uint32_t *seen, int sh, const float *world,
                                             const float *any_valid, size_t rstride, unsigned in4, int np,
                                             int np_lane, bool aligned, bool use_sensor, float4 *wslot) {
  unsigned vis = 0;
  const unsigned wsm = (unsigned)__cvta_generic_to_shared(wslot);
  if (in4) {
    if (sys) {
      if (aligned && in4 == 0xfu) {
#pragma unroll 1
        for (int j = 0; j < np_lane; ++j) cp_async4(wsm + j * 512, sys + j * rstride);
        cp_async_commit();
        cp_async_wait<0>();
#pragma unroll 1
        for (int j = 0; j < np_lane; ++j) {
          const uint32_t q = *reinterpret_cast<const uint32_t *>(wslot + j * 32);
          vis |= ((q & 1u) | ((q >> 7) & 2u) | ((q >> 14) & 4u) | ((q >> 21) & 8u)) << (4 * j);
        }
      } else {
#pragma unroll 1
        for (int j = 0; j < np_lane; ++j)
          for (int i = 0; i < 4; ++i)
            if (in4 & (1u << i)) vis |= (unsigned)(__ldg(sys + j * rstride + i) & 1u) << (4 * j + i);
      }
    } else if (seen) {
      if (aligned && in4 == 0xfu) {
#pragma unroll 1
        for (int j = 0; j < np_lane; ++j) cp_async16(wsm + j * 512, seen + j * rstride, 16);
        cp_async_commit();
        cp_async_wait<0>();
#pragma unroll 1
        for (int j = 0; j < np_lane; ++j) {
          const uint4 q = *reinterpret_cast<const uint4 *>(wslot + j * 32);
          vis |= (((q.x >> sh) & 1u) | (((q.y >> sh) & 1u) << 1) | (((q.z >> sh) & 1u) << 2) |
                  (((q.w >> sh) & 1u) << 3)) << (4 * j);
        }
      } else {
#pragma unroll 1
        for (int j = 0; j < np_lane; ++j)
          for (int i = 0; i < 4; ++i)
            if (in4 & (1u << i)) vis |= ((__ldg(seen + j * rstride + i) >> sh) & 1u) << (4 * j + i);
      }
    } else {  // unmasked source, partially inside the region
#pragma unroll 1
      for (int j = 0; j < np_lane; ++j) vis |= in4 << (4 * j);
    }
  }
  if (!use_sensor) {
#pragma unroll 1
    for (int j = 0; j < np; ++j) wslot[j * 32] = make_float4(0.f, 0.f, 0.f, 0.f);
  } else if (aligned && (in4 == 0xfu || in4 == 0)) {
#pragma unroll 1
    for (int j = 0; j < np; ++j) {
      const bool need = ((vis >> (4 * j)) & 0xfu) != 0;
      cp_async16(wsm + j * 512, need ? (const void *)(world + j * rstride) : (const void *)any_valid, need ? 16 : 0);
    }
  } else {
#pragma unroll 1
    for (int j = 0; j < np; ++j) {
      float t[4] = {0.f, 0.f, 0.f, 0.f};
      for (int i = 0; i < 4; ++i)
        if ((vis >> (4 * j + i)) & 1u) t[i] = __ldg(world + j * rstride + i);
      wslot[j * 32] = make_float4(t[0], t[1], t[2], t[3]);
    }
  }
  cp_async_commit();
  return vis;
}

// Stage chunk [t0, t0 + np) of the region into this lane's slots and commit ONE cp.async group
// (not waited for).  Returns the per-pass 4-bit visibility of this lane's cells (pass j -> bits
// 4j..4j+3): inside the region AND visible in the selected map; 0xffffffff = everything visible.
__device__ __forceinline__ unsigned stage_chunk(const StripCtx &c, int t0, int np, int lane, float4 *wslot) {
  const bool aligned = (c.N & 3) == 0;
  const int col = c.col_base + 4 * (lane & 15);
  unsigned in4 = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (col + i >= c.c0 && col + i < c.c1) in4 |= 1u << i;
  const int row0 = c.r0 + 2 * t0 + (lane >> 4);
  const size_t off0 = (size_t)row0 * c.N + col;
  const float *world = c.world + off0;
  const size_t rstride = (size_t)2 * c.N;
  int np_lane = np;  // rows beyond the region (odd row count): nothing visible
  if (row0 + 2 * (np - 1) >= c.r1) np_lane = np - 1;
  if (c.src == SRC_WORLD) {
    const bool full = __all_sync(0xffffffffu, in4 == 0xfu && np_lane == np);
    if (full && aligned && c.use_sensor) {  // hot path: one 16-byte cp.async per pass
      const unsigned wsm = (unsigned)__cvta_generic_to_shared(wslot);
#pragma unroll 1
      for (int j = 0; j < np; ++j) cp_async16(wsm + j * 512, world + j * rstride, 16);
      cp_async_commit();
      return 0xffffffffu;
    }
    return stage_masked(nullptr, nullptr, 0, world, c.world, rstride, in4, np, np_lane, aligned, c.use_sensor, wslot);
  }
  if (c.src == SRC_EXPLORED)
    return stage_masked(c.sys + off0, nullptr, 0, world, c.world, rstride, in4, np, np_lane, aligned, c.use_sensor, wslot);
  return stage_masked(nullptr, c.seen + off0, c.robot & 31, world, c.world, rstride, in4, np, np_lane, aligned,
                      c.use_sensor, wslot);
}

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
  double *dst = sums + ((size_t)e * R + slot) * 4;
  atomicAdd(dst + 0, m);
  atomicAdd(dst + 1, sx);
  atomicAdd(dst + 2, sy);
  atomicAdd(dst + 3, s2);
}

// the stated rule among nc <= 3 candidates k0 < k1 < k2 (site indices): index 0..nc-1 of the winner
__device__ __noinline__ int exact_label(const double *s_px, const double *s_py, int k0, int k1, int k2, int nc, int row,
                                        int col, double res) {
  const double cx = cell_centre(row, res), cy = cell_centre(col, res);
  int best = 0;
  double bd = d2_exact(cx, cy, s_px[k0], s_py[k0]);
  if (nc > 1) {
    double dd = d2_exact(cx, cy, s_px[k1], s_py[k1]);
    if (dd < bd) bd = dd, best = 1;
  }
  if (nc > 2) {
    double dd = d2_exact(cx, cy, s_px[k2], s_py[k2]);
    if (dd < bd) bd = dd, best = 2;
  }
  return best;
}

// Two owners: returns T in [0, rows], packed with two flags: bit 30 = candidate 1 owns the rows
// u < T (else the rows u >= T); bit 29 = every cell of the column needs the exact rule.
// h(u) = hc + u HA is d2_1 - d2_0 at row r0 + u of column `col`.
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
    const int lab = exact_label(s_px, s_py, k0, k1, k1, 2, r0 + T - 1, col, res);
    if ((lab == 1) != first1) T -= 1;
  }
  if (T >= 0 && T < rows && fabs(hc + T * HA) <= eps) {
    const int lab = exact_label(s_px, s_py, k0, k1, k1, 2, r0 + T, col, res);
    if ((lab == 1) == first1) T += 1;
  }
  return T | (first1 ? (1 << 30) : 0);
}

// Fast path: NC (1..VOR_MAXFAST) candidate sites k0 < k1 < k2 can own cells of this region.
// SINGLE: accumulate only the cells owned by candidate 0 (the robot itself).
// LABELS: also store the label of every cell (parity tests only; a separate instantiation).
// bars / phase: this warp's ring mbarriers (shared-memory addresses) and their parity bits.
template <int NC, bool SINGLE, bool LABELS>
__device__ __forceinline__ void region_fast(const StripCtx &c, int k0, int k1, int k2, int lane, float4 *wbuf,
                                            unsigned *svis, unsigned bars, unsigned &phase) {
  const double res = c.res;
  const int hp = lane >> 4;
  const int col = c.col_base + 4 * (lane & 15);
  const int rows = c.r1 - c.r0;
  const int npass = (rows + 1) >> 1;
  const int ksel[3] = {k0, k1, k2};

  // ---- label plan ---------------------------------------------------------------------------
  // NC == 2: integer row threshold per column.  NC == 3: fp32 affine differences per cell.
  int T0 = 0, T1 = 0, T2 = 0, T3 = 0;
  bool first1 = true;
  unsigned exact_cols = 0;
  float h0a = 0.f, haa = 0.f, hba = 0.f, h0b = 0.f, hab = 0.f, hbb = 0.f, margin = 0.f;
  if (NC >= 2) {
    const double cx0 = cell_centre(c.r0, res), cy0 = cell_centre(col, res);
    const double p0x = c.s_px[k0], p0y = c.s_py[k0];
    const double f0 = d2_exact(cx0, cy0, p0x, p0y);
    const double p1x = c.s_px[k1], p1y = c.s_py[k1];
    const double H0 = d2_exact(cx0, cy0, p1x, p1y) - f0;
    const double HA = 2.0 * res * (p0x - p1x), HB = 2.0 * res * (p0y - p1y);
    double hmax = fabs(H0) + rows * fabs(HA) + 4.0 * fabs(HB);
    if (NC == 2) {
      const double eps = 1e-9 * (1.0 + f0 + hmax);  // >> the fp64 rounding of d2 and of the affine form
      const int t0 = column_threshold(c.s_px, c.s_py, k0, k1, c.r0, col + 0, rows, H0, HA, eps, res);
      const int t1 = column_threshold(c.s_px, c.s_py, k0, k1, c.r0, col + 1, rows, H0 + HB, HA, eps, res);
      const int t2 = column_threshold(c.s_px, c.s_py, k0, k1, c.r0, col + 2, rows, H0 + 2.0 * HB, HA, eps, res);
      const int t3 = column_threshold(c.s_px, c.s_py, k0, k1, c.r0, col + 3, rows, H0 + 3.0 * HB, HA, eps, res);
      // the orientation flag depends on the sign of HA only, so it is the same for the 4 columns
      first1 = (t0 >> 30) & 1;
      T0 = t0 & 0xffff, T1 = t1 & 0xffff, T2 = t2 & 0xffff, T3 = t3 & 0xffff;
      exact_cols = ((t0 >> 29) & 1u) | (((t1 >> 29) & 1u) << 1) | (((t2 >> 29) & 1u) << 2) | (((t3 >> 29) & 1u) << 3);
    } else {
      const double p2x = c.s_px[k2], p2y = c.s_py[k2];
      const double G0 = d2_exact(cx0, cy0, p2x, p2y) - f0;
      const double GA = 2.0 * res * (p0x - p2x), GB = 2.0 * res * (p0y - p2y);
      hmax = fmax(hmax, fabs(G0) + rows * fabs(GA) + 4.0 * fabs(GB));
      h0a = (float)H0, haa = (float)HA, hba = (float)HB;
      h0b = (float)G0, hab = (float)GA, hbb = (float)GB;
      // fp32 evaluation error of h_k is < 2^-22 hmax; differences of two h's double it
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) hmax = fmax(hmax, __shfl_xor_sync(0xffffffffu, hmax, o));
      margin = (float)(hmax * (1.0 / 262144.0));
    }
  }

  constexpr int NSETS = SINGLE ? 1 : NC;
  Acc6 acc[NSETS];
#pragma unroll
  for (int s = 0; s < NSETS; ++s) acc[s].zero();
  const float fill = (c.src == SRC_ROBOT) ? c.fill : 0.f;

  // ---- stream the region through a ring of VOR_NBUF chunks (prefetch distance VOR_NBUF - 1) ------
  // Interior regions of the world map (`bulk`): lanes 0-15 each copy one 256-byte row of the chunk
  // with a bulk copy that completes on the slot's mbarrier -- one warp instruction per chunk.
  // Otherwise (masked sources, map edges): per-lane cp.async; every iteration commits exactly one
  // group (possibly empty), so "at most VOR_NBUF - 1 groups pending" means "chunk ch has landed".
  const int nch = (npass + VOR_CH - 1) / VOR_CH;
  float4 *wl = wbuf + lane;
  const bool bulk = c.src == SRC_WORLD && (c.N & 3) == 0 && c.use_sensor && c.c0 == c.col_base &&
                    c.c1 == c.col_base + VOR_TW && (rows & 1) == 0;
  const float *brow = c.world + (size_t)(c.r0 + (lane & 15)) * c.N + c.col_base;  // this lane's row of chunk 0
  const unsigned ring = (unsigned)__cvta_generic_to_shared(wbuf);
#pragma unroll 1
  for (int b = 0; b < VOR_NBUF - 1; ++b) {
    if (bulk) {
      if (b < nch) {
        const int nr = min(2 * VOR_CH, rows - 2 * VOR_CH * b);
        if (lane == 0) mbar_expect_tx(bars + 8 * b, nr * 256);
        __syncwarp();
        if (lane < nr) bulk_copy_g2s(ring + b * (VOR_CH * 512) + lane * 256, brow + (size_t)(2 * VOR_CH * b) * c.N, 256, bars + 8 * b);
      }
    } else if (b < nch) {
      svis[b * 32 + lane] = stage_chunk(c, b * VOR_CH, min(VOR_CH, npass - b * VOR_CH), lane, wl + b * (VOR_CH * 32));
    } else {
      cp_async_commit();
    }
  }
#pragma unroll 1
  for (int ch = 0; ch < nch; ++ch) {
    const int t0 = ch * VOR_CH;
    const int np = min(VOR_CH, npass - t0);
    const int cb = ch % VOR_NBUF;
    unsigned vis_cur = 0xffffffffu;
    {
      const int nx = ch + VOR_NBUF - 1, nb = nx % VOR_NBUF;
      if (bulk) {
        if (nx < nch) {
          const int nr = min(2 * VOR_CH, rows - 2 * VOR_CH * nx);
          if (lane == 0) mbar_expect_tx(bars + 8 * nb, nr * 256);
          __syncwarp();
          if (lane < nr) bulk_copy_g2s(ring + nb * (VOR_CH * 512) + lane * 256, brow + (size_t)(2 * VOR_CH * nx) * c.N, 256, bars + 8 * nb);
        }
        mbar_wait(bars + 8 * cb, (phase >> cb) & 1u);
        phase ^= 1u << cb;
      } else {
        if (nx < nch) svis[nb * 32 + lane] = stage_chunk(c, nx * VOR_CH, min(VOR_CH, npass - nx * VOR_CH), lane, wl + nb * (VOR_CH * 32));
        else cp_async_commit();
        cp_async_wait<VOR_NBUF - 1>();
        vis_cur = svis[cb * 32 + lane];
      }
    }
    const float4 *wp = wl + cb * (VOR_CH * 32);
    const bool plain_cur = bulk || __all_sync(0xffffffffu, vis_cur == 0xffffffffu);
    int u = 2 * t0 + hp;                         // row of this lane in this pass, relative to r0
    double xd = (double)(c.r0 + u) + 0.5;
#pragma unroll(VOR_UNROLL)
    for (int j = 0; j < np; ++j, u += 2, xd += 2.0) {
      const float4 w4 = wp[j * 32];
      float v0 = w4.x, v1 = w4.y, v2 = w4.z, v3 = w4.w;
      if (!plain_cur) {
        const unsigned vb = (vis_cur >> (4 * j)) & 0xfu;
        const bool row_in = c.r0 + u < c.r1;
        const float f0 = (fill != 0.f && row_in && col + 0 >= c.c0 && col + 0 < c.c1) ? fill : 0.f;
        const float f1 = (fill != 0.f && row_in && col + 1 >= c.c0 && col + 1 < c.c1) ? fill : 0.f;
        const float f2 = (fill != 0.f && row_in && col + 2 >= c.c0 && col + 2 < c.c1) ? fill : 0.f;
        const float f3 = (fill != 0.f && row_in && col + 3 >= c.c0 && col + 3 < c.c1) ? fill : 0.f;
        v0 = (vb & 1u) ? v0 : f0, v1 = (vb & 2u) ? v1 : f1, v2 = (vb & 4u) ? v2 : f2, v3 = (vb & 8u) ? v3 : f3;
      }
      int l0 = 0, l1 = 0, l2 = 0, l3 = 0;
      if (NC == 2) {
        l0 = ((u < T0) == first1) ? 1 : 0, l1 = ((u < T1) == first1) ? 1 : 0;
        l2 = ((u < T2) == first1) ? 1 : 0, l3 = ((u < T3) == first1) ? 1 : 0;
        if (exact_cols) {
          if (exact_cols & 1u) l0 = exact_label(c.s_px, c.s_py, k0, k1, k1, 2, c.r0 + u, col + 0, res);
          if (exact_cols & 2u) l1 = exact_label(c.s_px, c.s_py, k0, k1, k1, 2, c.r0 + u, col + 1, res);
          if (exact_cols & 4u) l2 = exact_label(c.s_px, c.s_py, k0, k1, k1, 2, c.r0 + u, col + 2, res);
          if (exact_cols & 8u) l3 = exact_label(c.s_px, c.s_py, k0, k1, k1, 2, c.r0 + u, col + 3, res);
        }
      } else if (NC == 3) {
        const float uf = (float)u;
        const float hr0 = fmaf(uf, haa, h0a), hr1 = fmaf(uf, hab, h0b);
        bool amb = false;
        int lab[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float g1 = fmaf((float)i, hba, hr0), g2 = fmaf((float)i, hbb, hr1);
          float best = 0.f, second;  // winner among {0: 0, 1: g1, 2: g2}
          int bl = 0;
          if (g1 < best) second = best, best = g1, bl = 1;
          else second = g1;
          if (g2 < best) second = best, best = g2, bl = 2;
          else second = fminf(second, g2);
          amb |= (second - best < margin);
          lab[i] = bl;
        }
        l0 = lab[0], l1 = lab[1], l2 = lab[2], l3 = lab[3];
        if (amb) {  // rare: a cell of this lane's four is within the fp32 error of a bisector
          l0 = exact_label(c.s_px, c.s_py, k0, k1, k2, 3, c.r0 + u, col + 0, res);
          l1 = exact_label(c.s_px, c.s_py, k0, k1, k2, 3, c.r0 + u, col + 1, res);
          l2 = exact_label(c.s_px, c.s_py, k0, k1, k2, 3, c.r0 + u, col + 2, res);
          l3 = exact_label(c.s_px, c.s_py, k0, k1, k2, 3, c.r0 + u, col + 3, res);
        }
      }
      if (LABELS && !SINGLE && c.labels && c.r0 + u < c.r1) {
        int *lp = c.labels + (size_t)(c.r0 + u) * c.N + col;
        if (col + 0 >= c.c0 && col + 0 < c.c1) lp[0] = c.s_out[ksel[l0]];
        if (col + 1 >= c.c0 && col + 1 < c.c1) lp[1] = c.s_out[ksel[l1]];
        if (col + 2 >= c.c0 && col + 2 < c.c1) lp[2] = c.s_out[ksel[l2]];
        if (col + 3 >= c.c0 && col + 3 < c.c1) lp[3] = c.s_out[ksel[l3]];
      }
      const double dv0 = (double)v0, dv1 = (double)v1, dv2 = (double)v2, dv3 = (double)v3;
      const double xd2 = xd * xd;
      if (!SINGLE) {
        Acc6 &t = acc[0];  // set 0: every cell of the region
        t.c0 += dv0, t.c1 += dv1, t.c2 += dv2, t.c3 += dv3;
        const double s = (dv0 + dv1) + (dv2 + dv3);
        t.sx = fma(xd, s, t.sx);
        t.sxx = fma(xd2, s, t.sxx);
#pragma unroll
        for (int k = 1; k < NC; ++k) {
          Acc6 &a = acc[k];
          double s_k = 0.0;
          if (l0 == k) a.c0 += dv0, s_k += dv0;
          if (l1 == k) a.c1 += dv1, s_k += dv1;
          if (l2 == k) a.c2 += dv2, s_k += dv2;
          if (l3 == k) a.c3 += dv3, s_k += dv3;
          a.sx = fma(xd, s_k, a.sx);
          a.sxx = fma(xd2, s_k, a.sxx);
        }
      } else {
        Acc6 &a = acc[0];
        double s_k = 0.0;
        if (l0 == 0) a.c0 += dv0, s_k += dv0;
        if (l1 == 0) a.c1 += dv1, s_k += dv1;
        if (l2 == 0) a.c2 += dv2, s_k += dv2;
        if (l3 == 0) a.c3 += dv3, s_k += dv3;
        a.sx = fma(xd, s_k, a.sx);
        a.sxx = fma(xd2, s_k, a.sxx);
      }
    }
    __syncwarp();  // every lane is done with this slot before it is refilled
  }

  Sum4 tot = fold_and_reduce(acc[0], col);
  if (SINGLE) {
    if (lane == 0) emit(c.sums, c.e, c.R, c.robot, tot.m, tot.sx, tot.sy, tot.s2);
    return;
  }
#pragma unroll
  for (int k = 1; k < NSETS; ++k) {
    const Sum4 part = fold_and_reduce(acc[k], col);
    if (lane == 0) emit(c.sums, c.e, c.R, c.s_out[ksel[k]], part.m, part.sx, part.sy, part.s2);
    tot.m -= part.m, tot.sx -= part.sx, tot.sy -= part.sy, tot.s2 -= part.s2;
  }
  if (lane == 0) emit(c.sums, c.e, c.R, c.s_out[k0], tot.m, tot.sx, tot.sy, tot.s2);
}

// direct (unstaged) loader for the generic path
__device__ __forceinline__ void load_cells_direct(const StripCtx &c, int row, int col, float v[4]) {
  const size_t off = (size_t)row * c.N + col;
  const float fill = (c.src == SRC_ROBOT) ? c.fill : 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[i] = 0.f;
    if (col + i < c.c0 || col + i >= c.c1) continue;
    bool visible = true;
    if (c.src == SRC_EXPLORED) visible = __ldg(c.sys + off + i) & 1;
    else if (c.src == SRC_ROBOT) visible = (__ldg(c.seen + off + i) >> (c.robot & 31)) & 1u;
    v[i] = visible ? (c.use_sensor ? __ldg(c.world + off + i) : 0.f) : fill;
  }
}

// Generic path: any number of candidates, fp64 rule per cell, per-lane label runs flushed with
// global atomics.  Only small regions that still contain more than VOR_MAXFAST owners come here.
// The context is handed over through shared memory so that its address never escapes in the caller
// (which would force the hot path's copy into local memory).
__device__ __noinline__ void region_slow(const StripCtx *cs, const int *cand, int nc, bool single, int lane) {
  const StripCtx c = *cs;
  const double res = c.res;
  const int col = c.col_base + 4 * (lane & 15);
  int cur = -1;
  double m = 0, sx = 0, sy = 0, s2 = 0;
  for (int row = c.r0 + (lane >> 4); row < c.r1; row += 2) {
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
      if (!single && c.labels) c.labels[(size_t)row * c.N + col + i] = slot;
      if (v[i] == 0.f) continue;
      if (slot != cur) {
        if (cur >= 0) {
          emit(c.sums, c.e, c.R, cur, m, sx, sy, s2);
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
    emit(c.sums, c.e, c.R, cur, m, sx, sy, s2);
  }
}

// Warp-level cull of the candidate list (lane k < nc holds candidate k) for the rectangle
// rows [r0, r1) x columns [c.c0, c.c1).  Two exact steps, both resting on h_k - h_j being affine:
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
  const double cx0 = cell_centre(r0, res), cx1 = cell_centre(r1 - 1, res);
  const double cy0 = cell_centre(c0, res), cy1 = cell_centre(c1 - 1, res);
  double f[4];
  f[0] = d2_exact(cx0, cy0, kx, ky), f[1] = d2_exact(cx1, cy0, kx, ky);
  f[2] = d2_exact(cx0, cy1, kx, ky), f[3] = d2_exact(cx1, cy1, kx, ky);
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
// The region is culled; if more than VOR_MAXFAST sites can own cells it is split in row halves and
// each half is culled again (fewer rows, fewer owners) before falling back to the generic path.
template <bool SINGLE, bool LABELS>
__device__ __forceinline__ void process_region(StripCtx &c, const int *s_cand, int nc, int *wcand, int lane,
                                               float4 *wbuf, unsigned *svis, StripCtx *sctx, unsigned bars,
                                               unsigned &phase) {
  if (nc > 32) {
    if (lane == 0) *sctx = c;
    __syncwarp();
    region_slow(sctx, s_cand, nc, SINGLE, lane);
    return;
  }
  const bool has = lane < nc;
  const int my_site = has ? s_cand[lane] : 0;
  const double kx = c.s_px[my_site], ky = c.s_py[my_site];
  const int region_r1 = c.r1;
  int r = c.r0;
  int len = region_r1 - r;
  while (r < region_r1) {
    if (len > region_r1 - r) len = region_r1 - r;
    const unsigned wmask = cull_rect(c.res, c.c0, c.c1, r, r + len, has, kx, ky, lane);
    const int wnc = __popc(wmask);
    if (wnc > VOR_MAXFAST && len > VOR_MINSPLIT) {
      len = ((len >> 1) + 1) & ~1;  // split (keep it even): retry the first half
      continue;
    }
    c.r0 = r, c.r1 = r + len;
    r += len;
    if (SINGLE && !(wmask & 1u)) continue;  // the robot owns no cell of these rows
    int ksite[VOR_MAXFAST];
    {
      unsigned mm = wmask;
#pragma unroll
      for (int q = 0; q < VOR_MAXFAST; ++q) {
        ksite[q] = 0;
        if (mm) {
          int src_lane = __ffs(mm) - 1;
          mm &= mm - 1;
          ksite[q] = __shfl_sync(0xffffffffu, my_site, src_lane);
        }
      }
    }
    if (wnc == 1) {
      region_fast<1, SINGLE, LABELS>(c, ksite[0], ksite[0], ksite[0], lane, wbuf, svis, bars, phase);
    } else if (wnc == 2) {
      region_fast<2, SINGLE, LABELS>(c, ksite[0], ksite[1], ksite[1], lane, wbuf, svis, bars, phase);
    } else if (wnc == 3) {
      region_fast<3, SINGLE, LABELS>(c, ksite[0], ksite[1], ksite[2], lane, wbuf, svis, bars, phase);
    } else {
      const bool keep = (wmask >> lane) & 1u;
      if (keep) wcand[__popc(wmask & ((1u << lane) - 1u))] = my_site;
      if (lane == 0) *sctx = c;
      __syncwarp();
      region_slow(sctx, wcand, wnc, SINGLE, lane);
    }
    __syncwarp();
  }
}

// dynamic shared memory: double px[R], py[R]; int out[R]; int cand[R]; int wcand[VOR_WARPS][32];
// then (16-byte aligned) per warp: a ring of VOR_NBUF world chunks [VOR_CH][32] float4 and the
// per-lane visibility word of each ring slot.
//
// SINGLE = false (global): grid = (tiles_x * tiles_y, 1, E); a CTA owns a (2 x 64)-column x
//   (4 x rpw)-row tile of the map, one region per warp; every robot is a site and every site's sums
//   are produced.
// SINGLE = true (window):  grid = (1, R, E); a CTA owns one robot's whole clipped L-window and its
//   warps loop over the window's regions; sites = the robot + its neighbours within communication
//   range (decentralized_cvt.h:100-120), only the robot's own cell is reduced.
constexpr size_t VOR_STAGE_BYTES_PER_WARP =
    (size_t)VOR_NBUF * VOR_CH * 32 * 16 + VOR_NBUF * 32 * sizeof(unsigned) + 128 + 32;  // ring, vis, ctx, mbarriers
static_assert(sizeof(StripCtx) <= 128, "per-warp context slot too small");

template <bool SINGLE, bool LABELS>
__global__ void __launch_bounds__(VOR_THREADS) k_voronoi(VorArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const EnvDev &d = a.d;
  const int ns_cap = d.R;
  double *s_px = reinterpret_cast<double *>(smem_raw);
  double *s_py = s_px + ns_cap;
  int *s_out = reinterpret_cast<int *>(s_py + ns_cap);
  int *s_cand = s_out + ns_cap;
  int *s_wcand = s_cand + ns_cap;
  const size_t stage_off = ((size_t)ns_cap * 24 + VOR_WARPS * 32 * sizeof(int) + 15) & ~(size_t)15;
  __shared__ int s_ns, s_nc;
  __shared__ double s_best[VOR_WARPS];
  __shared__ int s_besti[VOR_WARPS];
  __shared__ int s_woff[VOR_WARPS + 1];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float4 *wbuf = reinterpret_cast<float4 *>(smem_raw + stage_off + warp * VOR_STAGE_BYTES_PER_WARP);
  unsigned *svis = reinterpret_cast<unsigned *>(wbuf + VOR_NBUF * VOR_CH * 32);
  StripCtx *sctx = reinterpret_cast<StripCtx *>(svis + VOR_NBUF * 32);
  const unsigned bars = (unsigned)__cvta_generic_to_shared(reinterpret_cast<unsigned char *>(sctx) + 128);
  unsigned phase = 0;
  if (lane == 0) {
    for (int b = 0; b < VOR_NBUF; ++b) mbar_init(bars + 8 * b, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  }
  __syncwarp();
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
  StripCtx c;
  const size_t nn = (size_t)d.N * d.N;
  c.N = d.N, c.R = d.R, c.res = d.res, c.fill = d.fill, c.use_sensor = d.use_sensor;
  c.world = d.world + (size_t)e * nn;
  c.seen = d.seen + ((size_t)e * d.W + (robot >> 5)) * nn;
  c.sys = d.sys + (size_t)e * nn;
  c.e = e, c.robot = robot, c.src = a.src;
  c.s_px = s_px, c.s_py = s_py, c.s_out = s_out, c.sums = a.sums;
  c.labels = (LABELS && a.labels && !SINGLE) ? a.labels + (size_t)e * nn : nullptr;
  const int nrow = (tr1 - tr0 + a.rpw - 1) / a.rpw;
  const int ncol = (tc1 - 1 - col_base0) / VOR_TW + 1;
  for (int sidx = warp; sidx < nrow * ncol; sidx += VOR_WARPS) {
    const int ri = sidx / ncol, ci = sidx - ri * ncol;
    c.r0 = tr0 + ri * a.rpw;
    c.r1 = min(c.r0 + a.rpw, tr1);
    c.col_base = col_base0 + ci * VOR_TW;
    c.c0 = max(c.col_base, tc0);
    c.c1 = min(c.col_base + VOR_TW, tc1);
    if (c.c0 >= c.c1) continue;
    process_region<SINGLE, LABELS>(c, s_cand, nc, s_wcand + warp * 32, lane, wbuf, svis, sctx, bars, phase);
    __syncwarp();
  }
}

}  // namespace ccb
