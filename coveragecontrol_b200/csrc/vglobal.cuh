// vglobal.cuh -- whole-map Voronoi assignment + moment reduction organised by SITE instead of by tile.
//
// Replaces reference voronoi.cpp:55-327 for the global cells (ClairvoyantCVT, CentralizedCVT,
// CoverageSystem::ComputeVoronoiCells / GetObjectiveValue).  A Voronoi cell is convex, so on every map
// row it is ONE column interval.  Three kernels:
//
//   k_vg_partition  one thread per map row.  On row x the squared distances to the sites are parabolas
//                   in y of equal curvature, F_q(y) = (y - py_q)^2 + (x - px_q)^2; their lower envelope
//                   (the row's partition into owners) comes from the classic stack sweep over the sites
//                   sorted by py.  Every decision carries a rigorous bound of its fp64 rounding error:
//                   a hidden/visible decision, a breakpoint next to a cell centre, or a pair of parabolas
//                   that cannot be separated within that bound FLAGS the row instead of guessing.
//                   Output: table[e][site][row] = lo | hi << 16 (the site's cells on that row).
//   k_vg_items      one warp per environment: numbers the (site, 32-row chunk) pairs that hold cells.
//   k_vg_reduce     one warp per (site, 32-row chunk): streams the site's intervals with 16-byte loads
//                   (4 rows x 32 columns per pass, a batch of passes requested before the first is used;
//                   a lane asks for a pass iff its four cells touch its row's interval and masks only
//                   the passes that hold an end of it), no labels, no per-cell distance, one owner per
//                   warp: fp64 sums in cell units as in voronoi.cuh, one set of fp64 atomics per
//                   (site, chunk).  Every map cell is read once, by its owner.  SRC_EXPLORED ANDs the
//                   intervals with the explored bit plane.
//   k_vg_fixup      flagged rows only: the stated rule per cell (d2 in fp64 without FMA contraction,
//                   strict <, ascending site index), label runs flushed with atomics.
//
// A cell's label is therefore either decided by an envelope whose every comparison was outside its
// error bound (then it equals the stated rule's label) or by the stated rule itself.
// Restrictions (the caller falls back to k_voronoi otherwise): R <= VG_MAXR, N % 4 == 0, N < 65536.
#pragma once

#include "common.cuh"
#include "voronoi.cuh"

namespace ccb {

constexpr int VG_MAXR = 64;
constexpr int VG_PART_THREADS = 128;
constexpr int VG_RED_WARPS = 4;
constexpr int VG_CHUNK = 32;  // rows per warp item: lane <-> row when the intervals are fetched
constexpr int VG_U = 6;       // passes requested together (measured: 6 passes / 64 registers beats 8 / 72 and 4 / 48)

struct VgArgs {
  EnvDev d;
  int row0, row1;       // rows [row0, row1): the whole map or a row band
  uint32_t *table;      // [E][R][N]
  int *flag_list;       // flagged rows as e * N + row, *nflag entries (zeroed before k_vg_partition)
  int *nflag;
  int *span;            // [E][R][2]: first row / last row + 1 on which the site owns a cell (within the band)
  int *items;           // [E][R + 1]: first (site, 32-row chunk) item of every site, then the total (k_vg_items)
  double *sums;         // [E][R][4]
  int *labels;          // optional [E][N*N]
  const struct VgPfx *pfx;  // SUM mode: block prefix sums of the world maps, [E][N][N / 8 + 1]
  double *band;         // SUM mode: [E][bands][R][4] moments of every CTA's row band
  int pf;               // SUM mode: L2 prefetch distance of the walk in items (0: none)
};

template <bool SUM>
__host__ __device__ __forceinline__ void vg_partition_row(const VgArgs &a, int e, int row, bool active, int Rs, const int *s_ord,
                                                          const double *s_x, const double *s_y, const double *s_inv,
                                                          int *s_rmin, int *s_rmax, double *s_wacc);

// The row sweep also runs on the host (ccb_selftest_vg_partition, CPU tests): one thread, plain updates there.
__host__ __device__ __forceinline__ void vg_min(int *p, int v) {
#ifdef __CUDA_ARCH__
  atomicMin(p, v);
#else
  if (v < *p) *p = v;
#endif
}
__host__ __device__ __forceinline__ void vg_max(int *p, int v) {
#ifdef __CUDA_ARCH__
  atomicMax(p, v);
#else
  if (v > *p) *p = v;
#endif
}
__host__ __device__ __forceinline__ int vg_next(int *p) {
#ifdef __CUDA_ARCH__
  return atomicAdd(p, 1);
#else
  return (*p)++;
#endif
}

__global__ void k_vg_init_span(int *span, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) span[2 * i] = 0x7fffffff, span[2 * i + 1] = 0;
}

// ---- block prefix sums: the world-map moments without streaming the map ---------------------------------
// The world map of an environment is constant while its robots move, and a site's cells on a row are ONE column
// interval [lo, hi).  Next to every map row the library keeps the running sums of (v, y v, y^2 v) (y = col + 0.5,
// cell units) at every VG_PB-th column: entry b = the sum over the columns < VG_PB b, b = 0 .. N / VG_PB, one
// record each (struct VgPfx below).  The moments of a cell are then differences of entries plus the few cells
// between a boundary and its block edge, so a whole-map pass reads one record per Voronoi BOUNDARY of a row instead
// of every map cell (~1.1 MB of DRAM traffic instead of 4.2 MB per 1024^2 environment with 32 robots).
// Accuracy.  The entries are DOUBLE-DOUBLE (hi + lo, ~106 bits) built by a SEQUENTIAL scan
// E[b + 1] = E[b] (+) blocksum(b); the fused walk adds the partial block [8 b, T) of a boundary (a plain fp64 sum of
// at most 7 cells, like the block sums) to the entry in double-double too and takes P(T2) (-) P(T1): the sum of
// [T1, T2) up to ~1e-32 of the row's running sums and a few ulp of the two PARTIAL BLOCKS -- never an ulp of the
// running sums.  A cell of mass 1e-9 at the end of a row whose prefix is 1e3 keeps its digits, so the
// `mass < 1e-10` fallback (voronoi.cpp:41-53) fires on the same cells as with exact sums -- and a run of zero cells
// leaves P bitwise unchanged (the partial block is the scan's own expression with the other cells masked to zero), so
// it sums to exactly 0.  (k_vg_lookup adds only the cells inside the interval to a difference of entries.)
// Only for the world map (SRC_WORLD, no labels): the explored and robot maps change with every step.
constexpr int VG_PB = 8;  // columns per block: one 32-byte sector of the map row
// One 128-byte line per block: the running sums (48 bytes) and, at byte 64, the block's own 8 map values, so that a
// boundary costs ONE line of DRAM traffic (entry and cells come together; as two arrays every boundary pulled a
// 128-byte line of each).  16 bytes of HBM per map cell.
struct __align__(128) VgPfx {
  double s_hi, s_lo, sy_hi, sy_lo, syy_hi, syy_lo, pad0, pad1;
  float v[VG_PB];  // the map values of columns [VG_PB b, VG_PB b + VG_PB) (zeros in the row's last entry)
  float pad2[8];
};
static_assert(sizeof(VgPfx) == 128, "one line per block");
__host__ __device__ inline size_t vg_pfx_row_entries(int N) { return (size_t)N / VG_PB + 1; }

#ifdef __CUDACC__
// four cells y0 .. y0 + 3 (cell units) of one row into the row's sums
__device__ __forceinline__ void vg_add4(const float4 x, const double y0, double &rs, double &rsy, double &rsyy) {
  const double v0 = (double)x.x, v1 = (double)x.y, v2 = (double)x.z, v3 = (double)x.w;
  const double s = (v0 + v1) + (v2 + v3);
  const double t = fma(3.0, v3, fma(2.0, v2, v1));   // sum i v_i
  const double u = fma(9.0, v3, fma(4.0, v2, v1));   // sum i^2 v_i
  rs += s;
  rsy += fma(y0, s, t);
  rsyy += fma(y0 * y0, s, fma(y0 + y0, t, u));
}


struct DD {
  double hi, lo;
};
// error-free transformations (Knuth / Dekker); -fmad=false and the explicit _rn intrinsics keep them un-contracted
__device__ __forceinline__ DD dd_two_sum(double a, double b) {
  const double s = __dadd_rn(a, b), bb = __dsub_rn(s, a);
  return DD{s, __dadd_rn(__dsub_rn(a, __dsub_rn(s, bb)), __dsub_rn(b, bb))};
}
__device__ __forceinline__ DD dd_quick(double a, double b) {  // |a| >= |b|
  const double s = __dadd_rn(a, b);
  return DD{s, __dsub_rn(b, __dsub_rn(s, a))};
}
__device__ __forceinline__ DD dd_add(DD a, DD b) {
  DD s = dd_two_sum(a.hi, b.hi);
  const DD t = dd_two_sum(a.lo, b.lo);
  s = dd_quick(s.hi, __dadd_rn(s.lo, t.hi));
  return dd_quick(s.hi, __dadd_rn(s.lo, t.lo));
}
// a - b rounded to fp64: the high parts by an error-free subtraction, the low parts and its error added in fp64
// (|error| <= ~2^-104 |a|: the entries are running sums, the result is rounded to 53 bits anyway)
__device__ __forceinline__ double dd_diff(DD a, DD b) {
  const DD d = dd_two_sum(a.hi, -b.hi);
  return __dadd_rn(d.hi, __dadd_rn(d.lo, __dsub_rn(a.lo, b.lo)));
}

// eight cells y0 .. y0 + 7 (cell units) of one row, those with a bit in `keep` only
__device__ __forceinline__ void vg_add8(float4 a, float4 b, unsigned keep, double y0, double &rs, double &rsy, double &rsyy) {
  a.x = (keep & 1u) ? a.x : 0.f, a.y = (keep & 2u) ? a.y : 0.f, a.z = (keep & 4u) ? a.z : 0.f, a.w = (keep & 8u) ? a.w : 0.f;
  b.x = (keep & 16u) ? b.x : 0.f, b.y = (keep & 32u) ? b.y : 0.f, b.z = (keep & 64u) ? b.z : 0.f, b.w = (keep & 128u) ? b.w : 0.f;
  vg_add4(a, y0, rs, rsy, rsyy);
  vg_add4(b, y0 + 4.0, rs, rsy, rsyy);
}

// 256-bit read-only loads (sm_100: LDG.E.256): a prefix entry is one of these plus a 128-bit load, a map sector one
struct VgEntry {
  DD s, sy, syy;
};
__device__ __forceinline__ VgEntry vg_entry_load(const VgPfx *p) {
  VgEntry e;
  asm volatile("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(e.s.hi), "=d"(e.s.lo), "=d"(e.sy.hi), "=d"(e.sy.lo) : "l"(p));
  const double2 t = __ldg(reinterpret_cast<const double2 *>(p) + 2);
  e.syy.hi = t.x, e.syy.lo = t.y;
  return e;
}
__device__ __forceinline__ void vg_sector_load(const float *p, float4 &a, float4 &b) {  // 8 floats, 32-byte aligned
  asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
               : "l"(p));
}

// E[b + 1] from E[b] and the eight cells of block b: THE step of the sequential scan (k_vg_prefix); the partition
// kernel repeats it to get the entry after a boundary without a second look-up, bit for bit
__device__ __forceinline__ void vg_pfx_step(DD &s, DD &sy, DD &syy, float4 a, float4 b, int blk) {
  double bs = 0.0, bsy = 0.0, bsyy = 0.0;
  vg_add8(a, b, 0xffu, (double)(VG_PB * blk) + 0.5, bs, bsy, bsyy);
  s = dd_add(s, DD{bs, 0.0}), sy = dd_add(sy, DD{bsy, 0.0}), syy = dd_add(syy, DD{bsyy, 0.0});
}
#endif  // __CUDACC__

// RPT rows per thread: a CTA's band is RPT * VG_PART_THREADS rows (RPT > 1, opt-in through CCB_VG_PART_RPT: the
// sort / twin / cull / 1 / (2 dy) preamble is paid once per band, against a taller band culling fewer sites)
constexpr int VG_LMAX = 16;  // owners of one row the fused walk lists (rows with more go to k_vg_fixup)
struct VgRowLists {           // shared memory of k_vg_partition<1, true>: the rows' owner lists, then the accumulators
  double wacc[VG_PART_THREADS / 32][VG_MAXR][4];
  unsigned short tc[VG_LMAX + 1][VG_PART_THREADS];
  unsigned char slot[VG_LMAX][VG_PART_THREADS];
  unsigned char nown[VG_PART_THREADS];
  int maxown;
};

// SUM (world-map moments with valid block prefix sums, RPT = 1): instead of writing the interval table for a second
// kernel, the row threads leave their owners and breakpoints in shared memory and the CTA walks them as ITEMS (m-th
// owner of 32 consecutive rows): a lane turns its row's two breakpoints into moments -- one prefix entry and one map
// sector per boundary (the entries at the block edges just inside the interval, the cells between the boundaries
// and those edges), all loads of an item independent -- and a segmented warp reduction (by site, fixed shuffle
// tree: neighbouring rows mostly share their m-th owner) adds the 32 rows into per-warp accumulators; the CTA adds
// its four warps in order and STORES the band's moments; k_vg_bandsum then adds the bands in order.  No table, no
// spans, no item list, no atomics: run-to-run bit-stable (flagged rows still go through k_vg_fixup's atomics).
template <int RPT = 1, bool SUM = false>
__global__ void __launch_bounds__(VG_PART_THREADS, SUM ? 6 : 1) k_vg_partition(VgArgs a) {
  __shared__ int s_ord[VG_MAXR];
  __shared__ double s_x[VG_MAXR], s_y[VG_MAXR];
  __shared__ int s_rmin[VG_MAXR], s_rmax[VG_MAXR];
  extern __shared__ double s_inv[];  // [R][R]: 1 / (2 (y_q - y_p)) of the sorted sites, shared by the CTA's rows
  const EnvDev &d = a.d;
  const int e = blockIdx.y, tid = threadIdx.x, R = d.R, N = d.N;
  const double *pos = d.pos_g + (size_t)e * R * 2;
  // sites sorted by (py, index): rank by counting
  __shared__ int t_ord[VG_MAXR];
  __shared__ double t_x[VG_MAXR], t_y[VG_MAXR];
  __shared__ unsigned char t_keep[VG_MAXR];
  __shared__ int s_nkeep;
  __shared__ __align__(16) unsigned char s_lists_raw[SUM ? sizeof(VgRowLists) : 16];
  double *s_wacc = reinterpret_cast<double *>(s_lists_raw);  // (VgRowLists::wacc comes first)
  if (SUM) {
    for (int k = tid; k < VG_PART_THREADS / 32 * VG_MAXR * 4; k += VG_PART_THREADS) s_wacc[k] = 0.0;
    if (tid == 0) reinterpret_cast<VgRowLists *>(s_lists_raw)->maxown = 0;
  }
  for (int k = tid; k < R; k += VG_PART_THREADS) {
    const double yk = pos[2 * k + 1];
    int rank = 0;
    for (int j = 0; j < R; ++j) {
      const double yj = pos[2 * j + 1];
      rank += (yj < yk || (yj == yk && j < k)) ? 1 : 0;
    }
    t_ord[rank] = k, t_x[rank] = pos[2 * k], t_y[rank] = yk;
    s_rmin[k] = 0x7fffffff, s_rmax[k] = 0;
  }
  __syncthreads();
  // Band cull: d2_q - d2_p is affine in the cell centre, so a site q that is farther than one and the
  // same site p at the four corner cells of this CTA's row band owns no cell of the band: the sweeps
  // below run over the ~10 sites whose cells reach the band instead of all R.
  {
    const double res = d.res;
    const int rlo = a.row0 + blockIdx.x * (RPT * VG_PART_THREADS), rhi = min(rlo + RPT * VG_PART_THREADS, a.row1) - 1;
    const double cxs[2] = {cell_centre(rlo, res), cell_centre(rhi, res)};
    const double cys[2] = {cell_centre(0, res), cell_centre(N - 1, res)};
    // thread (q, part): site q against every parts-th site p
    const int parts = max(1, VG_PART_THREADS / R);
    // a site that coincides with a lower-indexed one never wins under the stated rule (equal d2, strict <):
    // dropped here, otherwise every row the pair reaches would be flagged as inseparable
    for (int q = tid; q < R; q += VG_PART_THREADS) {
      bool dup = false;
      for (int p = 0; p < R; ++p) dup = dup || (t_ord[p] < t_ord[q] && t_x[p] == t_x[q] && t_y[p] == t_y[q]);
      t_keep[q] = dup ? 0 : 1;
    }
    __syncthreads();
    for (int t = tid; t < R * parts; t += VG_PART_THREADS) {
      const int q = t % R, part = t / R;
      double fq[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) fq[c] = d2_exact(cxs[c & 1], cys[c >> 1], t_x[q], t_y[q]);
      bool dominated = false;
      for (int p = part; p < R && !dominated; p += parts) {
        if (p == q || (t_x[p] == t_x[q] && t_y[p] == t_y[q])) continue;  // a twin dominates nothing
        bool all = true;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const double fp = d2_exact(cxs[c & 1], cys[c >> 1], t_x[p], t_y[p]);
          all = all && (fq[c] - fp > 1e-7 * (1.0 + fp));
        }
        dominated = all;
      }
      if (dominated) t_keep[q] = 0;
    }
  }
  __syncthreads();
  if (tid == 0) {  // compact, keeping the sorted order
    int n = 0;
    for (int q = 0; q < R; ++q)
      if (t_keep[q]) s_ord[n] = t_ord[q], s_x[n] = t_x[q], s_y[n] = t_y[q], ++n;
    s_nkeep = n;
  }
  __syncthreads();
  const int Rs = s_nkeep;
  for (int idx = tid; idx < Rs * Rs; idx += VG_PART_THREADS) {
    const int p = idx / Rs, q = idx - p * Rs;
    const double dy = s_y[q] - s_y[p];
    s_inv[idx] = dy > 0.0 ? 1.0 / (2.0 * dy) : 0.0;
  }
  __syncthreads();
  if (SUM) {
    VgRowLists &L = *reinterpret_cast<VgRowLists *>(s_lists_raw);
    const int row_t = a.row0 + blockIdx.x * VG_PART_THREADS + tid;
    vg_partition_row<true>(a, e, min(row_t, a.row1 - 1), row_t < a.row1, Rs, s_ord, s_x, s_y, s_inv, s_rmin, s_rmax, s_wacc);
    // ---- the walk: item = the m-th owner of the warp's own 32 rows; lane <-> row.  Every warp walks the lists its
    //      own threads wrote, so it starts as soon as ITS rows are swept (no CTA barrier between sweep and walk).
    const int lane = tid & 31, warp = tid >> 5;
    const int t = tid;
    const int nown_t = (int)L.nown[t];
    const int nitems = __reduce_max_sync(0xffffffffu, nown_t);
    const size_t pe = vg_pfx_row_entries(N);
    const int row = a.row0 + blockIdx.x * VG_PART_THREADS + t;
    const VgPfx *prow = a.pfx + ((size_t)e * N + min(row, a.row1 - 1)) * pe;
    // Item m = [T_m, T_{m+1}) of the lane's row.  The row's owners partition [0, N), so the lane carries
    // P(T) = the row's running sums up to its last boundary -- the prefix entry of T's block (+) the cells between
    // the block edge and T, double-double -- and an item is P(T_{m+1}) (-) P(T_m): ONE entry and ONE map sector per
    // boundary (the boundary between two owners is loaded once, not once per side).
    auto request = [&](int m) {
      if (m < nown_t) {
        const int T = L.tc[m + 1][t], b = T >> 3;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(prow + b));  // the entry and the block's cells: one line
      }
    };
    const int pf = a.pf;
    for (int k = 0; k < pf; ++k) request(k);
    DD ps{0.0, 0.0}, psy{0.0, 0.0}, psyy{0.0, 0.0};  // P(T_m); T_0 = 0
#pragma unroll 1
    for (int m = 0; m < nitems; ++m) {
      if (pf) request(m + pf);  // pf items ahead are on their way to L2 while this one is worked on
      int slot = -1;
      double vm = 0.0, vsx = 0.0, vsy = 0.0, vs2 = 0.0;
      if (m < nown_t) {
        const int T = L.tc[m + 1][t], b = T >> 3, f = T & (VG_PB - 1);
        const VgEntry er = vg_entry_load(prow + b);
        DD ns = er.s, nsy = er.sy, nsyy = er.syy;
        if (f) {
          float4 RA, RB;
          vg_sector_load(prow[b].v, RA, RB);
          double hs = 0.0, hsy = 0.0, hsyy = 0.0;
          vg_add8(RA, RB, (1u << f) - 1u, (double)(VG_PB * b) + 0.5, hs, hsy, hsyy);  // [8 b, T)
          ns = dd_add(ns, DD{hs, 0.0}), nsy = dd_add(nsy, DD{hsy, 0.0}), nsyy = dd_add(nsyy, DD{hsyy, 0.0});
        }
        const double rs = dd_diff(ns, ps), rsy = dd_diff(nsy, psy), rsyy = dd_diff(nsyy, psyy);
        ps = ns, psy = nsy, psyy = nsyy;
        const double x = (double)row + 0.5;
        vm = rs, vsx = x * rs, vsy = rsy, vs2 = fma(x * x, rs, rsyy);
        slot = L.slot[m][t];
      }
      unsigned todo = __ballot_sync(0xffffffffu, slot >= 0);
      while (todo) {
        const int leader = __ffs(todo) - 1;
        const int sl = __shfl_sync(0xffffffffu, slot, leader);
        const bool mine = slot == sl;
        const Sum4 tot = reduce4(mine ? vm : 0.0, mine ? vsx : 0.0, mine ? vsy : 0.0, mine ? vs2 : 0.0);
        if (lane == leader) {
          double *acc = &L.wacc[warp][sl][0];
          acc[0] += tot.m, acc[1] += tot.sx, acc[2] += tot.sy, acc[3] += tot.s2;
        }
        __syncwarp();  // the next leader (another lane) may add to the same accumulator
        todo &= ~__ballot_sync(0xffffffffu, mine);
      }
    }
    __syncthreads();
    // band moments = the four warps' accumulators in order; sites without cells in the band store zeros
    double *dst = a.band + ((size_t)e * gridDim.x + blockIdx.x) * R * 4;
    for (int i = tid; i < R * 4; i += VG_PART_THREADS) {
      double tsum = 0.0;
#pragma unroll
      for (int wq = 0; wq < VG_PART_THREADS / 32; ++wq) tsum += s_wacc[wq * VG_MAXR * 4 + i];
      dst[i] = tsum;
    }
    return;
  }
  if (RPT == 1) {
    const int row = a.row0 + blockIdx.x * VG_PART_THREADS + tid;
    if (row < a.row1) vg_partition_row<false>(a, e, row, true, Rs, s_ord, s_x, s_y, s_inv, s_rmin, s_rmax, nullptr);
  } else {
#pragma unroll 1
    for (int i = 0; i < RPT; ++i) {
      const int row = a.row0 + blockIdx.x * (RPT * VG_PART_THREADS) + i * VG_PART_THREADS + tid;
      if (row < a.row1) vg_partition_row<false>(a, e, row, true, Rs, s_ord, s_x, s_y, s_inv, s_rmin, s_rmax, nullptr);
    }
  }
  __syncthreads();
  // the CTA's rows extend every site's row span
  for (int k = tid; k < R; k += VG_PART_THREADS)
    if (s_rmax[k] > s_rmin[k]) {
      atomicMin(a.span + ((size_t)e * R + k) * 2, s_rmin[k]);
      atomicMax(a.span + ((size_t)e * R + k) * 2 + 1, s_rmax[k]);
    }
}

// Rs sorted sites survive the band cull (s_ord: their site indices); the table is written for all R
template <bool SUM>
__host__ __device__ __forceinline__ void vg_partition_row(const VgArgs &a, int e, int row, bool active, int Rs, const int *s_ord,
                                                          const double *s_x, const double *s_y, const double *s_inv,
                                                          int *s_rmin, int *s_rmax, double *s_wacc) {
  const EnvDev &d = a.d;
  const int R = d.R, N = d.N;
  const double res = d.res;
  const double cx = cell_centre(row, res);
  const double y_first = cell_centre(0, res), y_last = cell_centre(N - 1, res);

  unsigned char v[VG_MAXR];  // envelope: sorted positions of the owners, left to right (bytes: the stacks live in
                             // local memory and every resident warp's share of L1 counts)
  double z[VG_MAXR];       // z[m]: boundary between v[m-1] and v[m]
  double zerr[VG_MAXR];    // bound of its rounding error
  bool flagged = false;
  int top = 0;
  v[0] = 0, z[0] = -INFINITY, zerr[0] = 0.0;
  for (int q = 1; q < Rs; ++q) {
    const double yq = s_y[q], dxq = cx - s_x[q];
    while (true) {
      const int p = v[top];
      const double yp = s_y[p], dxp = cx - s_x[p];
      const double dy = yq - yp;                     // >= 0 (sorted)
      const double num = (dxq - dxp) * (dxq + dxp);  // dxq^2 - dxp^2
      const double ysum = yq + yp;
      // F_q(y) - F_p(y) = num + dy (ysum - 2 y): decreasing in y
      const double scale = 1.0 + fabs(num) + (fabs(ysum) + 2.0 * y_last) * dy + dxq * dxq + dxp * dxp;
      const double eps_h = 1e-12 * scale;
      if (2.0 * dy * (y_last - y_first) <= eps_h) {
        // the two parabolas differ by a constant within rounding on this row
        const double d0 = num + dy * (ysum - 2.0 * y_first), d1 = num + dy * (ysum - 2.0 * y_last);
        if (d0 > eps_h && d1 > eps_h) break;  // q is above p everywhere: q owns nothing
        if (d0 < -eps_h && d1 < -eps_h) {     // p is above q everywhere: drop p
          if (top == 0) {
            v[0] = q;
            break;
          }
          --top;
          continue;
        }
        flagged = true;
        break;
      }
      const double inv = s_inv[p * Rs + q];
      const double s = fma(num, inv, 0.5 * ysum);  // q is nearer than p for y > s
      // bound of the rounding error of s, plus the distance (in y) within which the rounded d2 of the
      // stated rule may order p and q differently from exact arithmetic: 1e-14 is ~45 ulp
      const double fsum = dxp * dxp + dxq * dxq + (s - yp) * (s - yp) + (s - yq) * (s - yq);
      const double serr = 1e-14 * (fabs(0.5 * ysum) + res + (fabs(num) + fsum) * inv);
      const double slack = serr + zerr[top];
      if (s <= z[top] - slack) {  // p never owns anything: hidden between v[top-1] and q
        --top;                    // (top >= 1 here: z[0] = -inf)
        continue;
      }
      if (s < z[top] + slack) {   // cannot tell hidden from visible within rounding
        flagged = true;
        --top;
        continue;
      }
      ++top;
      v[top] = q, z[top] = s, zerr[top] = serr;
      break;
    }
  }

  // integer breakpoints: v[m] owns columns [tc[m], tc[m + 1]); a cell centre within the error bound of a
  // boundary flags the row
  unsigned short tc[VG_MAXR + 1];
  tc[0] = 0;
  {
    int lo = 0;
    for (int m = 0; m <= top && !flagged; ++m) {
      int hi = N;
      if (m < top) {
        const double b = z[m + 1], berr = zerr[m + 1] + 1e-13 * (fabs(b) + res);
        // first column whose centre is >= b (at the default resolution 1 the quotient is b itself: no fp64 division)
        const double t = ceil((res == 1.0 ? b : b / res) - 0.5);
        const double tcl = fmin(fmax(t, 0.0), (double)N);
        hi = (int)tcl;
        // centres of the cells next to the boundary
        if (t >= 1.0 && t <= (double)N && fabs(cell_centre((int)t - 1, res) - b) <= berr) flagged = true;
        if (t >= 0.0 && t < (double)N && fabs(cell_centre((int)t, res) - b) <= berr) flagged = true;
      }
      if (hi < lo) hi = lo;
      tc[m + 1] = (unsigned short)hi;
      lo = hi;
    }
  }
  if (SUM && top + 1 > VG_LMAX) flagged = true;  // more owners than the walk's lists hold: the per-cell rule takes the row
  // a flagged row is left to k_vg_fixup entirely
  if (flagged && active) a.flag_list[vg_next(a.nflag)] = e * N + row;
  if (!SUM) {
    uint32_t *t = a.table + (size_t)e * R * N + row;
    unsigned long long owned = 0ull;  // sites whose interval has been written (R <= VG_MAXR = 64)
    if (!flagged)
      for (int m = 0; m <= top; ++m) {
        const int lo = tc[m], hi = tc[m + 1];
        const int site = s_ord[v[m]];
        t[(size_t)site * N] = (uint32_t)lo | ((uint32_t)hi << 16);
        owned |= 1ull << site;
        if (hi > lo) vg_min(&s_rmin[site], row), vg_max(&s_rmax[site], row + 1);
      }
    // everyone else owns nothing on this row
    for (int k = 0; k < R; ++k)
      if (!((owned >> k) & 1ull)) t[(size_t)k * N] = 0u;
    return;
  }
#ifdef __CUDA_ARCH__
  if (SUM) {
    // the row's owners and breakpoints for the CTA's walk (k_vg_partition): m-major, a row per thread
    VgRowLists &L = *reinterpret_cast<VgRowLists *>(s_wacc);
    const int t = threadIdx.x;
    const int nown = (active && !flagged) ? top + 1 : 0;
    L.nown[t] = (unsigned char)nown;
    for (int m = 0; m < nown; ++m) L.tc[m][t] = tc[m], L.slot[m][t] = (unsigned char)s_ord[v[m]];
    if (nown > 0) L.tc[nown][t] = tc[nown];
  }
#endif
}

// ---- reduce ----------------------------------------------------------------------------------------
// Items = (site, 32-row chunk) pairs on which the site owns cells, numbered site by site from the row
// spans k_vg_partition left.  items[e][k] = number of the first item of site k, items[e][R] = their total:
// one warp per environment (an exclusive prefix sum over <= VG_MAXR sites).
__global__ void k_vg_items(const int *__restrict__ span, int *__restrict__ items, int E, int R, int row0) {
  const int e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (e >= E) return;
  int before = 0;
  for (int kb = 0; kb < R; kb += 32) {
    const int kk = kb + lane;
    int cnt = 0;
    if (kk < R) {
      const int2 sp = __ldg(reinterpret_cast<const int2 *>(span) + (size_t)e * R + kk);
      if (sp.y > sp.x) cnt = (sp.y - 1 - row0) / VG_CHUNK - (sp.x - row0) / VG_CHUNK + 1;
    }
    int incl = cnt;  // inclusive prefix sum over the lanes
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += up;
    }
    if (kk < R) items[(size_t)e * (R + 1) + kk] = before + incl - cnt;
    before += __shfl_sync(0xffffffffu, incl, 31);
  }
  if (lane == 0) items[(size_t)e * (R + 1) + R] = before;
}

// keep bits of a lane's four cells col .. col + 3 against [lo, hi): bit i set iff lo <= col + i < hi
__device__ __forceinline__ unsigned vg_keep4(int col, int lo, int hi) {
  const int a = min(max(lo - col, 0), 4), b = min(max(hi - col, 0), 4);
  return ((1u << b) - 1u) & ~((1u << a) - 1u);
}

constexpr int VG_RED_MINB = 8;  // resident CTAs of VG_RED_WARPS warps per SM the register allocation aims for (64 registers)
// LPR lanes share a row: a pass is 32 / LPR rows x 4 LPR columns (512 bytes either way).  The instantiations
// without labels run with LPR = 4 (8 rows x 16 columns): half as many row groups per item and less of a pass
// lost at the ends of an interval than with 4 rows x 32 columns (map: 0.93 -> 0.85 ms, explored map: -16 %).
template <int SRC, bool LABELS, int U = VG_U, int MINB = VG_RED_MINB, int LPR = 8>
__global__ void __launch_bounds__(VG_RED_WARPS * 32, MINB) k_vg_reduce(VgArgs a) {
  const EnvDev &d = a.d;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int e = blockIdx.y;
  const int N = d.N, R = d.R;
  const int *its = a.items + (size_t)e * (R + 1);
  const int total = __ldg(its + R);
  const float *world = d.world + (size_t)e * N * N;
  const uint32_t *plane = SRC == SRC_EXPLORED ? d.sysb + ((size_t)e * 2) * plane_words(d) : nullptr;
  int *labels = (LABELS && a.labels) ? a.labels + (size_t)e * N * N : nullptr;
  constexpr int P = 4 * LPR, RQ = 32 / LPR, SH = LPR == 8 ? 5 : (LPR == 4 ? 4 : 3);
  static_assert((1 << SH) == P, "pass width");
  const int g = lane / LPR, c8 = lane % LPR;
  // the grid is a fraction of the largest possible item count: a warp strides over the items
#pragma unroll 1
  for (int item = blockIdx.x * (blockDim.x >> 5) + warp; item < total; item += gridDim.x * (blockDim.x >> 5)) {
    // the item's site: the first one whose items end beyond `item` (sites without cells have none)
    int k;
    {
      const int end0 = lane < R ? __ldg(its + lane + 1) : 0x7fffffff;
      unsigned hit = __ballot_sync(0xffffffffu, item < end0);
      if (hit) {
        k = __ffs(hit) - 1;
      } else {
        const int end1 = lane + 32 < R ? __ldg(its + lane + 33) : 0x7fffffff;
        hit = __ballot_sync(0xffffffffu, item < end1);
        k = 32 + __ffs(hit) - 1;
      }
    }
    const int chunk = (__ldg(a.span + ((size_t)e * R + k) * 2) - a.row0) / VG_CHUNK + (item - __ldg(its + k));
    const int rbase = a.row0 + chunk * VG_CHUNK;
    const int myrow = rbase + lane;
    const uint32_t iv = myrow < a.row1 ? __ldg(a.table + ((size_t)e * R + k) * N + myrow) : 0u;
    const int lo_l = (int)(iv & 0xffffu), hi_l = (int)(iv >> 16);
    // ask for this lane's row interval to be brought into L2 now (a line per 32 columns): the whole
    // chunk is in flight while the warp streams it, and the loads below see L2 latency
    if (SRC == SRC_WORLD) {
      const float *prow = world + (size_t)myrow * N;
      for (int c = lo_l & ~31; c < hi_l; c += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(prow + c));
    }
    double M = 0.0, SX = 0.0, SY = 0.0, S2 = 0.0;
#pragma unroll 1
    for (int q = 0; q < VG_CHUNK / RQ; ++q) {
      const int lo = __shfl_sync(0xffffffffu, lo_l, RQ * q + g), hi = __shfl_sync(0xffffffffu, hi_l, RQ * q + g);
      const bool has = hi > lo;
      const int qlo = __reduce_min_sync(0xffffffffu, has ? lo : 0x7fffffff), qhi = __reduce_max_sync(0xffffffffu, has ? hi : 0);
      if (qhi <= qlo) continue;  // warp-uniform
      const int row = rbase + RQ * q + g;
      // pass j covers columns cb0 + 32 j ..; this lane's cells: col0 + 32 j .. + 3
      const int cb0 = qlo & ~(P - 1), npass = (qhi - cb0 + P - 1) >> SH;
      const int col0 = cb0 + 4 * c8;
      const float4 *p4 = reinterpret_cast<const float4 *>(world + (size_t)row * N + col0);
      const uint32_t *pw = SRC == SRC_EXPLORED ? plane + (size_t)row * d.NW : nullptr;  // the row's explored bits
      int *lrow = LABELS && labels ? labels + (size_t)row * N + col0 : nullptr;
      // the interval relative to the lane's first cell (an empty one for a row without cells of the site)
      const int A = has ? lo - col0 : 0, B = has ? hi - col0 : 0;
      // Pass j holds the lane's cells 32 j .. 32 j + 3.  They touch [A, B) for jl0 <= j <= jl1 (the passes
      // the lane requests) and lie inside it for jf0 <= j <= jf1 (added without a mask); only the passes
      // holding A or B in between need one.
      const int jl0 = (A + P - 4) >> SH, jl1 = (B - 1) >> SH;
      const int jf0 = A > 0 ? (A + P - 1) >> SH : 0, jf1 = (B - 4) >> SH;
      double rs = 0.0, rsy = 0.0, rsyy = 0.0;
#pragma unroll 1
      for (int base = 0; base < npass; base += U) {
        // request the whole batch, then consume it: U x 512 bytes in flight per warp
        float4 w[U];
        const int Ab = A - P * base, Bb = B - P * base;  // the interval relative to the batch's first pass
        const int l0 = jl0 - base, l1 = jl1 - base, f0 = jf0 - base, f1 = jf1 - base;
        unsigned kl = 0u, kp = 0u;  // LABELS / SRC_EXPLORED: 4 keep bits per pass (kl: inside the interval, kp: to add)
        if (LABELS || SRC == SRC_EXPLORED) {
#pragma unroll
          for (int j = 0; j < U; ++j) kl |= vg_keep4(P * j, Ab, Bb) << (4 * j);
          kp = kl;
          if (SRC == SRC_EXPLORED) {
            uint32_t pwv[U];
#pragma unroll
            for (int j = 0; j < U; ++j) pwv[j] = (((kl >> (4 * j)) & 0xfu) && base + j < npass) ? __ldg(pw + ((col0 + P * (base + j)) >> 5)) : 0u;
#pragma unroll
            for (int j = 0; j < U; ++j) kp &= ~(0xfu << (4 * j)) | (((pwv[j] >> ((col0 + P * (base + j)) & 31)) & 0xfu) << (4 * j));
          }
        }
#pragma unroll
        for (int j = 0; j < U; ++j) {
          // (jl1 < npass: no pass beyond the row group's last one is asked for)
          const bool want = (LABELS || SRC == SRC_EXPLORED) ? ((kp >> (4 * j)) & 0xfu) != 0u && base + j < npass : (j >= l0 && j <= l1);
          if (want) w[j] = __ldcs(p4 + LPR * (base + j));  // read once: streaming (evict-first) beats __ldg by 2 %
        }
        double y0 = (double)(col0 + P * base) + 0.5;
#pragma unroll
        for (int j = 0; j < U; ++j) {
          if (base + j >= npass) break;  // warp-uniform
          if (LABELS) {
            if (lrow) {
#pragma unroll
              for (int i = 0; i < 4; ++i)
                if ((kl >> (4 * j + i)) & 1u) lrow[P * (base + j) + i] = k;
            }
          }
          // w[j] is only defined where the lane asked for it
          if ((LABELS || SRC == SRC_EXPLORED) || j < f0 || j > f1) {  // an end of the interval, outside it, unexplored cells
            const unsigned keep = (LABELS || SRC == SRC_EXPLORED) ? (kp >> (4 * j)) & 0xfu : vg_keep4(P * j, Ab, Bb);
            w[j].x = (keep & 1u) ? w[j].x : 0.f, w[j].y = (keep & 2u) ? w[j].y : 0.f;
            w[j].z = (keep & 4u) ? w[j].z : 0.f, w[j].w = (keep & 8u) ? w[j].w : 0.f;
          }
          vg_add4(w[j], y0, rs, rsy, rsyy);
          y0 += (double)P;
        }
      }
      const double x = (double)row + 0.5;
      M += rs;
      SX = fma(x, rs, SX);
      SY += rsy;
      S2 += fma(x * x, rs, rsyy);
    }
    const Sum4 tot = reduce4(M, SX, SY, S2);
    if (lane == 0) emit(a.sums, e, d.R, k, tot.m, tot.sx, tot.sy, tot.s2);
  }
}

// ---- A/B: the same reduce with bulk-asynchronous copies (TMA engine) instead of per-lane loads ---------------
// k_vg_reduce_bulk<SRC_WORLD>: one warp per (site, 32-row chunk) item as above, but a row's interval travels as ONE
// `cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes` per (row, 256-column piece) into a per-warp
// ring of shared-memory stages (8 rows x 1 KB each, two stages): the copies of stage s + 1 are issued before stage s
// is consumed, an mbarrier per stage counts the bytes in, and the lanes then read their cells from shared memory
// (row stride padded by 16 bytes: conflict-free).  No per-lane address arithmetic for the loads, no L2 prefetch loop,
// no `want` predicates -- what VERDICT r01 (next 6) asked to try.  Selected with ccb_test_set_option("vg_bulk", 1);
// profiles/r02_tma_ab.md has the timing against k_vg_reduce and the SASS evidence (UBLKCP).
constexpr int VGB_ROWS = 8;                        // rows per stage (4 lanes per row when consumed)
constexpr int VGB_COLS = 256;                      // columns per piece
constexpr int VGB_STRIDE = VGB_COLS * 4 + 16;      // bytes between the rows of a stage
constexpr int VGB_STAGE = VGB_ROWS * VGB_STRIDE;   // 8.1 KB: 65 KB per CTA of four warps, three CTAs per SM
constexpr int VGB_STAGES = 2;
constexpr int VGB_WARPS = 4;
constexpr size_t VGB_SMEM = (size_t)VGB_WARPS * VGB_STAGES * VGB_STAGE + 16 * VGB_WARPS * VGB_STAGES;

__device__ __forceinline__ uint32_t vgb_smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void vgb_wait(uint32_t mbar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "VGB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra VGB_DONE;\n"
      "bra VGB_WAIT;\n"
      "VGB_DONE:\n"
      "}\n" ::"r"(mbar), "r"(parity)
      : "memory");
}

__global__ void __launch_bounds__(VGB_WARPS * 32) k_vg_reduce_bulk(VgArgs a) {
  extern __shared__ __align__(128) unsigned char vgb_smem[];
  const EnvDev &d = a.d;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int e = blockIdx.y;
  const int N = d.N, R = d.R;
  const int *its = a.items + (size_t)e * (R + 1);
  const int total = __ldg(its + R);
  const float *world = d.world + (size_t)e * N * N;
  unsigned char *stage_base = vgb_smem + (size_t)warp * VGB_STAGES * VGB_STAGE;
  const uint32_t mbar0 = vgb_smem_addr(vgb_smem + (size_t)VGB_WARPS * VGB_STAGES * VGB_STAGE + 16 * VGB_STAGES * warp);
  if (lane == 0) {
#pragma unroll
    for (int sidx = 0; sidx < VGB_STAGES; ++sidx)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar0 + 16 * sidx), "r"(VGB_ROWS));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  uint32_t phase[VGB_STAGES] = {0u, 0u};
  const int g = lane >> 2, c4 = lane & 3;  // consumer mapping: 4 lanes per row
#pragma unroll 1
  for (int item = blockIdx.x * VGB_WARPS + warp; item < total; item += gridDim.x * VGB_WARPS) {
    int k;
    {
      const int end0 = lane < R ? __ldg(its + lane + 1) : 0x7fffffff;
      unsigned hit = __ballot_sync(0xffffffffu, item < end0);
      if (hit) {
        k = __ffs(hit) - 1;
      } else {
        const int end1 = lane + 32 < R ? __ldg(its + lane + 33) : 0x7fffffff;
        hit = __ballot_sync(0xffffffffu, item < end1);
        k = 32 + __ffs(hit) - 1;
      }
    }
    const int chunk = (__ldg(a.span + ((size_t)e * R + k) * 2) - a.row0) / VG_CHUNK + (item - __ldg(its + k));
    const int rbase = a.row0 + chunk * VG_CHUNK;
    const int myrow = rbase + lane;
    const uint32_t iv = myrow < a.row1 ? __ldg(a.table + ((size_t)e * R + k) * N + myrow) : 0u;
    const int lo_l = (int)(iv & 0xffffu), hi_l = (int)(iv >> 16);
    // the item's column range, in pieces of VGB_COLS columns from an aligned start
    const bool has_l = hi_l > lo_l;
    const int ilo = __reduce_min_sync(0xffffffffu, has_l ? lo_l : 0x7fffffff), ihi = __reduce_max_sync(0xffffffffu, has_l ? hi_l : 0);
    double M = 0.0, SX = 0.0, SY = 0.0, S2 = 0.0;
    if (ihi > ilo) {
      const int cb0 = ilo & ~3, npiece = (ihi - cb0 + VGB_COLS - 1) / VGB_COLS;
      const int nstage = (VG_CHUNK / VGB_ROWS) * npiece;  // stage index = row group * npiece + piece
      // issue: lanes 0 .. 7 own the rows of the stage's group
      auto issue = [&](int st, int slot) {
        const int q = st / npiece, pc = st - q * npiece;
        const int lo = __shfl_sync(0xffffffffu, lo_l, VGB_ROWS * q + (lane & 7)), hi = __shfl_sync(0xffffffffu, hi_l, VGB_ROWS * q + (lane & 7));
        if (lane < VGB_ROWS) {
          const int p0 = cb0 + VGB_COLS * pc, p1 = p0 + VGB_COLS;
          const int s0 = max(lo & ~3, p0), s1 = min((hi + 3) & ~3, p1);  // 16-byte aligned column range of this row's piece
          const uint32_t mb = mbar0 + 16 * slot;
          if (hi > lo && s1 > s0) {
            const uint32_t bytes = (uint32_t)(s1 - s0) * 4u;
            const uint32_t dst = vgb_smem_addr(stage_base + (size_t)slot * VGB_STAGE + (size_t)lane * VGB_STRIDE + (size_t)(s0 - p0) * 4);
            const float *src = world + (size_t)(rbase + VGB_ROWS * q + lane) * N + s0;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                         "r"(bytes), "r"(mb)
                         : "memory");
          } else {
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mb) : "memory");
          }
        }
      };
      issue(0, 0);
#pragma unroll 1
      for (int st = 0; st < nstage; ++st) {
        const int slot = st & 1;
        if (st + 1 < nstage) issue(st + 1, slot ^ 1);
        vgb_wait(mbar0 + 16 * slot, phase[slot]);
        phase[slot] ^= 1u;
        // consume: lane (g, c4) adds the cells of row g of the stage, 4 at a time, 16 columns apart
        const int q = st / npiece, pc = st - q * npiece;
        const int lo = __shfl_sync(0xffffffffu, lo_l, VGB_ROWS * q + g), hi = __shfl_sync(0xffffffffu, hi_l, VGB_ROWS * q + g);
        const int p0 = cb0 + VGB_COLS * pc;
        const int s0 = max(lo & ~3, p0), s1 = min((hi + 3) & ~3, p0 + VGB_COLS);
        if (hi > lo && s1 > s0) {
          const float4 *srow = reinterpret_cast<const float4 *>(stage_base + (size_t)slot * VGB_STAGE + (size_t)g * VGB_STRIDE);
          double rs = 0.0, rsy = 0.0, rsyy = 0.0;
          for (int col = s0 + 4 * c4; col < s1; col += 16) {
            float4 w = srow[(col - p0) >> 2];
            if (col < lo || col + 4 > hi) {  // an end of the interval
              const unsigned keep = vg_keep4(col, lo, hi);
              w.x = (keep & 1u) ? w.x : 0.f, w.y = (keep & 2u) ? w.y : 0.f, w.z = (keep & 4u) ? w.z : 0.f, w.w = (keep & 8u) ? w.w : 0.f;
            }
            vg_add4(w, (double)col + 0.5, rs, rsy, rsyy);
          }
          const double x = (double)(rbase + VGB_ROWS * q + g) + 0.5;
          M += rs;
          SX = fma(x, rs, SX);
          SY += rsy;
          S2 += fma(x * x, rs, rsyy);
        }
        __syncwarp();  // every lane is done with this slot before it is refilled (next iteration's issue)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy reads before async-proxy writes
      }
    }
    const Sum4 tot = reduce4(M, SX, SY, S2);
    if (lane == 0) emit(a.sums, e, d.R, k, tot.m, tot.sx, tot.sy, tot.s2);
  }
}

// ---- block prefix sums: build, band sums, and the site-ordered (bit-stable) look-up kernel -------------------
// one thread per map row of the listed environments: the sequential double-double scan (see VgPfx above).
// envs: the environments to (re)build, nenv of them (null: 0 .. nenv - 1).
__global__ void __launch_bounds__(128) k_vg_prefix(const float *__restrict__ world_all, VgPfx *__restrict__ pfx_all, int N,
                                                   const int *__restrict__ envs, int nenv) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long)nenv * N) return;
  const int ei = (int)(t / N), row = (int)(t - (long)ei * N);
  const int e = envs ? envs[ei] : ei;
  const float4 *src = reinterpret_cast<const float4 *>(world_all + ((size_t)e * N + row) * N);
  const int nblk = N / VG_PB;
  double2 *dst = reinterpret_cast<double2 *>(pfx_all + ((size_t)e * N + row) * vg_pfx_row_entries(N));
  DD s{0.0, 0.0}, sy{0.0, 0.0}, syy{0.0, 0.0};
  float4 a = __ldg(src), b = __ldg(src + 1);
  for (int blk = 0; blk < nblk; ++blk) {
    dst[8 * blk] = make_double2(s.hi, s.lo), dst[8 * blk + 1] = make_double2(sy.hi, sy.lo);
    dst[8 * blk + 2] = make_double2(syy.hi, syy.lo);
    reinterpret_cast<float4 *>(dst + 8 * blk + 4)[0] = a, reinterpret_cast<float4 *>(dst + 8 * blk + 4)[1] = b;
    float4 na = a, nb = b;
    if (blk + 1 < nblk) na = __ldg(src + 2 * blk + 2), nb = __ldg(src + 2 * blk + 3);  // next block in flight during the scan step
    vg_pfx_step(s, sy, syy, a, b, blk);
    a = na, b = nb;
  }
  dst[8 * nblk] = make_double2(s.hi, s.lo), dst[8 * nblk + 1] = make_double2(sy.hi, sy.lo);
  dst[8 * nblk + 2] = make_double2(syy.hi, syy.lo);
  reinterpret_cast<float4 *>(dst + 8 * nblk + 4)[0] = make_float4(0.f, 0.f, 0.f, 0.f);
  reinterpret_cast<float4 *>(dst + 8 * nblk + 4)[1] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// the moments of every site = the sum of its band moments in ascending band order (plain stores: no memset, no atomics)
__global__ void k_vg_bandsum(const double *__restrict__ band, double *__restrict__ sums, int E, int R, int nbands) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // (e, k, c)
  if (i >= E * R * 4) return;
  const int e = i / (R * 4), kc = i - e * (R * 4);
  const double *src = band + (size_t)e * nbands * R * 4 + kc;
  double t = 0.0;
  for (int b = 0; b < nbands; ++b) t += src[(size_t)b * R * 4];
  sums[i] = t;
}

// entry hi - entry lo, rounded to fp64 once
__device__ __forceinline__ void vg_pfx_diff(const VgPfx *__restrict__ ph, const VgPfx *__restrict__ pl, double &s, double &sy,
                                            double &syy) {
  const double2 *h = reinterpret_cast<const double2 *>(ph), *l = reinterpret_cast<const double2 *>(pl);
  const double2 h0 = __ldg(h), h1 = __ldg(h + 1), h2 = __ldg(h + 2);
  const double2 l0 = __ldg(l), l1 = __ldg(l + 1), l2 = __ldg(l + 2);
  s = dd_diff(DD{h0.x, h0.y}, DD{l0.x, l0.y}), sy = dd_diff(DD{h1.x, h1.y}, DD{l1.x, l1.y}), syy = dd_diff(DD{h2.x, h2.y}, DD{l2.x, l2.y});
}

// Site-ordered variant (test option "ordered_sums" and the parity reference of the fused kernel): one warp per
// (environment, site) walks the site's row span of the partition TABLE 32 rows at a time, lane <-> row; two entry
// look-ups and at most two map sectors per row.  The per-lane sums and the shuffle tree have a fixed order and the
// result is STORED (no atomics, no memset) (k_vg_fixup adds the cells of flagged rows afterwards).
constexpr int VG_LOOK_WARPS = 4;
__global__ void __launch_bounds__(VG_LOOK_WARPS * 32) k_vg_lookup(VgArgs a, const VgPfx *__restrict__ pfx_all) {
  const EnvDev &d = a.d;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int e = blockIdx.y, k = blockIdx.x * VG_LOOK_WARPS + warp;
  const int N = d.N, R = d.R;
  if (k >= R) return;
  const int2 sp = __ldg(reinterpret_cast<const int2 *>(a.span) + (size_t)e * R + k);
  const size_t pe = vg_pfx_row_entries(N);
  const VgPfx *pfx = pfx_all + (size_t)e * N * pe;
  const float *world = d.world + (size_t)e * N * N;
  const uint32_t *tab = a.table + ((size_t)e * R + k) * N;
  double M = 0.0, SX = 0.0, SY = 0.0, S2 = 0.0;
#pragma unroll 2
  for (int row = (sp.y > sp.x ? sp.x : sp.y) + lane; row < sp.y; row += 32) {  // (an empty span is {INT_MAX, 0})
    const uint32_t iv = __ldg(tab + row);
    const int lo = (int)(iv & 0xffffu), hi = (int)(iv >> 16);
    if (hi <= lo) continue;
    const int blo = (lo + VG_PB - 1) / VG_PB, bhi = hi / VG_PB;
    const float *wrow = world + (size_t)row * N;
    double rs = 0.0, rsy = 0.0, rsyy = 0.0;
    if (blo <= bhi) {
      if (blo < bhi) vg_pfx_diff(pfx + (size_t)row * pe + bhi, pfx + (size_t)row * pe + blo, rs, rsy, rsyy);
      if (lo & (VG_PB - 1)) {  // head: the cells [lo, 8 blo) of block blo - 1
        const float4 *p4 = reinterpret_cast<const float4 *>(wrow + VG_PB * (blo - 1));
        vg_add8(__ldg(p4), __ldg(p4 + 1), 0xffu & ~((1u << (lo & (VG_PB - 1))) - 1u), (double)(VG_PB * (blo - 1)) + 0.5, rs, rsy, rsyy);
      }
      if (hi & (VG_PB - 1)) {  // tail: the cells [8 bhi, hi) of block bhi
        const float4 *p4 = reinterpret_cast<const float4 *>(wrow + VG_PB * bhi);
        vg_add8(__ldg(p4), __ldg(p4 + 1), (1u << (hi & (VG_PB - 1))) - 1u, (double)(VG_PB * bhi) + 0.5, rs, rsy, rsyy);
      }
    } else {  // the interval lies inside block bhi
      const float4 *p4 = reinterpret_cast<const float4 *>(wrow + VG_PB * bhi);
      const unsigned keep = ((1u << (hi & (VG_PB - 1))) - 1u) & ~((1u << (lo & (VG_PB - 1))) - 1u);
      vg_add8(__ldg(p4), __ldg(p4 + 1), keep, (double)(VG_PB * bhi) + 0.5, rs, rsy, rsyy);
    }
    const double x = (double)row + 0.5;
    M += rs;
    SX = fma(x, rs, SX);
    SY += rsy;
    S2 += fma(x * x, rs, rsyy);
  }
  const Sum4 tot = reduce4(M, SX, SY, S2);
  if (lane == 0) {
    double *dst = a.sums + ((size_t)e * R + k) * 4;
    dst[0] = tot.m, dst[1] = tot.sx, dst[2] = tot.sy, dst[3] = tot.s2;
  }
}

// ---- flagged rows: the stated rule per cell --------------------------------------------------------
// items = flagged (environment, row) pairs listed by k_vg_partition, each split in 256-column pieces;
// a fixed grid strides over them.  One cell per thread per piece: 32 x R exact distances per warp.
template <int SRC, bool LABELS>
__global__ void __launch_bounds__(256) k_vg_fixup(VgArgs a) {
  const EnvDev &d = a.d;
  const int N = d.N, R = d.R;
  const int nflag = *a.nflag;
  const int pieces = (N + 255) / 256;
  for (int item = blockIdx.x; item < nflag * pieces; item += gridDim.x) {
    const int entry = a.flag_list[item / pieces], piece = item - (item / pieces) * pieces;
    const int e = entry / N, row = entry - e * N;
    const int col = piece * 256 + threadIdx.x;
    const double *pos = d.pos_g + (size_t)e * R * 2;
    const float *world = d.world + (size_t)e * N * N;
    const uint32_t *plane = SRC == SRC_EXPLORED ? d.sysb + ((size_t)e * 2) * plane_words(d) : nullptr;
    int *labels = (LABELS && a.labels) ? a.labels + (size_t)e * N * N : nullptr;
    const double res = d.res, cx = cell_centre(row, res), cy = cell_centre(min(col, N - 1), res);
    int best = 0;
    double bd = d2_exact(cx, cy, __ldg(pos), __ldg(pos + 1));
    for (int k = 1; k < R; ++k) {
      const double dd = d2_exact(cx, cy, __ldg(pos + 2 * k), __ldg(pos + 2 * k + 1));
      if (dd < bd) bd = dd, best = k;
    }
    float v = 0.f;
    if (col < N) {
      if (labels) labels[(size_t)row * N + col] = best;
      if (SRC == SRC_WORLD || plane_bit(plane, d.NW, row, col)) v = __ldg(world + (size_t)row * N + col);
    }
    // one set of atomics per (warp, label) instead of per cell: 32 consecutive cells rarely have more
    // than two owners
    const double dv = (double)v, xd = (double)row + 0.5, yd = (double)col + 0.5;
    const int lane = threadIdx.x & 31;
    unsigned todo = __ballot_sync(0xffffffffu, v != 0.f);
    while (todo) {
      const int leader = __ffs(todo) - 1;
      const int lab = __shfl_sync(0xffffffffu, best, leader);
      const bool mine = v != 0.f && best == lab;
      const double w = mine ? dv : 0.0;
      const Sum4 tot = reduce4(w, xd * w, yd * w, (xd * xd + yd * yd) * w);
      if (lane == leader) emit(a.sums, e, R, lab, tot.m, tot.sx, tot.sy, tot.s2);
      todo &= ~__ballot_sync(0xffffffffu, mine);
    }
  }
}

}  // namespace ccb
