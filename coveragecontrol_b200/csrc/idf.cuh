// idf.cuh -- WorldIDF generation on the GPU.
//
// Replaces reference cppsrc/core/src/cuda/generate_world_map.cu:146-280 (kernel, normalize,
// thrust::max_element, host wrapper).  Same arithmetic, different machine mapping:
//   * threadIdx.x runs along y (the contiguous axis) and each thread owns 4 consecutive cells,
//     so stores are coalesced float4 (the reference strides by N);
//   * the features are culled per tile and their 1-D erfc factors are tabulated once per tile
//     in shared memory: every cell-feature pair costs one FFMA + compare for the truncation test
//     and 4 FMUL + 4 FFMA when in range, instead of 8 erfcf + 10 IEEE divides.
// Mode 0 reproduces the reference kernel's floating-point operation order exactly (the 1-D
// factors are the same function of the same fp32 arguments, so tabulating them changes no bit).
#pragma once

#include "common.cuh"

namespace ccb {

struct BndF {
  float mx, my, sx, sy, sc, rho;  // generate_world_map.h:42-46
};

struct PolyDev {
  const float *x, *y;   // concatenated piece vertices
  const int *start;     // piece k occupies [start[k], start[k]+sz[k])
  const int *sz;
  const float *imp;
  const float4 *bounds; // xmin, xmax, ymin, ymax
  int num;
  const int *bk_start;  // the pieces whose bounding box touches a bucket of the feature grid: bucket b lists
  const int *bk_items;  // bk_items[bk_start[b] .. bk_start[b + 1])
};

constexpr int IDF_TX = 32;   // tile rows (x)
constexpr int IDF_TY = 32;   // tile cols (y)
constexpr int IDF_FC = 16;   // features tabulated per chunk
constexpr int IDF_THREADS = 256;
constexpr int IDF_MAXPOLY = 64;  // polygon pieces kept per tile before falling back to the full list

// Point-in-piece test with the outcome of the reference's device predicate (geometry_utils.cuh:44-101), restated
// as a side mask.  Every edge whose closed y-range holds the point is directed upwards (from its lower to its
// upper end; an edge with equal ends runs from the second vertex to the first, as there) and classifies the
// point: bit 0 strictly left of it, bit 1 strictly right, both bits on its supporting line.  The two products of
// the cross product are rounded separately and compared (no FMA), so the classification is the reference's bit
// for bit.  The point is inside iff the OR over the edges has both bits: some edge holds it, or it has an edge
// on either side.
__device__ __forceinline__ unsigned idf_edge_side(float ax, float ay, float bx, float by, float rx, float ry) {
  const bool a_low = ay < by;
  const float lx = a_low ? ax : bx, ly = a_low ? ay : by;
  const float ux = a_low ? bx : ax, uy = a_low ? by : ay;
  if (ry < ly || ry > uy) return 0u;
  const float along = __fmul_rn(__fsub_rn(ux, lx), __fsub_rn(ry, ly));
  const float across = __fmul_rn(__fsub_rn(uy, ly), __fsub_rn(rx, lx));
  return along > across ? 1u : (along < across ? 2u : 3u);
}

__device__ __forceinline__ bool idf_point_in_piece(const float *x, const float *y, int n, float rx, float ry) {
  unsigned sides = 0u;
  float ax = x[n - 1], ay = y[n - 1];  // edge (n - 1) -> 0 first: the OR does not depend on the order
  for (int i = 0; i < n && sides != 3u; ++i) {
    const float bx = x[i], by = y[i];
    sides |= idf_edge_side(ax, ay, bx, by, rx, ry);
    ax = bx, ay = by;
  }
  return sides == 3u;
}

__device__ __forceinline__ int float_to_ordered(float f) {
  int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ordered_to_float(int i) {
  return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff);
}

// One CTA per IDF_TX x (NTY IDF_TY) tile: NTY = 1 for small maps (more CTAs than SMs matters more than the fixed
// cost per CTA), 4 for large ones (one feature cull, one set of barriers and one x-side tabulation per 32 x 128
// cells: the C4 map took 65 536 CTAs of ~4 us each).  A thread owns 4 consecutive cells in each of the NTY sub-tiles.
// MODE 0: reference operation order.  MODE 1: separable differences tabulated in fp64 (more accurate than the
// reference, same truncation support).
template <int MODE, int NTY>
__global__ void __launch_bounds__(IDF_THREADS)
k_idf_generate(const BndF *__restrict__ bnd, int F, const int *__restrict__ bk_start, const int *__restrict__ bk_items, int bk_nb,
               int bk_size, PolyDev poly, int N, int row0, int row1, float res, float trun_sq, float one_by_sqrt2,
               float *__restrict__ out, int *__restrict__ max_ordered) {
  __shared__ int s_list[IDF_FC];
  __shared__ int s_nlist;
  __shared__ int s_scan_base;
  __shared__ float s_tx[IDF_FC][IDF_TX];    // (mid.x - mu_x)/sigma_x
  __shared__ float s_axl[IDF_FC][IDF_TX];   // MODE0: scale*erfc(left x) ; MODE1: scale/4*(erfc l - erfc r)
  __shared__ float s_axr[IDF_FC][IDF_TX];   // MODE0: scale*erfc(right x)
  constexpr int TYW = IDF_TY * NTY;         // columns of the CTA's tile
  __shared__ float s_ty2[IDF_FC][TYW];      // ((mid.y - mu_y)/sigma_y)^2
  __shared__ float s_eyl[IDF_FC][TYW];      // MODE0: erfc(lower y)      ; MODE1: erfc l - erfc r
  __shared__ float s_eyr[IDF_FC][MODE == 0 ? TYW : 1];  // MODE0: erfc(upper y)
  __shared__ int s_poly[IDF_MAXPOLY];
  __shared__ int s_npoly;
  __shared__ float s_red[IDF_THREADS / 32];

  const int tid = threadIdx.x;
  // rows [row0, row1): the whole map or a row band.  Tiles sit on the global 32-row grid whatever row0 is, so a tile
  // never straddles two feature buckets (bk_size is a multiple of IDF_TX); rows below row0 are computed and dropped
  const int x0 = (row0 & ~(IDF_TX - 1)) + blockIdx.y * IDF_TX, y0 = blockIdx.x * TYW;
  const int lrow = tid >> 3, lcol = (tid & 7) * 4;
  const int gx = x0 + lrow, gy = y0 + lcol;

  // tile extent of the cell mid points, used for conservative culling
  const float tile_xlo = max(x0, row0) * res, tile_xhi = (min(x0 + IDF_TX, row1)) * res;
  const float tile_ylo = y0 * res, tile_yhi = (min(y0 + TYW, N)) * res;

  float acc[NTY][4];
#pragma unroll
  for (int u = 0; u < NTY; ++u) acc[u][0] = acc[u][1] = acc[u][2] = acc[u][3] = 0.f;

  // The host sorted the features into a coarse grid of bk_size x bk_size-cell buckets (ascending feature
  // index inside a bucket): the tile only scans the features whose truncation box touches its bucket.
  const int bucket = (x0 / bk_size) * bk_nb + (y0 / bk_size);
  const int *items = bk_items + bk_start[bucket];
  const int Fb = bk_start[bucket + 1] - bk_start[bucket];
  (void)F;

  if (tid == 0) s_scan_base = 0;
  __syncthreads();
  while (true) {
    // ---- gather the next <= IDF_FC features whose truncation box touches this tile (ascending
    //      feature index, so the accumulation order is the reference's loop order) -------------
    const int base = s_scan_base;
    __syncthreads();  // everyone has read the cursor before warp 0 advances it
    if (base >= Fb) break;
    if (tid < 32) {
      const float T = sqrtf(trun_sq) * 1.0001f + 1e-3f;
      int n = 0;
      int f = base;
      while (f < Fb && n < IDF_FC) {
        const int ff = f + tid < Fb ? items[f + tid] : -1;
        bool keep = false;
        if (ff >= 0) {
          BndF b = bnd[ff];
          float rx = T * b.sx + res, ry = T * b.sy + res;
          keep = !(b.mx + rx < tile_xlo || b.mx - rx > tile_xhi || b.my + ry < tile_ylo || b.my - ry > tile_yhi);
        }
        unsigned m = __ballot_sync(0xffffffffu, keep);
        const int room = IDF_FC - n;
        int next_f = f + 32;
        if (__popc(m) > room) {  // keep the first `room` hits, resume right after the last kept one
          unsigned mm = m;
          int last = 0;
          for (int k = 0; k < room; ++k) {
            last = __ffs(mm) - 1;
            mm &= mm - 1;
          }
          m &= (last == 31) ? 0xffffffffu : ((1u << (last + 1)) - 1u);
          next_f = f + last + 1;
        }
        if ((m >> tid) & 1u) s_list[n + __popc(m & ((1u << tid) - 1u))] = ff;
        n += __popc(m);
        f = next_f;
      }
      if (tid == 0) {
        s_nlist = n;
        s_scan_base = (n == IDF_FC) ? min(f, Fb) : Fb;
      }
    }
    __syncthreads();
    const int nl = s_nlist;
    // ---- tabulate the 1-D factors ------------------------------------------------------------
    for (int t = tid; t < nl * (IDF_TX + TYW); t += IDF_THREADS) {
      int k = t / (IDF_TX + TYW);
      int u = t - k * (IDF_TX + TYW);
      BndF b = bnd[s_list[k]];
      bool is_x = u < IDF_TX;
      int li = is_x ? u : u - IDF_TX;
      int gi = (is_x ? x0 : y0) + li;
      float mu = is_x ? b.mx : b.my, sg = is_x ? b.sx : b.sy;
      float idx = (float)gi;
      // generate_world_map.cu:155-159 as compiled: tr = fma(idx,res,res); mid = fma(idx,res,tr)*0.5
      float hi = fmaf(idx, res, res);
      float mid = __fmul_rn(fmaf(idx, res, hi), 0.5f);
      float tm = __fdiv_rn(__fsub_rn(mid, mu), sg);  // generate_world_map.cu:105-106
      if (MODE == 0) {
        float al = __fdiv_rn(fmaf(idx, res, -mu), sg);     // generate_world_map.cu:80-81
        float ar = __fdiv_rn(__fsub_rn(hi, mu), sg);
        float el = erfcf(__fmul_rn(al, one_by_sqrt2));     // generate_world_map.cu:92
        float er = erfcf(__fmul_rn(ar, one_by_sqrt2));
        if (is_x) {
          s_tx[k][li] = tm;
          s_axl[k][li] = __fmul_rn(b.sc, el);
          s_axr[k][li] = __fmul_rn(b.sc, er);
        } else {
          s_ty2[k][li] = __fmul_rn(tm, tm);
          s_eyl[k][li] = el;
          s_eyr[k][li] = er;
        }
      } else {
        const double c = 0.70710678118654752440;
        double lo = ((double)gi * (double)res - (double)mu) / (double)sg * c;
        double up = ((double)(gi + 1) * (double)res - (double)mu) / (double)sg * c;
        double diff = erfc(lo) - erfc(up);
        if (is_x) {
          s_tx[k][li] = tm;
          s_axl[k][li] = (float)(0.25 * (double)b.sc * diff);
        } else {
          s_ty2[k][li] = __fmul_rn(tm, tm);
          s_eyl[k][li] = (float)diff;
        }
      }
    }
    __syncthreads();
    // ---- accumulate ----------------------------------------------------------------------------
    for (int k = 0; k < nl; ++k) {
      float tx = s_tx[k][lrow];
#pragma unroll
      for (int u = 0; u < NTY; ++u) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int yc = IDF_TY * u + lcol + c;
          float r2 = fmaf(tx, tx, s_ty2[k][yc]);  // generate_world_map.cu:113 as compiled
          if (r2 > trun_sq) continue;
          if (MODE == 0) {
            float axl = s_axl[k][lrow], axr = s_axr[k][lrow];
            float eyl = s_eyl[k][yc], eyr = s_eyr[k][yc];
            // generate_world_map.cu:116-119: +Q(bl) -Q(br) -Q(tl) +Q(tr), Q = (sc*ex)*ey/4
            acc[u][c] = fmaf(__fmul_rn(axl, eyl), 0.25f, acc[u][c]);
            acc[u][c] = fmaf(__fmul_rn(axr, eyl), -0.25f, acc[u][c]);
            acc[u][c] = fmaf(__fmul_rn(axl, eyr), -0.25f, acc[u][c]);
            acc[u][c] = fmaf(__fmul_rn(axr, eyr), 0.25f, acc[u][c]);
          } else {
            acc[u][c] = fmaf(s_axl[k][lrow], s_eyl[k][yc], acc[u][c]);
          }
        }
      }
    }
    __syncthreads();
  }

  // ---- polygon features (generate_world_map.cu:124-144) ----------------------------------------
  if (poly.num > 0) {
    if (tid == 0) s_npoly = 0;
    __syncthreads();
    // only the pieces listed for this tile's bucket (a 64-polygon world has hundreds of pieces, a tile sees a few)
    const int pk0 = poly.bk_start[bucket], pk1 = poly.bk_start[bucket + 1];
    for (int kk = pk0 + tid; kk < pk1; kk += IDF_THREADS) {
      const int k = poly.bk_items[kk];
      float4 b = poly.bounds[k];
      // mid points of the tile span [lo + res/2, hi - res/2]; test against the closed bbox
      bool keep = !(tile_xhi < b.x || tile_xlo > b.y || tile_yhi < b.z || tile_ylo > b.w);
      if (keep) {
        int slot = atomicAdd(&s_npoly, 1);
        if (slot < IDF_MAXPOLY) s_poly[slot] = k;
      }
    }
    __syncthreads();
    const int np = s_npoly;
    if (np > 0 && gx >= row0 && gx < row1) {
#pragma unroll
      for (int u = 0; u < NTY; ++u) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int gyc = gy + IDF_TY * u + c;
          if (gyc >= N) continue;
          float fx = (float)gx, fy = (float)gyc;
          float midx = __fmul_rn(fmaf(fx, res, fmaf(fx, res, res)), 0.5f);
          float midy = __fmul_rn(fmaf(fy, res, fmaf(fy, res, res)), 0.5f);
          float pim = 0.f;
          int cnt = np <= IDF_MAXPOLY ? np : pk1 - pk0;
          for (int q = 0; q < cnt; ++q) {
            int k = np <= IDF_MAXPOLY ? s_poly[q] : poly.bk_items[pk0 + q];
            float4 b = poly.bounds[k];
            if (midx < b.x || midx > b.y || midy < b.z || midy > b.w) continue;
            int st = poly.start[k];
            if (idf_point_in_piece(poly.x + st, poly.y + st, poly.sz[k], midx, midy))
              pim = fmaxf(pim, poly.imp[k]);
          }
          acc[u][c] = __fadd_rn(acc[u][c], pim);  // generate_world_map.cu:161-163
        }
      }
    }
  }

  // ---- store + tile max ------------------------------------------------------------------------
  float vmax = -INFINITY;
  if (gx >= row0 && gx < row1) {
    float *row = out + (size_t)gx * N;
#pragma unroll
    for (int u = 0; u < NTY; ++u) {
      const int gyu = gy + IDF_TY * u;
      if ((N & 3) == 0 && gyu + 3 < N) {
        *reinterpret_cast<float4 *>(row + gyu) = make_float4(acc[u][0], acc[u][1], acc[u][2], acc[u][3]);
        vmax = fmaxf(vmax, fmaxf(fmaxf(acc[u][0], acc[u][1]), fmaxf(acc[u][2], acc[u][3])));
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (gyu + c < N) {
            row[gyu + c] = acc[u][c];
            vmax = fmaxf(vmax, acc[u][c]);
          }
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
  if ((tid & 31) == 0) s_red[tid >> 5] = vmax;
  __syncthreads();
  if (tid == 0) {
    float m = s_red[0];
#pragma unroll
    for (int w = 1; w < IDF_THREADS / 32; ++w) m = fmaxf(m, s_red[w]);
    if (m > -INFINITY) atomicMax(max_ordered, float_to_ordered(m));
  }
}

// generate_world_map.cu:252-259 (factor) + :166-174 (normalize); the factor is computed on the
// device from the map maximum so that no host round trip sits between the two kernels.
__global__ void k_idf_factor(const int *max_ordered, float norm, float *factor) {
  float mx = ordered_to_float(*max_ordered);
  float f;
  if ((double)mx < kEps) f = norm;
  else f = __fdiv_rn(norm, mx);
  *factor = f;
}

__global__ void k_idf_normalize(float *__restrict__ map, size_t n4, size_t n, const float *factor) {
  const float f = *factor;
  if (!((double)f > 1e-5)) return;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  float4 *m4 = reinterpret_cast<float4 *>(map);
  for (size_t k = i; k < n4; k += stride) {
    float4 v = m4[k];
    v.x = __fmul_rn(v.x, f), v.y = __fmul_rn(v.y, f), v.z = __fmul_rn(v.z, f), v.w = __fmul_rn(v.w, f);
    m4[k] = v;
  }
  for (size_t k = n4 * 4 + i; k < n; k += stride) map[k] = __fmul_rn(map[k], f);
}

}  // namespace ccb
