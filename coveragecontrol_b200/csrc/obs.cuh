// obs.cuh -- LPAC observation tensors produced directly on the device.
//
// Replaces the per-robot Python loop of reference python/coverage_control/coverage_env_utils.py:
// get_raw_local_maps (:97-117), get_raw_obstacle_maps (:119-141), get_communication_maps
// (:143-191), resize_maps (:200-225, torchvision antialias bilinear) and get_maps (:227-277),
// which call RobotModel::GetRobotLocalMap / GetObstacleMap (robot_model.h:303-345) R times per
// step, copy every L x L map twice and resize on the CPU.
#pragma once

#include "common.cuh"
#include "vwindow.cuh"
#include "maps.cuh"

namespace ccb {

// E x R x L x L raw local maps (kind CCB_MAP_LOCAL) or obstacle maps (CCB_MAP_OBSTACLE).
// grid = (L, R, E): one CTA per output row, coalesced along y.
__global__ void k_raw_window_maps(EnvDev d, int kind, float *__restrict__ out) {
  const int a = blockIdx.x, r = blockIdx.y, e = blockIdx.z;
  const size_t er = (size_t)e * d.R + r;
  const double *p = d.pos_g + er * 2;
  const Window w = make_window(d.res, p[0], p[1], d.L, d.N);
  const int gx = w.ox + a;
  float *dst = out + (er * d.L + a) * d.L;
  for (int b = threadIdx.x; b < d.L; b += blockDim.x) {
    const int gy = w.oy + b;
    const bool in = gx >= 0 && gx < d.N && gy >= 0 && gy < d.N;
    float v;
    if (kind == CCB_MAP_OBSTACLE) v = in ? 0.f : 1.f;
    else v = in ? robot_map_value(d, e, r, gx, gy) : 0.f;
    dst[b] = v;
  }
}

// ---- communication graph of LPAC's GNN -----------------------------------------------------------------
// coverage_env_utils.py:313-335 get_weights and :368-369 of get_torch_geometric_data, float32 as there:
//   w[i][j] = exp(-|p_i - p_j|^2 / Rc^2), set to 0 below 1/e (i.e. beyond the communication range) and on
//   the diagonal;  pos_norm = (p + N/2) / N.
// One CTA per environment; weights [E][R][R], pos_norm [E][R][2] (either may be null).
__global__ void k_graph(EnvDev d, float *__restrict__ weights, float *__restrict__ pos_norm) {
  const int e = blockIdx.x, R = d.R;
  const double *pos = d.pos_g + (size_t)e * R * 2;
  const float rc2 = (float)(d.comm_range * d.comm_range);
  const float onebyexp = (float)(1.0 / exp(1.0));
  if (weights) {
    for (int t = threadIdx.x; t < R * R; t += blockDim.x) {
      const int i = t / R, j = t - i * R;
      const float dx = __fsub_rn((float)pos[2 * i], (float)pos[2 * j]), dy = __fsub_rn((float)pos[2 * i + 1], (float)pos[2 * j + 1]);
      const float d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
      float w = expf(-d2 / rc2);
      if (w < onebyexp || i == j) w = 0.f;
      weights[(size_t)e * R * R + t] = w;
    }
  }
  if (pos_norm) {
    const float half = (float)d.N / 2.f, size = (float)d.N;
    for (int t = threadIdx.x; t < 2 * R; t += blockDim.x) pos_norm[(size_t)e * 2 * R + t] = __fadd_rn((float)pos[t], half) / size;
  }
}

// Communication maps with the semantics LPAC consumes (coverage_env_utils.py:143-191): float32
// arithmetic, torch.round = round-half-even, duplicates summed.  out: E x R x 2 x C x C, or a
// channel slice of a wider tensor (chan_stride = floats between robots, chan_off = first channel).
// One CTA per (robot, env).
// one thread scatters robot r's neighbours into its two (already zeroed) C x C planes, ascending neighbour index
__device__ __forceinline__ void comm_scatter(const EnvDev &d, int e, int r, int C, float *ox, float *oy) {
  const double *pos = d.pos_g + (size_t)e * d.R * 2;
  const float denom = (float)(d.comm_range * d.res * 2.0);
  const float center = (float)(C / 2.0 - d.res / 2.0);
  const float cr = (float)d.comm_range;
  const double sx = pos[2 * r], sy = pos[2 * r + 1];
  for (int j = 0; j < d.R; ++j) {
    if (j == r) continue;
    // coverage_system.cpp:274-291: relative position p_j - p_r (stored negated for j < r)
    double dx, dy;
    if (j > r) dx = pos[2 * j] - sx, dy = pos[2 * j + 1] - sy;
    else dx = -(sx - pos[2 * j]), dy = -(sy - pos[2 * j + 1]);
    if (!(sqrt(dx * dx + dy * dy) < d.comm_range)) continue;
    const float fx = (float)dx, fy = (float)dy;
    const float ix = rintf(__fadd_rn(__fdiv_rn(__fmul_rn(fx, (float)C), denom), center));
    const float iy = rintf(__fadd_rn(__fdiv_rn(__fmul_rn(fy, (float)C), denom), center));
    const int xi = (int)ix, yi = (int)iy;
    if (xi < 0 || xi >= C || yi < 0 || yi >= C) continue;
    ox[xi * C + yi] += __fdiv_rn(fx, cr);
    oy[xi * C + yi] += __fdiv_rn(fy, cr);
  }
}

__global__ void k_comm_maps(EnvDev d, int C, float *__restrict__ out, size_t robot_stride, size_t chan_off) {
  const int r = blockIdx.x, e = blockIdx.y;
  const size_t er = (size_t)e * d.R + r;
  float *ox = out + er * robot_stride + chan_off;
  float *oy = ox + (size_t)C * C;
  for (int t = threadIdx.x; t < 2 * C * C; t += blockDim.x) ox[t] = 0.f;
  __syncthreads();
  if (threadIdx.x == 0) comm_scatter(d, e, r, C, ox, oy);
}

// Separable antialiased bilinear down-sampling weights (what torchvision resize(antialias=True)
// applies): for output index o the taps are input indices [xmin[o], xmin[o]+xsize[o]).
struct ResizeTable {
  const int *xmin;    // [C]
  const int *xsize;   // [C]
  const float *w;     // [C][kmax]
  int kmax;
};

// Fused crop -> mask -> antialias resize for the local-map and obstacle channels, writing
// straight into the E x R x nch x C x C CNN input.  grid = (C, R, E): one CTA per output row ox.
//   pass 1 (width, as torch does first): tmp[a][oy] = sum_t w[oy][t] * in[xmin[ox]+a][ymin[oy]+t]
//   pass 2 (height):                      out[ox][oy] = sum_a w[ox][a] * tmp[a][oy]
// Dynamic smem: kmax * L floats (input rows) + kmax * C floats (tmp).
__global__ void k_cnn_local_obstacle(EnvDev d, int C, ResizeTable tab, float *__restrict__ out,
                                     size_t robot_stride, size_t local_off, size_t obst_off) {
  extern __shared__ float s_obs[];
  float *s_in = s_obs;                       // [kmax][L]
  float *s_tmp = s_obs + (size_t)tab.kmax * d.L;  // [kmax][C]
  const int ox = blockIdx.x, r = blockIdx.y, e = blockIdx.z;
  const size_t er = (size_t)e * d.R + r;
  const double *p = d.pos_g + er * 2;
  const Window w = make_window(d.res, p[0], p[1], d.L, d.N);
  const int a0 = tab.xmin[ox], na = tab.xsize[ox];
  const float *wrow = tab.w + (size_t)ox * tab.kmax;

  for (int pass = 0; pass < 2; ++pass) {  // 0: local map, 1: obstacle map
    for (int t = threadIdx.x; t < na * d.L; t += blockDim.x) {
      const int a = t / d.L, b = t - a * d.L;
      const int gx = w.ox + a0 + a, gy = w.oy + b;
      const bool in = gx >= 0 && gx < d.N && gy >= 0 && gy < d.N;
      float v;
      if (pass == 0) v = in ? robot_map_value(d, e, r, gx, gy) : 0.f;
      else v = in ? 0.f : 1.f;
      s_in[a * d.L + b] = v;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < na * C; t += blockDim.x) {
      const int a = t / C, oy = t - a * C;
      const int b0 = tab.xmin[oy], nb = tab.xsize[oy];
      const float *wc = tab.w + (size_t)oy * tab.kmax;
      float acc = 0.f;
      for (int q = 0; q < nb; ++q) acc = fmaf(wc[q], s_in[a * d.L + b0 + q], acc);
      s_tmp[a * C + oy] = acc;
    }
    __syncthreads();
    float *dst = out + er * robot_stride + (pass == 0 ? local_off : obst_off) + (size_t)ox * C;
    for (int oy = threadIdx.x; oy < C; oy += blockDim.x) {
      float acc = 0.f;
      for (int a = 0; a < na; ++a) acc = fmaf(wrow[a], s_tmp[a * C + oy], acc);
      dst[oy] = acc;
    }
    __syncthreads();
  }
}


// ---- specialised path for L = 8 C (LPAC: 256 -> 32), N % 4 == 0 -----------------------------------
// One CTA per (robot, env), 8 warps, lane = output column oy.  The robot's bit plane says where its
// local map is non-zero, and after a few hundred steps that is still a small part of the 256 x 256
// window, so the kernel is organised around the rows that have a seen cell:
//   1. the 264 x 9 mask words of the window rows a in [-4, 260) are staged in shared memory, clipped to
//      the window's columns (2.4 KB of global reads instead of 270 KB of floats), and the rows with a
//      set bit are compacted into a list;
//   2. width pass: the warps deal the listed rows between them in batches of CNN8_BATCH rows whose
//      world values are all requested before the first is used (one HBM latency per batch instead of
//      one per row).  A lane owns the 8 window columns 8 oy .. 8 oy + 7 of its output column: one byte
//      of the row's mask says which are seen, the values arrive as 16-byte loads (the window origin is
//      arbitrary, so the loads start at the aligned column below it and the misalignment S, uniform
//      over the CTA, is a template parameter), the 4 + 4 neighbour taps come from the adjacent lanes by
//      shuffle, and the 16-tap sum goes to s_tmp[a][oy];
//   3. height pass: warp w owns output rows 4w .. 4w+3 and sums its 16 taps in ascending row order,
//      the order torchvision's separable antialias kernel uses.
// The obstacle channel needs no map at all: obstacle(a, b) = 1 - in_x(a) in_y(b), so its width pass
// has only two possible values per output column.  wfull[o][k] = weight of input index 8 o - 4 + k
// (zero outside torch's tap range).
constexpr int CNN8_THREADS = 256;
constexpr int CNN8_WARPS = CNN8_THREADS / 32;
constexpr int CNN8_ROWS = 264;  // window rows -4 .. 259
constexpr int CNN8_WORDS = 9;   // mask words a 256-cell window row can touch
constexpr int CNN8_BATCH = 3;   // rows whose loads are in flight together

struct Cnn8Shared {
  float tmp[CNN8_ROWS][32];
  uint32_t mask[CNN8_ROWS * CNN8_WORDS + 1];
  float w[32][16];
  uint32_t flags[(CNN8_ROWS + 31) / 32 + 1];  // bit a + 4: window row a has a non-zero cell
  int act[CNN8_ROWS];
  int nact;
};
constexpr size_t CNN8_SMEM_BYTES = sizeof(Cnn8Shared);
// The fused variant (VOR) also reduces DecentralizedCVT's window Voronoi (vwindow.cuh) over the same
// window: same bit plane, same seen rows, same world values -- the controller of the NEXT step finds
// its moments ready.  Extra state: the robot's own-cell interval of every window row as bits.
struct Cnn8Vor {
  uint32_t own[CNN8_ROWS * CNN8_WORDS + 1];  // cells of the robot's own Voronoi cell (indexed like Cnn8Shared::mask)
  int lo[256], hi[256];                      // the same as map-column intervals, per window row
  uint32_t exact[8];                         // rows that need the per-cell rule
  float wt[16][32];                          // tap weights transposed (conflict-free): frees 16 registers
  int K;
  double red[CNN8_WARPS][4];
};
// dynamic shared memory of the fused variant: Cnn8Shared, Cnn8Vor, px[R], py[R]
__host__ __device__ inline size_t cnn8_vor_smem_bytes(int R) {
  return sizeof(Cnn8Shared) + sizeof(Cnn8Vor) + sizeof(double) * 2 * (size_t)R;
}

// width pass over the listed rows; S = (window origin column) & 3
template <int S, bool VOR, bool DENSE>
__device__ __forceinline__ void cnn8_width_pass(const EnvDev &d, Cnn8Shared &sh, const float *__restrict__ world, int ox_map,
                                                int oy_map, int wbase, const float (&wc)[16], const Cnn8Vor *vor, VwAcc &acc) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nact = sh.nact;
  const int gy0 = oy_map + 8 * lane;  // map column of this lane's first cell
  const int gyb = gy0 - S;            // 16-byte aligned (N % 4 == 0)
  constexpr int NV = S == 0 ? 2 : 3;  // aligned 4-cell groups covering the 8 cells
  bool grp_in[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) grp_in[j] = gyb + 4 * j >= 0 && gyb + 4 * j + 3 < d.N;
  unsigned inmap = 0u;  // bit i: column gy0 + i is inside the map
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (gy0 + i >= 0 && gy0 + i < d.N) inmap |= 1u << i;
  // the lane's 8 mask bits start at bit `off` of the row's 9 words
  const int off = gy0 - 32 * wbase;
  const int wi = off >> 5, shf = off & 31;
  const float fill = d.fill;
  constexpr bool dense = DENSE;  // unseen cells carry pUnknownImportance (fill != 0): rare, compiled apart
  constexpr int BATCH = VOR ? 2 : CNN8_BATCH;  // the fused variant also carries the fp64 moments in registers
#pragma unroll 1
  for (int base = warp * BATCH; base < nact; base += CNN8_WARPS * BATCH) {
    float4 v[BATCH][NV];
    unsigned vis[BATCH];
#pragma unroll
    for (int b = 0; b < BATCH; ++b) {
      vis[b] = 0u;
#pragma unroll
      for (int j = 0; j < NV; ++j) v[b][j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (base + b < nact) {  // warp-uniform
        const int ai = sh.act[base + b];
        const uint32_t lo = sh.mask[ai * CNN8_WORDS + wi], hi = sh.mask[ai * CNN8_WORDS + wi + 1];
        vis[b] = __funnelshift_r(lo, hi, shf) & 0xffu;
        if (vis[b] != 0u && d.use_sensor) {
          const float4 *src = reinterpret_cast<const float4 *>(world + (size_t)(ox_map + ai - 4) * d.N + gyb);
#pragma unroll
          for (int j = 0; j < NV; ++j)
            if (grp_in[j]) v[b][j] = __ldcs(src + j);  // read once per CTA: streaming load, 3-5 % over __ldg
        }
      }
    }
#pragma unroll
    for (int b = 0; b < BATCH; ++b) {
      if (base + b >= nact) break;  // warp-uniform
      float x[4 * NV];
#pragma unroll
      for (int j = 0; j < NV; ++j) x[4 * j] = v[b][j].x, x[4 * j + 1] = v[b][j].y, x[4 * j + 2] = v[b][j].z, x[4 * j + 3] = v[b][j].w;
      float c[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float unseen = (dense && ((inmap >> i) & 1u)) ? fill : 0.f;
        c[i] = ((vis[b] >> i) & 1u) ? x[S + i] : unseen;
      }
      if (VOR) {
        // the robot's own Voronoi cell on this row: fp64 moments in cell units, as vwindow.cuh
        const int ai = sh.act[base + b];
        const unsigned ob = __funnelshift_r(vor->own[ai * CNN8_WORDS + wi], vor->own[ai * CNN8_WORDS + wi + 1], shf) & 0xffu;
        if (__any_sync(0xffffffffu, ob != 0u)) {
          const double xd = (double)(ox_map + ai - 4) + 0.5;
          double s = 0.0;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const double dv = (double)(((ob >> i) & 1u) ? c[i] : 0.f);
            acc.cs[i] += dv;
            s += dv;
          }
          acc.sx = fma(xd, s, acc.sx);
          acc.sxx = fma(xd * xd, s, acc.sxx);
        }
      }
      // taps q = 0..15 are window columns 8 lane - 4 + q: the left neighbour's c[4..7], c[0..7], the
      // right neighbour's c[0..3]; beyond the window (lanes 0 / 31) the padding is zero
      float tmp = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float l = __shfl_up_sync(0xffffffffu, c[4 + q], 1);
        if (lane == 0) l = 0.f;
        tmp = fmaf(VOR ? vor->wt[q][lane] : wc[q], l, tmp);
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) tmp = fmaf(VOR ? vor->wt[4 + q][lane] : wc[4 + q], c[q], tmp);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float r = __shfl_down_sync(0xffffffffu, c[q], 1);
        if (lane == 31) r = 0.f;
        tmp = fmaf(VOR ? vor->wt[12 + q][lane] : wc[12 + q], r, tmp);
      }
      sh.tmp[sh.act[base + b]][lane] = tmp;
    }
  }
}

template <bool VOR, bool DENSE>
__global__ void __launch_bounds__(CNN8_THREADS, 3)
k_cnn_fused8(EnvDev d, const float *__restrict__ wfull /* [32][16] */, float *__restrict__ out,
             size_t robot_stride, size_t local_off, size_t obst_off, double *__restrict__ wsums /* VOR: [E][R][4] */) {
  extern __shared__ __align__(16) unsigned char cnn8_smem[];  // CNN8_SMEM_BYTES (opt-in: above 48 KB)
  Cnn8Shared &sh = *reinterpret_cast<Cnn8Shared *>(cnn8_smem);
  Cnn8Vor *vor = VOR ? reinterpret_cast<Cnn8Vor *>(cnn8_smem + sizeof(Cnn8Shared)) : nullptr;
  double *s_px = reinterpret_cast<double *>(cnn8_smem + sizeof(Cnn8Shared) + sizeof(Cnn8Vor));
  double *s_py = s_px + d.R;
  const int r = blockIdx.x, e = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t er = (size_t)e * d.R + r;
  const size_t nn = cell_count(d);
  for (int t = tid; t < 32 * 16; t += CNN8_THREADS) {
    const float wv = wfull[t];
    sh.w[t >> 4][t & 15] = wv;
    if (VOR) vor->wt[t & 15][t >> 4] = wv;
  }
  if (tid < (CNN8_ROWS + 31) / 32 + 1) sh.flags[tid] = 0u;
  if (tid == 0) sh.nact = 0, sh.mask[CNN8_ROWS * CNN8_WORDS] = 0u;
  const double *p = d.pos_g + er * 2;
  const Window w = make_window(d.res, p[0], p[1], 256, d.N);
  const float *world = d.world + (size_t)e * nn;
  const uint32_t *plane = robot_plane(d, d.seenb, e, r);
  constexpr bool dense = DENSE;  // unseen cells carry pUnknownImportance: every in-map row counts
  // window rows / mask words that can hold a non-zero cell: inside the map and (unless unseen cells carry a
  // value) inside the robot's seen box
  int alo = max(0, -w.ox), ahi = min(256, d.N - w.ox);
  int wlo = 0, whi = CNN8_WORDS;
  if (!dense) {
    const int *box = d.seen_box + 4 * er;
    alo = max(alo, box[0] - w.ox), ahi = min(ahi, box[1] - w.ox);
    wlo = max(0, (box[2] >> 5) - (w.oy >> 5)), whi = min(CNN8_WORDS, ((box[3] - 1) >> 5) - (w.oy >> 5) + 1);
  }
  if (VOR) {
    if (tid < 8) vor->exact[tid] = 0u;
    if (tid == 0) vor->own[CNN8_ROWS * CNN8_WORDS] = 0u;
    if (warp == 0) {
      // sites: self, then the neighbours within communication range in ascending index
      // (decentralized_cvt.h:100-120), exactly as k_voronoi_window
      const double *pos = d.pos_g + (size_t)e * d.R * 2;
      const double sx = p[0], sy = p[1];
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
        const unsigned km = __ballot_sync(0xffffffffu, keep);
        if (keep) {
          const int slot = n + __popc(km & ((1u << lane) - 1u));
          s_px[slot] = jx, s_py[slot] = jy;
        }
        n += __popc(km);
      }
      if (lane == 0) vor->K = n - 1;
    }
  }
  __syncthreads();
  if (VOR) {
    // own-cell interval of every window row (thread a <-> window row a), as k_voronoi_window
    const int a = tid, row = w.ox + a, ncols = w.y1 - w.y0;
    int lo = 0, hi = 0;
    if (row >= w.x0 && row < w.x1 && ncols > 0 && a >= alo && a < ahi) {
      if (vor->K == 0) lo = 0, hi = ncols;
      else if (vw_row_interval(d.res, row, w.y0, ncols, vor->K, s_px, s_py, lo, hi)) atomicOr(&vor->exact[a >> 5], 1u << (a & 31));
    }
    vor->lo[a] = w.y0 + lo, vor->hi[a] = w.y0 + max(hi, lo);
    __syncthreads();
  }

  // ---- 1. mask words, clipped to the window's map columns [cy0, cy1).  Only the words inside the robot's
  //         seen box can hold a bit (EnvDev::seen_box, kept by k_mark_windows): a few steps into an episode
  //         that is a sixth of the window's words.  Batches of CNN8_STAGE loads are issued before use.
  const int wbase = w.oy >> 5;  // arithmetic shift: floor for negative origins
  const int cy0 = max(w.oy, 0), cy1 = min(w.oy + 256, d.N);
  const int nrow_s = max(ahi - alo, 0), nw_s = max(whi - wlo, 0);
  const int nstage = nrow_s * nw_s;
  // the words of those rows outside the box hold nothing: zeros without a load
  for (int t = tid; t < nrow_s * CNN8_WORDS; t += CNN8_THREADS) {
    const int ar = t / CNN8_WORDS, j = t - ar * CNN8_WORDS;
    if (j >= wlo && j < whi) continue;
    sh.mask[(alo + ar + 4) * CNN8_WORDS + j] = 0u;
    if (VOR) vor->own[(alo + ar + 4) * CNN8_WORDS + j] = 0u;
  }
  constexpr int CNN8_STAGE = 4;
#pragma unroll 1
  for (int t0 = 0; t0 < nstage; t0 += CNN8_STAGE * CNN8_THREADS) {
    uint32_t mw[CNN8_STAGE], kw[CNN8_STAGE];
#pragma unroll
    for (int it = 0; it < CNN8_STAGE; ++it) {
      const int t = t0 + tid + it * CNN8_THREADS;
      const int ar = t / nw_s, j = wlo + t - ar * nw_s;
      const int a = alo + ar, wd = wbase + j;
      uint32_t m = 0u, keep = 0u;
      if (t < nstage && wd >= 0 && wd < d.NW) {
        keep = vw_range_bits(cy0, cy1, wd << 5);
        if (keep) m = __ldg(plane + (size_t)(w.ox + a) * d.NW + wd) & keep;
      }
      mw[it] = m, kw[it] = keep;
    }
#pragma unroll
    for (int it = 0; it < CNN8_STAGE; ++it) {
      const int t = t0 + tid + it * CNN8_THREADS;
      if (t >= nstage) continue;
      const int ar = t / nw_s, j = wlo + t - ar * nw_s;
      const int a = alo + ar, ai = a + 4;
      sh.mask[ai * CNN8_WORDS + j] = mw[it];
      if (mw[it] || (dense && kw[it])) atomicOr(&sh.flags[ai >> 5], 1u << (ai & 31));
      if (VOR) {
        uint32_t o = kw[it] ? vw_range_bits(vor->lo[a], vor->hi[a], (wbase + j) << 5) : 0u;
        uint32_t todo = dense ? o : (o & mw[it]);
        if (todo && ((vor->exact[a >> 5] >> (a & 31)) & 1u)) {
          // per-cell rule on this row: drop the cells where some neighbour is strictly nearer
          while (todo) {
            const int bit = __ffs(todo) - 1;
            todo &= todo - 1;
            const int col = ((wbase + j) << 5) + bit;
            bool lost = false;
            for (int k = 1; k <= vor->K && !lost; ++k) lost = vw_beats(d.res, w.ox + a, col, s_px[k], s_py[k], s_px[0], s_py[0]);
            if (lost) o &= ~(1u << bit);
          }
        }
        vor->own[ai * CNN8_WORDS + j] = o;
      }
    }
  }
  __syncthreads();
  // rows with work, in ascending order (a fixed order keeps the fp64 sums reproducible run to run)
  for (int ai = tid; ai < CNN8_ROWS; ai += CNN8_THREADS) {
    int before = 0;
#pragma unroll
    for (int wq = 0; wq < (CNN8_ROWS + 31) / 32; ++wq) {
      const uint32_t f = sh.flags[wq];
      before += wq < (ai >> 5) ? __popc(f) : (wq == (ai >> 5) ? __popc(f & ((1u << (ai & 31)) - 1u)) : 0);
    }
    const bool on = (sh.flags[ai >> 5] >> (ai & 31)) & 1u;
    if (on) sh.act[before] = ai;
    if (ai == CNN8_ROWS - 1) sh.nact = before + (on ? 1 : 0);
  }
  __syncthreads();

  // ---- 2. width pass over the listed rows ----------------------------------------------------------
  float wc[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) wc[q] = sh.w[lane][q];
  VwAcc acc;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc.cs[i] = 0.0;
  acc.sx = acc.sxx = 0.0;
  switch (w.oy & 3) {
    case 0: cnn8_width_pass<0, VOR, DENSE>(d, sh, world, w.ox, w.oy, wbase, wc, vor, acc); break;
    case 1: cnn8_width_pass<1, VOR, DENSE>(d, sh, world, w.ox, w.oy, wbase, wc, vor, acc); break;
    case 2: cnn8_width_pass<2, VOR, DENSE>(d, sh, world, w.ox, w.oy, wbase, wc, vor, acc); break;
    default: cnn8_width_pass<3, VOR, DENSE>(d, sh, world, w.ox, w.oy, wbase, wc, vor, acc); break;
  }
  if (VOR) {
    // fold with the column coordinates; one deterministic CTA reduction below
    double m = 0.0, smy = 0.0, s2 = acc.sxx;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const double y = (double)(w.oy + 8 * lane + i) + 0.5;
      m += acc.cs[i];
      smy = fma(y, acc.cs[i], smy);
      s2 = fma(y * y, acc.cs[i], s2);
    }
    const Sum4 tot = reduce4(m, acc.sx, smy, s2);
    if (lane == 0) vor->red[warp][0] = tot.m, vor->red[warp][1] = tot.sx, vor->red[warp][2] = tot.sy, vor->red[warp][3] = tot.s2;
  }

  // width pass of the obstacle indicator for this lane's output column
  float oy_out = 0.f, oy_all = 0.f;
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    const int gy = w.oy + 8 * lane - 4 + q;
    const bool in_y = gy >= 0 && gy < d.N;
    oy_all = fmaf(wc[q], 1.f, oy_all);
    oy_out = fmaf(wc[q], in_y ? 0.f : 1.f, oy_out);
  }
  __syncthreads();
  if (VOR && tid < 4) {
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < CNN8_WARPS; ++k) t += vor->red[k][tid];
    wsums[er * 4 + tid] = t;
  }

  // ---- 3. height pass: output rows 4 warp .. 4 warp + 3, taps in ascending row order --------------
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ox = warp * 4 + i;
    float acc = 0.f, acc_obs = 0.f;
    const int ai0 = 8 * ox;  // a + 4 of tap 0
    // the 16 row flags of this output row: bits ai0 .. ai0 + 15
    const uint32_t fl = __funnelshift_r(sh.flags[ai0 >> 5], sh.flags[(ai0 >> 5) + 1], ai0 & 31) & 0xffffu;
    // taps whose window row is inside the window and the map: a in [alo, ahi)
    const int a0 = ai0 - 4;
    const int alo = max(0, -w.ox), ahi = min(256, d.N - w.ox);
    const int tl = min(max(alo - a0, 0), 16), th = min(max(ahi - a0, 0), 16);
    const uint32_t inx = th > tl ? (((1u << th) - 1u) & ~((1u << tl) - 1u)) : 0u;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const float wr = sh.w[ox][k];  // zero outside torch's tap range
      acc_obs = fmaf(wr, ((inx >> k) & 1u) ? oy_out : oy_all, acc_obs);
      if ((fl >> k) & 1u) acc = fmaf(wr, sh.tmp[ai0 + k][lane], acc);
    }
    float *dst = out + er * robot_stride + (size_t)ox * 32 + lane;
    dst[local_off] = acc;
    dst[obst_off] = acc_obs;
  }
}


// ---- item-ordered fused observation kernel (L = 256 -> C = 32, N % 4 == 0): the product path --------------------
// k_cnn_fused8 above deals whole window ROWS to warps, a lane per output column: a few steps into an episode a robot
// has seen ~70 x 70 cells of its 256 x 256 window, so three lanes out of four carry zeros through every select, FMA
// and shuffle (ncu r02b: 24.7 K warp instructions per robot, issue-bound at 56 % issue-active, 12 % of HBM).  Here
// the unit of work is an ITEM = one non-zero width-pass output (window row a, output column oy): the 16 taps
// [8 oy - 4, 8 oy + 12) of row a hold a seen cell.  Items are listed in (row, oy) order and dealt to the lanes, so
// every lane of every warp does useful work whatever the shape of the seen region.
// What bounds such a kernel is the chain of dependent memory latencies of ONE robot (position -> window and seen box
// -> mask words -> map values), not its arithmetic (ncu r02d: 15 K warp instructions, 3 resident CTAs of 256 threads
// per SM, 24 % warps active): the CTA is therefore small -- 128 threads, ~24 KB of shared memory, 6+ per SM -- and
// walks the window in four CHUNKS of 8 output rows (64 window rows + 12 rows of tap halo); chunks without a seen
// row cost nothing.  Per chunk:
//   1. seen bits of the chunk's rows inside the robot's seen box -> shared memory (11 words per row: a pad word
//      either side so that any 20-bit field can be fetched with one funnel shift);
//   2. one thread per row: which oy have a seen tap (nibble arithmetic on the mask words), prefix sum over the
//      rows, item list (row << 5 | oy);
//   3. items: a lane loads the 4 or 5 aligned float4 groups under its 16 taps (only those with a seen bit), selects
//      the seen cells, applies the 16 weights in ascending tap order (the order torchvision's separable antialias
//      kernel uses) and stores tmp[row][oy].  VOR: the 8 cells of the item's own column group (taps 4..11: every
//      window cell is the own group of exactly one item of the chunk that owns its row) are also reduced into the
//      robot's window-Voronoi moments: fp64 (s, t = sum i v_i, u = sum i^2 v_i) per item, folded with the group's
//      first column and the row into five lane-private accumulators; fixed item -> lane assignment, fixed
//      reduction tree: bit-stable sums;
//   4. height pass: warp w owns two of the chunk's output rows, sums the taps in ascending row order.
// Around the chunks: the own-cell column interval of every window row (VOR; vw_row_interval, as k_voronoi_window),
// the obstacle channel in closed form -- obstacle(a, b) = 1 - in_x(a) in_y(b) is separable, so its resized value
// is A(ox) cy_out(oy) + B(ox) cy_all(oy) with A / B = the tap weights of the rows inside / outside the map and
// cy_out / cy_all = the tap weights of the columns outside the map / all columns (exactly 0 when the window lies
// inside the map) -- and (use_comm) the two communication-map channels by one warp: a lane per neighbour, cells hit
// by several neighbours summed in ascending neighbour order.
// wfull[o][k] = weight of input index 8 o - 4 + k (zero outside torch's tap range); rows 1 .. 30 are identical
// (the host checks this before it picks the kernel): their 16 weights live in registers.
// the 16 tap weights of the interior output columns (rows 1 .. 30 of wfull, identical): FFMA reads them straight from
// the constant bank, no register each (set once per device by the host when the resize table is built)
__constant__ float c_ob_w[16];

constexpr int OB_THREADS = 128;
constexpr int OB_WARPS = OB_THREADS / 32;
constexpr int OB_CHUNKS = 4;                  // 8 output rows = 64 window rows each
constexpr int OB_CROWS = 64 + 12;             // chunk rows incl. the tap halo: window rows 64 c - 4 .. 64 c + 71
constexpr int OB_WORDS = 11;                  // pad + 9 mask words + pad
constexpr int OB_CAP = OB_CROWS * 32;         // items of a chunk at most

struct ObsShared {
  float tmp[OB_CROWS][32];              // width-pass results; entry (row, oy) valid iff need[row] bit oy
  uint32_t mask[OB_CROWS * OB_WORDS];   // seen & in-map & in-window bits; word 1 + j <-> map word wbase + j
  uint32_t need[OB_CROWS];
  unsigned short items[OB_CAP];
  float w[32][16];
  float oA[32], oB[32];
  int lo[256], hi[256];                 // VOR: own-cell interval of every window row, map columns
  uint32_t exact[8];                    // VOR: rows that need the per-cell rule
  int wtot[OB_WARPS];
  int total;
  int K;
  double red[OB_WARPS][4];
};
__host__ __device__ inline size_t obs_smem_bytes(int R) { return sizeof(ObsShared) + sizeof(double) * 2 * (size_t)R; }

// bits [lo, hi) of an n-bit field whose bit 0 is column c0
__device__ __forceinline__ unsigned ob_range(int lo, int hi, int c0, int n) {
  const int a = min(max(lo - c0, 0), n), b = min(max(hi - c0, 0), n);
  return ((1u << b) - 1u) & ~((1u << a) - 1u);
}

// per-cell rule on a row a bisector runs along (rare: kept out of line, the item loop is instantiated four times)
__device__ __noinline__ unsigned ob_exact_rule(double res, int row, int c0, unsigned ob, int K, const double *s_px, const double *s_py) {
  unsigned todo = ob;
  while (todo) {
    const int bit = __ffs(todo) - 1;
    todo &= todo - 1;
    bool lost = false;
    for (int k = 1; k <= K && !lost; ++k) lost = vw_beats(res, row, c0 + bit, s_px[k], s_py[k], s_px[0], s_py[0]);
    if (lost) ob &= ~(1u << bit);
  }
  return ob;
}

// items of one chunk; rbase = window row of the chunk's row 0, [own0, own1) = window rows whose moments this chunk adds
template <int S, bool VOR, bool DENSE>
__device__ __forceinline__ void ob_items(const EnvDev &d, ObsShared &sh, const float *__restrict__ world, const Window &w,
                                         int off0, int rbase, int own0, int own1, int amap0, int amap1, int cy0, int cy1,
                                         const double *s_px, const double *s_py, double (&acc)[5]) {
  constexpr int NV = S == 0 ? 4 : 5;  // aligned 4-cell groups under the 16 taps
  const int total = sh.total;
  const float fill = d.fill;
#pragma unroll 1
  for (int i = threadIdx.x; i < total; i += OB_THREADS) {
    const int it = sh.items[i];
    const int cr = it >> 5, oy = it & 31;         // chunk row, output column
    const int a = rbase + cr, row = w.ox + a;
    const int gyb = w.oy + 8 * oy - 4 - S;        // 16-byte aligned map column of the first group
    const int pos = off0 + 8 * oy - 4 - S + 32;   // its bit in the row's 11 mask words
    const uint32_t *m = sh.mask + cr * OB_WORDS + (pos >> 5);
    const unsigned vis = __funnelshift_r(m[0], m[1], pos & 31) & ((1u << (4 * NV)) - 1u);
    float4 v[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (d.use_sensor) {
      const float4 *src = reinterpret_cast<const float4 *>(world + (size_t)row * d.N + gyb);
#pragma unroll
      for (int j = 0; j < NV; ++j)
        if ((vis >> (4 * j)) & 0xfu) v[j] = __ldcs(src + j);  // a set bit is an in-map cell: the whole aligned group is
    }
    float x[4 * NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) x[4 * j] = v[j].x, x[4 * j + 1] = v[j].y, x[4 * j + 2] = v[j].z, x[4 * j + 3] = v[j].w;
    unsigned inm = 0u;  // DENSE: in-map, in-window cells of an in-map row carry `fill` until seen
    if (DENSE) inm = (a >= amap0 && a < amap1) ? ob_range(cy0, cy1, gyb, 4 * NV) : 0u;
    float c[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const float unseen = (DENSE && ((inm >> (S + q)) & 1u)) ? fill : 0.f;
      c[q] = ((vis >> (S + q)) & 1u) ? x[S + q] : unseen;
    }
    float t = 0.f;
    if (oy != 0 && oy != 31) {
#pragma unroll
      for (int q = 0; q < 16; ++q) t = fmaf(c_ob_w[q], c[q], t);
    } else {  // the two border columns: torch truncates and renormalises their taps
#pragma unroll
      for (int q = 0; q < 16; ++q) t = fmaf(sh.w[oy][q], c[q], t);
    }
    sh.tmp[cr][oy] = t;
    if (VOR && a >= own0 && a < own1) {
      // the item's own column group (taps 4 .. 11 = window columns 8 oy .. 8 oy + 7) inside the robot's cell
      const int c0 = w.oy + 8 * oy;
      unsigned ob = ob_range(sh.lo[a], sh.hi[a], c0, 8) & ((DENSE ? inm : vis) >> (S + 4)) & 0xffu;
      if (ob != 0u) {
        if ((sh.exact[a >> 5] >> (a & 31)) & 1u) ob = ob_exact_rule(d.res, row, c0, ob, sh.K, s_px, s_py);
        // (s, t, u) = sum (1, q, q^2) v_q over the group's own cells: pairwise trees, three short dependency chains
        double dv[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) dv[q] = (double)(((ob >> q) & 1u) ? c[4 + q] : 0.f);
        const double s = ((dv[0] + dv[1]) + (dv[2] + dv[3])) + ((dv[4] + dv[5]) + (dv[6] + dv[7]));
        const double tt = (fma(3.0, dv[3], fma(2.0, dv[2], dv[1]))) + (fma(7.0, dv[7], fma(6.0, dv[6], fma(5.0, dv[5], 4.0 * dv[4]))));
        const double u = (fma(9.0, dv[3], fma(4.0, dv[2], dv[1]))) + (fma(49.0, dv[7], fma(36.0, dv[6], fma(25.0, dv[5], 16.0 * dv[4]))));
        const double y0 = (double)c0 + 0.5, xd = (double)row + 0.5;
        acc[0] += s;
        acc[1] = fma(xd, s, acc[1]);
        acc[2] += fma(y0, s, tt);
        acc[3] = fma(xd * xd, s, acc[3]);
        acc[4] += fma(y0 * y0, s, fma(y0 + y0, tt, u));
      }
    }
  }
}

// the two communication-map planes of robot r (already zeroed), by ONE warp: a lane per other robot, 32 at a time;
// a cell hit by several robots of a group of 32 receives their sum in ascending index (what a sequential loop adds)
__device__ __forceinline__ void comm_scatter_warp(const EnvDev &d, int e, int r, int C, float *ox, float *oy) {
  const int lane = threadIdx.x & 31;
  const double *pos = d.pos_g + (size_t)e * d.R * 2;
  const float denom = (float)(d.comm_range * d.res * 2.0);
  const float center = (float)(C / 2.0 - d.res / 2.0);
  const float cr = (float)d.comm_range;
  const double sx = pos[2 * r], sy = pos[2 * r + 1];
  for (int base = 0; base < d.R; base += 32) {
    const int j = base + lane;
    int idx = -1;
    float vx = 0.f, vy = 0.f;
    if (j < d.R && j != r) {
      // coverage_system.cpp:274-291: relative position p_j - p_r (stored negated for j < r)
      double dx, dy;
      if (j > r) dx = pos[2 * j] - sx, dy = pos[2 * j + 1] - sy;
      else dx = -(sx - pos[2 * j]), dy = -(sy - pos[2 * j + 1]);
      if (sqrt(dx * dx + dy * dy) < d.comm_range) {
        const float fx = (float)dx, fy = (float)dy;
        const float ix = rintf(__fadd_rn(__fdiv_rn(__fmul_rn(fx, (float)C), denom), center));
        const float iy = rintf(__fadd_rn(__fdiv_rn(__fmul_rn(fy, (float)C), denom), center));
        const int xi = (int)ix, yi = (int)iy;
        if (xi >= 0 && xi < C && yi >= 0 && yi < C) idx = xi * C + yi, vx = __fdiv_rn(fx, cr), vy = __fdiv_rn(fy, cr);
      }
    }
    const unsigned any = __ballot_sync(0xffffffffu, idx >= 0);
    if (any == 0u) continue;  // warp-uniform
    // ordered sum over the lanes that hit my cell; the lowest of them writes
    float ax = 0.f, ay = 0.f;
    bool first = true;
    for (unsigned todo = any; todo; todo &= todo - 1) {
      const int k = __ffs(todo) - 1;
      const int ik = __shfl_sync(0xffffffffu, idx, k);
      const float xk = __shfl_sync(0xffffffffu, vx, k), yk = __shfl_sync(0xffffffffu, vy, k);
      if (ik == idx) {
        if (k < lane) first = false;
        ax = __fadd_rn(ax, xk), ay = __fadd_rn(ay, yk);
      }
    }
    if (idx >= 0 && first) ox[idx] = __fadd_rn(ox[idx], ax), oy[idx] = __fadd_rn(oy[idx], ay);
    __syncwarp();
  }
}

// The item loop is instantiated once per 16-byte misalignment S = (window origin column) & 3, so a CTA runs one of four
// code paths; CTAs resident on an SM at the same time share its instruction cache.  These two kernels order the CTAs by
// S (robots of one class are launched next to each other): cls[e R + r] = S | slot << 2, then order[base(S) + slot].
__global__ void k_obs_classify(EnvDev d, int *__restrict__ cls, int *__restrict__ counts /* [4], zeroed */) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d.E * d.R) return;
  const Window w = make_window(d.res, d.pos_g[2 * (size_t)i], d.pos_g[2 * (size_t)i + 1], 256, d.N);
  const int S = w.oy & 3;
  cls[i] = S | (atomicAdd(counts + S, 1) << 2);
}
__global__ void k_obs_order(int n, const int *__restrict__ cls, const int *__restrict__ counts, int *__restrict__ order) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = cls[i], S = c & 3;
  int base = 0;
  for (int k = 0; k < S; ++k) base += counts[k];
  order[base + (c >> 2)] = i;
}

template <bool VOR, bool DENSE>
__global__ void __launch_bounds__(OB_THREADS, 8)
k_obs_fused(EnvDev d, const float *__restrict__ wfull /* [32][16] */, float *__restrict__ out, size_t robot_stride,
            size_t local_off, size_t obst_off, size_t comm_off, int use_comm, double *__restrict__ wsums /* VOR: [E][R][4] */,
            const int *__restrict__ order /* optional: CTA -> e R + r, robots of one window misalignment together */) {
  extern __shared__ __align__(16) unsigned char ob_smem[];
  ObsShared &sh = *reinterpret_cast<ObsShared *>(ob_smem);
  double *s_px = reinterpret_cast<double *>(ob_smem + sizeof(ObsShared));
  double *s_py = s_px + d.R;
  const int cta = blockIdx.y * gridDim.x + blockIdx.x;
  const int er_i = order ? __ldg(order + cta) : cta;
  const int r = er_i % d.R, e = er_i / d.R;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t er = (size_t)e * d.R + r;
  float *dst = out + er * robot_stride;
  for (int t = tid; t < 32 * 16; t += OB_THREADS) sh.w[t >> 4][t & 15] = wfull[t];
  if (use_comm) {  // the communication channels start as zeros
    float4 *cz = reinterpret_cast<float4 *>(dst + comm_off);
    for (int t = tid; t < 2 * 32 * 32 / 4; t += OB_THREADS) cz[t] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const double *p = d.pos_g + er * 2;
  const Window w = make_window(d.res, p[0], p[1], 256, d.N);
  const float *world = d.world + (size_t)e * cell_count(d);
  const uint32_t *plane = robot_plane(d, d.seenb, e, r);
  // window rows / columns inside the map, and (unless unseen cells carry a value) inside the robot's seen box
  const int amap0 = max(0, -w.ox), amap1 = min(256, d.N - w.ox);
  const int cy0 = max(w.oy, 0), cy1 = min(w.oy + 256, d.N);
  const int wbase = w.oy >> 5;  // arithmetic shift: floor for negative origins
  const int off0 = w.oy - 32 * wbase;
  int alo = amap0, ahi = amap1, wlo = 0, whi = 9;
  if (!DENSE) {
    const int *box = d.seen_box + 4 * er;
    alo = max(alo, box[0] - w.ox), ahi = min(ahi, box[1] - w.ox);
    wlo = max(0, (box[2] >> 5) - wbase), whi = min(9, ((box[3] - 1) >> 5) - wbase + 1);
  }
  const int nw_s = max(whi - wlo, 0);
  if (d.obs_prefetch) {
    // The map values and mask words under the seen box go DRAM -> L2 while the masks are staged and the items
    // listed: the tap loads of the width pass then wait an L2 hit, not a DRAM access.
    int c0 = cy0, c1 = cy1;
    if (!DENSE) c0 = max(c0, d.seen_box[4 * er + 2]), c1 = min(c1, d.seen_box[4 * er + 3]);
    const int nrow = ahi - alo, per = c1 > c0 ? ((c1 - c0 + 31) >> 5) + 1 : 0;
    for (int t = tid; t < nrow * per; t += OB_THREADS) {
      const int ar = t / per, k = t - ar * per;
      const float *rowp = world + (size_t)(w.ox + alo + ar) * d.N;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(rowp + min(c0 + 32 * k, c1 - 1)));
    }
    if (!DENSE && per)
      for (int a = alo + tid; a < ahi; a += OB_THREADS)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(plane + (size_t)(w.ox + a) * d.NW + max(wbase + wlo, 0)));
  }
  if (VOR) {
    if (tid < 8) sh.exact[tid] = 0u;
    if (warp == 0) {
      // sites: self, then the neighbours within communication range in ascending index (decentralized_cvt.h:100-120)
      const double *pos = d.pos_g + (size_t)e * d.R * 2;
      const double sx = p[0], sy = p[1];
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
        const unsigned km = __ballot_sync(0xffffffffu, keep);
        if (keep) {
          const int slot = n + __popc(km & ((1u << lane) - 1u));
          s_px[slot] = jx, s_py[slot] = jy;
        }
        n += __popc(km);
      }
      if (lane == 0) sh.K = n - 1;
    }
  }
  __syncthreads();  // weights, neighbour list, zeroed communication planes
  if (use_comm && warp == OB_WARPS - 1) comm_scatter_warp(d, e, r, 32, dst + comm_off, dst + comm_off + 32 * 32);
  if (VOR) {  // own-cell interval of the rows that can hold a cell (the warp that scattered the neighbours sits out)
    const int ncols = w.y1 - w.y0;
    const int nthr = use_comm ? OB_THREADS - 32 : OB_THREADS;
    for (int a = alo + tid; a < ahi && tid < nthr; a += nthr) {
      int lo = 0, hi = 0;
      if (ncols > 0) {
        if (sh.K == 0) lo = 0, hi = ncols;
        else if (vw_row_interval(d.res, w.ox + a, w.y0, ncols, sh.K, s_px, s_py, lo, hi)) atomicOr(&sh.exact[a >> 5], 1u << (a & 31));
      }
      sh.lo[a] = w.y0 + lo, sh.hi[a] = w.y0 + max(hi, lo);
    }
  }
  // obstacle channel: A = weight of the tap rows inside the map, B = of those outside it (per output row);
  // cy_out / cy_all = weight of the tap columns outside the map / of all of them (per output column = lane)
  const bool inside = w.ox >= 0 && w.ox + 256 <= d.N && w.oy >= 0 && w.oy + 256 <= d.N;
  if (tid < 32) {
    float A = 0.f, B = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int a = 8 * tid - 4 + k;
      const float wk = sh.w[tid][k];  // zero outside torch's tap range (in particular outside the window)
      if (a >= amap0 && a < amap1) A += wk;
      else B += wk;
    }
    sh.oA[tid] = A, sh.oB[tid] = B;
  }
  float cy_out = 0.f, cy_all = 0.f;
  if (!inside) {
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const int gy = w.oy + 8 * lane - 4 + q;
      const float wq = sh.w[lane][q];
      cy_all += wq;
      if (!(gy >= 0 && gy < d.N)) cy_out += wq;
    }
  }
  double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  __syncthreads();  // oA / oB, the row intervals

#pragma unroll 1
  bool carry = false;  // the previous chunk ran: its last 12 rows (this chunk's tap halo) are kept, not recomputed
  for (int ch = 0; ch < OB_CHUNKS; ++ch) {
    const int rbase = 64 * ch - 4;                                   // window row of chunk row 0
    // window rows this chunk computes (every window row is computed by exactly one chunk)
    const int r0 = max(alo, rbase + (carry ? 12 : 0)), r1 = min(ahi, rbase + OB_CROWS);
    const int ox0 = 8 * ch;
    if (r1 <= r0 && !carry) {  // CTA-uniform: nothing seen under these 8 output rows
      carry = false;
      for (int t = tid; t < 8 * 32; t += OB_THREADS) {
        const int ox = ox0 + (t >> 5);  // (t & 31 == lane: the thread's own output column)
        float *o = dst + (size_t)ox * 32 + lane;
        o[local_off] = 0.f;
        o[obst_off] = inside ? 0.f : fmaf(sh.oA[ox], cy_out, sh.oB[ox] * cy_all);
      }
      continue;
    }
    __syncthreads();  // the previous chunk's height pass is done with tmp / need / mask
    if (carry) {  // chunk rows 64 .. 75 of the previous chunk are rows 0 .. 11 of this one
      for (int t = tid; t < 12 * 32; t += OB_THREADS) sh.tmp[t >> 5][t & 31] = sh.tmp[64 + (t >> 5)][t & 31];
      if (tid < 12) sh.need[tid] = sh.need[64 + tid];
    }
    // ---- 1. mask words of rows r0 .. r1 - 1 (chunk rows r0 - rbase ..); pad words and words outside the box are zero
    {
      const int nrow = max(r1 - r0, 0);
      for (int t = tid; t < nrow * OB_WORDS; t += OB_THREADS) {
        const int ar = t / OB_WORDS, j = t - ar * OB_WORDS;  // j = 0 / 10: pads; 1 + wlo .. whi: loaded below
        if (j < 1 + wlo || j >= 1 + whi) sh.mask[(r0 - rbase + ar) * OB_WORDS + j] = 0u;
      }
      const int nstage = nrow * nw_s;
      const float inv_nw = 1.f / (float)max(nw_s, 1);  // t / nw_s for t < 76 * 9 through a float product: (t + 0.5) / nw_s
      constexpr int STAGE = 4;                          // is at least 0.05 away from an integer, the product is good to 1e-4
#pragma unroll 1
      for (int t0 = 0; t0 < nstage; t0 += STAGE * OB_THREADS) {
        uint32_t mw[STAGE];
#pragma unroll
        for (int k = 0; k < STAGE; ++k) {
          const int t = t0 + tid + k * OB_THREADS;
          const int ar = (int)(((float)t + 0.5f) * inv_nw), j = wlo + t - ar * nw_s;
          const int wd = wbase + j;
          uint32_t mv = 0u;
          if (t < nstage && wd >= 0 && wd < d.NW) {
            const uint32_t keep = vw_range_bits(cy0, cy1, wd << 5);
            if (keep) mv = __ldg(plane + (size_t)(w.ox + r0 + ar) * d.NW + wd) & keep;
          }
          mw[k] = mv;
        }
#pragma unroll
        for (int k = 0; k < STAGE; ++k) {
          const int t = t0 + tid + k * OB_THREADS;
          if (t >= nstage) continue;
          const int ar = (int)(((float)t + 0.5f) * inv_nw), j = wlo + t - ar * nw_s;
          sh.mask[(r0 - rbase + ar) * OB_WORDS + 1 + j] = mw[k];
        }
      }
    }
    __syncthreads();
    // ---- 2. which outputs of a row are non-zero; item list (thread <-> chunk row) ---------------------------
    {
      const int cr = tid, a = rbase + cr;
      uint32_t nd = 0u;
      if (cr < OB_CROWS && a >= r0 && a < r1) {
        if (DENSE) {
          // every output whose 16 taps touch an in-map column of this in-map row
          for (int oy = 0; oy < 32; ++oy)
            if (ob_range(cy0, cy1, w.oy + 8 * oy - 4, 16)) nd |= 1u << oy;
        } else {
          const uint32_t *mrow = sh.mask + cr * OB_WORDS + 1;
          uint32_t tprev = 0u, tcur, tnext;
          {
            const uint32_t x0 = __funnelshift_r(mrow[0], mrow[1], off0);
            uint32_t t = x0 | (x0 >> 1);
            tcur = (t | (t >> 2)) & 0x11111111u;  // bit 4 n: nibble n (4 window columns) holds a seen cell
          }
#pragma unroll
          for (int mI = 0; mI < 8; ++mI) {
            tnext = 0u;
            if (mI < 7) {
              const uint32_t x1 = __funnelshift_r(mrow[mI + 1], mrow[mI + 2], off0);
              uint32_t t = x1 | (x1 >> 1);
              tnext = (t | (t >> 2)) & 0x11111111u;
            }
            // output 4 mI + g reads nibbles 2 g - 1 .. 2 g + 2 of this word (and of its neighbours at the ends)
            const uint32_t u = tcur | (tcur >> 4) | (tcur << 4) | (tcur >> 8) | (tprev >> 28) | (tnext << 24);
            nd |= ((((u & 0x01010101u) * 0x01020408u) >> 24) & 0xfu) << (4 * mI);
            tprev = tcur, tcur = tnext;
          }
        }
      }
      if (cr < OB_CROWS && !(carry && cr < 12)) sh.need[cr] = nd;
      // exclusive prefix sum of the item counts over the chunk's rows
      const int cnt = __popc(nd);
      int incl = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
      }
      if (lane == 31) sh.wtot[warp] = incl;
      __syncthreads();
      int before = 0;
#pragma unroll
      for (int k = 0; k < OB_WARPS; ++k) before += k < warp ? sh.wtot[k] : 0;
      int off = before + incl - cnt;
      if (tid == OB_THREADS - 1) sh.total = off + cnt;
      while (nd) {
        const int oy = __ffs(nd) - 1;
        nd &= nd - 1;
        sh.items[off++] = (unsigned short)((cr << 5) | oy);
      }
    }
    __syncthreads();
    // ---- 3. items (every window row is computed, and its own-cell moments added, by exactly one chunk)
    {
      const int own0 = -4, own1 = 260;
      switch (w.oy & 3) {
        case 0: ob_items<0, VOR, DENSE>(d, sh, world, w, off0, rbase, own0, own1, amap0, amap1, cy0, cy1, s_px, s_py, acc); break;
        case 1: ob_items<1, VOR, DENSE>(d, sh, world, w, off0, rbase, own0, own1, amap0, amap1, cy0, cy1, s_px, s_py, acc); break;
        case 2: ob_items<2, VOR, DENSE>(d, sh, world, w, off0, rbase, own0, own1, amap0, amap1, cy0, cy1, s_px, s_py, acc); break;
        default: ob_items<3, VOR, DENSE>(d, sh, world, w, off0, rbase, own0, own1, amap0, amap1, cy0, cy1, s_px, s_py, acc); break;
      }
    }
    __syncthreads();
    // ---- 4. height pass: lane = output column, warp -> output rows ox0 + 2 warp, + 1 ---------------------------
#pragma unroll
    for (int i = 0; i < 8 / OB_WARPS; ++i) {
      const int oxl = warp * (8 / OB_WARPS) + i, ox = ox0 + oxl;
      const int c0r = 8 * oxl;  // chunk row of tap 0 (window row 8 ox - 4)
      const uint32_t ndl = sh.need[c0r + (lane & 15)];
      float acc_l = 0.f;
      if (__any_sync(0xffffffffu, ndl != 0u)) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const uint32_t nk = __shfl_sync(0xffffffffu, ndl, k);
          if ((nk >> lane) & 1u) acc_l = fmaf(sh.w[ox][k], sh.tmp[c0r + k][lane], acc_l);
        }
      }
      float *o = dst + (size_t)ox * 32 + lane;
      o[local_off] = acc_l;
      o[obst_off] = inside ? 0.f : fmaf(sh.oA[ox], cy_out, sh.oB[ox] * cy_all);
    }
    carry = r1 > r0;  // (no new row in this chunk: everything from its row 12 on is unseen, so is the next chunk's halo)
  }
  if (VOR) {
    const Sum4 tot = reduce4(acc[0], acc[1], acc[2], acc[3] + acc[4]);
    if (lane == 0) sh.red[warp][0] = tot.m, sh.red[warp][1] = tot.sx, sh.red[warp][2] = tot.sy, sh.red[warp][3] = tot.s2;
    __syncthreads();
    if (tid < 4) {
      double t = 0.0;
#pragma unroll
      for (int k = 0; k < OB_WARPS; ++k) t += sh.red[k][tid];
      wsums[er * 4 + tid] = t;
    }
  }
}

}  // namespace ccb
